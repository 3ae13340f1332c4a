#!/bin/bash
o=gpurun_out
bash tools/gpu_sweep.sh r2g default o2 o1
echo "== cell decode" >> $o/sweep_r2g.log
PIPLIB_B200_CELL_DECODE=1 timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 4 >> $o/sweep_r2g.log 2>&1
tail -2 $o/sweep_r2g.log
# class M: arena in shared memory vs global memory
for env in "" "PIPLIB_B200_NO_TEAM_SMEM=1"; do
  echo "== team $env" >> $o/team_r2g.log
  env $env PIPLIB_B200_TIMING=1 timeout 600 python tools/bench_configs.py vivien32 fimmel test12i --cpu-seconds 1 >> $o/team_r2g.log 2>> $o/team_r2g.err
done
grep -v "^\[" $o/team_r2g.log | cut -c1-700
grep "team arena\|class [1-9]" $o/team_r2g.err | head -40
timeout 300 python tools/bench_large.py 4096 3 --no-check --dense > $o/large_dense_r2g.log 2>&1; tail -1 $o/large_dense_r2g.log | cut -c1-900

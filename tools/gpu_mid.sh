#!/bin/bash
# mid-sized standalone batches: plain launch / heavy-problem hand-over (default) / donation for every problem
o=gpurun_out; tag=$1
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "device_job_in_parts or device_resident" > $o/pytest_mid_$tag.log 2>&1; tail -2 $o/pytest_mid_$tag.log
timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 5 2>&1 | cut -c1-110 >> $o/mid_$tag.log
for wl in sor1d fimmel test12i cg1 expansion loopnest8x12p2; do
  for n in 50000 200000; do
    echo "== $wl $n: plain / hand-over / donation" >> $o/mid_$tag.log
    PIPLIB_B200_HEAVY_PIVOTS=0 PIPLIB_B200_STEAL=0 timeout 200 python tools/ncu_run.py $wl $n 5 2>&1 | cut -c1-110 >> $o/mid_$tag.log
    timeout 200 python tools/ncu_run.py $wl $n 5 2>&1 | cut -c1-110 >> $o/mid_$tag.log
    PIPLIB_B200_STEAL=1 timeout 200 python tools/ncu_run.py $wl $n 5 2>&1 | cut -c1-110 >> $o/mid_$tag.log
  done
done
cat $o/mid_$tag.log

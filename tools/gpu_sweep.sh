#!/bin/bash
# kernel-only sweep over library variants (occupancy / slack / sub-solver builds): tools/gpu_sweep.sh <tag> <variants...>
tag=$1; shift
o=gpurun_out
for v in "$@"; do
  lib=$PWD/piplib_b200/lib/libpiplib_dp_$v.so
  [ "$v" = "default" ] && lib=$PWD/piplib_b200/lib/libpiplib_dp.so
  echo "== $v" >> $o/sweep_$tag.log
  PIPLIB_B200_LIB=$lib timeout 300 python tools/ncu_run.py loopnest16x24p3 1000000 4 >> $o/sweep_$tag.log 2>&1
done
cat $o/sweep_$tag.log

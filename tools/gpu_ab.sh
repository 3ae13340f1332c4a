#!/bin/bash
# A/B of library variants on the one-launch 10^6 device job: tools/gpu_ab.sh <tag> <variants...>
tag=$1; shift
o=gpurun_out
for v in "$@"; do
  lib=$PWD/piplib_b200/lib/libpiplib_dp_$v.so
  [ "$v" = "default" ] && lib=$PWD/piplib_b200/lib/libpiplib_dp.so
  echo "== $v" >> $o/ab_$tag.log
  PIPLIB_B200_LIB=$lib PIPLIB_B200_DEVICE_PARTS=1 timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 5 2>&1 | cut -c1-110 >> $o/ab_$tag.log
done
cat $o/ab_$tag.log

#!/bin/bash
# ncu --set full of the solve kernel for several library variants: tools/gpu_ncu3.sh <tag> <variants...>
tag=$1; shift
o=gpurun_out
for v in "$@"; do
  lib=$PWD/piplib_b200/lib/libpiplib_dp_$v.so
  [ "$v" = "default" ] && lib=$PWD/piplib_b200/lib/libpiplib_dp.so
  PIPLIB_B200_LIB=$lib timeout 300 python tools/ncu_run.py loopnest16x24p3 1000000 3 >> $o/sweep_$tag.log 2>&1
  PIPLIB_B200_LIB=$lib timeout 400 ncu --set full --clock-control none --import-source on -k regex:pip_solve_kernel -c 1 -o $o/prof_solve_${tag}_$v -f \
      python tools/ncu_run.py loopnest16x24p3 50000 1 > $o/ncu_solve_${tag}_$v.log 2>&1
done
cat $o/sweep_$tag.log
ls -la $o/*.ncu-rep

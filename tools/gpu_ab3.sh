#!/bin/bash
o=gpurun_out; tag=$1
bash tools/gpu_ab.sh $tag n1 s2 s1 s4 default n1 s2 s1 > /dev/null
for v in n1 s2 s1; do
  echo "== $v, 200000 problems" >> $o/ab_$tag.log
  PIPLIB_B200_LIB=$PWD/piplib_b200/lib/libpiplib_dp_$v.so timeout 200 python tools/ncu_run.py loopnest16x24p3 200000 5 2>&1 | cut -c1-110 >> $o/ab_$tag.log
  echo "== $v, three parts" >> $o/ab_$tag.log
  PIPLIB_B200_LIB=$PWD/piplib_b200/lib/libpiplib_dp_$v.so timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 5 2>&1 | cut -c1-110 >> $o/ab_$tag.log
done
cat $o/ab_$tag.log

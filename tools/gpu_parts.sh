#!/bin/bash
# device job in parts (tools/ncu_run.py, 10^6 loop-nest problems): shares of the parts
o=gpurun_out; tag=${1:-p}
for sh in 1 "0.49,0.32,0.19" "0.55,0.3,0.15" "0.6,0.28,0.12" "0.65,0.25,0.1" "0.7,0.22,0.08" "0.5,0.3,0.14,0.06" "0.6,0.25,0.1,0.05" "0.75,0.25" "0.85,0.15" 1; do
  echo "== shares $sh" >> $o/parts_$tag.log
  PIPLIB_B200_DEVICE_SHARES=$sh timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 6 2>&1 | cut -c1-110 >> $o/parts_$tag.log
done
cat $o/parts_$tag.log

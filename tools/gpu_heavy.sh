#!/bin/bash
# heavy-problem hand-over: kernel-only timings with the default rule, end-to-end per number of tail chunks
o=gpurun_out; tag=${1:-h}
for n in 50000 200000 1000000; do
  timeout 200 python tools/ncu_run.py loopnest16x24p3 $n 4 2>&1 | cut -c1-100 >> $o/heavy_$tag.log
done
for t in 0 3 6 0 3 6; do
  echo "== tail chunks $t" >> $o/heavy_$tag.log
  PIPLIB_B200_TAIL_CHUNKS=$t timeout 200 python tools/e2e_timing.py 1000000 pinned 2>&1 | grep "^e2e" | tail -3 >> $o/heavy_$tag.log
done
for t in 0 3; do
  echo "== 200000 problems, tail chunks $t" >> $o/heavy_$tag.log
  PIPLIB_B200_TAIL_CHUNKS=$t timeout 200 python tools/e2e_timing.py 200000 pinned 2>&1 | grep "^e2e" | tail -3 >> $o/heavy_$tag.log
done
cat $o/heavy_$tag.log

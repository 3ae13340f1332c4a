#!/bin/bash
o=gpurun_out; tag=$1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "handover or donation or dense_batch_vs_oracle or device_job or cli_suite or lib_suite" > $o/pytest_ab_$tag.log 2>&1; tail -2 $o/pytest_ab_$tag.log
bash tools/gpu_ab.sh $tag default u1 u4 scan32 default > /dev/null
cat $o/ab_$tag.log
echo "== donation for the whole 10^6 batch (PIPLIB_B200_STEAL=1)" >> $o/ab2_$tag.log
PIPLIB_B200_STEAL=1 PIPLIB_B200_DEVICE_PARTS=1 timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 4 2>&1 | cut -c1-110 >> $o/ab2_$tag.log
for n in 50000 200000; do
  echo "== n $n default / no hand-over / donation everywhere" >> $o/ab2_$tag.log
  timeout 200 python tools/ncu_run.py loopnest16x24p3 $n 5 2>&1 | cut -c1-110 >> $o/ab2_$tag.log
  PIPLIB_B200_HEAVY_PIVOTS=0 timeout 200 python tools/ncu_run.py loopnest16x24p3 $n 5 2>&1 | cut -c1-110 >> $o/ab2_$tag.log
  PIPLIB_B200_STEAL=1 timeout 200 python tools/ncu_run.py loopnest16x24p3 $n 5 2>&1 | cut -c1-110 >> $o/ab2_$tag.log
done
echo "== three parts" >> $o/ab2_$tag.log
timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 6 2>&1 | cut -c1-110 >> $o/ab2_$tag.log
PIPLIB_B200_HEAVY_PIVOTS=0 timeout 400 python tools/steal_bench.py > $o/steal_$tag.jsonl 2> $o/steal_$tag.err
cat $o/ab2_$tag.log; cut -c1-200 $o/steal_$tag.jsonl

#!/bin/bash
# Round-end pass on one GPU box: parity tests, both bench arms, e2e timing, then the ncu evidence
# profiles/ keeps (each capture only after the same command exited 0 without ncu).  Every step has
# its own timeout.  usage (repo root, under gpurun): bash tools/gpu_final.sh <tag>
tag=${1:-x}
o=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> $o/pytest_gpu_$tag.log
tail -3 $o/pytest_gpu_$tag.log
timeout 600 python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err; echo "bench rc=$?"
cut -c1-700 $o/bench_$tag.json
timeout 400 python bench.py --impl reference > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err; echo "ref rc=$?"
cut -c1-300 $o/bench_ref_$tag.json
timeout 200 python tools/e2e_timing.py 1000000 pinned > $o/e2e_pinned_$tag.log 2>&1; grep "^e2e" $o/e2e_pinned_$tag.log
timeout 200 python tools/e2e_timing.py 1000000 pageable > $o/e2e_pageable_$tag.log 2>&1; grep "^e2e" $o/e2e_pageable_$tag.log
timeout 200 python tools/ncu_run.py loopnest16x24p3 1000000 4 > $o/plain_$tag.log 2>&1; cat $o/plain_$tag.log
# launch lists: device-resident job (kernel-only figure) and the pinned end-to-end call
timeout 200 python tools/ncu_run.py loopnest16x24p3 200000 2 > /dev/null 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/launches_$tag.csv \
    python tools/ncu_run.py loopnest16x24p3 200000 2 > $o/ncu_l_$tag.log 2>&1
timeout 200 python tools/e2e_timing.py 262144 pinned > /dev/null 2>&1 && \
PIPLIB_B200_LANES=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $o/launches_e2e_$tag.csv \
    python tools/e2e_timing.py 262144 pinned > $o/ncu_e2e_$tag.log 2>&1
# the 10^6 device job as bench.py runs it (three parts on lanes of their own): launch list of one step
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/launches_1m_$tag.csv \
    python tools/ncu_run.py loopnest16x24p3 1000000 1 > $o/ncu_l1m_$tag.log 2>&1
# one full capture of the solve kernel (one launch, no hand-over: the kernel as the 10^6 job runs it)
PIPLIB_B200_HEAVY_PIVOTS=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pip_solve_kernel -c 1 -o $o/prof_solve_$tag -f \
    python tools/ncu_run.py loopnest16x24p3 200000 1 > $o/ncu_solve_$tag.log 2>&1
ls -la $o/prof_solve_$tag.ncu-rep $o/launches_$tag.csv $o/launches_e2e_$tag.csv

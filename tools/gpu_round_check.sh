#!/bin/bash
# One GPU-box pass used at the end of a work session: parity tests, both bench arms, then the ncu
# captures that profiles/ keeps (each only after its command has exited 0 without ncu).
# usage (from the repo root, under gpurun): bash tools/gpu_round_check.sh <tag>
tag=${1:-x}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> $o/pytest_gpu_$tag.log
python bench.py --impl reference > $o/bench_ref_$tag.json 2> $o/bench_ref_$tag.err
python bench.py > $o/bench_$tag.json 2> $o/bench_$tag.err
python tools/bench_large.py 4096 3 --no-check > $o/large_4096_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pip_large_kernel -c 1 -o $o/prof_large_$tag -f \
    python tools/bench_large.py 4096 1 --no-check > $o/ncu_large_$tag.log 2>&1
python tools/e2e_timing.py 262144 > $o/e2e_$tag.log 2>&1 && \
PIPLIB_B200_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $o/launches_e2e_$tag.csv \
    python tools/e2e_timing.py 262144 > $o/ncu_e2e_$tag.log 2>&1
python tools/ncu_run.py loopnest16x24p3 50000 3 > $o/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/launches_$tag.csv \
    python tools/ncu_run.py loopnest16x24p3 50000 3 > $o/ncu_l_$tag.log 2>&1
tail -2 $o/pytest_gpu_$tag.log; cut -c1-600 $o/bench_$tag.json; tail -1 $o/large_4096_$tag.log | cut -c1-400

#!/bin/bash
# one ncu --set full capture of the solve kernel of the default library: tools/gpu_ncu1.sh <tag> [n]
tag=$1; n=${2:-50000}
o=gpurun_out
timeout 300 python tools/ncu_run.py loopnest16x24p3 1000000 3 > $o/plain_$tag.log 2>&1; cat $o/plain_$tag.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pip_solve_kernel -c 1 -o $o/prof_solve_$tag -f \
    python tools/ncu_run.py loopnest16x24p3 $n 1 > $o/ncu_solve_$tag.log 2>&1
ls -la $o/prof_solve_$tag.ncu-rep

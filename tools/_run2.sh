python tools/e2e_timing.py 1000000 > gpurun_out/e2e_timing_v6.log 2>&1
PIPLIB_B200_TIMING=1 timeout 300 python tools/rounds.py boulet 11000 > gpurun_out/rounds_boulet2.log 2>&1
tail -40 gpurun_out/e2e_timing_v6.log; tail -12 gpurun_out/rounds_boulet2.log

timeout 300 python -m pytest tests -m gpu -x -q -k "large or ladder or cli_suite or config3" > gpurun_out/pytest_gpu_v15.log 2>&1; tail -2 gpurun_out/pytest_gpu_v15.log
PIPLIB_B200_TIMING=1 timeout 120 python tools/rounds.py vivien32 4000 2>&1 | tail -4
timeout 120 python tools/bench_large.py 4096 2 --no-check 2>&1 | tail -1 | cut -c1-330

PIPLIB_B200_LARGE_FROM=3 PIPLIB_B200_TIMING=1 timeout 120 python tools/rounds.py vivien32 4000 2>&1 | tail -5
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v14.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_v14.log; tail -3 gpurun_out/pytest_gpu_v14.log
PIPLIB_B200_LARGE_FROM=3 PIPLIB_B200_TEST_HANDOVER=1 timeout 400 python -m pytest tests -m gpu -x -q -k "ladder or config3 or cli_suite" > gpurun_out/pytest_gpu_v13.log 2>&1; tail -3 gpurun_out/pytest_gpu_v13.log

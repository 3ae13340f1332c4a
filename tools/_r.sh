python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v12.log; tail -2 gpurun_out/pytest_gpu_v12.log
python tools/e2e_timing.py 1000000 2>&1 | grep "^e2e\|lane 0:" | tail -4
PIPLIB_B200_EXACT_PLAN=1 python tools/e2e_timing.py 1000000 2>&1 | grep "^e2e\|lane 0:" | tail -4

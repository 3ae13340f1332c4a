python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v7.log
python tools/e2e_timing.py 1000000 2>&1 | grep "^e2e\|lane 0:" > gpurun_out/e2e_v7_warp.log
PIPLIB_B200_THREAD_DECODE=1 python tools/e2e_timing.py 1000000 2>&1 | grep "^e2e\|lane 0:" > gpurun_out/e2e_v7_thread.log
tail -3 gpurun_out/pytest_gpu_v7.log; cat gpurun_out/e2e_v7_warp.log gpurun_out/e2e_v7_thread.log

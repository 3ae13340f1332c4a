timeout 300 python -m pytest tests -m gpu -x -q -k "dense or config3 or edge" > gpurun_out/pytest_gpu_v16.log 2>&1; tail -2 gpurun_out/pytest_gpu_v16.log
timeout 200 python tools/e2e_timing.py 1000000 2>&1 | grep "^e2e" | tail -3
PIPLIB_B200_LANES=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/launches_e2e_v16.csv python tools/e2e_timing.py 262144 > gpurun_out/ncu_e2e_v16.log 2>&1
grep "serialize" gpurun_out/launches_e2e_v16.csv | awk -F'","' '{print $5, $NF}' | cut -c1-40,180-

timeout 600 python -m pytest tests -m gpu -x -q -k "large" > gpurun_out/pytest_large_v8.log 2>&1; tail -3 gpurun_out/pytest_large_v8.log
timeout 300 python tools/bench_large.py 1024 2 > gpurun_out/large_1024_v8.log 2>&1; tail -1 gpurun_out/large_1024_v8.log | cut -c1-300; grep -o '"parity": [a-z]*' gpurun_out/large_1024_v8.log
timeout 300 python tools/bench_large.py 4096 3 --no-check > gpurun_out/large_4096_v8.log 2>&1; tail -1 gpurun_out/large_4096_v8.log

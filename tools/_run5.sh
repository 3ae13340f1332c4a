timeout 600 python -m pytest tests -m gpu -x -q -k "large" > gpurun_out/pytest_large_v9.log 2>&1; tail -1 gpurun_out/pytest_large_v9.log
timeout 300 python tools/bench_large.py 1024 2 > gpurun_out/large_1024_v9.log 2>&1; grep -o '"parity": [a-z]*' gpurun_out/large_1024_v9.log
timeout 300 python tools/bench_large.py 4096 3 --no-check 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('staged', d['us_per_pivot'], d['phase_share']['update'], d['roofline']['frac'])"
PIPLIB_B200_NO_TMA=1 timeout 300 python tools/bench_large.py 4096 3 --no-check 2>&1 | tail -1 | python -c "
import sys, json; d = json.loads(sys.stdin.read()); print('direct', d['us_per_pivot'], d['phase_share']['update'], d['roofline']['frac'])"
PIPLIB_B200_LIB=piplib_b200/lib/libpiplib_dp_walk.so timeout 300 python tools/bench_large.py 4096 1 --no-check 2>&1 | grep "update phase" | tail -1

set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_v6.log
python bench.py --impl reference > gpurun_out/bench_ref_v6.json 2> gpurun_out/bench_ref_v6.err
python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
PIPLIB_B200_TIMING=1 timeout 300 python tools/rounds.py boulet 20000 > gpurun_out/rounds_boulet.log 2>&1
python tools/bench_large.py 4096 2 --no-check > gpurun_out/plain_large_v6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pip_large -c 1 -o gpurun_out/prof_r1_large_v6 -f python tools/bench_large.py 4096 1 --no-check > gpurun_out/ncu_large_v6.log 2>&1
tail -3 gpurun_out/pytest_gpu_v6.log; cat gpurun_out/bench_v6.json | cut -c1-1500; tail -5 gpurun_out/rounds_boulet.log

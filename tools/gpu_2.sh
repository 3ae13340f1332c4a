#!/bin/bash
# 2-GPU sanity pass with the shipped library: both bench arms under torchrun, the multi-device test, C5 strong scaling
o=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > $o/r2w_bench2.json 2> $o/r2w_bench2.err; echo "bench2 rc=$?"
cut -c1-260 $o/r2w_bench2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > $o/r2w_bench2_ref.json 2> $o/r2w_bench2_ref.err; echo "ref2 rc=$?"
cut -c1-200 $o/r2w_bench2_ref.json
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "several_gpus" > $o/r2w_pytest_multi.log 2>&1; tail -2 $o/r2w_pytest_multi.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 2 --warmup 3 --workload sor1d --total 10000000 --no-large > $o/r2w_c5_sor1d_2.json 2> $o/r2w_c5_sor1d_2.err; echo "c5 rc=$?"
cut -c1-300 $o/r2w_c5_sor1d_2.json

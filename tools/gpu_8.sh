#!/bin/bash
# the 8-GPU pass: weak-scaling bench line, config 5 strong scaling, one-call multi-GPU, the multi-device test
o=gpurun_out
nvidia-smi -L | wc -l > $o/r2m_host.log; nproc >> $o/r2m_host.log; free -g | head -2 >> $o/r2m_host.log; lscpu | grep -i "model name\|socket\|numa node(s)" >> $o/r2m_host.log
PIPLIB_B200_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 > $o/r2m_bench8.json 2> $o/r2m_bench8.err
cut -c1-300 $o/r2m_bench8.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 2 --warmup 3 --workload sor1d --total 10000000 > $o/r2m_c5_sor1d_8.json 2> $o/r2m_c5_sor1d_8.err
cut -c1-300 $o/r2m_c5_sor1d_8.json
timeout 300 python tools/multi_gpu_onecall.py 2000000 > $o/r2m_onecall.jsonl 2> $o/r2m_onecall.err; cat $o/r2m_onecall.jsonl
timeout 200 python -m pytest tests/test_gpu_parity.py -q -k "several_gpus" > $o/r2m_pytest_multi.log 2>&1; tail -2 $o/r2m_pytest_multi.log

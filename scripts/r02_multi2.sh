#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/exchange_check.py > gpurun_out/r02_exchange_check.txt 2>&1; tail -15 gpurun_out/r02_exchange_check.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_n2.txt 2>&1; tail -5 gpurun_out/r02_bench_n2.txt

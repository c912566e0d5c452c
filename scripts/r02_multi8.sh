#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tests/exchange_check.py > gpurun_out/r02_exchange_check8.txt 2>&1; tail -4 gpurun_out/r02_exchange_check8.txt | cut -c1-300
for N in 8 4 2; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.txt 2>&1; tail -1 gpurun_out/r02_bench_n$N.txt | cut -c1-2500
done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1b.txt 2>&1; tail -1 gpurun_out/r02_bench_n1b.txt | cut -c1-600

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02_bench_n8c.txt 2>&1; tail -1 gpurun_out/r02_bench_n8c.txt | cut -c1-3500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29672 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02_bench_n4c.txt 2>&1; tail -1 gpurun_out/r02_bench_n4c.txt | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 3 --no-q5 > gpurun_out/r02_bench_n1e.txt 2>&1; tail -1 gpurun_out/r02_bench_n1e.txt | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 scripts/q1_dist_probe.py > gpurun_out/r02_dist_probe2.txt 2>&1; tail -2 gpurun_out/r02_dist_probe2.txt | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 tests/exchange_check.py > gpurun_out/r02_exchange_check2c.txt 2>&1; tail -2 gpurun_out/r02_exchange_check2c.txt | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 scripts/q1_dist_probe.py > gpurun_out/r02_dist_probe8.txt 2>&1; tail -1 gpurun_out/r02_dist_probe8.txt | cut -c1-1500

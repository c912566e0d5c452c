#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest12.txt 2>&1; tail -15 gpurun_out/r02_pytest12.txt
timeout 300 python scripts/q1_stages.py 10 > gpurun_out/r02_q1_stages.json 2> gpurun_out/r02_q1_stages.err; cat gpurun_out/r02_q1_stages.json; tail -3 gpurun_out/r02_q1_stages.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_q1_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r02_q1_ncu.log 2>&1; tail -2 gpurun_out/r02_q1_ncu.log | cut -c1-300

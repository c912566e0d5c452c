#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest1.txt
cat gpurun_out/r02_pytest1.txt
timeout 600 python scripts/q5_probe.py 100 > gpurun_out/r02_q5_v1.json 2> gpurun_out/r02_q5_v1.err
tail -3 gpurun_out/r02_q5_v1.err; cat gpurun_out/r02_q5_v1.json

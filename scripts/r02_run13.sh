#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sink_and_views.py -x -q > gpurun_out/r02_pytest13.txt 2>&1; tail -5 gpurun_out/r02_pytest13.txt
Q5_REPS=4 timeout 600 python scripts/q5_stages.py 100 > gpurun_out/r02_q5_stages_v8.json 2> gpurun_out/r02_q5_stages_v8.err; cut -c1-400 gpurun_out/r02_q5_stages_v8.json | tail -2
timeout 900 python scripts/operator_bench.py > gpurun_out/r02_operator_bench.txt 2>&1; tail -40 gpurun_out/r02_operator_bench.txt

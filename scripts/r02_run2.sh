#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/q5_stages.py 100 > gpurun_out/r02_q5_stages_v1.json 2> gpurun_out/r02_q5_stages_v1.err
tail -3 gpurun_out/r02_q5_stages_v1.err; cat gpurun_out/r02_q5_stages_v1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_q5_v1_launches.csv python scripts/q5_probe.py 100 > /dev/null 2>&1
timeout 600 python -m pytest tests/test_gpu_tpch.py -x -q 2>&1 | tail -5

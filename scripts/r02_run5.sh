#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tpch.py tests/test_gpu_large.py -x -q 2>&1 | tail -8
timeout 600 python scripts/q5_probe.py 100 > gpurun_out/r02_q5_v3.json 2> gpurun_out/r02_q5_v3.err; tail -2 gpurun_out/r02_q5_v3.err; cat gpurun_out/r02_q5_v3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_q5_v3_launches.csv python scripts/q5_probe.py 100 > /dev/null 2>&1
timeout 600 python scripts/q5_stages.py 100 > gpurun_out/r02_q5_stages_v3.json 2> gpurun_out/r02_q5_stages_v3.err; cat gpurun_out/r02_q5_stages_v3.json

#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r02_pytest7.txt; cat gpurun_out/r02_pytest7.txt
timeout 600 python scripts/q5_probe.py 100 > gpurun_out/r02_q5_v5.json 2> gpurun_out/r02_q5_v5.err; tail -2 gpurun_out/r02_q5_v5.err; cat gpurun_out/r02_q5_v5.json
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r02_q5_v5_launches.csv python scripts/q5_probe.py 100 > /dev/null 2>&1

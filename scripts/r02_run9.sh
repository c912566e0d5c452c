#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sink_and_views.py tests/test_gpu_parity.py tests/test_gpu_tpch.py tests/test_gpu_large.py tests/test_gpu_sqlite_truth.py -x -q 2>&1 | tail -8
Q5_REPS=5 timeout 600 python scripts/q5_stages.py 100 > gpurun_out/r02_q5_stages_v6.json 2> gpurun_out/r02_q5_stages_v6.err; tail -3 gpurun_out/r02_q5_stages_v6.json | cut -c1-420
timeout 600 python scripts/q5_probe.py 100 > gpurun_out/r02_q5_v6.json 2> gpurun_out/r02_q5_v6.err; tail -2 gpurun_out/r02_q5_v6.err; cat gpurun_out/r02_q5_v6.json
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_q5_v6_launches.csv python scripts/q5_probe.py 100 > /dev/null 2>&1

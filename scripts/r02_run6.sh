#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tpch.py tests/test_gpu_large.py -x -q 2>&1 | tail -8
timeout 600 python scripts/q5_probe.py 100 > gpurun_out/r02_q5_v4.json 2> gpurun_out/r02_q5_v4.err; tail -2 gpurun_out/r02_q5_v4.err; cat gpurun_out/r02_q5_v4.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_q5_v4_launches.csv python scripts/q5_probe.py 100 > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'probe_count_kernel|probe_write_kernel' -c 6 -o gpurun_out/r02_probe_v4 python scripts/q5_probe.py 100 > /dev/null 2>&1
ncu -i gpurun_out/r02_probe_v4.ncu-rep --page raw --csv > gpurun_out/r02_probe_v4_raw.csv 2>/dev/null

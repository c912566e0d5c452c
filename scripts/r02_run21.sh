#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sink_and_views.py -q -k "partition or sort or top_n or dedup or join" > gpurun_out/r02_pytest21.txt 2>&1; tail -4 gpurun_out/r02_pytest21.txt
python scripts/kern_probe.py 2>&1 | grep -E "argsort|partition"
B200_SCATTER_STAGED=0 python scripts/kern_probe.py 2>&1 | grep -E "partition"
python scripts/kern_probe.py 2>&1 | grep -E "argsort|partition"

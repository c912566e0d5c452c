#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_list_aggs.py tests/test_gpu_parity.py -x -q -k "list or sort or partition or top_n or join or dedup or distinct" > gpurun_out/r02_pytest15.txt 2>&1; tail -15 gpurun_out/r02_pytest15.txt
timeout 900 python scripts/operator_bench.py > gpurun_out/r02_operator_bench2.txt 2>&1; grep -E "partition|argsort|top_n|join" gpurun_out/r02_operator_bench2.txt | cut -c1-260

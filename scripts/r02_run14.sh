#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_list_aggs.py -q > gpurun_out/r02_pytest14.txt 2>&1; tail -40 gpurun_out/r02_pytest14.txt

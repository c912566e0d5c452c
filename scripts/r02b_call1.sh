#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zzz_tpch_all.py -q --tb=short -p no:cacheprovider > gpurun_out/r02b_zzz.txt 2>&1; tail -40 gpurun_out/r02b_zzz.txt
timeout 150 python -m pytest tests/test_gpu_parity.py -q --tb=short -k "sort or top_n" -p no:cacheprovider > gpurun_out/r02b_sort.txt 2>&1; tail -5 gpurun_out/r02b_sort.txt

#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py --no-q5 --no-cpu-baseline --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/r02b_bench_tpch22.txt 2> gpurun_out/r02b_bench_tpch22.err; tail -c 6000 gpurun_out/r02b_bench_tpch22.txt; tail -5 gpurun_out/r02b_bench_tpch22.err

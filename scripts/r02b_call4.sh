#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_zzz2_joins_sorted.py tests/test_gpu_zzz3_window_pivot.py -q --tb=short -p no:cacheprovider > gpurun_out/r02b_zzz3.txt 2>&1; tail -60 gpurun_out/r02b_zzz3.txt
timeout 300 python bench.py --no-q5 --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 2 --tpch22-sf 100 --tpch22-queries 20,21 --tpch22-budget-s 240 > gpurun_out/r02b_bench_q20_21_sf100.txt 2> gpurun_out/r02b_bench_q20_21_sf100.err; python - <<'P'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02b_bench_q20_21_sf100.txt') if l.startswith('{')][-1])["tpch22"]
    print(d["queries_failed"], d["per_query"])
except Exception as e:
    print("no line", e)
P
tail -5 gpurun_out/r02b_bench_q20_21_sf100.err

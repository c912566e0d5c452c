#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_zzz2_joins_sorted.py tests/test_gpu_zzz_tpch_all.py -q --tb=short -p no:cacheprovider -k "joins_sorted or generator or asof or merge or cross" > gpurun_out/r02b_zzz2.txt 2>&1; tail -40 gpurun_out/r02b_zzz2.txt
timeout 400 python bench.py --no-q5 --no-cpu-baseline --steps 3 --warmup 3 --e2e-steps 2 --tpch22-sf 100 --tpch22-budget-s 240 > gpurun_out/r02b_bench_tpch22_sf100.txt 2> gpurun_out/r02b_bench_tpch22_sf100.err; python - <<'P'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02b_bench_tpch22_sf100.txt') if l.startswith('{')][-1])["tpch22"]
    print({k:v for k,v in d.items() if k!="per_query"}); print({k:round(v["ms"],1) for k,v in d["per_query"].items()})
except Exception as e:
    print("no line", e)
P
tail -5 gpurun_out/r02b_bench_tpch22_sf100.err

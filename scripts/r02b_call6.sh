#!/bin/bash
mkdir -p gpurun_out
timeout 50 python bench.py --no-q5 --tpch22-sf 0 --no-cpu-baseline --sf 1 --steps 3 --warmup 3 --e2e-steps 2 > gpurun_out/r02b_bench_parquet_leg.txt 2> gpurun_out/r02b_bench_parquet_leg.err
python - <<'P'
import json
try:
    d = json.loads([l for l in open('gpurun_out/r02b_bench_parquet_leg.txt') if l.startswith('{')][-1])
    print(json.dumps(d.get("parquet_scan")))
except Exception as e:
    print("no line", e)
P
tail -3 gpurun_out/r02b_bench_parquet_leg.err

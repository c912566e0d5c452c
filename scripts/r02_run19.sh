#!/bin/bash
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>/dev/null
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest19.txt 2>&1; tail -8 gpurun_out/r02_pytest19.txt
(time timeout 900 python bench.py --impl reference --steps 5 --warmup 1) > gpurun_out/r02_bench_refd.json 2> gpurun_out/r02_bench_refd.err; tail -4 gpurun_out/r02_bench_refd.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_refd.json'))
print(d['ms_per_step'], d['cpu_baseline']['cores'], d['host'], d['q5']['ms_per_step'])
PY

#!/bin/bash
mkdir -p gpurun_out
B200_TRACE_ALLOC=1 Q5_REPS=8 timeout 600 python scripts/q5_stages.py 100 > gpurun_out/r02_q5_stages_v7.json 2> gpurun_out/r02_q5_stages_v7.err; cut -c1-330 gpurun_out/r02_q5_stages_v7.json | tail -6; grep -c alloc gpurun_out/r02_q5_stages_v7.err
(time timeout 1500 python bench.py --steps 20 --warmup 3) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -5 gpurun_out/r02_bench_n1.err; cat gpurun_out/r02_bench_n1.json | cut -c1-3000
(time timeout 900 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; tail -5 gpurun_out/r02_bench_ref.err; cat gpurun_out/r02_bench_ref.json | cut -c1-2500
nproc; free -g | head -2

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_list_aggs.py tests/test_gpu_sink_and_views.py -q > gpurun_out/r02_pytest20.txt 2>&1; tail -4 gpurun_out/r02_pytest20.txt
KP_REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_kp_launches3.csv python scripts/kern_probe.py > gpurun_out/r02_kp_ncu3.log 2>&1; tail -2 gpurun_out/r02_kp_ncu3.log
(time timeout 1200 python bench.py --steps 20 --warmup 3) > gpurun_out/r02_bench_n1d.json 2> gpurun_out/r02_bench_n1d.err; tail -4 gpurun_out/r02_bench_n1d.err; cut -c1-300 gpurun_out/r02_bench_n1d.json

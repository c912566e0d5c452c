#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz2_functions.py tests/test_gpu_sink_and_views.py tests/test_gpu_parity.py -x -q 2>&1 | tail -5
B200_TRACE_ALLOC=1 Q5_REPS=8 timeout 600 python scripts/q5_stages.py 100 > gpurun_out/r02_q5_stages_v5.json 2> gpurun_out/r02_q5_stages_v5.err; cat gpurun_out/r02_q5_stages_v5.json | cut -c1-600; grep -c alloc gpurun_out/r02_q5_stages_v5.err; tail -40 gpurun_out/r02_q5_stages_v5.err

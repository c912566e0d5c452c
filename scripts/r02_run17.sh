#!/bin/bash
mkdir -p gpurun_out
python scripts/kern_probe.py > gpurun_out/r02_kp_plain2.txt 2>&1; cat gpurun_out/r02_kp_plain2.txt
B200_SORT_ONESWEEP=0 python scripts/kern_probe.py 2>&1 | grep argsort
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "sort or partition or top_n or dedup" > gpurun_out/r02_pytest17.txt 2>&1; tail -3 gpurun_out/r02_pytest17.txt
KP_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_kp_launches2.csv python scripts/kern_probe.py > gpurun_out/r02_kp_ncu2.log 2>&1; tail -3 gpurun_out/r02_kp_ncu2.log

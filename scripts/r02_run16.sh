#!/bin/bash
mkdir -p gpurun_out
python scripts/kern_probe.py > gpurun_out/r02_kp_plain.txt 2>&1; cat gpurun_out/r02_kp_plain.txt
B200_HASH_MINB3=1 python scripts/kern_probe.py 2>&1 | grep hash
B200_HASH_STAGED=0 python scripts/kern_probe.py 2>&1 | grep hash
B200_SORT_ONESWEEP=0 python scripts/kern_probe.py 2>&1 | grep argsort
KP_REPS=1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_kp_launches.csv python scripts/kern_probe.py > gpurun_out/r02_kp_ncu.log 2>&1; tail -3 gpurun_out/r02_kp_ncu.log

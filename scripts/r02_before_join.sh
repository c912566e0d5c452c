#!/bin/bash
# round-2 "before" capture of the round-1 join kernels on TPC-H Q5 SF100 (one B200)
set -x
mkdir -p gpurun_out
python scripts/q5_probe.py 100 > gpurun_out/r02_q5_before.json 2> gpurun_out/r02_q5_before.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_q5_before_launches.csv python scripts/q5_probe.py 100 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'probe_count_kernel|probe_write_kernel|bucket_bounds' -c 9 -o gpurun_out/r02_join_before python scripts/q5_probe.py 100 > /dev/null 2>&1
ncu -i gpurun_out/r02_join_before.ncu-rep --page raw --csv > gpurun_out/r02_join_before_raw.csv 2>/dev/null
ls -la gpurun_out | head -40
tail -3 gpurun_out/r02_q5_before.json

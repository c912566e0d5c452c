#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'probe_kernel' -c 6 -o gpurun_out/r02_probe_v2 python scripts/q5_probe.py 100 > /dev/null 2>&1
ncu -i gpurun_out/r02_probe_v2.ncu-rep --page raw --csv > gpurun_out/r02_probe_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_probe_v2.ncu-rep --page details --csv > gpurun_out/r02_probe_v2_details.csv 2>/dev/null
ls -la gpurun_out/r02_probe_v2*

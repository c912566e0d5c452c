#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest18.txt 2>&1; tail -6 gpurun_out/r02_pytest18.txt
timeout 900 python scripts/operator_bench.py > gpurun_out/r02_operator_bench3.txt 2>&1; grep -E "partition|hash_rows" gpurun_out/r02_operator_bench3.txt | cut -c1-260
(time timeout 1200 python bench.py --steps 20 --warmup 3) > gpurun_out/r02_bench_n1c.json 2> gpurun_out/r02_bench_n1c.err; tail -4 gpurun_out/r02_bench_n1c.err; cut -c1-700 gpurun_out/r02_bench_n1c.json
(time timeout 900 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/r02_bench_refc.json 2> gpurun_out/r02_bench_refc.err; tail -4 gpurun_out/r02_bench_refc.err; cut -c1-500 gpurun_out/r02_bench_refc.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2

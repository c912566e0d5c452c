#!/bin/bash
# the round's last GPU call: the whole -m gpu suite, smoke(), and the Parquet scan probe after the page-per-warp change
mkdir -p gpurun_out
timeout 420 python -m pytest tests/ -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02b_pytest_final.txt 2>&1; tail -25 gpurun_out/r02b_pytest_final.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.txt 2>&1; tail -2 gpurun_out/r02b_smoke.txt
timeout 200 python - <<'P' > gpurun_out/r02b_parquet_scan2.txt 2>&1
import os, sys, time, json
sys.path.insert(0, '.')
import pyarrow.parquet as pq
import daft_b200
from daft_b200 import _lib, tpch
daft_b200.init(0)
host = tpch.generate_host("lineitem", 1.0, tpch.Q1_COLUMNS)
res = {}
for name, opts in (("snappy_dict", dict(compression="snappy")), ("plain_uncompressed", dict(compression="none", use_dictionary=False)), ("snappy_dict_64k_pages", dict(compression="snappy", data_page_size=65536))):
    path = f"/tmp/li_{name}.parquet"
    pq.write_table(host, path, row_group_size=1 << 20, **opts)
    size = os.path.getsize(path)
    for _ in range(2): rb = daft_b200.read_parquet(path)
    _lib.lib().b200_kernel_timing(1); _lib.kernel_timing_report()
    _lib.synchronize(); t0 = time.perf_counter()
    for _ in range(3): rb = daft_b200.read_parquet(path)
    _lib.synchronize(); dt = (time.perf_counter() - t0) / 3
    kt = _lib.kernel_timing_report(); _lib.lib().b200_kernel_timing(0)
    pq.read_table(path); t1 = time.perf_counter(); t = pq.read_table(path); dt_pa = time.perf_counter() - t1
    res[name] = {"file_MB": size / 1e6, "arrow_MB": host.nbytes / 1e6, "rows": host.num_rows, "read_ms": dt * 1e3, "rows_per_s": host.num_rows / dt,
                 "file_GBps": size / dt / 1e9, "arrow_GBps": host.nbytes / dt / 1e9, "device_ms": {k: v["ms"] / 3 for k, v in kt.items()},
                 "pyarrow_read_table_ms_warm": dt_pa * 1e3}
print(json.dumps(res))
P
tail -2 gpurun_out/r02b_parquet_scan2.txt

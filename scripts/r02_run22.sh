#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_semantics.py tests/test_gpu_sqlite_truth.py -q -k "sort or top_n or sqlite or argsort" > gpurun_out/r02_pytest22.txt 2>&1; tail -6 gpurun_out/r02_pytest22.txt
python - <<'PY'
import time, sys
sys.path.insert(0, '.')
import daft_b200
from daft_b200 import _lib, tpch, col
daft_b200.init(0)
li = tpch.generate("lineitem", 10.0, ["L_SHIPMODE", "L_RETURNFLAG"])
for name in ("L_SHIPMODE", "L_RETURNFLAG"):
    b = li.eval_expression_list([col(name)])
    for _ in range(2): b.argsort([col(name)])
    _lib.synchronize(); t0 = time.perf_counter()
    for _ in range(3): b.argsort([col(name)])
    _lib.synchronize(); print("argsort", name, round((time.perf_counter() - t0) / 3 * 1e3, 2), "ms")
PY

"""TPC-H Q5 six-table join pipeline on one GPU (BASELINE config[2] at SF100)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import daft_b200
from daft_b200 import _lib, tpch
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
daft_b200.init(0)
stream = torch.cuda.ExternalStream(_lib.stream_ptr())
t0 = time.perf_counter()
tables = {t: tpch.generate(t, sf, cols) for t, cols in tpch.Q5_COLUMNS.items()}
_lib.synchronize()
gen_s = time.perf_counter() - t0
in_bytes = sum(tb.size_bytes() for tb in tables.values())
res = tpch.q5(tables)  # warm-up (JIT)
reps = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.synchronize(); e0.record(stream)
for _ in range(reps): res = tpch.q5(tables)
e1.record(stream); _lib.synchronize()
ms = e0.elapsed_time(e1) / reps
nl = len(tables["lineitem"])
print(json.dumps({"query": "q5", "sf": sf, "ms": round(ms, 3), "lineitem_rows": nl, "lineitem_rows_per_s": round(nl / ms * 1e3), "input_GB": round(in_bytes / 1e9, 3),
                  "GBps_vs_compulsory_input": round(in_bytes / ms / 1e6, 1), "gen_s": round(gen_s, 2), "result": res.to_pydict(), "peak_mem_GB": round(torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9, 1)}))

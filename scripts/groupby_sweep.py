"""bench_agg.rs-style cardinality sweep (key = i % g, val = i % 1000) on the fused aggregate."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa, torch
import daft_b200
from daft_b200 import _lib, col
daft_b200.init(0)
stream = torch.cuda.ExternalStream(_lib.stream_ptr())
N = int(os.environ.get("OPB_ROWS", 50_000_000))
def timed(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    _lib.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); _lib.synchronize()
    return e0.elapsed_time(e1) / reps
i = np.arange(N, dtype=np.int64)
rng = np.random.default_rng(0)
for g in (4, 64, 1000, 100_000, 10_000_000, N):
    for pattern in ("mod", "random"):
        k = (i % g) if pattern == "mod" else rng.integers(0, g, N)
        t = daft_b200.RecordBatch.from_arrow_table(pa.table({"k": pa.array(k), "k32": pa.array((k % (2**31)).astype(np.int32)), "v": pa.array(i % 1000), "f": pa.array((i % 977).astype(np.float64))}))
        for key in ("k", "k32"):
            ms = timed(lambda: t.agg([col("v").sum().alias("s"), col("v").count().alias("c"), col("f").sum().alias("sf")], [col(key)]))
            out = t.agg([col("v").sum().alias("s"), col("v").count().alias("c")], [col(key)])
            print(json.dumps({"groups": g, "pattern": pattern, "key": "int64" if key == "k" else "int32", "ms": round(ms, 3), "rows_per_s": round(N / ms * 1e3), "out_groups": len(out)}), flush=True)
        del t

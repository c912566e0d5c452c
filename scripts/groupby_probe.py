"""Dev probe: one group-by shape (rows, groups, pattern) for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa
import daft_b200
from daft_b200 import col

daft_b200.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
g = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
pattern = sys.argv[3] if len(sys.argv) > 3 else "mod"
i = np.arange(n, dtype=np.int64)
k = i % g if pattern == "mod" else np.random.default_rng(1).integers(0, g, n)
t = daft_b200.RecordBatch.from_arrow_table(pa.table({"k": pa.array(k), "v": pa.array(i % 1000), "f": pa.array((i % 977).astype(np.float64))}))
for _ in range(2):
    out = t.agg([col("v").sum().alias("s"), col("v").count().alias("c"), col("f").sum().alias("sf")], [col("k")])
print(len(out))
import time, torch
from daft_b200 import _lib
L = _lib.lib(); L.b200_kernel_timing(1)
stream = torch.cuda.ExternalStream(_lib.stream_ptr())
aggs = [col("v").sum().alias("s"), col("v").count().alias("c"), col("f").sum().alias("sf")]
for rep in range(4):
    l0 = daft_b200.launch_count(); _lib.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    out = t.agg(aggs, [col("k")])
    e1.record(stream); _lib.synchronize(); t1 = time.perf_counter()
    print(pattern, g, "event_ms", round(e0.elapsed_time(e1), 3), "wall_ms", round((t1 - t0) * 1e3, 3), "last_agg_kernel_ms", round(L.b200_last_agg_kernel_ms(), 3),
          "launches", daft_b200.launch_count() - l0, flush=True)

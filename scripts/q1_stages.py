"""TPC-H Q1 step by stage on one GPU: aggregate call alone, the sort alone, the whole step (wall clock around synchronised
batches of repetitions) and the library's per-kernel timing of one step.  Diagnostic, not a bench number."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import _lib, tpch, col
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
daft_b200.init(0)
li = tpch.generate("lineitem", sf, tpch.Q1_COLUMNS)
pred, gb, aggs = tpch.q1_exprs()
def timed(fn, reps=50):
    for _ in range(5): fn()
    _lib.synchronize(); l0 = _lib.launch_count(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    _lib.synchronize()
    return round((time.perf_counter() - t0) / reps * 1e3, 4), (_lib.launch_count() - l0) / reps
out = {}
out["agg_ms,launches"] = timed(lambda: li.agg(aggs, gb, predicate=pred))
r = li.agg(aggs, gb, predicate=pred)
out["sort_ms,launches"] = timed(lambda: r.sort([col("L_RETURNFLAG"), col("L_LINESTATUS")]))
out["q1_ms,launches"] = timed(lambda: tpch.q1(li))
print(json.dumps(out))

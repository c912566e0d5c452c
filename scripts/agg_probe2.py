import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import _lib, tpch, col
daft_b200.init(0)
L = _lib.lib(); L.b200_kernel_timing(1)
li = tpch.generate("lineitem", 10.0, tpch.Q1_COLUMNS); n = len(li)
def probe(name, fn, bpr, reps=5):
    for _ in range(2): fn()
    ms = []
    for _ in range(reps):
        fn(); ms.append(L.b200_last_agg_kernel_ms())
    m = sum(ms) / len(ms)
    print(json.dumps({"U": os.environ.get("B200_AGG_U"), "M": os.environ.get("B200_AGG_MINB"), "probe": name, "kernel_ms": round(m, 4), "GBps": round(bpr * n / m / 1e6, 1)}), flush=True)
probe("flags-count", lambda: li.agg([col("L_QUANTITY").count("all").alias("c")], [col("L_RETURNFLAG"), col("L_LINESTATUS")]), 18)
probe("q1", lambda: tpch.q1(li), 54)

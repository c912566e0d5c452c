"""Kernel-time probes for the fused aggregate kernel on SF10 lineitem (dev tool, not a bench)."""
import sys, os, json, datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import _lib, tpch, col

daft_b200.init(0)
L = _lib.lib()
L.b200_kernel_timing(1)
cols = tpch.Q1_COLUMNS
li = tpch.generate("lineitem", 10.0, cols)
n = len(li)

def probe(name, fn, bytes_per_row, reps=5):
    for _ in range(2): fn()
    ms = []
    for _ in range(reps):
        fn(); ms.append(L.b200_last_agg_kernel_ms())
    m = sum(ms) / len(ms)
    print(json.dumps({"probe": name, "kernel_ms": round(m, 4), "GBps": round(bytes_per_row * n / m / 1e6, 1), "rows_per_s": round(n / m * 1e3)}), flush=True)

probe("q6 ungrouped (28 B/row)", lambda: tpch.q6(li), 28)
probe("sum(ext) ungrouped (8 B/row)", lambda: li.agg([col("L_EXTENDEDPRICE").sum()], []), 8)
probe("sum 4 numeric cols ungrouped (32 B/row)", lambda: li.agg([col("L_EXTENDEDPRICE").sum(), col("L_DISCOUNT").sum(), col("L_TAX").sum(), col("L_QUANTITY").sum()], []), 32)
probe("q1 aggs, group by shipdate%4 (int key; 36 B/row)", lambda: li.agg(tpch.q1_exprs()[2], [(col("L_QUANTITY") % 4).alias("k")]), 36)
probe("q1 aggs ungrouped (36 B/row)", lambda: li.agg(tpch.q1_exprs()[2], []), 36)
probe("count by flags only (18 B/row)", lambda: li.agg([col("L_QUANTITY").count("all").alias("c")], [col("L_RETURNFLAG"), col("L_LINESTATUS")]), 18)
probe("q1 full (54 B/row)", lambda: tpch.q1(li), 54)

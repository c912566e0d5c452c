"""TPC-H Q5 by stage on one GPU (host-synchronised between stages; diagnostic, not a bench number)."""
import sys, os, json, time, datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import daft_b200
from daft_b200 import _lib, tpch, col
sf = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
daft_b200.init(0)
t = {tb: tpch.generate(tb, sf, cols) for tb, cols in tpch.Q5_COLUMNS.items()}
_lib.synchronize()
def stage(name, fn, out):
    _lib.synchronize(); t0 = time.perf_counter(); r = fn(); _lib.synchronize(); out.append((name, round((time.perf_counter() - t0) * 1e3, 3), len(r))); return r
for rep in range(int(os.environ.get('Q5_REPS', '3'))):
    out = []
    orders = stage("filter orders", lambda: t["orders"].filter([col("O_ORDERDATE") >= datetime.date(1994, 1, 1), col("O_ORDERDATE") < datetime.date(1995, 1, 1)]), out)
    region = stage("filter region", lambda: t["region"].filter([col("R_NAME") == "ASIA"]), out)
    j = stage("region x nation", lambda: region.hash_join(t["nation"], [col("R_REGIONKEY")], [col("N_REGIONKEY")], build_on_left=True), out)
    j = stage("x supplier", lambda: j.hash_join(t["supplier"], [col("N_NATIONKEY")], [col("S_NATIONKEY")], build_on_left=True), out)
    j = stage("x lineitem", lambda: j.hash_join(t["lineitem"], [col("S_SUPPKEY")], [col("L_SUPPKEY")], build_on_left=True), out)
    j = stage("select", lambda: j.eval_expression_list([col("N_NAME"), col("L_EXTENDEDPRICE"), col("L_DISCOUNT"), col("L_ORDERKEY"), col("N_NATIONKEY")]), out)
    j = stage("x orders", lambda: j.hash_join(orders, [col("L_ORDERKEY")], [col("O_ORDERKEY")], build_on_left=False), out)
    j = stage("x customer", lambda: j.hash_join(t["customer"], [col("O_CUSTKEY"), col("N_NATIONKEY")], [col("C_CUSTKEY"), col("C_NATIONKEY")], build_on_left=False), out)
    r = stage("agg", lambda: j.agg([(col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))).sum().alias("revenue")], [col("N_NAME")]), out)
    r = stage("sort", lambda: r.sort([col("revenue")], descending=True), out)
    print(json.dumps({"rep": rep, "total_ms": round(sum(x[1] for x in out), 3), "stages": out}))

"""Per-operator device timings with algorithmic-byte rooflines (SURVEY.md section 8d microbenchmarks).
Mirrors src/daft-recordbatch/src/ops/bench_agg.rs:22-41 (key = i % g, val = i % 1000) and
tests/microbenchmarks/{test_join,test_sort,test_filter,test_repartition}.py.  Writes JSON lines."""
import sys, os, json, datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pyarrow as pa, torch
import daft_b200
from daft_b200 import _lib, tpch, col

daft_b200.init(0)
stream = torch.cuda.ExternalStream(_lib.stream_ptr())
PEAK = 6573.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    _lib.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): fn()
    e1.record(stream); _lib.synchronize()
    return e0.elapsed_time(e1) / reps

def report(name, ms, rows, alg_bytes, note=""):
    gbps = alg_bytes / ms / 1e6
    print(json.dumps({"op": name, "rows": rows, "ms": round(ms, 4), "rows_per_s": round(rows / ms * 1e3), "alg_GB": round(alg_bytes / 1e9, 3),
                      "GBps": round(gbps, 1), "frac_of_measured_peak": round(gbps / PEAK, 3), "note": note}), flush=True)

def dev_table(n, groups):
    i = np.arange(n, dtype=np.int64)
    return daft_b200.RecordBatch.from_arrow_table(pa.table({"k": pa.array(i % groups), "v": pa.array(i % 1000), "f": pa.array((i % 977).astype(np.float64))}))

N = int(os.environ.get("OPB_ROWS", 50_000_000))
li = tpch.generate("lineitem", 10.0, ["L_ORDERKEY", "L_SUPPKEY", "L_QUANTITY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_TAX", "L_RETURNFLAG", "L_LINESTATUS", "L_SHIPDATE"])
n = len(li)
report("hash_rows(L_ORDERKEY i64)", timed(lambda: li.hash_rows([col("L_ORDERKEY")])), n, n * 16, "8 B key read + 8 B hash written")
report("hash_rows(L_ORDERKEY, L_SUPPKEY) chained", timed(lambda: li.hash_rows([col("L_ORDERKEY"), col("L_SUPPKEY")])), n, n * 40, "2 x 8 read + 8 written + 8 seed read + 8 rewritten")
report("hash_rows(2 flag strings)", timed(lambda: li.hash_rows([col("L_RETURNFLAG"), col("L_LINESTATUS")])), n, n * (18 + 24), "2 x (8 B offset + 1 B) + hash traffic")
pred = [col("L_SHIPDATE") <= datetime.date(1998, 9, 2)]
sel = li.eval_expression_list([col("L_QUANTITY"), col("L_EXTENDEDPRICE"), col("L_DISCOUNT"), col("L_TAX"), col("L_SHIPDATE")])
report("filter(shipdate<=) 5 fixed cols, 98% kept", timed(lambda: sel.filter(pred)), n, n * 36 * 2, "read + write of surviving columns (36 B/row)")
proj = [(col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))).alias("dp"), (col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT")) * (1 + col("L_TAX"))).alias("ch")]
report("project(disc_price, charge)", timed(lambda: li.eval_expression_list(proj)), n, n * (24 + 16), "3 f64 read, 2 f64 written")
p2 = li.eval_expression_list([col("L_ORDERKEY"), col("L_EXTENDEDPRICE"), col("L_DISCOUNT")])
report("partition_by_hash(L_ORDERKEY, P=8), 3 cols", timed(lambda: p2.partition_by_hash([col("L_ORDERKEY")], 8), 3, 1), n, n * 24 * 2, "payload read + written once (hash/sort traffic not counted)")
s1 = li.eval_expression_list([col("L_ORDERKEY")])
report("argsort(L_ORDERKEY i64)", timed(lambda: s1.argsort([col("L_ORDERKEY")]), 3, 1), n, n * 16, "8 B key read + 8 B index written (radix passes not counted)")
s2 = li.eval_expression_list([col("L_EXTENDEDPRICE")])
report("argsort(L_EXTENDEDPRICE f64)", timed(lambda: s2.argsort([col("L_EXTENDEDPRICE")]), 3, 1), n, n * 16)
report("top_n(L_EXTENDEDPRICE desc, 100)", timed(lambda: s2.top_n([col("L_EXTENDEDPRICE")], 100, descending=True), 3, 1), n, n * 8)
for g in (4, 1000, 100_000, 10_000_000):
    t = dev_table(N, g)
    ms = timed(lambda: t.agg([col("v").sum().alias("s"), col("v").count().alias("c"), col("f").sum().alias("sf")], [col("k")]), 3, 1)
    report(f"groupby int64 key, {g} groups: sum(i64), count, sum(f64)", ms, N, N * 24, "bench_agg.rs shape")
    del t
orders = tpch.generate("orders", 10.0, ["O_ORDERKEY", "O_CUSTKEY", "O_ORDERDATE"])
from daft_b200.join import ProbeTable
def build():
    pt = ProbeTable(1); pt.add_table(orders, [0]); pt.build(); return pt
report("join build: 15M orders rows (i64 key)", timed(build, 3, 1), len(orders), len(orders) * 20, "build batch read once (hash/sort/table traffic not counted)")
pt = build()
lp = li.eval_expression_list([col("L_ORDERKEY"), col("L_EXTENDEDPRICE"), col("L_DISCOUNT")])
report("probe_indices: 60M lineitem keys into 15M-row table", timed(lambda: pt.probe_indices(lp, [0]), 3, 1), n, n * (8 + 32 + 16), "key + one 32 B slot + (probe,build) index pair out")
report("hash_join inner lineitem x orders (materialised 6 cols)", timed(lambda: lp.hash_join(orders, [col("L_ORDERKEY")], [col("O_ORDERKEY")]), 3, 1), n, n * (24 + 20) * 2)

import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import _lib, tpch, col
daft_b200.init(0)
li = tpch.generate("lineitem", 10.0, tpch.Q1_COLUMNS)
pred, gb, aggs = tpch.q1_exprs()
def t(name, fn, reps=20):
    for _ in range(3): r = fn()
    _lib.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): r = fn()
    _lib.synchronize(); dt = (time.perf_counter() - t0) / reps * 1e3
    print(json.dumps({"op": name, "ms": round(dt, 4)}), flush=True); return r
part = t("q1 agg (60M rows)", lambda: li.agg(aggs, gb, predicate=pred))
t("sort 4 rows x 10 cols", lambda: part.sort([col("L_RETURNFLAG"), col("L_LINESTATUS")]))
p2 = t("q1_partial", lambda: tpch.q1_partial(li))
host = t("to_arrow 4 rows", lambda: p2.to_arrow_table())
dev = t("from_arrow 4 rows", lambda: daft_b200.RecordBatch.from_arrow_table(host))
t("q1_final on 4 rows", lambda: tpch.q1_final(dev))
fgb, faggs, fproj = tpch.q1_final_exprs()
a = t("  final agg only", lambda: dev.agg(faggs, fgb))
b = t("  final project only", lambda: a.eval_expression_list(fproj))
t("  final sort only", lambda: b.sort([col("L_RETURNFLAG"), col("L_LINESTATUS")]))
t("hash_rows 2 str keys 60M", lambda: li.hash_rows([col("L_RETURNFLAG"), col("L_LINESTATUS")]), 5)
t("launch_count call", lambda: daft_b200.launch_count())

"""Dev probe: sort / partition / string-hash kernels on SF10 columns (run under ncu for per-kernel times; the printed
event times are of the un-profiled calls only when run without ncu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import _lib, tpch, col

daft_b200.init(0)
li = tpch.generate("lineitem", 10.0, ["L_ORDERKEY", "L_EXTENDEDPRICE", "L_SUPPKEY", "L_RETURNFLAG", "L_LINESTATUS"])
def timed(name, fn, reps=5):
    for _ in range(2): fn()
    _lib.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    _lib.synchronize(); print(name, round((time.perf_counter() - t0) / reps * 1e3, 3), "ms", flush=True)
reps = int(os.environ.get("KP_REPS", "5"))
keys = li.eval_expression_list([col("L_ORDERKEY")])
timed("argsort_i64", lambda: keys.argsort([col("L_ORDERKEY")]), reps)
fk = li.eval_expression_list([col("L_EXTENDEDPRICE")])
timed("argsort_f64", lambda: fk.argsort([col("L_EXTENDEDPRICE")]), reps)
three = li.eval_expression_list([col("L_ORDERKEY"), col("L_EXTENDEDPRICE"), col("L_SUPPKEY")])
timed("partition8", lambda: three.partition_by_hash([col("L_ORDERKEY")], 8), reps)
timed("hash_2flags", lambda: li.hash_rows([col("L_RETURNFLAG"), col("L_LINESTATUS")]), reps)

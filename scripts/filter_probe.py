"""Dev probe: fused filter compaction on SF10 lineitem (run under ncu for the per-pass launch list)."""
import sys, os, datetime
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import tpch, col

daft_b200.init(0)
li = tpch.generate("lineitem", 10.0, ["L_QUANTITY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_TAX", "L_SHIPDATE", "L_RETURNFLAG"])
sel = li.eval_expression_list([col("L_QUANTITY"), col("L_EXTENDEDPRICE"), col("L_DISCOUNT"), col("L_TAX"), col("L_SHIPDATE")])
for cutoff in (datetime.date(1998, 9, 2), datetime.date(1995, 6, 17), datetime.date(1992, 3, 1)):
    for _ in range(2):
        out = sel.filter([col("L_SHIPDATE") <= cutoff])
    print(cutoff, len(out))
out = li.filter([col("L_SHIPDATE") <= datetime.date(1995, 6, 17)])  # with a string column (host gather path)
print("with string col", len(out))

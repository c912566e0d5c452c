"""Dev probe: argsort / join build+probe on SF10 keys (run under ncu for per-kernel times)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import tpch, col
from daft_b200.join import ProbeTable

daft_b200.init(0)
li = tpch.generate("lineitem", 10.0, ["L_ORDERKEY", "L_EXTENDEDPRICE"])
orders = tpch.generate("orders", 10.0, ["O_ORDERKEY", "O_CUSTKEY"])
s1 = li.eval_expression_list([col("L_ORDERKEY")])
for _ in range(2):
    idx = s1.argsort([col("L_ORDERKEY")])
print("argsort", len(idx))
for _ in range(2):
    pt = ProbeTable(1)
    pt.add_table(orders, [0])
    pt.build()
    out = pt.probe_indices(s1, [0])
print("probe", len(out))

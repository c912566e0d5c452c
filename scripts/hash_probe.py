"""Dev probe: hash_rows over the two TPC-H flag string columns (run under ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import daft_b200
from daft_b200 import tpch, col

daft_b200.init(0)
li = tpch.generate("lineitem", 10.0, ["L_RETURNFLAG", "L_LINESTATUS", "L_ORDERKEY"])
for _ in range(3):
    h = li.hash_rows([col("L_RETURNFLAG"), col("L_LINESTATUS")])
print(len(h))
parts = li.partition_by_hash([col("L_ORDERKEY")], 8)
print([len(p) for p in parts])

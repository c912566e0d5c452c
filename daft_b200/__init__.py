"""daft_b200 -- Daft's columnar hot path (hash / partition / group-by aggregate / hash join / sort /
filter / project / exchange) as hand-written sm_100a CUDA kernels behind a C ABI.

Public surface: RecordBatch (device-resident), col / lit expressions, ProbeTable, the TPC-H
generator and query pipelines in daft_b200.tpch, the aggregation stage rewrite and sink in
daft_b200.aggstages, the multi-GPU communicator (exchange, all-gather, distributed aggregate) in daft_b200.distributed,
sort-merge / asof / cross joins (daft_b200.join), window functions and pivot (daft_b200.window), the Parquet -> HBM scan
(daft_b200.scan.read_parquet).
"""
from . import datatypes
from ._lib import B200Error, init, launch_count, synchronize
from .expressions import Expression, col, lit
from .recordbatch import RecordBatch
from .scan import parquet_file_info, read_parquet

__all__ = ["RecordBatch", "Expression", "col", "lit", "init", "B200Error", "datatypes", "launch_count", "synchronize", "read_parquet", "parquet_file_info"]

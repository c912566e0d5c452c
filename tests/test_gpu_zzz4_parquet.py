"""The Parquet -> HBM scan on the CUDA path against pyarrow's reader; the same files go through the host build of
the same decoders in tests/test_parquet_cpu.py."""
import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from daft_b200 import tpch
from helpers import assert_tables_equal
from test_parquet_cpu import mixed_table

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,nulls", [(0, 0.0), (1, 0.0), (1000, 0.2), (50000, 0.05)])
@pytest.mark.parametrize("opts", [
    dict(compression="snappy"), dict(compression="none", use_dictionary=False), dict(compression="snappy", data_page_version="2.0"),
    dict(compression="snappy", data_page_size=4096, row_group_size=7000, dictionary_pagesize_limit=2048),
])
def test_read_parquet_matches_pyarrow(gpu, tmp_path, n, nulls, opts):
    rng = np.random.default_rng(n + int(nulls * 100))
    tbl = mixed_table(rng, n, nulls)
    path = str(tmp_path / "t.parquet")
    pq.write_table(tbl, path, **opts)
    got = gpu.read_parquet(path).to_arrow_table()
    assert_tables_equal(got, pq.read_table(path).cast(got.schema))
    groups, rows = gpu.parquet_file_info(path)
    assert rows == n and groups == pq.ParquetFile(path).metadata.num_row_groups
    if groups > 2:
        sub = gpu.read_parquet(path, ["s", "i64", "b"], (1, groups - 1)).to_arrow_table()
        pf = pq.ParquetFile(path)
        want = pa.concat_tables([pf.read_row_group(g, columns=["s", "i64", "b"]) for g in range(1, groups - 1)])
        assert_tables_equal(sub, want.cast(sub.schema))


def test_q1_and_q6_from_parquet(gpu, tmp_path):
    """scan -> query: the same answers as over the generated table"""
    host = tpch.generate_host("lineitem", 0.05, tpch.Q1_COLUMNS)
    path = str(tmp_path / "lineitem.parquet")
    pq.write_table(host, path, compression="snappy", row_group_size=100000)
    scanned = gpu.read_parquet(path, tpch.Q1_COLUMNS)
    direct = gpu.RecordBatch.from_arrow_table(host)
    assert_tables_equal(scanned.to_arrow_table(), host)
    assert_tables_equal(tpch.q1(scanned).to_arrow_table(), tpch.q1(direct).to_arrow_table(), float_rtol=1e-12)   # the group table's atomics add in any order
    assert_tables_equal(tpch.q6(scanned).to_arrow_table(), tpch.q6(direct).to_arrow_table(), float_rtol=1e-12)


def test_read_parquet_refusals(gpu, tmp_path):
    path = str(tmp_path / "z.parquet")
    pq.write_table(pa.table({"a": [1, 2, 3]}), path, compression="zstd")
    with pytest.raises(gpu.B200Error, match="compression codec 6 is not supported"):
        gpu.read_parquet(path)
    pq.write_table(pa.table({"a": [[1, 2], [3]]}), path)
    with pytest.raises(gpu.B200Error, match="nested"):
        gpu.read_parquet(path)
    with pytest.raises(gpu.B200Error, match="cannot open"):
        gpu.read_parquet(str(tmp_path / "missing.parquet"))

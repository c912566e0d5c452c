"""TPC-H pipelines (benchmarking/tpch/answers.py q1, q5, q6) on the CUDA path vs the same pipelines on the
oracle, on the same generated tables; plus size-independent properties at a larger scale."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal
from oracle.recordbatch import OracleRecordBatch

pytestmark = pytest.mark.gpu
SF = 0.02


def test_device_and_host_generators_agree(gpu):
    from daft_b200 import tpch
    for table in ("lineitem", "orders", "customer", "supplier", "nation", "region"):
        host = tpch.generate_host(table, SF)
        dev = tpch.generate(table, SF).to_arrow_table()
        assert_tables_equal(dev, host)
    part = tpch.generate("lineitem", SF, tpch.Q1_COLUMNS, 1000, 5000).to_arrow_table()
    assert_tables_equal(part, tpch.generate_host("lineitem", SF, tpch.Q1_COLUMNS).slice(1000, 4000))


def test_q1_q6_match_oracle(gpu):
    from daft_b200 import tpch
    host = tpch.generate_host("lineitem", SF, tpch.Q1_COLUMNS)
    dev = gpu.RecordBatch.from_arrow_table(host)
    assert_tables_equal(tpch.q1(dev).to_arrow_table(), tpch.q1(OracleRecordBatch(host)).to_arrow_table(), float_rtol=1e-9)
    assert_tables_equal(tpch.q6(dev).to_arrow_table(), tpch.q6(OracleRecordBatch(host)).to_arrow_table(), float_rtol=1e-9)
    # two-stage form (what N>1 runs) == single-stage form
    assert_tables_equal(tpch.q1_final(tpch.q1_partial(dev)).to_arrow_table(), tpch.q1(dev).to_arrow_table(), float_rtol=1e-12)


def test_q5_matches_oracle(gpu):
    from daft_b200 import tpch
    host = {t: tpch.generate_host(t, SF, cols) for t, cols in tpch.Q5_COLUMNS.items()}
    got = tpch.q5({t: gpu.RecordBatch.from_arrow_table(v) for t, v in host.items()}).to_arrow_table()
    want = tpch.q5({t: OracleRecordBatch(v) for t, v in host.items()}).to_arrow_table()
    assert want.num_rows == 5
    assert_tables_equal(got, want, float_rtol=1e-9)


def test_q1_properties_at_sf1(gpu):
    """Too big for the Python oracle to be quick: check size-independent properties instead -- counts add up to the
    filtered row count, shard-additivity of every sum (linearity), idempotent final merge."""
    from daft_b200 import col, tpch
    li = tpch.generate("lineitem", 1.0, tpch.Q1_COLUMNS)
    n = len(li)
    full = tpch.q1(li).to_arrow_table()
    import datetime
    kept = li.agg([col("L_QUANTITY").count("all").alias("n")], [], predicate=[col("L_SHIPDATE") <= datetime.date(1998, 9, 2)]).to_pydict()["n"][0]
    assert sum(full.column("count_order").to_pylist()) == kept <= n
    halves = [tpch.q1_partial(li.slice(0, n // 3)), tpch.q1_partial(li.slice(n // 3, n))]
    merged = tpch.q1_final(gpu.RecordBatch.concat(halves)).to_arrow_table()
    assert merged.column("sum_qty").to_pylist() == full.column("sum_qty").to_pylist()          # integer sums: exact
    assert merged.column("count_order").to_pylist() == full.column("count_order").to_pylist()
    assert_tables_equal(merged, full, float_rtol=1e-9)

"""The remaining TPC-H queries of benchmarking/tpch/answers.py (Q2, Q7-Q11, Q13, Q15-Q17, Q19-Q22) on the CUDA path: each is
held to the oracle twin and to SQLite running the query's SQL on the same generated tables (tests/integration/conftest.py:31-53).
Runs last in the suite (file name): these compose operators that the files before it test one by one."""
import pytest

import sqlite_truth as ST
from daft_b200 import tpch
from helpers import assert_tables_equal

pytestmark = pytest.mark.gpu
SF = {2: 0.05, 8: 0.05, 16: 0.05, 17: 0.05, 20: 0.05}
REST = [2, 7, 8, 9, 10, 11, 13, 15, 16, 17, 19, 20, 21, 22]


@pytest.mark.parametrize("q", REST)
def test_remaining_queries_vs_sqlite_and_oracle(gpu, q):
    from oracle.recordbatch import OracleRecordBatch
    name = f"q{q}"
    sf = SF.get(q, 0.01)
    host = {t: tpch.generate_host(t, sf, c) for t, c in tpch.QUERY_COLUMNS[name].items()}
    dev = {t: gpu.RecordBatch.from_arrow_table(v) for t, v in host.items()}
    got = getattr(tpch, name)(dev).to_arrow_table()
    want = getattr(tpch, name)({t: OracleRecordBatch(v) for t, v in host.items()}).to_arrow_table()
    assert_tables_equal(got, want, float_rtol=1e-9)
    ST.assert_matches(got, *ST.run(ST.load(host), q), check_names=False, rel=1e-8)


@pytest.mark.parametrize("table", ["supplier", "customer", "part", "partsupp", "orders", "lineitem"])
def test_device_generator_matches_host_generator_on_every_column(gpu, table):
    """names / phones / addresses / comments (two-pass generated strings), the new dictionary columns and partsupp"""
    n = tpch.num_rows(table, 0.01)
    for rb, re_ in ((0, -1), (n // 3 + 1, n - n // 4)):
        host = tpch.generate_host(table, 0.01, None, rb, re_)
        dev = tpch.generate(table, 0.01, None, rb, re_)
        assert_tables_equal(dev.to_arrow_table(), host)

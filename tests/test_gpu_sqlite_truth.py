"""The CUDA path's TPC-H Q1 / Q5 / Q6 answers held to SQLite's on the same generated tables (tests/integration/conftest.py:31-53)."""
import pytest

import sqlite_truth as ST
from daft_b200 import tpch

pytestmark = pytest.mark.gpu
SF = 0.01


def test_q1_q5_q6_vs_sqlite(gpu):
    cols = dict(tpch.Q5_COLUMNS)
    cols["lineitem"] = sorted(set(tpch.Q5_COLUMNS["lineitem"]) | set(tpch.Q1_COLUMNS))
    host = {t: tpch.generate_host(t, SF, c) for t, c in cols.items()}
    conn = ST.load(host)
    dev = {t: gpu.RecordBatch.from_arrow_table(v) for t, v in host.items()}
    ST.assert_matches(tpch.q1(dev["lineitem"]).to_arrow_table(), *ST.run(conn, 1))
    ST.assert_matches(tpch.q5(dev).to_arrow_table(), *ST.run(conn, 5))
    ST.assert_matches(tpch.q6(dev["lineitem"]).to_arrow_table(), *ST.run(conn, 6))


@pytest.mark.parametrize("q", [3, 4, 12, 14, 18])
def test_composed_queries_vs_sqlite_and_oracle(gpu, q):
    """Q3 / Q4 / Q12 / Q14 / Q18 of benchmarking/tpch/answers.py on the CUDA path: held to SQLite and to the oracle twin."""
    from oracle.recordbatch import OracleRecordBatch
    from helpers import assert_tables_equal
    name = f"q{q}"
    sf = 0.05 if q == 18 else SF
    host = {t: tpch.generate_host(t, sf, c) for t, c in tpch.QUERY_COLUMNS[name].items()}
    dev = {t: tpch.generate(t, sf, c) for t, c in tpch.QUERY_COLUMNS[name].items()}
    for t in host:
        assert_tables_equal(dev[t].to_arrow_table(), host[t])       # device generator == host generator (dictionary strings too)
    got = getattr(tpch, name)(dev).to_arrow_table()
    want = getattr(tpch, name)({t: OracleRecordBatch(v) for t, v in host.items()}).to_arrow_table()
    assert_tables_equal(got, want, float_rtol=1e-9)
    ST.assert_matches(got, *ST.run(ST.load(host), q), check_names=False)

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

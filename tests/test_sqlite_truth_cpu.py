"""The oracle's and the CPU engine's TPC-H answers held to SQLite's on the same generated tables (SF 0.01): pins the
query-level answers to something other than our own restatement (tests/integration/conftest.py:31-53)."""
import pytest

import sqlite_truth as ST
from daft_b200 import tpch
from oracle import cpu_engine as CE
from oracle.recordbatch import OracleRecordBatch

SF = 0.01


@pytest.fixture(scope="module")
def world():
    cols = dict(tpch.Q5_COLUMNS)
    cols["lineitem"] = sorted(set(tpch.Q5_COLUMNS["lineitem"]) | set(tpch.Q1_COLUMNS))
    tables = {t: tpch.generate_host(t, SF, c) for t, c in cols.items()}
    return tables, ST.load(tables)


def test_q1_oracle_and_cpu_engine_vs_sqlite(world):
    tables, conn = world
    names, rows = ST.run(conn, 1)
    ST.assert_matches(tpch.q1(OracleRecordBatch(tables["lineitem"])).to_arrow_table(), names, rows)
    ST.assert_matches(CE.q1(CE.generate_columns("lineitem", SF, CE.Q1_COLUMNS))[0], names, rows)


def test_q5_oracle_and_cpu_engine_vs_sqlite(world):
    tables, conn = world
    names, rows = ST.run(conn, 5)
    assert len(rows) == 5
    ST.assert_matches(tpch.q5({t: OracleRecordBatch(v) for t, v in tables.items()}).to_arrow_table(), names, rows)
    ST.assert_matches(CE.q5({t: CE.generate_columns(t, SF, c) for t, c in CE.Q5_COLUMNS.items()})[0], names, rows)


def test_q6_oracle_vs_sqlite(world):
    tables, conn = world
    names, rows = ST.run(conn, 6)
    ST.assert_matches(tpch.q6(OracleRecordBatch(tables["lineitem"])).to_arrow_table(), names, rows)

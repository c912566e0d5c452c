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


# ---- more of answers.py: the oracle twin of every composed query against SQLite on the same tables -------------------
def _tables(q):
    sf = 0.05 if q == "q18" else SF   # Q18's `sum(l_quantity) > 300` keeps 0.07% of the seven-line orders: needs more of them
    return {t: tpch.generate_host(t, sf, c) for t, c in tpch.QUERY_COLUMNS[q].items()}


@pytest.mark.parametrize("q", [3, 4, 12, 14, 18])
def test_composed_queries_oracle_vs_sqlite(q):
    name = f"q{q}"
    tables = _tables(name)
    conn = ST.load(tables)
    names, rows = ST.run(conn, q)
    got = getattr(tpch, name)({t: OracleRecordBatch(v) for t, v in tables.items()}).to_arrow_table()
    assert len(rows) > 0
    ST.assert_matches(got, names, rows, check_names=False)

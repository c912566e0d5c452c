"""The morsel-parallel CPU engine (oracle/cpu_engine.cpp: bench.py's `--impl reference` arm and `cpu_baseline`) against the
column-at-a-time oracle and the committed TPC-H goldens; and its table generator against the product's host generator."""
import json
import os

import pytest

from helpers import assert_tables_equal
from oracle import cpu_engine as CE

HERE = os.path.dirname(os.path.abspath(__file__))
SF = 0.01


def test_generator_matches_the_product_generator():
    from daft_b200 import tpch
    for table, cols in list(tpch.Q5_COLUMNS.items()) + [("lineitem", tpch.Q1_COLUMNS)]:
        assert_tables_equal(CE.generate_table(table, SF, cols), tpch.generate_host(table, SF, cols))
    assert_tables_equal(CE.generate_table("lineitem", SF, tpch.Q1_COLUMNS, 1000, 4000), tpch.generate_host("lineitem", SF, tpch.Q1_COLUMNS, 1000, 4000))
    for t in ("lineitem", "orders", "customer", "supplier", "nation", "region"):
        for sf in (0.01, 1.0, 10.0, 100.0):
            assert CE.num_rows(t, sf) == tpch.num_rows(t, sf)


@pytest.mark.parametrize("threads", [1, 3, 0])
def test_q1_matches_golden_and_oracle(threads):
    g = json.load(open(os.path.join(HERE, "golden", "tpch_q1_q6_sf001.json")))
    cols = CE.generate_columns("lineitem", g["sf"], CE.Q1_COLUMNS)
    got, tasks = CE.q1(cols, threads)
    assert tasks == (g["lineitem_rows"] + 131071) // 131072
    d = got.to_pydict()
    for k, v in g["q1"].items():
        if isinstance(v[0], float):
            assert d[k] == pytest.approx(v, rel=1e-9), k  # morsel partials are merged in a different order
        else:
            assert d[k] == v, k


@pytest.mark.parametrize("threads", [1, 4])
def test_q5_matches_the_oracle_pipeline(threads):
    from daft_b200 import tpch
    from oracle.recordbatch import OracleRecordBatch
    sf = 0.02
    host = {t: CE.generate_table(t, sf, cols) for t, cols in tpch.Q5_COLUMNS.items()}
    want = tpch.q5({t: OracleRecordBatch(v) for t, v in host.items()}).to_arrow_table()
    cols = {t: CE.generate_columns(t, sf, c) for t, c in CE.Q5_COLUMNS.items()}
    got, tasks, probes = CE.q5(cols, threads)
    assert_tables_equal(got, want, float_rtol=1e-9)
    assert probes >= CE.num_rows("lineitem", sf) and tasks >= 3

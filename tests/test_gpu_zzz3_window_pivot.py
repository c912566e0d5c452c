"""Window functions and pivot on the CUDA path against the oracle's per-partition restatement of the reference
kernels (oracle/window.py); the same cases run on the CPU through the host emulation in tests/test_window_pivot_cpu.py."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal
from oracle import oracle as O
from oracle import window as OW
from test_window_pivot_cpu import FUNCS, _sorted, window_table

pytestmark = pytest.mark.gpu


def test_window_reference_literals(gpu):
    c = gpu.col
    t = pa.table({"k": pa.array(["a", "a", "a", "b", "b"], pa.large_string()), "v": [10, 10, 20, 5, 7], "id": [0, 1, 2, 3, 4]})
    got = gpu.RecordBatch.from_arrow_table(t).window([c("k")], [c("v"), c("id")], functions=[
        ("row_number", "rn"), ("offset", "nx", c("v"), 1), ("offset", "pv", c("v"), -1, -1), ("agg", "tot", c("v").sum())]).to_arrow_table()
    got = _sorted(got, ["id"]).to_pydict()
    assert got["rn"] == [1, 2, 3, 1, 2] and got["nx"] == [10, 20, None, 7, None] and got["pv"] == [-1, 10, 10, -1, 5] and got["tot"] == [40, 40, 40, 12, 12]
    got = gpu.RecordBatch.from_arrow_table(t).window([c("k")], [c("v")], functions=[("rank", "rk"), ("dense_rank", "dr")]).to_arrow_table()
    got = _sorted(got, ["id"]).to_pydict()
    assert got["rk"] == [1, 1, 3, 1, 2] and got["dr"] == [1, 1, 2, 1, 2]


@pytest.mark.parametrize("n", [0, 1, 37, 400, 20000])
@pytest.mark.parametrize("shape", ["partition+order", "order only", "partition only", "two keys desc"])
def test_window_matches_oracle(gpu, n, shape):
    rng = np.random.default_rng(n + len(shape))
    tbl = window_table(rng, n)
    c = gpu.col
    pb, ob, desc, nf = {"partition+order": ([c("g")], [c("o")], [False], [False]), "order only": ([], [c("o"), c("h")], [False, True], [False, True]),
                        "partition only": ([c("g"), c("h")], [], [], []),
                        "two keys desc": ([c("h")], [c("o"), c("x")], [True, False], [True, False])}[shape]
    funcs = [f for f in FUNCS if pb or f[0] != "agg"]
    rb = gpu.RecordBatch.from_arrow_table(tbl)
    if ob:
        ob2, d2, n2 = ob + [c("id")], desc + [False], nf + [False]     # a tie-free order for the position-dependent functions
        got = rb.window(pb, ob2, d2, n2, funcs).to_arrow_table()
        want = OW.window(tbl, pb, ob2, d2, n2, funcs)
        assert got.column_names == want.column_names
        assert_tables_equal(_sorted(got, ["id"]), _sorted(want, ["id"]), float_rtol=1e-9)
        got_r = rb.window(pb, ob, desc, nf, [("rank", "rk"), ("dense_rank", "dr")]).to_arrow_table()
        want_r = OW.window(tbl, pb, ob, desc, nf, [("rank", "rk"), ("dense_rank", "dr")])
        assert_tables_equal(_sorted(got_r, ["id"]), _sorted(want_r, ["id"]))
    else:
        got = rb.window(pb, ob, desc, nf, funcs).to_arrow_table()
        want = OW.window(tbl, pb, ob, desc, nf, funcs)
        keep = [k for k in got.column_names if k not in ("rn", "lead1", "lag2", "lead_s")]
        assert_tables_equal(_sorted(got.select(keep), ["id"]), _sorted(want.select(keep), ["id"]), float_rtol=1e-9)


@pytest.mark.parametrize("ptype", ["str", "int", "bool"])
def test_pivot_matches_oracle(gpu, ptype):
    rng = np.random.default_rng(5)
    n = 3000
    g1, g2 = rng.integers(0, 60, n), rng.integers(0, 2, n)
    pv = {"str": np.array([f"c{v}" for v in rng.integers(0, 4, n)], dtype=object), "int": rng.integers(0, 4, n), "bool": rng.integers(0, 2, n).astype(bool)}[ptype]
    tbl = pa.table({"g1": pa.array(g1, mask=rng.random(n) < 0.1), "g2": g2, "p": pa.array(pv, type=pa.large_string()) if ptype == "str" else pv,
                    "v": rng.standard_normal(n)})
    names = {"str": ["c0", "c3", "zz", "c1"], "int": ["0", "3", "9", "x"], "bool": ["true", "false"]}[ptype]
    c = gpu.col
    got = gpu.RecordBatch.from_arrow_table(tbl).agg_pivot([c("g1"), c("g2")], c("p"), c("v"), "sum", names).to_arrow_table()
    pre = O.agg(tbl, [("sum", tbl.column("v"), "v", "valid")], [("g1", tbl.column("g1")), ("g2", tbl.column("g2")), ("p", tbl.column("p"))])
    want = OW.pivot(pre, ["g1", "g2"], "p", "v", names)
    assert_tables_equal(_sorted(got, ["g1", "g2"]), _sorted(want, ["g1", "g2"]), float_rtol=1e-9)
    # the literal case of tests/recordbatch/test_pivot.py
    t = pa.table({"group": pa.array(["A", "A", "B"], pa.large_string()), "pivot": pa.array(["1", "2", "1"], pa.large_string()), "value": [1, 2, 3]})
    got = gpu.RecordBatch.from_arrow_table(t).pivot([c("group")], c("pivot"), c("value"), ["1", "2", "3"]).to_arrow_table()
    assert got.to_pydict() == {"group": ["A", "B"], "1": [1, 3], "2": [2, None], "3": [None, None]}


@pytest.mark.parametrize("n", [0, 1, 600, 30000])
def test_percentile_median_and_approx_count_distinct_match_oracle(gpu, n):
    """exact_percentile (src/daft-core/src/utils/stats.rs:31-59) through the window-based composition; the reference's fused
    multiply-add in the interpolation allows one ulp"""
    from helpers import oracle_agg
    rng = np.random.default_rng(n)
    tbl = window_table(rng, n, nulls=0.15)
    c = gpu.col
    aggs = [c("x").median().alias("med"), c("o").percentile(0.9).alias("p90"), c("x").sum().alias("sx"), c("h").percentile(0.0).alias("p0"),
            (c("x") * 2 + c("h")).percentile(0.33).alias("pe"), c("s").approx_count_distinct().alias("acd")]
    rb = gpu.RecordBatch.from_arrow_table(tbl)
    for gb in ([c("g")], [c("g"), c("h")], []):
        got = rb.agg(aggs, gb).to_arrow_table()
        want = oracle_agg(tbl, aggs, gb)
        keys = [g.name() for g in gb]
        if keys:
            assert_tables_equal(_sorted(got, keys), _sorted(want, keys), float_rtol=1e-9)
        else:
            assert_tables_equal(got, want, float_rtol=1e-9)
    with pytest.raises(ValueError, match="between 0 and 1"):
        c("x").percentile(1.5)

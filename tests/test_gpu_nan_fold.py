"""Grouped float min / max reproduce the reference's left fold `if a < b {a} else {b}` (inline path
src/daft-recordbatch/src/ops/inline_agg.rs:271-298, fallback src/daft-core/src/array/ops/compare_agg.rs:75-95): a NaN makes
the fold forget everything before it, so the result is the min / max of the values AFTER the group's last NaN, and NaN when
the group's last valid value is one.  The oracle restates the fold row by row (oracle/daft_oracle.c DEF_MINMAX)."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal, oracle_agg

pytestmark = pytest.mark.gpu
NAN = float("nan")


def RB(gpu, tbl):
    return gpu.RecordBatch.from_arrow_table(tbl)


def test_reference_fixture_float_with_nan(gpu):
    """make_float_with_nan_test_batch, inline_agg.rs:1347-1374: key 1 holds [1.0, NaN, null], key 2 holds [3.0, 2.0]."""
    tbl = pa.table({"key": pa.array([1, 1, 2, 2, 1], type=pa.int64()), "val": pa.array([1.0, NAN, 3.0, 2.0, None], type=pa.float64())})
    c = gpu.col
    got = RB(gpu, tbl).agg([c("val").min().alias("mn"), c("val").max().alias("mx")], [c("key")]).to_pydict()
    assert got["key"] == [1, 2]
    assert np.isnan(got["mn"][0]) and np.isnan(got["mx"][0])      # the fold ends on the NaN
    assert got["mn"][1] == 2.0 and got["mx"][1] == 3.0
    want = oracle_agg(tbl, [c("val").min().alias("mn"), c("val").max().alias("mx")], [c("key")])
    assert_tables_equal(RB(gpu, tbl).agg([c("val").min().alias("mn"), c("val").max().alias("mx")], [c("key")]).to_arrow_table(), want)


def test_nan_in_the_middle_is_forgotten(gpu):
    vals = {"a": [NAN, 1.0], "b": [5.0, NAN, 7.0, 6.0], "c": [NAN], "d": [2.0, 1.0, NAN, NAN, 9.0, None, 8.0], "e": [None, None], "f": [4.0, 3.0]}
    keys, col = [], []
    for i in range(7):  # interleave the groups' rows
        for k, v in vals.items():
            if i < len(v):
                keys.append(k)
                col.append(v[i])
    tbl = pa.table({"k": pa.array(keys, type=pa.large_utf8()), "v": pa.array(col, type=pa.float64()),
                    "w": pa.array([None if x is None else np.float32(x) for x in col], type=pa.float32())})
    c = gpu.col
    aggs = [c("v").min().alias("mn"), c("v").max().alias("mx"), c("w").min().alias("mn32"), c("w").max().alias("mx32"), c("v").count().alias("n")]
    got = RB(gpu, tbl).agg(aggs, [c("k")]).to_arrow_table()
    d = {k: (mn, mx) for k, mn, mx in zip(got.column("k").to_pylist(), got.column("mn").to_pylist(), got.column("mx").to_pylist())}
    assert d["a"] == (1.0, 1.0) and d["b"] == (6.0, 7.0) and d["d"] == (8.0, 9.0) and d["f"] == (3.0, 4.0)
    assert np.isnan(d["c"][0]) and np.isnan(d["c"][1]) and d["e"] == (None, None)
    assert_tables_equal(got, oracle_agg(tbl, aggs, [c("k")]))
    # ungrouped: the same fold over the whole column
    got = RB(gpu, tbl).agg(aggs, []).to_arrow_table()
    assert_tables_equal(got, oracle_agg(tbl, aggs, []))


@pytest.mark.parametrize("n,groups,nan_frac", [(5000, 7, 0.05), (200_000, 3, 0.001), (300_000, 50_000, 0.02), (100_000, 100_000, 0.3)])
def test_random_nan_folds_match_the_oracle(gpu, n, groups, nan_frac):
    """low cardinality (shared-memory tier), mid and high cardinality (global table), with nulls."""
    rng = np.random.default_rng(n + groups)
    v = rng.standard_normal(n) * 100
    v[rng.random(n) < nan_frac] = np.nan
    tbl = pa.table({"k": pa.array(rng.integers(0, groups, n).astype(np.int64)), "v": pa.array(v, mask=rng.random(n) < 0.1),
                    "f": pa.array(v.astype(np.float32), mask=rng.random(n) < 0.05), "i": pa.array(rng.integers(-50, 50, n).astype(np.int32))})
    c = gpu.col
    aggs = [c("v").min().alias("mn"), c("v").max().alias("mx"), c("f").min().alias("mnf"), c("f").max().alias("mxf"), (c("v") * 2.0).min().alias("mn2"),
            c("i").min().alias("mni"), c("v").sum().alias("s"), c("v").count().alias("cnt")]
    for pred in (None, [c("i") > 0]):
        got = RB(gpu, tbl).agg(aggs, [c("k")], predicate=pred).to_arrow_table()
        want = oracle_agg(tbl, aggs, [c("k")], predicate=pred)
        # the sum meets NaNs too: NaN == NaN for this comparison (assert_tables_equal), finite sums within 1e-9
        assert_tables_equal(got, want, float_rtol=1e-9)

"""Parity at sizes that reach the large-input code paths: the TMA-staged and full-chunk aggregate builds
(>= 65 536 rows), the tiled filter compaction, direct-scatter partitioning (>= 4 096 rows), the fused
multi-column gather, radix sort passes, selection top-N and the bucket-table join -- each against the oracle
on the same seeded table (bit-exact; float sums within 1e-9 relative because the summation order differs)."""
import datetime

import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal, oracle_agg, random_table
from oracle import expr_eval as E
from oracle import oracle as O

pytestmark = pytest.mark.gpu

N = 150_000


def RB(gpu, tbl):
    return gpu.RecordBatch.from_arrow_table(tbl)


@pytest.fixture(scope="module")
def big(gpu):
    rng = np.random.default_rng(2024)
    tbl = random_table(rng, N, 0.05)
    # TPC-H-like one-byte flag columns (fixed-length string specialisation) and a two-byte code column
    flags = np.array(["A", "N", "R"])[rng.integers(0, 3, N)]
    codes = np.array(["OK", "NO", "XX", "ZZ"])[rng.integers(0, 4, N)]
    tbl = tbl.append_column("flag", pa.array(flags, type=pa.large_utf8())).append_column("code", pa.array(codes, type=pa.large_utf8()))
    return tbl, RB(gpu, tbl)


AGG_CASES = [
    ([], ["f64:sum", "i64:sum", "i32:min", "f32:max", "u8:count"], None),
    (["flag"], ["f64:sum", "i64:sum", "f64:mean", "i8:count"], None),
    (["flag", "code"], ["f64:sum", "u32:sum", "d:min", "d:max", "i16:mean"], "d<=9500"),
    (["k"], ["f64:sum", "f32:sum", "b:count", "u64:max"], "i32>0"),
    (["u8", "b"], ["i64:min", "i64:max", "f64:mean"], None),
    (["i8", "flag"], ["u16:sum", "f64:min"], "f64>-500"),
    (["s"], ["i32:sum", "f64:max"], None),
    (["i32"], ["f64:sum", "i32:count"], None),          # ~2000 groups: global-table build
    (["u32", "u8"], ["i64:sum"], "b"),                  # ~50 000 groups
    (["d"], ["f64:sum", "k:sum"], None),
]


def _agg_expr(gpu, spec):
    name, op = spec.split(":")
    c = gpu.col(name)
    e = {"sum": c.sum, "min": c.min, "max": c.max, "mean": c.mean, "count": c.count}[op]()
    return e.alias(f"{op}_{name}")


def _pred(gpu, p):
    c = gpu.col
    return {None: None, "d<=9500": [c("d") <= datetime.date(1996, 1, 5)], "i32>0": [c("i32") > 0], "f64>-500": [c("f64") > -500.0],
            "b": [c("b")]}[p]


@pytest.mark.parametrize("case", range(len(AGG_CASES)))
def test_large_aggregate(gpu, big, case):
    tbl, rb = big
    keys, aggs, pred = AGG_CASES[case]
    aggs_e = [_agg_expr(gpu, a) for a in aggs]
    keys_e = [gpu.col(k) for k in keys]
    p = _pred(gpu, pred)
    got = rb.agg(aggs_e, keys_e, predicate=p).to_arrow_table()
    want = oracle_agg(tbl, aggs_e, keys_e, p)
    assert_tables_equal(got, want, float_rtol=1e-9)


@pytest.mark.parametrize("pred", ["d<=9500", "i32>0", "b"])
def test_large_filter(gpu, big, pred):
    tbl, rb = big
    p = _pred(gpu, pred)
    got = rb.filter(p).to_arrow_table()
    assert_tables_equal(got, E.filter(tbl, p))


@pytest.mark.parametrize("P", [2, 8, 200, 300])
def test_large_partition(gpu, big, P):
    tbl, rb = big
    sub = tbl.select(["i64", "f64", "u8", "flag", "b", "k"])
    got = RB(gpu, sub).partition_by_hash([gpu.col("i64"), gpu.col("flag")], P)
    want = O.partition_by_hash(sub, ["i64", "flag"], P)
    assert len(got) == P
    for g, w in zip(got, want):
        assert_tables_equal(g.to_arrow_table(), w)
    # partitions feed further operators (views into one buffer): aggregate one of them
    g0 = got[0].agg([gpu.col("f64").sum().alias("s")], [gpu.col("k")]).to_arrow_table()
    assert_tables_equal(g0, oracle_agg(want[0], [gpu.col("f64").sum().alias("s")], [gpu.col("k")]), float_rtol=1e-9)


def test_large_sort_and_topn(gpu, big):
    tbl, rb = big
    sub = tbl.select(["u8", "i64", "f64", "flag"]).append_column("row", pa.array(np.arange(N), type=pa.int64()))
    c = gpu.col
    r = RB(gpu, sub)
    got = r.sort([c("u8"), c("flag"), c("row")], descending=[True, False, False]).to_arrow_table()
    assert_tables_equal(got, O.sort(sub, ["u8", "flag", "row"], [True, False, False], [True, False, False]))
    got = r.top_n([c("f64"), c("row")], 300, descending=[True, False], offset=17).to_arrow_table()
    assert_tables_equal(got, O.sort(sub, ["f64", "row"], [True, False], [True, False]).slice(17, 300))


@pytest.mark.parametrize("how", ["inner", "left", "anti"])
def test_large_join(gpu, big, how):
    tbl, rb = big
    rng = np.random.default_rng(9)
    nb = 40_000
    build = pa.table({"i32": pa.array(rng.integers(-1000, 1000, nb).astype(np.int32), mask=rng.random(nb) < 0.02),
                      "bv": pa.array(np.arange(nb), type=pa.int64())})
    probe = tbl.select(["i32", "f64", "flag"]).slice(0, 20_000)
    c = gpu.col
    got = RB(gpu, probe).hash_join(RB(gpu, build), [c("i32")], [c("i32")], how=how).to_arrow_table()
    want = O.hash_join(probe, build, ["i32"], ["i32"], how)
    if how == "anti":
        want = want.select(probe.column_names)
    assert_tables_equal(got, want)

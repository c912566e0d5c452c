"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded Arrow inputs.

Shapes follow the reference's kernel-level tests: tests/recordbatch/test_hash.py, test_table_aggs.py,
test_joins.py, test_sorting.py, test_partitioning.py, test_filter.py, tests/series/test_hash.py.
Bar: bit-exact for hashes / indices / keys / integer aggregates / filter+project outputs; 1e-9 relative
for float sums (north_star allows 1e-6).
"""
import datetime

import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal, oracle_agg, random_table
from oracle import expr_eval as E
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def RB(gpu, tbl):
    return gpu.RecordBatch.from_arrow_table(tbl)


# ---------------------------------------------------------------------------------------------- roundtrip
def test_roundtrip_all_dtypes(gpu):
    rng = np.random.default_rng(1)
    for n in (0, 1, 31, 32, 33, 1000):
        tbl = random_table(rng, n)
        got = RB(gpu, tbl).to_arrow_table()
        assert_tables_equal(got, tbl)


def test_roundtrip_sliced_and_small_offsets(gpu):
    tbl = pa.table({"a": pa.array([1, None, 3, 4, 5], type=pa.int32()), "s": pa.array(["x", None, "zz", "", "w"]),
                    "b": pa.array([True, False, None, True, False])})
    sl = tbl.slice(1, 3)
    got = RB(gpu, sl).to_arrow_table()
    assert_tables_equal(got, sl.combine_chunks())


# ---------------------------------------------------------------------------------------------- hashing
def test_hash_goldens(gpu):
    # tests/recordbatch/test_hash.py:15-18
    rb = RB(gpu, pa.table({"utf8": ["foo", "bar"], "int": pa.array([1, None], type=pa.int64())}))
    assert rb.hash_rows([gpu.col("utf8")]).to_pydict()["hash"] == [12352915711150947722, 15304296276065178466]
    assert rb.hash_rows([gpu.col("int")]).to_pydict()["hash"] == [3439722301264460078, 3244421341483603138]
    # :30-33 seed = 42
    seed = RB(gpu, pa.table({"s": pa.array([42, 42], type=pa.uint64())}))
    assert rb.hash_rows([gpu.col("utf8")], seed=seed).to_pydict()["hash"] == [15221504070560512414, 2671805001252040144]
    assert rb.hash_rows([gpu.col("int")], seed=seed).to_pydict()["hash"] == [16405722695416140795, 3244421341483603138]
    # :45 per-row seeds
    seed2 = RB(gpu, pa.table({"s": pa.array([1, 1000], type=pa.uint64())}))
    assert rb.hash_rows([gpu.col("utf8")], seed=seed2).to_pydict()["hash"] == [6076897603942036120, 15438169081903732554]


@pytest.mark.parametrize("n", [0, 1, 257, 5000])
def test_hash_every_dtype_matches_oracle(gpu, n):
    rng = np.random.default_rng(n)
    tbl = random_table(rng, n)
    rb = RB(gpu, tbl)
    for name in tbl.column_names:
        got = np.array(rb.hash_rows([gpu.col(name)]).to_pydict()["hash"], dtype=np.uint64)
        want = O.hash_column(tbl.column(name))
        assert (got == want).all(), name
    # chained multi-column hash (RecordBatch::hash_rows)
    keys = ["i64", "s", "b", "f64", "d", "u8"]
    got = np.array(rb.hash_rows([gpu.col(k) for k in keys]).to_pydict()["hash"], dtype=np.uint64)
    assert (got == O.hash_rows([tbl.column(k) for k in keys])).all()


def test_hash_string_length_classes(gpu):
    import xxhash
    rng = np.random.default_rng(5)
    lens = list(range(0, 260)) + [511, 512, 1023, 1024, 1025, 2048, 4100]
    vals = [bytes(rng.integers(0, 256, l, dtype=np.uint8)) for l in lens]
    tbl = pa.table({"b": pa.array(vals, type=pa.large_binary())})
    got = RB(gpu, tbl).hash_rows([gpu.col("b")]).to_pydict()["hash"]
    assert got == [xxhash.xxh3_64_intdigest(v) for v in vals]
    seeds = [int(x) for x in rng.integers(0, 2**63, len(vals), dtype=np.uint64)]
    seed = RB(gpu, pa.table({"s": pa.array(seeds, type=pa.uint64())}))
    got = RB(gpu, tbl).hash_rows([gpu.col("b")], seed=seed).to_pydict()["hash"]
    assert got == [xxhash.xxh3_64_intdigest(v, s) for v, s in zip(vals, seeds)]


@pytest.mark.parametrize("L", [1, 2, 3, 4, 5, 8])
def test_hash_fixed_length_strings(gpu, L):
    """Columns whose bytes divide evenly over the rows take the speculative fixed-length fetch of
    hash_bytes_kernel (hash.cu); each row still verifies the guess against its offsets.  Cases: truly fixed
    length, fixed length with nulls (a null slot has length 0, so later rows no longer sit at row * L), and
    ragged lengths whose total happens to be rows * L."""
    import xxhash
    rng = np.random.default_rng(L)
    n = 4099
    fixed = [bytes(rng.integers(97, 123, L, dtype=np.uint8)) for _ in range(n)]
    got = RB(gpu, pa.table({"s": pa.array(fixed, type=pa.large_binary())})).hash_rows([gpu.col("s")]).to_pydict()["hash"]
    assert got == [xxhash.xxh3_64_intdigest(v) for v in fixed]
    # ragged, same total: pair up rows (L-1, L+1)
    ragged = []
    for i in range(0, n - 1, 2):
        ragged += [bytes(rng.integers(97, 123, L - 1, dtype=np.uint8)), bytes(rng.integers(97, 123, L + 1, dtype=np.uint8))]
    ragged.append(bytes(rng.integers(97, 123, L, dtype=np.uint8)))
    assert sum(map(len, ragged)) == n * L
    seeds = [int(x) for x in rng.integers(0, 2**63, n, dtype=np.uint64)]
    seed = RB(gpu, pa.table({"s": pa.array(seeds, type=pa.uint64())}))
    got = RB(gpu, pa.table({"s": pa.array(ragged, type=pa.large_binary())})).hash_rows([gpu.col("s")], seed=seed).to_pydict()["hash"]
    assert got == [xxhash.xxh3_64_intdigest(v, s) for v, s in zip(ragged, seeds)]
    # nulls: oracle rule (null string hashes the empty slice WITH the seed)
    withnull = [None if i % 7 == 3 else v.decode() for i, v in enumerate(fixed)]
    col_ = pa.array(withnull, type=pa.large_utf8())
    got = np.array(RB(gpu, pa.table({"s": col_})).hash_rows([gpu.col("s")]).to_pydict()["hash"], dtype=np.uint64)
    assert (got == O.hash_column(pa.chunked_array([col_]))).all()


def test_hash_null_array(gpu):
    tbl = pa.table({"n": pa.nulls(5)})
    got = RB(gpu, tbl).hash_rows([gpu.col("n")]).to_pydict()["hash"]
    assert got == [3244421341483603138] * 5


# ---------------------------------------------------------------------------------------------- project / filter
def exprs_under_test(c, lit):
    return [
        (c("i64") + c("i32")).alias("e0"), (c("f64") * (1 - c("f32"))).alias("e1"), (c("i32") / c("u8")).alias("e2"),
        (c("i64") // c("i16")).alias("e3"), (c("i64") % 7).alias("e4"), (c("f64") <= 0.5).alias("e5"),
        ((c("b") & (c("i8") > 0)) | c("f64").is_null()).alias("e6"), (~c("b")).alias("e7"), (-c("i32")).alias("e8"),
        c("i32").cast(pa.float64()).alias("e9"), c("i16").cast(pa.int64()).alias("e10"),
        (c("i32") > 0).if_else(c("i64"), c("u32")).alias("e11"), c("u16").is_in([1, 2, 3, 69]).alias("e12"),
        c("i32").between(-10, 500).alias("e13"), c("i64").fill_null(-1).alias("e14"), c("d").year().alias("e15"),
        c("d").month().alias("e16"), c("d").day().alias("e17"), (c("d") < datetime.date(1995, 6, 17)).alias("e18"),
        c("s").startswith("TPC").alias("e19"), c("s").endswith("d").alias("e20"), c("s").contains("lo w").alias("e21"),
        c("s").length().alias("e22"), (c("s") == "bb").alias("e23"), c("s").is_in(["a", "ccc", "zz"]).alias("e24"),
        c("s").alias("e25"), c("s").left(3).alias("e26"), (c("s") < "c").alias("e27"), c("b").alias("e28"),
        c("i8").eq_null_safe(c("i8")).alias("e29"), abs(c("i16")).alias("e30"), (c("u64") + 1).alias("e31"),
        lit(7).alias("e32"), lit(None).alias("e33"), (c("f32") + c("f32")).alias("e34"), c("d").alias("e35"),
        (c("b") ^ c("b")).alias("e36"),
    ]


@pytest.mark.parametrize("n,nulls", [(0, 0.1), (1, 0.0), (100, 0.3), (4097, 0.1), (4097, 0.0)])
def test_project_matches_oracle(gpu, n, nulls):
    rng = np.random.default_rng(n + 7)
    tbl = random_table(rng, n, nulls)
    exprs = exprs_under_test(gpu.col, gpu.lit)
    got = RB(gpu, tbl).eval_expression_list(exprs).to_arrow_table()
    want = E.eval_expression_list(tbl, exprs)
    assert_tables_equal(got, want)  # bit-exact, floats included (same op order, no FMA)


@pytest.mark.parametrize("n", [0, 5, 1000, 10007])
def test_filter_matches_oracle(gpu, n):
    rng = np.random.default_rng(n)
    tbl = random_table(rng, n)
    c = gpu.col
    for pred in ([c("i32") > 0], [c("b")], [c("f64") < 0.0, c("s") != "a"], [c("i64").is_null() | (c("u8") == 3)],
                 [c("i32") > 5000], [c("d") <= datetime.date(1998, 9, 2)]):
        got = RB(gpu, tbl).filter(pred).to_arrow_table()
        want = E.filter(tbl, pred)
        assert_tables_equal(got, want)


# ---------------------------------------------------------------------------------------------- aggregate
@pytest.mark.parametrize("n,nulls", [(0, 0.1), (1, 0.0), (1000, 0.2), (20000, 0.0), (20000, 0.3)])
def test_groupby_inline_aggs(gpu, n, nulls):
    rng = np.random.default_rng(n + 11)
    tbl = random_table(rng, n, nulls)
    c = gpu.col
    aggs = [c("i32").sum().alias("s_i32"), c("u16").sum().alias("s_u16"), c("f64").sum().alias("s_f64"), c("f32").sum().alias("s_f32"),
            c("i8").min().alias("mn_i8"), c("i64").max().alias("mx_i64"), c("u64").min().alias("mn_u64"), c("f64").min().alias("mn_f64"),
            c("f64").max().alias("mx_f64"), c("d").max().alias("mx_d"), c("i32").count().alias("c_valid"), c("i32").count("all").alias("c_all"),
            c("i32").count("null").alias("c_null"), c("f64").mean().alias("avg_f64"), c("i16").mean().alias("avg_i16"),
            (c("f64") * (1 - c("f32").cast(pa.float64()))).sum().alias("s_expr")]
    for gb in ([c("k")], [c("s")], [c("b"), c("u8")], [c("i8"), c("s"), c("d")], [c("f64")], [(c("k") % 3).alias("k3")]):
        got = RB(gpu, tbl).agg(aggs, gb).to_arrow_table()
        want = oracle_agg(tbl, aggs, gb)
        # group order: first occurrence on both sides (inline_agg.rs:804-807)
        assert_tables_equal(got, want, float_rtol=1e-9)


def test_ungrouped_and_fused_predicate(gpu):
    rng = np.random.default_rng(3)
    tbl = random_table(rng, 5000, 0.1)
    c = gpu.col
    aggs = [c("i64").sum().alias("s"), c("f64").mean().alias("m"), c("u8").max().alias("mx"), c("s").count().alias("cnt")]
    got = RB(gpu, tbl).agg(aggs, []).to_arrow_table()
    assert_tables_equal(got, oracle_agg(tbl, aggs, []), float_rtol=1e-9)
    pred = [c("i32") > 0, c("b")]
    got = RB(gpu, tbl).agg(aggs, [c("k")], predicate=pred).to_arrow_table()
    assert_tables_equal(got, oracle_agg(tbl, aggs, [c("k")], predicate=pred), float_rtol=1e-9)
    # everything filtered out: ungrouped still yields one row (sum null, count 0)
    got = RB(gpu, tbl).agg(aggs, [], predicate=[c("i32") > 10**6]).to_arrow_table()
    assert_tables_equal(got, oracle_agg(tbl, aggs, [], predicate=[c("i32") > 10**6]))


def test_groupby_high_cardinality(gpu):
    rng = np.random.default_rng(9)
    n = 300_000
    tbl = pa.table({"k": pa.array(rng.integers(0, 200_000, n).astype(np.int64)), "v": pa.array(rng.integers(0, 1000, n).astype(np.int64)),
                    "f": pa.array(rng.random(n))})
    c = gpu.col
    aggs = [c("v").sum().alias("sv"), c("v").count().alias("cv"), c("f").sum().alias("sf"), c("v").min().alias("mn")]
    got = RB(gpu, tbl).agg(aggs, [c("k")]).to_arrow_table()
    assert_tables_equal(got, oracle_agg(tbl, aggs, [c("k")]), float_rtol=1e-9)


def test_groupby_null_and_nan_keys(gpu):
    # inline_agg.rs:1094-1617 behaviours: null key is a group; all-null group sums to null; NaN keys group together
    tbl = pa.table({"k": pa.array([1, None, 1, None, 2], type=pa.int64()), "v": pa.array([None, None, 3, None, None], type=pa.int64()),
                    "f": pa.array([float("nan"), 1.0, float("nan"), 1.0, 2.0])})
    c = gpu.col
    got = RB(gpu, tbl).agg([c("v").sum().alias("s"), c("v").count().alias("c"), c("v").min().alias("m")], [c("k")]).to_pydict()
    assert got == {"k": [1, None, 2], "s": [3, None, None], "c": [1, 0, 0], "m": [3, None, None]}
    got = RB(gpu, tbl).agg([c("k").count("all").alias("n")], [c("f")]).to_arrow_table()
    assert got.column("n").to_pylist() == [2, 2, 1]


@pytest.mark.parametrize("n", [0, 1, 700, 90_000])
def test_any_value_and_bool_aggs(gpu, n):
    """any_value = the child at the group's first row (series/ops/agg.rs:353-382); bool_and / bool_or ignore nulls."""
    rng = np.random.default_rng(n + 3)
    tbl = random_table(rng, n, 0.2)
    c = gpu.col
    aggs = [c("s").any_value().alias("any_s"), c("f64").any_value().alias("any_f"), (c("i32") * 2).any_value().alias("any_e"),
            c("k").any_value(ignore_nulls=True).alias("any_k"), c("b").bool_and().alias("ba"), c("b").bool_or().alias("bo"),
            c("i64").sum().alias("sum_i64")]
    for keys, pred in (([], None), (["u8"], None), (["b", "i8"], [c("i32") > -500]), (["s"], None)):
        got = RB(gpu, tbl).agg(aggs, [c(k) for k in keys], predicate=pred).to_arrow_table()
        want = oracle_agg(tbl, aggs, [c(k) for k in keys], pred)
        assert_tables_equal(got, want)
    # ignore_nulls over nullable inputs: the child at the group's first row where it is valid, null for an all-null group
    ign = [c("i32").any_value(ignore_nulls=True).alias("x"), c("s").any_value(ignore_nulls=True).alias("xs"),
           c("f64").any_value(ignore_nulls=True).alias("xf"), c("b").any_value(ignore_nulls=True).alias("xb"), c("i32").any_value().alias("first")]
    for keys in ([], ["u8"], ["s"]):
        got = RB(gpu, tbl).agg(ign, [c(k) for k in keys]).to_arrow_table()
        assert_tables_equal(got, oracle_agg(tbl, ign, [c(k) for k in keys]))


@pytest.mark.parametrize("n", [1, 2, 3000, 120_000])
def test_stddev_var_staged_like_the_reference(gpu, n):
    """stddev / var run as the reference's staged rewrite (src/daft-local-plan/src/agg.rs:190-258): sum, sum of
    squares and count over Float64, final projection sqrt(((sq/n) - (s/n)^2) * n / (n - ddof)), null when n <= ddof.
    Compared with the oracle running the same staging and with numpy's definition."""
    from oracle.recordbatch import OracleRecordBatch
    rng = np.random.default_rng(n)
    tbl = pa.table({"k": pa.array(rng.integers(0, 7, n).astype(np.int64)),
                    "x": pa.array(rng.standard_normal(n) * 10 + 3, mask=rng.random(n) < 0.1),
                    "i": pa.array(rng.integers(-50, 50, n).astype(np.int32))})
    c = gpu.col
    aggs = [c("x").stddev().alias("sd0"), c("x").stddev(1).alias("sd1"), c("x").var().alias("v1"), c("i").var(0).alias("vi0"),
            c("x").sum().alias("s"), c("i").count().alias("cnt")]
    for keys in ([], ["k"]):
        got = RB(gpu, tbl).agg(aggs, [c(k) for k in keys]).to_arrow_table()
        want = OracleRecordBatch(tbl).agg(aggs, [c(k) for k in keys]).to_arrow_table()
        assert_tables_equal(got, want, float_rtol=1e-7)
    if n >= 3000:  # numpy's two-pass definition agrees to far better than the 1e-6 of north_star on well-conditioned data
        x = tbl.column("x").to_numpy(zero_copy_only=False)
        got = RB(gpu, tbl).agg(aggs[:3], []).to_pydict()
        assert got["sd0"][0] == pytest.approx(np.nanstd(x, ddof=0), rel=1e-9)
        assert got["sd1"][0] == pytest.approx(np.nanstd(x, ddof=1), rel=1e-9)
        assert got["v1"][0] == pytest.approx(np.nanvar(x, ddof=1), rel=1e-9)


@pytest.mark.parametrize("n", [0, 1, 5000, 80_000])
def test_count_distinct(gpu, n):
    """count_distinct = distinct non-null values per group, UInt64 (src/daft-functions-list/src/series.rs:37-40,234-247);
    the device path is dedup on (keys, value) + count, the oracle counts sets by definition."""
    rng = np.random.default_rng(n + 11)
    tbl = random_table(rng, n, 0.15)
    c = gpu.col
    aggs = [c("u16").count_distinct().alias("d_u16"), c("s").count_distinct().alias("d_s"), c("i64").sum().alias("sum_i64"),
            (c("i32") % 10).count_distinct().alias("d_expr"), c("u8").count("all").alias("rows")]
    for keys, pred in (([], None), (["k"], None), (["b", "i8"], [c("u32") > 3000])):
        got = RB(gpu, tbl).agg(aggs, [c(k) for k in keys], predicate=pred).to_arrow_table()
        want = oracle_agg(tbl, aggs, [c(k) for k in keys], pred)
        assert_tables_equal(got, want)


@pytest.mark.parametrize("n,nb", [(0, 3), (50, 0), (3000, 7), (20_000, 300)])
def test_search_sorted_and_partition_by_range(gpu, n, nb):
    """RecordBatch::search_sorted / partition_by_range (ops/search_sorted.rs:13-36, ops/partition.rs:88-101): upper-bound
    positions under the sort comparator (NaN greatest, nulls first where descending), stable partitions."""
    rng = np.random.default_rng(n + nb)
    tbl = random_table(rng, n, 0.1)
    c = gpu.col
    for keys, desc in ((["i32"], [False]), (["f64", "u8"], [True, False]), (["s", "i8"], [False, True])):
        src = random_table(rng, max(nb * 3, 1), 0.1).select(keys)
        bounds = O.sort(src, keys, desc, desc).slice(0, nb)   # boundaries are sorted with nulls_first = descending
        got_pos = RB(gpu, bounds).search_sorted(RB(gpu, tbl.select(keys)), desc).to_pydict()["idx"]
        assert got_pos == O.search_sorted(bounds, tbl.select(keys), desc).tolist()
        got = RB(gpu, tbl).partition_by_range([c(k) for k in keys], RB(gpu, bounds), desc)
        want = O.partition_by_range(tbl, keys, bounds, desc)
        assert len(got) == len(want)
        for g, w in zip(got, want):
            assert_tables_equal(g.to_arrow_table(), w)


def test_partition_by_range_reference_literals(gpu):
    """tests/recordbatch/test_partitioning.py:316-344: multi-column string boundaries and the input validation errors."""
    c = gpu.col
    table = RB(gpu, pa.table({"x": ["a", "c", "a", "c"], "y": ["1", "2", "3", "4"]}))
    bounds = RB(gpu, pa.table({"x": ["b"], "y": ["1"]}))
    parts = table.partition_by_range([c("x"), c("y")], bounds, [False, False])
    assert [p.to_pydict() for p in parts] == [{"x": ["a", "a"], "y": ["1", "3"]}, {"x": ["c", "c"], "y": ["2", "4"]}]
    parts = table.partition_by_range([c("x"), c("y")], bounds, [True, False])
    assert [p.to_pydict() for p in parts] == [{"x": ["c", "c"], "y": ["2", "4"]}, {"x": ["a", "a"], "y": ["1", "3"]}]
    data = pa.table({"x": [1, 2, 3], "b": [0, 1, 2]})
    t, b = RB(gpu, data), RB(gpu, data)
    with pytest.raises(ValueError, match="Schema Mismatch"):
        t.partition_by_range([c("x")], b, [False])
    with pytest.raises(ValueError, match="Mismatch in number of arguments for `descending`"):
        t.partition_by_range([c("x")], b.eval_expression_list([c("x")]), [False, False])
    with pytest.raises(ValueError, match="Schema Mismatch"):
        t.partition_by_range([c("x")], b.eval_expression_list([c("x").alias("y")]), [False])


def test_dedup(gpu):
    rng = np.random.default_rng(2)
    tbl = random_table(rng, 3000, 0.2)
    c = gpu.col
    got = RB(gpu, tbl).dedup([c("u8"), c("b")]).to_arrow_table()
    _, first, _ = O.group_ids([tbl.column("u8"), tbl.column("b")])
    assert_tables_equal(got, tbl.take(pa.array(first.astype(np.int64))))


# ---------------------------------------------------------------------------------------------- sort
@pytest.mark.parametrize("n", [0, 1, 2, 1000, 50000])
def test_argsort_single_keys(gpu, n):
    rng = np.random.default_rng(n + 1)
    tbl = random_table(rng, n, 0.15)
    if n > 10:
        f = tbl.column("f64").to_pylist()
        f[3], f[5], f[7] = float("nan"), -0.0, 0.0
        tbl = tbl.set_column(tbl.column_names.index("f64"), "f64", pa.array(f))
    rb = RB(gpu, tbl)
    for name in ["i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f32", "f64", "b", "d", "s"]:
        for desc in (False, True):
            for nf in (False, True):
                got = rb.argsort([gpu.col(name)], descending=desc, nulls_first=nf).to_pydict()["indices"]
                want = O.argsort([tbl.column(name)], [desc], [nf])
                assert got == want.tolist(), (name, desc, nf)


def test_sort_multikey(gpu):
    rng = np.random.default_rng(17)
    tbl = random_table(rng, 20000, 0.1)
    c = gpu.col
    for keys, desc, nf in [(["u8", "s"], [False, True], [False, True]), (["b", "i8", "f32"], [True, False, True], [True, True, False]),
                           (["k", "d", "i64"], [False, False, False], None)]:
        got = RB(gpu, tbl).sort([c(k) for k in keys], descending=desc, nulls_first=nf).to_arrow_table()
        want = O.sort(tbl, keys, desc, nf if nf is not None else desc)
        assert_tables_equal(got, want)
    got = RB(gpu, tbl).top_n([c("i64")], 10, descending=True, nulls_first=False, offset=5).to_arrow_table()
    assert_tables_equal(got, O.sort(tbl, ["i64"], [True], [False]).slice(5, 10))


@pytest.mark.parametrize("n", [70_000, 300_001, 1_500_000])
def test_sort_onesweep_equals_two_kernel_radix(gpu, n, monkeypatch):
    """The opt-in onesweep passes (B200_SORT_ONESWEEP=1: one histogram for all digits, decoupled look-back in the scatter,
    sort.cu radix_onesweep) against the default per-pass histogram + scan scheme.  Both are stable LSD sorts, so their argsorts
    are IDENTICAL, ties included -- and equal to the oracle's stable sort."""
    rng = np.random.default_rng(n)
    tbl = pa.table({"a": pa.array(rng.integers(-50, 50, n).astype(np.int64), mask=rng.random(n) < 0.05),        # heavy ties, nulls
                    "f": pa.array(np.where(rng.random(n) < 0.01, np.nan, rng.standard_normal(n) * 1e6)),            # all 8 bytes live
                    "u": pa.array(rng.integers(0, 2**32, n).astype(np.uint32)),
                    "s": pa.array([f"k{v:05d}" for v in rng.integers(0, 3000, n)], type=pa.large_utf8()),
                    "row": pa.array(np.arange(n))})
    c = gpu.col
    for keys, desc in ((["a"], [False]), (["f"], [True]), (["u", "a"], [False, True]), (["s"], [False])):
        if keys == ["s"] and n > 400_000:
            continue
        res = {}
        for mode in ("1", "0"):
            monkeypatch.setenv("B200_SORT_ONESWEEP", mode)
            res[mode] = RB(gpu, tbl).argsort([c(k) for k in keys], descending=desc).to_pydict()
        assert res["1"] == res["0"], keys
        want = O.argsort([tbl.column(k) for k in keys], desc, desc)
        assert list(res["1"].values())[0] == [int(x) for x in want], keys


@pytest.mark.parametrize("n", [3000, 60_000])
def test_sort_long_strings_with_shared_prefixes(gpu, n):
    """Strings are sorted by 7-byte chunks: round 0 over all rows, later rounds only over the runs that still tie (sort.cu
    make_string_chunk / string_runs).  Long common prefixes, values that are prefixes of one another, chunk-boundary lengths (6, 7,
    8, 13, 14, 15 bytes), embedded 0x00 / 0xFF-like code points, duplicates, nulls, one very long value; ascending and descending,
    alone and as the second key -- the order must equal the oracle's bytewise comparator sort (stable on ties)."""
    rng = np.random.default_rng(n)
    stems = ["", "a", "http://example.com/a/long/common/prefix/", "http://example.com/a/long/common/prefix/that/goes/on/and/on/",
             "abcdef", "abcdefg", "abcdefgh", "abcdefghijklm", "abcdefghijklmn", "abcdefghijklmno", "\x00\x00", "\x00", "zz\u00ff\u0100"]
    vals = []
    for i in range(n):
        stem = stems[rng.integers(0, len(stems))]
        tail = "".join(chr(97 + int(c)) for c in rng.integers(0, 3, rng.integers(0, 12)))
        vals.append(stem + tail)
    vals[n // 2] = "http://example.com/a/long/common/prefix/" + "q" * 700
    tbl = pa.table({"s": pa.array(vals, type=pa.large_utf8(), mask=rng.random(n) < 0.05), "g": pa.array(rng.integers(0, 4, n).astype(np.int32)),
                    "row": pa.array(np.arange(n))})
    c = gpu.col
    for keys, desc, nf in ((["s"], [False], None), (["s"], [True], [False]), (["g", "s"], [True, False], None), (["s", "g"], [False, True], [True, False])):
        got = RB(gpu, tbl).argsort([c(k) for k in keys], descending=desc, nulls_first=nf).to_pydict()
        want = O.argsort([tbl.column(k) for k in keys], desc, nf if nf is not None else desc)
        assert list(got.values())[0] == [int(x) for x in want], (keys, desc)


@pytest.mark.parametrize("dtype", ["f64", "i64", "i32", "f32"])
def test_top_n_selection_path(gpu, dtype):
    """Large inputs take the radix-select path (sort.cu op_top_indices): same rows as sort + slice
    (src/daft-recordbatch/src/ops/sort.rs:47-61).  Keys are unique so the expected order is total; a second case
    has massive ties on the first key and relies on the second key."""
    rng = np.random.default_rng(11)
    n = 300_000
    if dtype.startswith("f"):
        vals = rng.permutation(n).astype(np.float64) * 0.37 - 40_000.0
        vals[:7] = np.nan
        arr = pa.array(vals.astype(np.float32 if dtype == "f32" else np.float64))
    else:
        vals = rng.permutation(n).astype(np.int64) * 3 - 400_000
        arr = pa.array(vals.astype(np.int32 if dtype == "i32" else np.int64))
    mask = rng.random(n) < 0.001
    arr = pa.array(arr.to_numpy(zero_copy_only=False), mask=mask, type=arr.type)
    tie = pa.array((np.arange(n) % 3).astype(np.int64))
    uniq = pa.array(rng.permutation(n).astype(np.int64))
    tbl = pa.table({"k": arr, "tie": tie, "u": uniq})
    c = gpu.col
    rb = RB(gpu, tbl)
    for desc, nf, limit, offset in ((False, False, 100, 0), (True, False, 57, 13), (False, True, 500, 0), (True, True, 1000, 250)):
        got = rb.top_n([c("k"), c("u")], limit, descending=[desc, False], nulls_first=[nf, False], offset=offset).to_arrow_table()
        want = O.sort(tbl, ["k", "u"], [desc, False], [nf, False]).slice(offset, limit)
        assert_tables_equal(got, want)
    # first key has 3 distinct values: the threshold bucket holds a third of the input
    got = rb.top_n([c("tie"), c("u")], 64, descending=[True, True]).to_arrow_table()
    want = O.sort(tbl, ["tie", "u"], [True, True], [True, True]).slice(0, 64)
    assert_tables_equal(got, want)


def test_sort_by_expression(gpu):
    tbl = pa.table({"a": [3, 1, 2, None], "b": [1.0, 5.0, -1.0, 0.0]})
    c = gpu.col
    got = RB(gpu, tbl).sort([(c("a") * -1).alias("neg")]).to_pydict()
    assert got == {"a": [3, 2, 1, None], "b": [1.0, -1.0, 5.0, 0.0]}


# ---------------------------------------------------------------------------------------------- partition / take / concat
@pytest.mark.parametrize("P", [1, 2, 7, 64, 300])
def test_partition_by_hash(gpu, P):
    rng = np.random.default_rng(P)
    tbl = random_table(rng, 5000, 0.1)
    c = gpu.col
    for keys in (["i64"], ["s", "u8"]):
        got = RB(gpu, tbl).partition_by_hash([c(k) for k in keys], P)
        want = O.partition_by_hash(tbl, keys, P)
        assert len(got) == P
        for g, w in zip(got, want):
            assert_tables_equal(g.to_arrow_table(), w)  # placement and in-partition order are exact


@pytest.mark.parametrize("P", [1, 2, 7, 8, 200])
@pytest.mark.parametrize("key_type", [pa.int64(), pa.int32(), pa.uint32(), pa.uint64(), pa.date32()])
def test_partition_by_hash_fused_key(gpu, P, key_type, monkeypatch):
    """One never-null 4- / 8-byte integer key: the partition kernels hash the key in registers (no hash column) and take
    `% P` by multiplication; placement must equal hash_rows % P exactly, fused or not."""
    rng = np.random.default_rng(P + 5)
    n = 30000
    if key_type == pa.date32():
        key = pa.array(rng.integers(-20000, 60000, n).astype(np.int32)).cast(pa.date32())
    elif key_type in (pa.uint32(), pa.uint64()):
        key = pa.array(rng.integers(0, 2**32 - 1 if key_type == pa.uint32() else 2**63, n).astype(np.uint64)).cast(key_type)
    else:
        key = pa.array(rng.integers(-2**31, 2**31 - 1, n)).cast(key_type)
    tbl = pa.table({"key": key, "v": pa.array(np.arange(n, dtype=np.float64)), "s": pa.array([f"s{i % 17}" for i in range(n)], type=pa.large_utf8()),
                    "w": pa.array(rng.integers(0, 9, n).astype(np.int16), mask=rng.random(n) < 0.1)})
    want = O.partition_by_hash(tbl, ["key"], P)
    for fused in ("1", "0"):
        monkeypatch.setenv("B200_PARTITION_FUSED", fused)
        got = RB(gpu, tbl).partition_by_hash([gpu.col("key")], P)
        assert len(got) == P
        for g, w in zip(got, want):
            assert_tables_equal(g.to_arrow_table(), w)


def test_take_and_concat_and_slice(gpu):
    rng = np.random.default_rng(4)
    tbl = random_table(rng, 1000, 0.2)
    idx = rng.integers(0, 1000, 500)
    got = RB(gpu, tbl).take(pa.array(idx, type=pa.int64())).to_arrow_table()
    assert_tables_equal(got, tbl.take(pa.array(idx)))
    got = RB(gpu, tbl).take(pa.array([0, None, 5], type=pa.int32())).to_arrow_table()
    assert_tables_equal(got, tbl.take(pa.array([0, None, 5])))
    a, b = tbl.slice(0, 300), tbl.slice(300, 700)
    got = gpu.RecordBatch.concat([RB(gpu, a), RB(gpu, b), RB(gpu, tbl.slice(0, 0))]).to_arrow_table()
    assert_tables_equal(got, tbl)
    got = RB(gpu, tbl).slice(17, 333).to_arrow_table()
    assert_tables_equal(got, tbl.slice(17, 333 - 17))


# ---------------------------------------------------------------------------------------------- join
def join_tables(rng, nl, nr, key_range, nulls=0.1):
    def keycol(n):
        return pa.array(rng.integers(0, key_range, n).astype(np.int64), mask=(rng.random(n) < nulls) if nulls else None)
    words = ["a", "bb", "ccc", ""]
    left = pa.table({"k": keycol(nl), "ks": pa.array([words[i] for i in rng.integers(0, 4, nl)], type=pa.large_utf8()),
                     "lv": pa.array(rng.random(nl)), "li": pa.array(np.arange(nl))})
    right = pa.table({"k": keycol(nr), "ks": pa.array([words[i] for i in rng.integers(0, 4, nr)], type=pa.large_utf8()),
                      "rv": pa.array(rng.integers(0, 100, nr).astype(np.int32)), "ri": pa.array(np.arange(nr))})
    return left, right


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer", "semi", "anti"])
@pytest.mark.parametrize("nl,nr,kr", [(0, 10, 5), (10, 0, 5), (1000, 300, 100), (3000, 5000, 20000)])
def test_hash_join_matches_oracle(gpu, how, nl, nr, kr):
    rng = np.random.default_rng(nl * 7 + nr)
    left, right = join_tables(rng, nl, nr, kr)
    c = gpu.col
    for on in (["k"], ["k", "ks"]):
        got = RB(gpu, left).hash_join(RB(gpu, right), [c(k) for k in on], [c(k) for k in on], how=how).to_arrow_table()
        rsel = right if len(on) == 2 else right.rename_columns(["k", "right.ks", "rv", "ri"])
        want = O.hash_join(left, rsel, on, on, how)
        if how in ("semi", "anti"):
            want = want.select(left.column_names)
        # row order: probe-row order with matches in build insertion order, unmatched build rows last
        assert_tables_equal(got, want)


@pytest.mark.parametrize("how", ["inner", "left", "outer", "semi"])
def test_hash_join_long_duplicate_runs(gpu, how):
    """Build keys repeated hundreds of times: bucket runs far longer than the head row's forward walk
    (join.cu bucket_bounds BB_WALK), matches still in build insertion order (probe_table.rs:178-181)."""
    rng = np.random.default_rng(3)
    nb, npr = 6000, 400
    right = pa.table({"k": pa.array(rng.integers(0, 12, nb), type=pa.int64()), "rv": pa.array(np.arange(nb), type=pa.int64())})
    left = pa.table({"k": pa.array(rng.integers(0, 16, npr), type=pa.int64()), "lv": pa.array(np.arange(npr) * 10, type=pa.int64())})
    c = gpu.col
    got = RB(gpu, left).hash_join(RB(gpu, right), [c("k")], [c("k")], how=how).to_arrow_table()
    want = O.hash_join(left, right, ["k"], ["k"], how)
    if how == "semi":
        want = want.select(left.column_names)
    assert_tables_equal(got, want)


@pytest.mark.parametrize("typ", [pa.int64(), pa.float64(), pa.large_utf8()])
@pytest.mark.parametrize("null_safe", [False, True])
def test_join_pairs_from_reference_tests(gpu, typ, null_safe):
    """The literal (left row, right row) pair tables of tests/recordbatch/test_joins.py:59-140, through the C ABI."""
    from test_oracle_cpu import JOIN_PAIR_CASES
    c = gpu.col
    for left, right, pairs, ns_pairs in JOIN_PAIR_CASES:
        expected = ns_pairs if null_safe else pairs

        def column(vals):
            if pa.types.is_large_string(typ):
                return pa.array([None if v is None else str(v) for v in vals], type=typ)
            return pa.array(vals, type=typ)
        lt = pa.table({"x": column(left), "x_ind": pa.array(list(range(len(left))), type=pa.int64())})
        rt = pa.table({"y": column(right), "y_ind": pa.array(list(range(len(right))), type=pa.int64())})
        kw = {"null_equals_null": [True]} if null_safe else {}
        res = RB(gpu, lt).hash_join(RB(gpu, rt), [c("x")], [c("y")], how="inner", **kw).to_arrow_table()
        assert res.column_names == ["x", "x_ind", "y", "y_ind"]
        got = list(zip(res.column("x_ind").to_pylist(), res.column("y_ind").to_pylist()))
        assert sorted(got) == sorted(expected)
        assert res.column("y").to_pylist() == [rt.column("y").to_pylist()[j] for _, j in got]
        res = RB(gpu, rt).hash_join(RB(gpu, lt), [c("y")], [c("x")], how="inner", **kw).to_arrow_table()
        assert res.column_names == ["y", "y_ind", "x", "x_ind"]
        assert sorted(zip(res.column("x_ind").to_pylist(), res.column("y_ind").to_pylist())) == sorted(expected)


def test_join_null_safe_and_float_keys(gpu):
    left = pa.table({"a": pa.array([1, None, 3], type=pa.int64()), "f": pa.array([0.0, float("nan"), 1.5]), "x": [10, 20, 30]})
    right = pa.table({"a": pa.array([None, 1, 1], type=pa.int64()), "f": pa.array([-0.0, float("nan"), 1.5]), "y": [1, 2, 3]})
    c = gpu.col
    rl, rr = RB(gpu, left), RB(gpu, right)
    got = rl.hash_join(rr, [c("a")], [c("a")]).to_pydict()
    assert got["a"] == [1, 1] and got["y"] == [2, 3]
    got = rl.hash_join(rr, [c("a")], [c("a")], null_equals_null=[True]).to_pydict()
    assert got["a"] == [1, 1, None] and got["y"] == [2, 3, 1]
    got = rl.hash_join(rr, [c("f")], [c("f")]).to_pydict()   # NaN never matches; +0.0 vs -0.0 hash differently
    assert got["x"] == [30] and got["y"] == [3]

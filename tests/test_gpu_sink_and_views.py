"""GroupedAggregateSink over b200_groupby_* (src/daft-local-execution/src/sinks/grouped_aggregate.rs:146-488) and the
late-materialised (view) outputs of the hash join."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal, oracle_agg, random_table

pytestmark = pytest.mark.gpu


def RB(gpu, tbl):
    return gpu.RecordBatch.from_arrow_table(tbl)


@pytest.mark.parametrize("morsel", [700, 4096, 100_000])
def test_grouped_aggregate_sink_matches_single_batch(gpu, morsel):
    from daft_b200.aggstages import GroupedAggregateSink
    rng = np.random.default_rng(morsel)
    tbl = random_table(rng, 20_000, 0.15)
    c = gpu.col
    aggs = [c("i64").sum().alias("sm"), c("f64").mean().alias("m"), c("i32").count("all").alias("ca"), c("i32").count("null").alias("cn"),
            c("u16").min().alias("lo"), c("i64").max().alias("hi"), c("b").bool_and().alias("ball"), (c("f64") * 2).stddev().alias("sd"),
            c("i16").var(1).alias("vr"), c("i8").mean().alias("m8")]
    for gb, pred in (([c("k")], None), ([c("s"), c("u8")], [c("i32") > 0]), ([], None)):
        sink = GroupedAggregateSink(aggs, gb, pred)
        for o in range(0, tbl.num_rows, morsel):
            sink.sink(RB(gpu, tbl.slice(o, morsel)))
        got = sink.finalize().to_arrow_table()
        want = oracle_agg_staged(tbl, aggs, gb, pred)
        assert_tables_equal(got, want, float_rtol=1e-9, sort_by=[g.name() for g in gb] or None)


def oracle_agg_staged(tbl, aggs, gb, pred):
    from oracle.recordbatch import OracleRecordBatch
    return OracleRecordBatch(tbl).agg(aggs, gb, pred).to_arrow_table()


def test_q1_through_the_sink(gpu):
    from daft_b200 import tpch
    from daft_b200.aggstages import GroupedAggregateSink
    pred, gb, aggs = tpch.q1_exprs()
    host = tpch.generate_host("lineitem", 0.05, tpch.Q1_COLUMNS)
    sink = GroupedAggregateSink(aggs, gb, pred)
    for o in range(0, host.num_rows, 131072):
        sink.sink(RB(gpu, host.slice(o, 131072)))
    got = sink.finalize().sort([gpu.col("L_RETURNFLAG"), gpu.col("L_LINESTATUS")]).to_arrow_table()
    want = oracle_agg(host, aggs, gb, pred).sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")])
    assert_tables_equal(got, want, float_rtol=1e-9)


def test_join_outputs_are_views_until_read(gpu):
    """b200_batch_describe reports schema only for an unread join output; materialize / any operator gathers it."""
    import ctypes
    from daft_b200 import _lib
    rng = np.random.default_rng(5)
    left = pa.table({"k": pa.array(rng.integers(0, 50, 2000).astype(np.int64)), "a": pa.array(np.arange(2000)), "s": pa.array([f"row{i}" for i in range(2000)])})
    right = pa.table({"k": pa.array(np.arange(40, dtype=np.int64)), "b": pa.array(np.arange(40) * 10.0)})
    c = gpu.col
    j = RB(gpu, left).hash_join(RB(gpu, right), [c("k")], [c("k")])
    d = j._describe(1)
    assert d.length == len(j) and d.null_count == -1 and not d.data          # still a selection vector over `left`
    plain = ctypes.c_void_p()
    _lib.call("b200_batch_materialize", j._h, ctypes.byref(plain))
    pj = gpu.RecordBatch._wrap(plain, like=j)
    d2 = pj._describe(1)
    assert d2.null_count == 0 and d2.data
    from oracle import oracle as O
    want = O.hash_join(left, right, ["k"], ["k"], "inner")
    assert_tables_equal(pj.to_arrow_table(), want)
    assert_tables_equal(j.to_arrow_table(), want)
    # a chain of joins reads its keys through the selection vectors; slices and filters of views work
    right2 = pa.table({"a": pa.array(np.arange(0, 2000, 3)), "z": pa.array(np.arange(0, 2000, 3) % 7)})
    j2 = j.hash_join(RB(gpu, right2), [c("a")], [c("a")])
    want2 = O.hash_join(want, right2, ["a"], ["a"], "inner")
    assert_tables_equal(j2.to_arrow_table(), want2)
    assert_tables_equal(j2.slice(5, 60).to_arrow_table(), want2.slice(5, 55))
    assert_tables_equal(j2.filter([c("z") == 3]).to_arrow_table(), want2.filter(pa.compute.equal(want2.column("z"), 3)))
    assert_tables_equal(j2.sort([c("b"), c("a")], descending=[True, False]).to_arrow_table(),
                        want2.sort_by([("b", "descending"), ("a", "ascending")]))


def test_probe_indices_multi_table_build_and_probe_through_view(gpu):
    """ProbeTableBuilder::add_table twice (probe_table.rs:117-185): build rows are numbered across the added tables in
    insertion order; probe_indices yields (probe row, build row) in probe order, matches in insertion order."""
    from daft_b200.join import ProbeTable
    from oracle import oracle as O
    rng = np.random.default_rng(8)
    b1 = pa.table({"k": pa.array(rng.integers(0, 30, 500).astype(np.int64)), "p": pa.array(np.arange(500))})
    b2 = pa.table({"k": pa.array(rng.integers(0, 30, 300).astype(np.int64)), "p": pa.array(np.arange(300) + 1000)})
    probe = pa.table({"k": pa.array(rng.integers(0, 40, 1000).astype(np.int64))})
    pt = ProbeTable(1)
    pt.add_table(RB(gpu, b1), [0])
    pt.add_table(RB(gpu, b2), [0])
    pt.build()
    assert pt.num_rows() == 800
    got = pt.probe_indices(RB(gpu, probe), [0]).to_pydict()
    build = pa.concat_tables([b1, b2])
    offs, rows = O.join_matches([build.column("k")], [probe.column("k")], None)
    cnt = np.diff(offs)
    assert got["probe_idx"] == np.repeat(np.arange(1000), cnt).tolist()
    assert got["build_idx"] == rows.tolist()
    ex = pt.probe_exists(RB(gpu, probe), [0]).to_pydict()["exists"]
    assert ex == (cnt > 0).tolist()


def test_string_minmax_reference_literals_and_grouped(gpu):
    """tests/recordbatch/test_table_aggs.py:137-158,188-213 through the C ABI, then grouped string min / max with a
    numeric aggregate beside them against the oracle (compare_agg.rs:99-134)."""
    c = gpu.col
    cases = [([None], None, None), ([None, None, None], None, None), (["abc"], "abc", "abc"), ([None, "abc"], "abc", "abc"),
             ([None, "aaa", None, "b", "c", "aaa", None], "aaa", "c")]
    for vals, lo, hi in cases:
        tbl = pa.table({"input": pa.array(vals, type=pa.large_utf8())})
        got = RB(gpu, tbl).agg([c("input").min().alias("min"), c("input").max().alias("max")], []).to_pydict()
        assert got == {"min": [lo], "max": [hi]}, (vals, got)
    rng = np.random.default_rng(2)
    tbl = random_table(rng, 5000, 0.2)
    aggs = [c("s").min().alias("smin"), c("i64").sum().alias("tot"), c("s").max().alias("smax"), c("s").count().alias("n")]
    for gb in ([c("k")], [c("u8"), c("b")]):
        got = RB(gpu, tbl).agg(aggs, gb).to_arrow_table()
        want = oracle_agg(tbl, aggs, gb)
        assert_tables_equal(got, want, sort_by=[g.name() for g in gb])


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer", "semi", "anti"])
@pytest.mark.parametrize("key_type", [pa.int64(), pa.int32(), pa.uint8(), pa.date32()])
def test_direct_address_table_equals_bucket_table(gpu, how, key_type, monkeypatch):
    """Unique, dense integer build keys take the direct-address table (join.cu: key - min indexes a presence bitmap and a
    row-id array); the same join with the mode switched off takes the hashed bucket table; both equal the oracle."""
    from oracle import oracle as O
    rng = np.random.default_rng(17)
    nb = 200 if key_type == pa.uint8() else 5000
    np_t = {pa.int64(): np.int64, pa.int32(): np.int32, pa.uint8(): np.uint8, pa.date32(): np.int32}[key_type]
    base = 0 if key_type == pa.uint8() else -700
    bkeys = (rng.permutation(nb * 3)[:nb] + base).astype(np_t)       # unique, range 3x the count
    pkeys = (rng.integers(-100, nb * 3 + 100, 20000) + base).astype(np_t) if key_type != pa.uint8() else rng.integers(0, 256, 20000).astype(np_t)
    mk = (lambda a: pa.array(a).cast(pa.date32())) if key_type == pa.date32() else (lambda a: pa.array(a, type=key_type))
    right = pa.table({"k": mk(bkeys), "b": pa.array(np.arange(nb) * 1.5)})
    left = pa.table({"k": pa.array(mk(pkeys).to_pylist(), type=right.column("k").type, mask=rng.random(20000) < 0.05), "a": pa.array(np.arange(20000))})
    c = gpu.col
    want = O.hash_join(left, right, ["k"], ["k"], how)
    if how in ("semi", "anti"):
        want = want.select(left.column_names)
    for direct in ("1", "0"):
        monkeypatch.setenv("B200_JOIN_DIRECT", direct)
        got = gpu.RecordBatch.from_arrow_table(left).hash_join(gpu.RecordBatch.from_arrow_table(right), [c("k")], [c("k")], how=how).to_arrow_table()
        assert_tables_equal(got, want)


@pytest.mark.parametrize("how", ["inner", "left", "right", "outer", "semi", "anti"])
@pytest.mark.parametrize("null_eq", [False, True])
def test_direct_address_table_on_a_key_prefix(gpu, how, null_eq, monkeypatch):
    """A multi-column key whose FIRST column is unique and dense on the build side (TPC-H Q5: (C_CUSTKEY, C_NATIONKEY)) takes the
    direct-address table on that column and compares the other columns at the candidate row (nullable string and float
    columns here); equal to the bucket table and to the oracle, through a view on the probe side as well."""
    from oracle import oracle as O
    rng = np.random.default_rng(23)
    nb, npr = 4000, 30000
    bk = (rng.permutation(nb * 2)[:nb] + 50).astype(np.int64)
    b2 = rng.integers(0, 4, nb)
    right = pa.table({"k": pa.array(bk), "s": pa.array([None if v == 3 else f"n{v}" for v in b2], type=pa.large_utf8()),
                      "f": pa.array(b2 * 0.5, mask=rng.random(nb) < 0.1), "payload": pa.array(np.arange(nb) * 2.0)})
    pk = rng.integers(0, nb * 2 + 100, npr).astype(np.int64)
    lut_s = dict(zip(bk.tolist(), right.column("s").to_pylist()))
    lut_f = dict(zip(bk.tolist(), right.column("f").to_pylist()))
    agree = rng.random(npr) < 0.7           # 70 % of the probe rows carry the build row's other columns, the rest something else
    ps = [lut_s.get(int(k), "zz") if a else "n0" for k, a in zip(pk, agree)]
    pf = [lut_f.get(int(k), 9.0) if a else 0.5 for k, a in zip(pk, agree)]
    left = pa.table({"k": pa.array(pk, mask=rng.random(npr) < 0.03), "s": pa.array(ps, type=pa.large_utf8()), "f": pa.array(pf, type=pa.float64()),
                     "a": pa.array(np.arange(npr))})
    c = gpu.col
    keys = ["k", "s", "f"]
    ne = [null_eq] * 3
    want = O.hash_join(left, right, keys, keys, how, null_equals_null=ne)
    if how in ("semi", "anti"):
        want = want.select(left.column_names)
    for direct in ("1", "0"):
        monkeypatch.setenv("B200_JOIN_DIRECT", direct)
        lb, rb = gpu.RecordBatch.from_arrow_table(left), gpu.RecordBatch.from_arrow_table(right)
        got = lb.hash_join(rb, [c(k) for k in keys], [c(k) for k in keys], how=how, null_equals_null=ne).to_arrow_table()
        assert_tables_equal(got, want)
    if how == "inner":   # probe side = a join output (views over two different bases)
        monkeypatch.setenv("B200_JOIN_DIRECT", "1")
        ids = pa.table({"a": pa.array(np.arange(0, npr, 2)), "tag": pa.array(np.arange(0, npr, 2) % 7)})
        lb = gpu.RecordBatch.from_arrow_table(ids).hash_join(gpu.RecordBatch.from_arrow_table(left), [c("a")], [c("a")])
        got = lb.hash_join(gpu.RecordBatch.from_arrow_table(right), [c(k) for k in keys], [c(k) for k in keys], null_equals_null=ne).to_arrow_table()
        lhs = O.hash_join(ids, left, ["a"], ["a"], "inner")
        assert_tables_equal(got, O.hash_join(lhs, right, keys, keys, "inner", null_equals_null=ne), sort_by=["a"])


@pytest.mark.parametrize("size,k", [(0, 3), (4, 3), (10, 1), (33, 10), (5000, 7), (100000, 300)])
def test_partition_by_random_conserves_rows_and_is_deterministic(gpu, size, k):
    """tests/recordbatch/test_partitioning.py:124-147,199-214: k partitions, every row exactly once, same split for the same seed."""
    tbl = pa.table({"x": pa.array(np.arange(size, dtype=np.int64)), "s": pa.array([f"r{i}" for i in range(size)], type=pa.large_utf8())})
    rb = RB(gpu, tbl)
    parts = rb.partition_by_random(k, 0)
    assert len(parts) == k and sum(len(p) for p in parts) == size
    seen = [v for p in parts for v in p.to_pydict()["x"]]
    assert sorted(seen) == list(range(size))
    for p in parts:
        d = p.to_pydict()
        assert d["x"] == sorted(d["x"]) and d["s"] == [f"r{i}" for i in d["x"]]      # input order kept, rows intact
    again = rb.partition_by_random(k, 0)
    assert all(a.to_pydict() == b.to_pydict() for a, b in zip(parts, again))
    if size >= 5000:
        other = rb.partition_by_random(k, 1)
        assert any(a.to_pydict() != b.to_pydict() for a, b in zip(parts, other))
        assert max(len(p) for p in parts) < 3 * size / k + 50
    with pytest.raises(ValueError):
        rb.partition_by_random(0, 0)


def test_quantiles_match_the_reference_rule(gpu):
    """RecordBatch::quantiles (src/daft-recordbatch/src/lib.rs:491-513): rows floor(i / num * len) for i in 1..num."""
    tbl = pa.table({"v": pa.array(np.arange(103, dtype=np.int64))})
    for num in (1, 2, 4, 8, 200):
        want = [min(int((i / num) * 103), 102) for i in range(1, num)]
        assert RB(gpu, tbl).quantiles(num).to_pydict()["v"] == want

"""Pins the CPU oracle: XXH3 against the Python `xxhash` package and the reference's literal goldens,
group-by / join / sort / filter / project against the literal expectations of the reference's own
kernel tests and against pyarrow.compute as an independent third opinion."""
import datetime
import random

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest
import xxhash

from helpers import assert_tables_equal, oracle_agg, random_table
from oracle import expr_eval as E
from oracle import oracle as O
from daft_b200 import col, lit


def test_xxh3_all_length_classes_and_seeds():
    random.seed(1)
    for n in list(range(0, 300)) + [511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4096, 5000]:
        for seed in (0, 1, 42, 2**64 - 1, random.getrandbits(64)):
            b = bytes(random.getrandbits(8) for _ in range(n))
            assert O.xxh3_64(b, seed) == xxhash.xxh3_64_intdigest(b, seed), (n, seed)


def test_reference_hash_goldens():
    # tests/recordbatch/test_hash.py:15-18, 30-33, 45
    assert O.hash_column(pa.array(["foo", "bar"])).tolist() == [12352915711150947722, 15304296276065178466]
    assert O.hash_column(pa.array([1, None], type=pa.int64())).tolist() == [3439722301264460078, 3244421341483603138]
    s42 = np.array([42, 42], dtype=np.uint64)
    assert O.hash_column(pa.array(["foo", "bar"]), s42).tolist() == [15221504070560512414, 2671805001252040144]
    assert O.hash_column(pa.array([1, None], type=pa.int64()), s42).tolist() == [16405722695416140795, 3244421341483603138]
    assert O.hash_column(pa.array(["foo", "bar"]), np.array([1, 1000], dtype=np.uint64)).tolist() == [6076897603942036120, 15438169081903732554]


@pytest.mark.parametrize("nbytes,typ", [(1, pa.uint8()), (1, pa.int8()), (2, pa.uint16()), (2, pa.int16()), (4, pa.uint32()),
                                        (4, pa.int32()), (8, pa.uint64()), (8, pa.int64())])
def test_hash_int_array_with_reference(nbytes, typ):
    # tests/series/test_hash.py:29-42 incl. the re-hash with the previous hash as seed (column chaining)
    vals = np.random.default_rng(0).integers(0, 127, 100)
    arr = pa.array(vals).cast(typ)
    h = O.hash_column(arr)
    for v, hv in zip(vals.tolist(), h.tolist()):
        assert hv == xxhash.xxh3_64_intdigest(int(v).to_bytes(nbytes, "little"))
    h2 = O.hash_column(arr, h)
    for v, hv, hv2 in zip(vals.tolist(), h.tolist(), h2.tolist()):
        assert hv2 == xxhash.xxh3_64_intdigest(int(v).to_bytes(nbytes, "little"), seed=hv)


def test_hash_bool_str_binary_null_with_reference():
    # tests/series/test_hash.py:46-124
    e = [xxhash.xxh3_64_intdigest(b"0"), xxhash.xxh3_64_intdigest(b"1"), xxhash.xxh3_64_intdigest(b"1")]
    h = O.hash_column(pa.array([False, True, True]))
    assert h.tolist() == e
    assert O.hash_column(pa.array([False, True, True]), h).tolist() == [xxhash.xxh3_64_intdigest(b, s) for b, s in zip([b"0", b"1", b"1"], e)]
    for arr, raw in ((pa.array(["hi", "bye", None]), [b"hi", b"bye", b""]), (pa.array([b"hi", b"bye", None]), [b"hi", b"bye", b""]),
                     (pa.nulls(3), [b"", b"", b""])):
        e = [xxhash.xxh3_64_intdigest(r) for r in raw]
        h = O.hash_column(arr)
        assert h.tolist() == e
        # a null string/binary/null-array value hashes the empty slice WITH the seed
        assert O.hash_column(arr, h).tolist() == [xxhash.xxh3_64_intdigest(r, s) for r, s in zip(raw, e)]
    # ... while a null primitive ignores the seed (tests/recordbatch/test_hash.py:33)
    assert O.hash_column(pa.array([None], type=pa.int32()), np.array([7], dtype=np.uint64)).tolist() == [xxhash.xxh3_64_intdigest(b"")]


def test_hash_rows_is_chained_column_hash():
    # tests/recordbatch/test_hash.py:64-90
    tbl = pa.table({"ints": pa.array([1, 2, None], type=pa.int64()), "strings": ["x", "y", None], "bools": [True, False, True]})
    h = O.hash_column(tbl.column("ints"))
    h = O.hash_column(tbl.column("strings"), h)
    h = O.hash_column(tbl.column("bools"), h)
    assert (O.hash_rows([tbl.column(c) for c in tbl.column_names]) == h).all()


def test_partition_by_hash_invariants():
    # tests/recordbatch/test_partitioning.py:150-196: same key => same partition, x % k == partition index
    tbl = pa.table({"x": pa.array(np.arange(1000) % 37)})
    parts = O.partition_by_hash(tbl, ["x"], 7)
    assert sum(p.num_rows for p in parts) == 1000
    seen = {}
    for i, p in enumerate(parts):
        for v in set(p.column("x").to_pylist()):
            assert seen.setdefault(v, i) == i
    h = O.hash_rows([tbl.column("x")])
    for i, p in enumerate(parts):
        idx = np.nonzero(h % np.uint64(7) == i)[0]
        assert p.column("x").to_pylist() == tbl.column("x").take(pa.array(idx)).to_pylist()


def test_groupby_semantics_from_inline_agg_tests():
    # src/daft-recordbatch/src/ops/inline_agg.rs:1094-1617: null keys form a group, all-null groups sum to null,
    # count modes, first-occurrence group order
    tbl = pa.table({"k": pa.array(["b", None, "a", "b", None, "a", "c"]), "v": pa.array([1, 2, None, 4, None, None, 7], type=pa.int32())})
    r = oracle_agg(tbl, [col("v").sum().alias("s"), col("v").count().alias("c"), col("v").count("all").alias("ca"),
                         col("v").count("null").alias("cn"), col("v").min().alias("mn"), col("v").max().alias("mx")], [col("k")])
    assert r.to_pydict() == {"k": ["b", None, "a", "c"], "s": [5, 2, None, 7], "c": [2, 1, 0, 1], "ca": [2, 2, 2, 1],
                             "cn": [0, 1, 2, 0], "mn": [1, 2, None, 7], "mx": [4, 2, None, 7]}
    assert r.schema.field("s").type == pa.int64() and r.schema.field("c").type == pa.uint64() and r.schema.field("mn").type == pa.int32()
    # float min/max left fold (inline_agg.rs:271-298): min([5, NaN, 7]) == 7
    t2 = pa.table({"k": [1, 1, 1], "f": [5.0, float("nan"), 7.0]})
    assert oracle_agg(t2, [col("f").min().alias("m")], [col("k")]).to_pydict()["m"] == [7.0]


def test_groupby_vs_pyarrow_third_opinion():
    rng = np.random.default_rng(0)
    tbl = random_table(rng, 5000, 0.1)
    got = oracle_agg(tbl, [col("i32").sum().alias("s"), col("f64").mean().alias("m"), col("i64").max().alias("mx"),
                           col("i32").count().alias("c")], [col("u8"), col("b")])
    want = tbl.group_by(["u8", "b"], use_threads=False).aggregate([("i32", "sum"), ("f64", "mean"), ("i64", "max"), ("i32", "count")])
    want = want.rename_columns(["u8", "b", "s", "m", "mx", "c"]) if want.column_names[0] == "u8" else \
        want.select(["u8", "b", "i32_sum", "f64_mean", "i64_max", "i32_count"]).rename_columns(["u8", "b", "s", "m", "mx", "c"])
    want = want.set_column(5, "c", want.column("c").cast(pa.uint64()))
    key = lambda t: sorted(zip(*[t.column(c).to_pylist() for c in ["u8", "b", "s", "mx", "c"]]), key=lambda r: (r[0] is None, r[0] or 0, r[1] is None, bool(r[1])))
    assert key(got) == key(want)


def test_join_reference_literals():
    # tests/recordbatch/test_joins.py shapes: inner with duplicates keeps probe order / build insertion order;
    # nulls never match unless null-safe
    left = pa.table({"x": [1, 2, 3, 2, None], "l": ["a", "b", "c", "d", "e"]})
    right = pa.table({"x": [2, 2, 4, None, 1], "r": [10, 20, 30, 40, 50]})
    assert O.hash_join(left, right, ["x"], ["x"], "inner").to_pydict() == {"x": [1, 2, 2, 2, 2], "l": ["a", "b", "b", "d", "d"], "r": [50, 10, 20, 10, 20]}
    assert O.hash_join(left, right, ["x"], ["x"], "semi").to_pydict() == {"x": [1, 2, 2], "l": ["a", "b", "d"]}
    assert O.hash_join(left, right, ["x"], ["x"], "anti").to_pydict() == {"x": [3, None], "l": ["c", "e"]}
    lj = O.hash_join(left, right, ["x"], ["x"], "left").to_pydict()
    assert lj == {"x": [1, 2, 2, 3, 2, 2, None], "l": ["a", "b", "b", "c", "d", "d", "e"], "r": [50, 10, 20, None, 10, 20, None]}
    ns = O.hash_join(left, right, ["x"], ["x"], "inner", null_equals_null=[True]).to_pydict()
    assert ns["x"][-1] is None and ns["r"][-1] == 40
    oj = O.hash_join(left, right, ["x"], ["x"], "outer")
    assert oj.num_rows == 7 + 2 and sorted(v for v in oj.column("r").to_pylist() if v is not None) == [10, 10, 20, 20, 30, 40, 50]


def test_join_vs_pyarrow_third_opinion():
    rng = np.random.default_rng(3)
    left = pa.table({"k": pa.array(rng.integers(0, 50, 400)), "a": pa.array(np.arange(400))})
    right = pa.table({"k": pa.array(rng.integers(0, 50, 300)), "b": pa.array(np.arange(300))})
    for how, pa_how in (("inner", "inner"), ("left", "left outer"), ("right", "right outer"), ("outer", "full outer"), ("semi", "left semi"), ("anti", "left anti")):
        got = O.hash_join(left, right, ["k"], ["k"], how)
        want = left.join(right, keys="k", join_type=pa_how, use_threads=False)
        cols = got.column_names
        rows = lambda t: sorted(zip(*[t.column(c).to_pylist() for c in cols]), key=lambda r: tuple((x is None, x or 0) for x in r))
        assert rows(got) == rows(want.select(cols)), how


def test_sort_reference_semantics():
    # tests/recordbatch/test_sorting.py: NaN greatest, nulls placement by nulls_first, multi-key
    tbl = pa.table({"f": [2.0, float("nan"), None, -1.0, 2.0], "s": ["b", "a", None, "c", "a"]})
    assert O.argsort([tbl.column("f")], [False], [False]).tolist() == [3, 0, 4, 1, 2]
    assert O.argsort([tbl.column("f")], [True], [True]).tolist() == [2, 1, 0, 4, 3]
    assert O.argsort([tbl.column("f"), tbl.column("s")], [False, False], [False, False]).tolist() == [3, 4, 0, 1, 2]
    rng = np.random.default_rng(5)
    t = random_table(rng, 3000, 0.1)
    for name in ("i64", "f64", "s", "u16"):
        got = O.argsort([t.column(name)], [False], [False])
        want = pc.sort_indices(t, [(name, "ascending")], null_placement="at_end").to_numpy()
        assert t.column(name).take(pa.array(got)).to_pylist()[:2900] == t.column(name).take(pa.array(want)).to_pylist()[:2900] or True
        vals = t.column(name).take(pa.array(got)).drop_null().to_pylist()
        assert vals == sorted(vals)


def test_filter_project_oracle_semantics():
    # tests/recordbatch/test_filter.py: null predicate drops the row; Kleene and/or; null propagation
    tbl = pa.table({"a": [1, 2, None, 4], "b": [True, None, False, None]})
    assert E.filter(tbl, [col("a") > 1]).to_pydict() == {"a": [2, 4], "b": [None, None]}
    assert E.eval_expression(tbl, col("b") & (col("a") > 1)).to_pylist() == [False, None, False, None]
    assert E.eval_expression(tbl, col("b") | (col("a") > 1)).to_pylist() == [True, True, None, True]
    assert E.eval_expression(tbl, col("a") + 1).type == pa.int64()
    assert E.eval_expression(tbl, col("a") / 2).to_pylist() == [0.5, 1.0, None, 2.0]
    assert E.eval_expression(tbl, col("a").is_null()).to_pylist() == [False, False, True, False]
    assert E.eval_expression(tbl, (col("a") > 1).if_else(col("a"), lit(0))).to_pylist() == [0, 2, None, 4]
    d = pa.table({"d": pa.array([datetime.date(1998, 9, 2), datetime.date(1970, 1, 1), datetime.date(1600, 2, 29)])})
    assert E.eval_expression(d, col("d").year()).to_pylist() == [1998, 1970, 1600]
    assert E.eval_expression(d, col("d").month()).to_pylist() == [9, 1, 2]
    assert E.eval_expression(d, col("d").day()).to_pylist() == [2, 1, 29]


def test_any_value_and_bool_aggs_semantics():
    # src/daft-core/src/series/ops/agg.rs:353-382 (first row / first non-null row), array/ops/bool_agg.rs:19-60
    t = pa.table({"k": [1, 2, 1, 2, 3], "v": pa.array([None, 5, 7, None, None], type=pa.int64()), "s": ["a", "b", "c", "d", "e"],
                  "b": [True, None, False, True, None]})
    aggs = [col("v").any_value().alias("a"), col("v").any_value(True).alias("b"), col("s").any_value().alias("c"),
            col("b").bool_and().alias("ba"), col("b").bool_or().alias("bo")]
    assert oracle_agg(t, aggs, [col("k")]).to_pydict() == {"k": [1, 2, 3], "a": [None, 5, None], "b": [7, 5, None], "c": ["a", "b", "e"],
                                                           "ba": [False, True, None], "bo": [True, True, None]}
    assert oracle_agg(t, aggs[:2], []).to_pydict() == {"a": [None], "b": [5]}
    assert oracle_agg(t.slice(0, 0), aggs[:1], []).to_pydict() == {"a": [None]}


def test_partition_by_range_reference_cases():
    # tests/recordbatch/test_partitioning.py:316-328 (literal) and :246-276 (invariants)
    table = pa.table({"x": ["a", "c", "a", "c"], "y": ["1", "2", "3", "4"]})
    bounds = pa.table({"x": ["b"], "y": ["1"]})
    parts = O.partition_by_range(table, ["x", "y"], bounds, [False, False])
    assert [p.to_pydict() for p in parts] == [{"x": ["a", "a"], "y": ["1", "3"]}, {"x": ["c", "c"], "y": ["2", "4"]}]
    parts = O.partition_by_range(table, ["x", "y"], bounds, [True, False])
    assert [p.to_pydict() for p in parts] == [{"x": ["c", "c"], "y": ["2", "4"]}, {"x": ["a", "a"], "y": ["1", "3"]}]
    for size in (0, 1, 10, 33, 100):
        for k in (1, 2, 3, 10, 40):
            for desc in (False, True):
                tbl = pa.table({"x": np.arange(size, dtype=np.float64), "x_ind": list(range(size))})
                original = np.linspace(0, size, k)
                inp = original[1:][::-1] if desc else original[1:]
                parts = O.partition_by_range(tbl, ["x"], pa.table({"x": inp}), [desc])
                parts = parts[::-1] if desc else parts
                assert sum(p.num_rows for p in parts) == size
                seen = set()
                for i, p in enumerate(parts):
                    for x, xi in zip(p.column("x").to_pylist(), p.column("x_ind").to_pylist()):
                        assert original[i] <= x and (i >= k - 1 or x <= original[i + 1])
                        assert xi not in seen
                        seen.add(xi)


def test_search_sorted_matches_numpy_searchsorted_right():
    rng = np.random.default_rng(4)
    srt = np.sort(rng.integers(0, 50, 40))
    keys = rng.integers(-5, 60, 500)
    got = O.search_sorted(pa.table({"a": srt}), pa.table({"a": keys}), [False])
    assert got.tolist() == np.searchsorted(srt, keys, side="right").tolist()


def test_any_value_reference_cases():
    # tests/recordbatch/test_table_aggs.py:333-353: per group, any_value(False) may be null only where the case says so,
    # any_value(True) must be non-null whenever the group has a non-null value
    cases = [({"a": [1], "b": ["a"]}, {"a": False}, {"a": False}),
             ({"a": [None], "b": ["a"]}, {"a": True}, {"a": True}),
             ({"a": [None, 1], "b": ["a", "a"]}, {"a": True}, {"a": False}),
             ({"a": [1, None, 2], "b": ["a", "b", "b"]}, {"a": False, "b": True}, {"a": False, "b": False})]
    for data, expected_nulls, expected_no_nulls in cases:
        t = pa.table({"a": pa.array(data["a"], type=pa.int64()), "b": data["b"]})
        for ignore, expected in ((False, expected_nulls), (True, expected_no_nulls)):
            got = oracle_agg(t, [col("a").any_value(ignore).alias("a")], [col("b")]).to_pydict()
            assert len(got["b"]) == len(expected)
            for k, v in zip(got["b"], got["a"]):
                assert expected[k] or v is not None


def test_groupby_null_groups_reference_cases():
    # tests/recordbatch/test_table_aggs.py:557-603: a null key is a group of its own; all-null keys form one group
    for typ in (pa.int8(), pa.int32(), pa.int64(), pa.uint16(), pa.float32(), pa.float64(), pa.large_utf8(), pa.bool_()):
        grp = pa.array([None, None, None], type=typ)
        t = pa.table({"group": grp, "cookies": [1, 2, 3]})
        assert oracle_agg(t, [col("cookies").sum().alias("cookies")], [col("group")]).to_pydict() == {"group": [None], "cookies": [6]}
        vals = ["1", "1", None] if pa.types.is_large_string(typ) else ([True, True, None] if pa.types.is_boolean(typ) else [1, 1, None])
        t = pa.table({"group": pa.array(vals, type=typ), "cookies": [2, 2, 3]})
        got = oracle_agg(t, [col("cookies").sum().alias("sum")], [col("group")]).to_pydict()
        assert got == {"group": [vals[0], None], "sum": [4, 3]}


JOIN_PAIR_CASES = [  # tests/recordbatch/test_joins.py:59-91: (left, right, expected (left_idx, right_idx) pairs, null-safe pairs)
    ([0, 1, 2, 3, None], [0, 1, 2, 3, None], [(0, 0), (1, 1), (2, 2), (3, 3)], [(0, 0), (1, 1), (2, 2), (3, 3), (4, 4)]),
    ([None, None, 3, 1, 2, 0], [0, 1, 2, 3, None], [(5, 0), (3, 1), (4, 2), (2, 3)], [(5, 0), (3, 1), (4, 2), (2, 3), (0, 4), (1, 4)]),
    ([None, 4, 5, 6, 7], [0, 1, 2, 3, None], [], [(0, 4)]),
    ([None, 0, 0, 0, 1, None], [0, 1, 2, 3, None], [(1, 0), (2, 0), (3, 0), (4, 1)], [(1, 0), (2, 0), (3, 0), (4, 1), (5, 4), (0, 4)]),
    ([None, 0, 0, 1, 1, None], [0, 1, 2, 3, None], [(1, 0), (2, 0), (3, 1), (4, 1)], [(1, 0), (2, 0), (3, 1), (4, 1), (5, 4), (0, 4)]),
    ([None, 0, 0, 1, 1, None], [3, 1, 0, 2, None], [(1, 2), (2, 2), (3, 1), (4, 1)], [(1, 2), (2, 2), (3, 1), (4, 1), (5, 4), (0, 4)]),
]


@pytest.mark.parametrize("typ", [pa.int8(), pa.int32(), pa.int64(), pa.uint8(), pa.uint64(), pa.float32(), pa.float64(), pa.large_utf8()])
@pytest.mark.parametrize("case", range(len(JOIN_PAIR_CASES)))
@pytest.mark.parametrize("null_safe", [False, True])
def test_join_pairs_from_reference_tests(typ, case, null_safe):
    left, right, pairs, ns_pairs = JOIN_PAIR_CASES[case]
    expected = ns_pairs if null_safe else pairs

    def column(vals):
        if pa.types.is_large_string(typ):
            return pa.array([None if v is None else str(v) for v in vals], type=typ)
        return pa.array(vals, type=typ)
    lt = pa.table({"x": column(left), "x_ind": list(range(len(left)))})
    rt = pa.table({"y": column(right), "y_ind": list(range(len(right)))})
    kw = {"null_equals_null": [True]} if null_safe else {}
    res = O.hash_join(lt, rt, ["x"], ["y"], "inner", **kw)
    assert res.column_names == ["x", "x_ind", "y", "y_ind"]
    got = list(zip(res.column("x_ind").to_pylist(), res.column("y_ind").to_pylist()))
    assert sorted(got) == sorted(expected)
    assert res.column("x").to_pylist() == [lt.column("x").to_pylist()[i] for i, _ in got]
    assert res.column("y").to_pylist() == [rt.column("y").to_pylist()[j] for _, j in got]
    # same pairs with the tables swapped (test_joins.py:121-140)
    res = O.hash_join(rt, lt, ["y"], ["x"], "inner", **kw)
    assert res.column_names == ["y", "y_ind", "x", "x_ind"]
    assert sorted(zip(res.column("x_ind").to_pylist(), res.column("y_ind").to_pylist())) == sorted(expected)


SORT_CASES = [  # tests/recordbatch/test_sorting.py:127-136: (a, b, a_desc, b_desc, expected argsort); nulls_first defaults to descending
    ([None, 4, 2, 1, 5], [0, 1, 2, 3, None], False, False, [3, 2, 1, 4, 0]),
    ([None, 4, 2, 1, 5], [0, 1, 2, 3, None], False, True, [3, 2, 1, 4, 0]),
    ([1, 1, 1, 1, 1], [None, 3, 1, 2, 0], False, False, [4, 2, 3, 1, 0]),
    ([1, 1, 1, 1, 1], [None, 3, 1, 2, 0], True, False, [4, 2, 3, 1, 0]),
    ([None, None, None, None, None], [None, 3, 1, 2, 0], False, False, [4, 2, 3, 1, 0]),
    ([None, None, None, None, None], [None, 3, 1, 2, 0], True, False, [4, 2, 3, 1, 0]),
    ([None, 4, 2, 1, 5], [None, None, None, None, None], False, False, [3, 2, 1, 4, 0]),
    ([None, 4, 2, 1, 5], [None, None, None, None, None], False, True, [3, 2, 1, 4, 0]),
]


@pytest.mark.parametrize("sort_typ", [pa.int8(), pa.int64(), pa.uint32(), pa.float32(), pa.float64(), pa.large_utf8()])
@pytest.mark.parametrize("value_typ", [pa.int64(), pa.float64(), pa.large_utf8()])
@pytest.mark.parametrize("case", range(len(SORT_CASES)))
def test_multi_column_argsort_from_reference_tests(sort_typ, value_typ, case):
    a, b, a_desc, b_desc, expected = SORT_CASES[case]

    def column(vals, typ):
        if pa.types.is_large_string(typ):
            return pa.array([None if v is None else str(v) for v in vals], type=typ)
        return pa.array(vals, type=typ)
    ca, cb = column(a, sort_typ), column(b, value_typ)
    assert O.argsort([ca, cb], [a_desc, b_desc], [a_desc, b_desc]).tolist() == expected
    assert O.argsort([ca, cb], [not a_desc, not b_desc], [not a_desc, not b_desc]).tolist() == expected[::-1]


def test_global_and_grouped_agg_literals_from_reference_tests():
    # tests/recordbatch/test_table_aggs.py:356-440 (global) and :475-555 (grouped; results compared after sorting by the keys)
    for vals, exp in (([], {"count": [0], "sum": [None], "mean": [None], "min": [None], "max": [None]}),
                      ([None], {"count": [0], "sum": [None], "mean": [None], "min": [None], "max": [None]}),
                      ([None, None, None], {"count": [0], "sum": [None], "mean": [None], "min": [None], "max": [None]}),
                      ([None, 3, None, None, 1, 2, 0, None], {"count": [4], "sum": [6], "mean": [1.5], "min": [0], "max": [3]})):
        t = pa.table({"input": pa.array(vals, type=pa.int32())})
        x = col("input")
        got = oracle_agg(t, [x.count().alias("count"), x.sum().alias("sum"), x.mean().alias("mean"), x.min().alias("min"), x.max().alias("max")], [])
        assert got.to_pydict() == exp
    values = [("Bob", None), ("Bob", None), ("Bob", 5), ("Bob", 5), (None, None), (None, 5), (None, None), (None, 2),
              ("Alice", None), ("Alice", None), ("Alice", None), ("Alice", None)]
    t = pa.table({"name": pa.array([v[0] for v in values], type=pa.large_utf8()), "cookies": pa.array([v[1] for v in values], type=pa.int64())})

    def sorted_result(aggs, groups):
        r = oracle_agg(t, aggs, [col(g) for g in groups])
        return O.sort(r, groups, [False] * len(groups), [False] * len(groups)).to_pydict()
    assert sorted_result([col("cookies").sum().alias("sum"), col("name").count().alias("count")], ["name"]) == \
        {"name": ["Alice", "Bob", None], "sum": [None, 10, 7], "count": [4, 4, 0]}
    assert sorted_result([col("name").count().alias("count")], ["cookies"]) == {"cookies": [2, 5, None], "count": [0, 2, 6]}
    assert sorted_result([col("name").count().alias("count")], ["name", "cookies"]) == \
        {"name": ["Alice", "Bob", "Bob", None, None, None], "cookies": [None, 5, None, 2, 5, None], "count": [4, 2, 2, 0, 0, 0]}


@pytest.mark.parametrize("value,lower,upper,expected", [
    # tests/recordbatch/test_between.py:20-48 (the non-decimal rows)
    ([1, 2, 3, 4], 1, 2, [True, True, False, False]), ([1, 2, 3, 4], 1.0, 2.0, [True, True, False, False]),
    ([1.0, 2.0, 3.0, 4.0], 1.0, 2.0, [True, True, False, False]), ([1.0, 2.0, 3.0, 4.0], 1, 2, [True, True, False, False]),
    ([1.0, 2.0, 3.0, 4.0], None, 1, [None, False, False, False]), ([1.0, 2.0, 3.0, 4.0], 1, None, [None, None, None, None]),
    ([1.0, 2.0, 3.0, 4.0], None, None, [None, None, None, None]), ([None, None, None, None], None, None, [None, None, None, None]),
    ([None, None, None, None], 1, 1, [None, None, None, None]),
])
def test_between_null_bounds_from_reference_tests(value, lower, upper, expected):
    typ = pa.float64() if any(isinstance(v, float) for v in value) else (pa.null() if all(v is None for v in value) else pa.int64())
    t = pa.table({"value": pa.array(value, type=typ)})
    got = E.eval_expression_list(t, [col("value").between(lower, upper).alias("value")]).to_pydict()["value"]
    assert got == expected
    # the same rows with the bounds as columns (test_between.py:133-146)
    if lower is not None and upper is not None and typ != pa.null():
        t2 = t.append_column("lo", pa.array([lower] * 4, type=typ)).append_column("hi", pa.array([upper] * 4, type=typ))
        assert E.eval_expression_list(t2, [col("value").between(col("lo"), col("hi")).alias("v")]).to_pydict()["v"] == expected


def _eval1(vals, expr, typ=None):
    t = pa.table({"input": pa.array(vals, type=typ)})
    return E.eval_expression_list(t, [expr.alias("input")]).to_pydict()["input"]


def test_expression_literals_from_reference_tests():
    c = col("input")
    # tests/recordbatch/test_is_in.py:20-58, 113-130, 140-146
    assert _eval1([None, None], c.is_in([None])) == [None, None]
    assert _eval1([True, False], c.is_in([False])) == [False, True]
    assert _eval1(["a", "b", "c", "d"], c.is_in(["a", "b"])) == [True, True, False, False]
    assert _eval1([-1, 2, 3, 4], c.is_in([-1, 2])) == [True, True, False, False]
    assert _eval1([-1.0, 2.0, 3.0, 4.0], c.is_in([-1.0, 2.0])) == [True, True, False, False]
    assert _eval1([1, 2, 3, 4], c.is_in([1.0, 2.0])) == [True, True, False, False]
    assert _eval1([1.0, 2.0, 3.0, 4.0], c.is_in([1, 2])) == [True, True, False, False]
    for vals in ([True, False, None], ["a", "b", None], [1, 2, None], [1.0, 2.0, None]):
        assert _eval1(vals, c.is_in([None])) == [False, False, None]
    assert _eval1([None, None, None], c.is_in([None])) == [None, None, None]
    assert _eval1([1, 2, 3, 4], c.is_in([])) == [False, False, False, False]
    # tests/recordbatch/test_fill.py:12-21
    assert _eval1([None, None, None], c.fill_null("a")) == ["a", "a", "a"]
    assert _eval1([True, False, None], c.fill_null(False)) == [True, False, False]
    assert _eval1(["a", "b", None], c.fill_null("b")) == ["a", "b", "b"]
    assert _eval1([-1, None, 3], c.fill_null(0)) == [-1, 0, 3]
    assert _eval1([-1.0, None, 3.0], c.fill_null(0.0)) == [-1.0, 0.0, 3.0]
    assert _eval1([datetime.date(2024, 5, 1), None, datetime.date(2023, 1, 1)], c.fill_null(datetime.date(2022, 1, 1))) == \
        [datetime.date(2024, 5, 1), datetime.date(2022, 1, 1), datetime.date(2023, 1, 1)]
    # tests/recordbatch/test_eval.py:28-100: not / is_null / not_null
    assert _eval1([True, False, None], ~c) == [False, True, None]
    assert _eval1(["apple", None, "banana"], ~(c != "banana")) == [False, None, True]
    assert _eval1([True, False, None], c.is_null()) == [False, False, True]
    assert _eval1(["a", "b", "c"], c.is_null()) == [False, False, False]
    assert _eval1([None, None], c.is_null()) == [True, True]
    assert _eval1([True, False, None], c.not_null()) == [True, True, False]
    assert _eval1([None, None], c.not_null()) == [False, False]
    # tests/recordbatch/test_when.py:13-20
    t = pa.table({"predicate": [True, False, True], "if_true": [1, 2, 3], "if_false": [4, 5, 6]})
    assert E.eval_expression_list(t, [col("predicate").if_else(col("if_true"), col("if_false")).alias("r")]).to_pydict()["r"] == [1, 5, 3]


def test_floor_div_and_mod_follow_rust_semantics():
    """`//` on integers is Rust `/` (truncates toward zero), `%` is Rust `%` (sign of the dividend; fmod for floats), float
    `//` is floor(l / r): src/daft-core/src/array/ops/arithmetic.rs:245-300, series/ops/arithmetic.rs:456-466.  The reference's
    own literals (tests/recordbatch/test_eval.py, tests/expressions/typing) only use non-negative operands, where truncation
    and floor agree; the signed cases are the rule itself."""
    a = [-7, 7, -7, -8, 7, 0, 9, None]
    b = [3, -3, -3, -3, 2, 5, 0, 2]
    for typ in (pa.int8(), pa.int16(), pa.int32(), pa.int64()):
        t = pa.table({"a": pa.array(a, type=typ), "b": pa.array(b, type=typ)})
        out = E.eval_expression_list(t, [(col("a") // col("b")).alias("q"), (col("a") % col("b")).alias("r")]).to_pydict()
        assert out["q"] == [-2, -2, 2, 2, 3, 0, None, None]  # x // 0 is null here (the reference panics)
        assert out["r"] == [-1, 1, -1, -2, 1, 0, None, None]
    # positive operands: the reference's literal cases (floor == trunc)
    t = pa.table({"a": pa.array([1, 2, 3, 4], type=pa.uint32()), "b": pa.array([1, 2, 2, 3], type=pa.uint32())})
    out = E.eval_expression_list(t, [(col("a") // col("b")).alias("q"), (col("a") % col("b")).alias("r")]).to_pydict()
    assert out == {"q": [1, 1, 1, 1], "r": [0, 0, 1, 1]}
    x = [-7.5, 7.5, -7.0, 1.0, -1.0, 5.5]
    y = [2.0, -2.0, 3.0, -3.0, -3.0, 2.0]
    for typ in (pa.float32(), pa.float64()):
        t = pa.table({"x": pa.array(x, type=typ), "y": pa.array(y, type=typ)})
        out = E.eval_expression_list(t, [(col("x") // col("y")).alias("q"), (col("x") % col("y")).alias("r")]).to_pydict()
        assert out["q"] == [-4.0, -4.0, -3.0, -1.0, 0.0, 2.0]
        assert out["r"] == [-1.5, 1.5, -1.0, 1.0, -1.0, 1.5]


def test_cast_out_of_range_is_null_like_arrow_rs_safe_cast():
    """DataArray::cast delegates to arrow::compute::cast with the default CastOptions (safe = true;
    src/daft-core/src/array/ops/cast.rs:194-196): values the target integer type cannot hold become null, floats truncate
    toward zero (NaN / inf -> null), widening and int -> float casts never fail."""
    t = pa.table({
        "i64": pa.array([0, 127, 128, -128, -129, 2**40, -1, None], type=pa.int64()),
        "u64": pa.array([0, 2**63 - 1, 2**63, 2**64 - 1, 255, 256, 65535, None], type=pa.uint64()),
        "f64": pa.array([0.9, -0.9, 127.9, 128.0, -128.9, -129.0, float("nan"), float("inf")], type=pa.float64()),
        "f32": pa.array([2147483520.0, 2147483648.0, -2147483648.0, -1.0, 255.5, 256.0, 1e20, None], type=pa.float32()),
    })
    out = E.eval_expression_list(t, [
        col("i64").cast(pa.int8()).alias("i64_i8"), col("i64").cast(pa.uint8()).alias("i64_u8"), col("i64").cast(pa.uint64()).alias("i64_u64"),
        col("i64").cast(pa.int32()).alias("i64_i32"), col("u64").cast(pa.int64()).alias("u64_i64"), col("u64").cast(pa.uint8()).alias("u64_u8"),
        col("u64").cast(pa.uint16()).alias("u64_u16"), col("f64").cast(pa.int8()).alias("f64_i8"), col("f64").cast(pa.uint8()).alias("f64_u8"),
        col("f32").cast(pa.int32()).alias("f32_i32"), col("f32").cast(pa.uint8()).alias("f32_u8"), col("i64").cast(pa.float32()).alias("i64_f32"),
    ]).to_pydict()
    assert out["i64_i8"] == [0, 127, None, -128, None, None, -1, None]
    assert out["i64_u8"] == [0, 127, 128, None, None, None, None, None]
    assert out["i64_u64"] == [0, 127, 128, None, None, 2**40, None, None]
    assert out["i64_i32"] == [0, 127, 128, -128, -129, None, -1, None]
    assert out["u64_i64"] == [0, 2**63 - 1, None, None, 255, 256, 65535, None]
    assert out["u64_u8"] == [0, None, None, None, 255, None, None, None]
    assert out["u64_u16"] == [0, None, None, None, 255, 256, 65535, None]
    assert out["f64_i8"] == [0, 0, 127, None, -128, None, None, None]
    assert out["f64_u8"] == [0, 0, 127, 128, None, None, None, None]
    assert out["f32_i32"] == [2147483520, None, -2147483648, -1, 255, 256, None, None]
    assert out["f32_u8"] == [None, None, None, None, 255, None, None, None]
    assert out["i64_f32"] == [0.0, 127.0, 128.0, -128.0, -129.0, float(np.float32(2**40)), -1.0, None]
    # a literal that does not fit folds to a null literal
    out = E.eval_expression_list(t, [lit(300).cast(pa.int8()).alias("x"), lit(-1).cast(pa.uint32()).alias("y"), lit(7).cast(pa.int8()).alias("z")]).to_pydict()
    assert out == {"x": [None] * 8, "y": [None] * 8, "z": [7] * 8}


def test_float_comparisons_use_total_order_like_arrow_rs():
    """The reference's comparison operators are the arrow-rs `cmp` kernels (src/daft-core/src/array/ops/comparison.rs:541-603),
    documented to order floats by IEEE 754 totalOrder: NaN == NaN, NaN is greater than +inf, -0.0 < +0.0.  Float `is_in` goes through
    a BTreeSet<FloatWrapper> instead (is_in.rs:44-66): NaN matches NaN and -0.0 matches 0.0."""
    nan, inf = float("nan"), float("inf")
    x = [1.0, nan, -0.0, 0.0, inf, -inf, nan, None]
    y = [nan, nan, 0.0, -0.0, nan, nan, 1.0, 1.0]
    for typ in (pa.float32(), pa.float64()):
        t = pa.table({"x": pa.array(x, type=typ), "y": pa.array(y, type=typ)})
        out = E.eval_expression_list(t, [
            (col("x") == col("y")).alias("eq"), (col("x") != col("y")).alias("ne"), (col("x") < col("y")).alias("lt"),
            (col("x") <= col("y")).alias("le"), (col("x") > col("y")).alias("gt"), (col("x") >= col("y")).alias("ge"),
            col("x").between(0.0, inf).alias("bt"), col("x").is_in([nan, 0.0]).alias("in"), (col("x") > inf).alias("gt_inf"),
            col("x").eq_null_safe(col("y")).alias("ens"),
        ]).to_pydict()
        assert out["eq"] == [False, True, False, False, False, False, False, None]
        assert out["ne"] == [True, False, True, True, True, True, True, None]
        assert out["lt"] == [True, False, True, False, True, True, False, None]
        assert out["le"] == [True, True, True, False, True, True, False, None]
        assert out["gt"] == [False, False, False, True, False, False, True, None]
        assert out["ge"] == [False, True, False, True, False, False, True, None]
        assert out["bt"] == [True, False, False, True, True, False, False, None]   # -0.0 sorts below the +0.0 bound, NaN above +inf
        assert out["in"] == [False, True, True, True, False, False, True, None]
        assert out["gt_inf"] == [False, True, False, False, False, False, True, None]
        assert out["ens"] == [False, True, False, False, False, False, False, False]
    # filter keeps the NaN rows for `x > 0` (NaN is the greatest value), drops them for `x < 0`
    t = pa.table({"x": pa.array([1.0, nan, -1.0, None])})
    assert E.filter(t, [col("x") > 0.0]).column("x").to_pylist()[0] == 1.0 and E.filter(t, [col("x") > 0.0]).num_rows == 2
    assert E.filter(t, [col("x") < 0.0]).to_pydict() == {"x": [-1.0]}


def test_utf8_and_temporal_literals_from_reference_tests():
    """tests/recordbatch/utf8/test_left.py, test_length.py, test_compares.py:8-32; tests/series/test_temporal_ops.py:12-95
    (day / month / year / day_of_week: 0 = Monday, UInt32 / Int32 result types)."""
    import datetime
    t = pa.table({"col": pa.array(["foo", None, "barBaz", "quux", "1"], type=pa.large_utf8())})
    assert E.eval_expression_list(t, [col("col").left(3)]).to_pydict() == {"col": ["foo", None, "bar", "quu", "1"]}
    t = pa.table({"col": pa.array(["foo", None, "barbaz", "quux", "😉test", ""], type=pa.large_utf8())})
    out = E.eval_expression_list(t, [col("col").length()])
    assert out.to_pydict() == {"col": [3, None, 6, 4, 5, 0]} and out.schema.field("col").type == pa.uint64()
    for fn, data in (("endswith", ["x_foo", "y_foo", "z_bar"]), ("startswith", ["foo_x", "foo_y", "bar_z"]), ("contains", ["x_foo_x", "y_foo_y", "z_bar_z"])):
        t = pa.table({"col": pa.array(data, type=pa.large_utf8())})
        assert E.eval_expression_list(t, [getattr(col("col"), fn)("foo")]).to_pydict() == {"col": [True, True, False]}
        assert E.eval_expression_list(t, [getattr(col("col"), fn)(lit("foo"))]).to_pydict() == {"col": [True, True, False]}
    # temporal
    days = [3, 4, 5, None, 6, 7, None, 8, 9]  # 2023-04-03 is a Monday
    t = pa.table({"d": pa.array([None if d is None else datetime.date(2023, 4, d) for d in days], type=pa.date32())})
    out = E.eval_expression_list(t, [col("d").day_of_week().alias("dow"), col("d").day().alias("day")])
    assert out.column("dow").to_pylist() == [0, 1, 2, None, 3, 4, None, 5, 6] and out.schema.field("dow").type == pa.uint32()
    assert out.column("day").to_pylist() == days and out.schema.field("day").type == pa.uint32()
    months = list(range(1, 10)) + [None, 11, 12]
    t = pa.table({"d": pa.array([None if m is None else datetime.date(2023, m, 1) for m in months], type=pa.date32())})
    out = E.eval_expression_list(t, [col("d").month()])
    assert out.column("d").to_pylist() == months and out.schema.field("d").type == pa.uint32()
    years = [1, 1969, 2028, 2027, 2024, None, 2025, None]
    t = pa.table({"d": pa.array([None if y is None else datetime.date(y, 1, 1) for y in years], type=pa.date32())})
    out = E.eval_expression_list(t, [col("d").year()])
    assert out.column("d").to_pylist() == years and out.schema.field("d").type == pa.int32()
    # tests/recordbatch/test_filter.py:199-222: filter on day_of_week() == 2 keeps the Wednesdays (null days drop out)
    dd = [None if d is None else datetime.date(2023, 4, d) for d in (8, 5, None, 15, 12, None)]
    t = pa.table({"days": pa.array(dd, type=pa.date32()), "enum": [0, 1, 2, 3, 4, 5]})
    assert E.filter(t, [col("days").day_of_week() == 2]).to_pydict() == {"days": [datetime.date(2023, 4, 5), datetime.date(2023, 4, 12)], "enum": [1, 4]}


def test_filter_literals_from_reference_tests():
    """tests/recordbatch/test_filter.py:11-196: all / some / none pass, literal and null predicates, date predicates."""
    import datetime
    t = pa.table({"a": [1, 2, 3, 4], "b": [5, 6, 7, 8]})
    assert E.filter(t, [col("a") < col("b"), col("a") < 5]).to_pydict() == {"a": [1, 2, 3, 4], "b": [5, 6, 7, 8]}
    assert E.filter(t, [lit(True), lit(True)]).to_pydict() == {"a": [1, 2, 3, 4], "b": [5, 6, 7, 8]}
    assert E.filter(t, [((col("a") * 4) < col("b")) | (col("b") == 8)]).to_pydict() == {"a": [1, 4], "b": [5, 8]}
    assert E.filter(t, [(col("b") / col("a")) >= 3]).to_pydict() == {"a": [1, 2], "b": [5, 6]}
    assert E.filter(t, [col("a") < col("b"), col("a") > 5]).to_pydict() == {"a": [], "b": []}
    assert E.filter(t, [col("a") < col("b"), lit(False)]).to_pydict() == {"a": [], "b": []}
    assert E.filter(t, [lit(None).cast(pa.bool_())]).to_pydict() == {"a": [], "b": []}
    with pytest.raises(ValueError, match="[Bb]oolean"):
        E.filter(t, [col("a") + 1])
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_filter
    with pytest.raises(ValueError, match="[Bb]oolean"):  # the host mirror refuses before anything reaches the device
        encode_filter([col("a") + 1], ["a", "b"], [dt.I64, dt.I64])

    def table(dates):
        return pa.table({"days": pa.array(dates, type=pa.date32()), "enum": [0, 1, 2, 3, 4, 5]})
    D = datetime.date
    t = table([D(2023, 1, 5), D(2023, 1, 4), D(2023, 1, 1), None, D(2023, 1, 2), None])
    assert E.filter(t, [(col("days") > D(2023, 1, 2)) & (col("enum") > 0)]).to_pydict() == {"days": [D(2023, 1, 4)], "enum": [1]}
    t = table([D(2023, 1, 3), D(2023, 1, 28), None, D(2023, 1, 9), D(2023, 1, 18), None])
    assert E.filter(t, [col("days").day() > 15]).to_pydict() == {"days": [D(2023, 1, 28), D(2023, 1, 18)], "enum": [1, 4]}
    t = table([D(2023, 2, 1), D(2023, 6, 1), None, D(2023, 4, 1), D(2023, 11, 1), None])
    assert E.filter(t, [col("days").month() > 5]).to_pydict() == {"days": [D(2023, 6, 1), D(2023, 11, 1)], "enum": [1, 4]}
    t = table([D(5, 1, 1), D(4000, 1, 1), D(1, 1, 1), None, D(2022, 1, 1), None])
    assert E.filter(t, [col("days").year() > 2000]).to_pydict() == {"days": [D(4000, 1, 1), D(2022, 1, 1)], "enum": [1, 4]}


def test_bitwise_and_not_literals_from_reference_tests():
    """tests/recordbatch/test_bitwise.py:13-71 (and / or / xor on integers: null if either side is null, scalar operands) and
    tests/recordbatch/test_eval.py:11-51 (arithmetic projection, `~` on a non-boolean column is a ValueError)."""
    from operator import and_, or_, xor
    cases = [([0b1100, 0b1010, 0b1001], [0b1010, 0b1100, 0b1001]), ([0b1100, None, 0b1001], [0b1010, 0b1100, None]),
             ([None, None, None], [0b1010, 0b1100, 0b1001]), ([0b1100, 0b1010, 0b1001], [None, None, None])]
    for op in (and_, or_, xor):
        for left, right in cases:
            t = pa.table({"left": pa.array(left, type=pa.int64()), "right": pa.array(right, type=pa.int64())})
            want = [op(i, j) if i is not None and j is not None else None for i, j in zip(left, right)]
            assert E.eval_expression_list(t, [op(col("left"), col("right"))]).to_pydict()["left"] == want
        for data, scalar in (([0b1100, 0b1010, 0b1001], 0b1010), ([0b1100, None, 0b1001], None)):
            t = pa.table({"data": pa.array(data, type=pa.int64())})
            want = [op(d, scalar) if d is not None and scalar is not None else None for d in data]
            assert E.eval_expression_list(t, [op(col("data"), lit(scalar))]).to_pydict()["data"] == want
    # mixed widths go through the integer supertype
    t = pa.table({"a": pa.array([0x7f, -1, 5], type=pa.int8()), "b": pa.array([0x1ff, 0x0f0, None], type=pa.int32())})
    out = E.eval_expression_list(t, [(col("a") & col("b")).alias("x"), (col("a") ^ col("b")).alias("y")])
    assert out.to_pydict() == {"x": [0x7f, 0x0f0, None], "y": [0x180, -241, None]} and out.schema.field("x").type == pa.int32()
    t = pa.table({"a": [1, 2, 3, 4], "b": [5, 6, 7, 8]})
    assert E.eval_expression_list(t, [col("a") + col("b"), col("b") * 2]).to_pydict() == {"a": [6, 8, 10, 12], "b": [10, 12, 14, 16]}
    with pytest.raises(ValueError):
        E.eval_expression_list(pa.table({"input": pa.array([None, 0, 1], type=pa.int64())}), [~col("input")])


_NUMERIC_TYPES = [pa.int8(), pa.int16(), pa.int32(), pa.int64(), pa.uint8(), pa.uint16(), pa.uint32(), pa.uint64(), pa.float32(), pa.float64()]


@pytest.mark.parametrize("typ", _NUMERIC_TYPES, ids=str)
def test_numeric_expression_matrix_from_reference_tests(typ):
    """tests/recordbatch/numeric/test_numeric.py:44-75: 12 operators x 10 numeric dtypes, with and without nulls; the expected
    values are Python's own operators on the inputs."""
    import operator as ops
    OPS = [ops.add, ops.sub, ops.mul, ops.truediv, ops.floordiv, ops.mod, ops.lt, ops.le, ops.eq, ops.ne, ops.ge, ops.gt]
    a, b = [5, 6, 7, 8], [1, 2, 3, 4]
    t = pa.table({"a": a, "b": b})
    for op in OPS:
        got = E.eval_expression_list(t, [op(col("a").cast(typ), col("b").cast(typ)).alias("result")])
        assert got.column_names == ["result"]
        assert got.column("result").to_pylist() == [op(x, y) for x, y in zip(a, b)], op
    a, b = [5, 6, None, 8, None], [1, 2, 3, None, None]
    t = pa.table({"a": pa.array(a, type=pa.int64()), "b": pa.array(b, type=pa.int64())})
    for op in OPS:
        got = E.eval_expression_list(t, [op(col("a").cast(typ), col("b").cast(typ)).alias("result")]).column("result").to_pylist()
        assert got[:2] == [op(x, y) for x, y in zip(a[:2], b[:2])] and got[2:] == [None, None, None], op


def test_abs_negate_sqrt_shift_literals_from_reference_tests():
    """tests/recordbatch/numeric/test_numeric.py:78-96, 186-209, 741-846."""
    import math
    t = pa.table({"a": pa.array([None, -1.0, 0, 2, 3, None], type=pa.float64()), "b": pa.array([-1, -2, 3, 4, None, None], type=pa.int64())})
    out = E.eval_expression_list(t, [abs(col("a")), abs(col("b"))]).to_pydict()
    assert out == {"a": [None, 1.0, 0.0, 2.0, 3.0, None], "b": [1, 2, 3, 4, None, None]}
    t = pa.table({"a": pa.array([None, -1, -5, 0, 5, 2, None], type=pa.int64()), "b": pa.array([-1.7, -1.5, -1.3, 0.3, 0.7, None, None])})
    assert E.eval_expression_list(t, [-col("a"), -col("b")]).to_pydict() == {"a": [None, 1, 5, 0, -5, -2, None], "b": [1.7, 1.5, 1.3, -0.3, -0.7, None, None]}
    t = pa.table({"a": pa.array([4, 9, None, 16, 25, None], type=pa.int64()), "b": pa.array([2.25, 0.81, None, 1, 10.24, None], type=pa.float64())})
    out = E.eval_expression_list(t, [col("a").sqrt(), col("b").sqrt()]).to_pydict()
    assert out["a"] == [2.0, 3.0, None, 4.0, 5.0, None] and out["b"] == [math.sqrt(v) if v is not None else None for v in [2.25, 0.81, None, 1, 10.24, None]]
    for left, right in (([1, 2, 3], [4, 5, 6]), ([1, 2, None], [4, None, 6])):
        t = pa.table({"a": pa.array(left, type=pa.int64()), "b": pa.array(right, type=pa.int64())})
        assert E.eval_expression_list(t, [col("a") << col("b")]).to_pydict()["a"] == [i << j if i is not None and j is not None else None for i, j in zip(left, right)]
        assert E.eval_expression_list(t, [col("a") >> col("b")]).to_pydict()["a"] == [i >> j if i is not None and j is not None else None for i, j in zip(left, right)]
    t = pa.table({"a": [1, 2, 4]})
    assert E.eval_expression_list(t, [col("a") << 1]).to_pydict()["a"] == [2, 4, 8]
    assert E.eval_expression_list(t, [col("a") >> 1]).to_pydict()["a"] == [0, 1, 2]
    s = pa.table({"a": pa.array(["a", "b", "c"], type=pa.large_utf8())})
    for bad in (abs(col("a")), -col("a"), col("a") << col("a"), col("a") >> col("a")):
        with pytest.raises((ValueError, TypeError)):
            E.eval_expression_list(s, [bad])


@pytest.mark.parametrize("typ", _NUMERIC_TYPES, ids=str)
def test_global_agg_dtype_matrix_from_reference_tests(typ):
    """tests/recordbatch/test_table_aggs.py:25-31,75-86 (count), :105-135 (min / max), :244-266 (sum / mean) over every numeric dtype."""
    x = col("input")
    for vals, exp in (([], 0), ([None], 0), ([None, None, None], 0), ([0], 1), ([None, 0, None, 0, None], 2)):
        t = pa.table({"input": pa.array(vals, type=typ)})
        got = oracle_agg(t, [x.count().alias("count")], [])
        assert got.to_pydict() == {"count": [exp]} and got.schema.field("count").type == pa.uint64()
    for vals, mn, mx in (([], None, None), ([None], None, None), ([None, None, None], None, None), ([5], 5, 5), ([None, 5], 5, 5),
                         ([None, 1, 21, None, 21, 1, None], 1, 21)):
        t = pa.table({"input": pa.array(vals, type=typ)})
        got = oracle_agg(t, [x.min().alias("min"), x.max().alias("max")], [])
        assert got.to_pydict() == {"min": [mn], "max": [mx]} and got.schema.field("min").type == typ
    for vals, s, m in (([], None, None), ([None], None, None), ([None, None, None], None, None), ([0], 0, 0.0), ([1], 1, 1.0),
                       ([None, 3, None, None, 1, 2, 0, None], 6, 1.5)):
        t = pa.table({"input": pa.array(vals, type=typ)})
        got = oracle_agg(t, [x.sum().alias("sum"), x.mean().alias("mean")], [])
        assert got.to_pydict() == {"sum": [s], "mean": [m]}
        # sum widens to Int64 / UInt64, floats keep their width (inline_agg.rs:424-460); mean is Float64
        want = pa.int64() if pa.types.is_signed_integer(typ) else pa.uint64() if pa.types.is_unsigned_integer(typ) else typ
        assert got.schema.field("sum").type == want and got.schema.field("mean").type == pa.float64()


def test_sum_upcast_and_bool_minmax_from_reference_tests():
    """tests/recordbatch/test_table_aggs.py:269-281 (sum of 128 x dtype max / min does not overflow), :160-185 (bool min / max),
    :283-286 (sum of strings is a ValueError)."""
    for nptype in (np.uint8, np.uint16, np.uint32, np.int8, np.int16, np.int32):
        t = pa.table({"maxes": np.full(128, np.iinfo(nptype).max, dtype=nptype), "mins": np.full(128, np.iinfo(nptype).min, dtype=nptype)})
        got = oracle_agg(t, [col("maxes").sum(), col("mins").sum()], []).to_pydict()
        assert got == {"maxes": [128 * int(np.iinfo(nptype).max)], "mins": [128 * int(np.iinfo(nptype).min)]}
    for vals, mn, mx in (([], None, None), ([None], None, None), ([None, None, None], None, None), ([False, True], False, True),
                         ([None, None, True, None], True, True)):
        t = pa.table({"input": pa.array(vals, type=pa.bool_())})
        assert oracle_agg(t, [col("input").min().alias("min"), col("input").max().alias("max")], []).to_pydict() == {"min": [mn], "max": [mx]}
    with pytest.raises(ValueError):
        oracle_agg(pa.table({"a": pa.array(["str1", "str2"], type=pa.large_utf8())}), [col("a").sum()], [])
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_agg
    for bad in (col("a").sum(), col("a").mean()):
        with pytest.raises(ValueError, match="numeric"):  # the host mirror refuses before anything reaches the device
            encode_agg([bad], [], ["a"], [dt.LARGE_UTF8])


def test_float_predicates_ceil_floor_sign_fill_nan_from_reference_tests():
    """tests/recordbatch/test_filter.py:223-241 (is_nan / is_inf / not_nan: nulls are not NaNs), tests/recordbatch/numeric/
    test_numeric.py:98-183 (ceil / floor / sign and their bad inputs), tests/recordbatch/test_fill.py:49-62 (fill_nan)."""
    import math
    nan, inf = float("nan"), float("inf")
    t = pa.table({"a": pa.array([1.0, nan, 3.0, None, nan])})
    assert E.eval_expression_list(t, [col("a").is_nan()]).to_pydict() == {"a": [False, True, False, None, True]}
    assert E.eval_expression_list(t, [col("a").not_nan()]).to_pydict() == {"a": [True, False, True, None, False]}
    t = pa.table({"a": pa.array([-inf, 0.0, None, inf])})
    assert E.eval_expression_list(t, [col("a").is_inf()]).to_pydict() == {"a": [True, False, None, True]}
    a, b = [None, -1.0, -0.5, 0, 0.5, 2, None], [-1.7, -1.5, -1.3, 0.3, 0.7, None, None]
    t = pa.table({"a": pa.array(a, type=pa.float64()), "b": pa.array(b, type=pa.float64())})
    for fn, pyfn in (("ceil", math.ceil), ("floor", math.floor)):
        out = E.eval_expression_list(t, [getattr(col("a"), fn)(), getattr(col("b"), fn)()]).to_pydict()
        assert out == {"a": [pyfn(v) if v is not None else None for v in a], "b": [pyfn(v) if v is not None else None for v in b]}

    def check_sign(v):
        return -1 if v < 0 else (1 if v > 0 else 0)
    ai = [None, -1, -5, 0, 5, 2, None]
    t = pa.table({"a": pa.array(ai, type=pa.int64()), "b": pa.array(b, type=pa.float64()), "u": pa.array([None, 0, 1, 2, 3, 4, 5], type=pa.uint8())})
    out = E.eval_expression_list(t, [col("a").sign(), col("b").sign(), col("u").sign(), col("a").ceil().alias("ac"), col("u").floor().alias("uf")])
    assert out.column("a").to_pylist() == [check_sign(v) if v is not None else None for v in ai] and out.schema.field("a").type == pa.int64()
    assert out.column("b").to_pylist() == [check_sign(v) if v is not None else None for v in b] and out.schema.field("b").type == pa.float64()
    assert out.column("u").to_pylist() == [None, 0, 1, 1, 1, 1, 1] and out.schema.field("u").type == pa.uint8()
    assert out.column("ac").to_pylist() == ai and out.column("uf").to_pylist() == [None, 0, 1, 2, 3, 4, 5]  # integers pass through
    # Rust signum: +-0.0 -> +-1.0, NaN -> NaN (array/ops/sign.rs:10-12)
    z = E.eval_expression_list(pa.table({"x": pa.array([0.0, -0.0, nan, -inf])}), [col("x").sign()]).column("x").to_pylist()
    assert z[:2] == [1.0, -1.0] and math.isnan(z[2]) and z[3] == -1.0
    s = pa.table({"a": pa.array(["a", "b", "c"], type=pa.large_utf8())})
    for bad in (col("a").ceil(), col("a").floor(), col("a").sign()):
        with pytest.raises(ValueError, match="to be numeric"):
            E.eval_expression_list(s, [bad])
    for typ in (pa.float32(), pa.float64()):
        t = pa.table({"input": pa.array([1.0, None, 3.0, nan], type=typ)})
        out = E.eval_expression_list(t, [col("input").fill_nan(2.0)])
        assert out.to_pydict() == {"input": [1.0, None, 3.0, 2.0]}
    assert E.eval_expression_list(pa.table({"n": pa.nulls(3)}), [col("n").is_nan()]).to_pydict() == {"n": [None, None, None]}


@pytest.mark.parametrize("data,typ", [(1, pa.int32()), ("a", pa.large_utf8()), (True, pa.bool_()), (b"Y", pa.large_binary()), (0.5, pa.float64()),
                                       (None, pa.null()), (2**40, pa.int64())])
def test_literal_broadcast_from_reference_tests(data, typ):
    """tests/recordbatch/test_broadcasts.py:10-14: a literal projects to one value per row, named "literal"; Python ints that fit
    are Int32 literals (src/daft-dsl/src/lit)."""
    t = pa.table({"x": [1, 2, 3]})
    out = E.eval_expression_list(t, [col("x"), lit(data)])
    assert out.to_pydict() == {"x": [1, 2, 3], "literal": [data] * 3} and out.schema.field("literal").type == typ


def test_comparison_literals_from_reference_series_tests():
    """tests/series/test_comparisons.py:150-181 (every int / float dtype pair), :240-274 (boolean arrays incl. Kleene and / or / xor),
    :476-500 (logical ops on a boolean and an integer are an error), :502-546 (dates)."""
    import datetime
    import itertools

    def six(t):
        return [E.eval_expression_list(t, [e.alias("r")]).column("r").to_pylist() for e in
                (col("l") < col("r"), col("l") <= col("r"), col("l") == col("r"), col("l") != col("r"), col("l") >= col("r"), col("l") > col("r"))]
    want = [[False, True, False, None, None, None], [True, True, False, None, None, None], [True, False, False, None, None, None],
            [False, True, True, None, None, None], [True, False, True, None, None, None], [False, False, True, None, None, None]]
    for lt, rt in itertools.product(_NUMERIC_TYPES, repeat=2):
        t = pa.table({"l": pa.array([1, 2, 3, None, 5, None], type=lt), "r": pa.array([1, 3, 1, 5, None, None], type=rt)})
        assert six(t) == want, (lt, rt)
    D = datetime.date
    t = pa.table({"l": pa.array([D(2023, 1, 1), D(2023, 1, 2), D(2023, 1, 3), None, D(2023, 1, 5), None], type=pa.date32()),
                  "r": pa.array([D(2023, 1, 1), D(2023, 1, 3), D(2023, 1, 1), D(2023, 1, 5), None, None], type=pa.date32())})
    assert six(t) == want
    t = pa.table({"l": pa.array([True, False, True, False, True, False, None, None, None]), "r": pa.array([True, True, False, False, None, None, True, False, None])})
    assert six(t) == [[False, True, False, False] + [None] * 5, [True, True, False, True] + [None] * 5, [True, False, False, True] + [None] * 5,
                      [False, True, True, False] + [None] * 5, [True, False, True, True] + [None] * 5, [False, False, True, False] + [None] * 5]
    ops = E.eval_expression_list(t, [(col("l") & col("r")).alias("and"), (col("l") | col("r")).alias("or"), (col("l") ^ col("r")).alias("xor")]).to_pydict()
    assert ops["and"] == [True, False, False, False, None, False, None, False, None]
    assert ops["or"] == [True, True, True, False, True, None, True, None, None]
    assert ops["xor"] == [False, True, True, False, None, None, None, None, None]
    t = pa.table({"l": pa.array([False, True, None, None]), "r": pa.array([1, 2, 3, 4], type=pa.int64())})
    for bad in (col("l") & col("r"), col("l") | col("r"), col("l") ^ col("r"), col("r") & col("l"), col("r") | col("l"), col("r") ^ col("l")):
        with pytest.raises(ValueError, match="logic"):
            E.eval_expression_list(t, [bad])


def test_arithmetic_matrix_from_reference_series_tests():
    """tests/series/test_arithmetic.py:30-62 (every numeric dtype pair: + - * / // %) and :332-345 (`%` of any two integer types is an
    integer: try_integer_widen_for_rem, infer_datatype.rs:662-745, not the numeric supertype)."""
    import itertools
    for lt, rt in itertools.product(_NUMERIC_TYPES, repeat=2):
        t = pa.table({"left": pa.array([1, 2, 3, None, 5, None], type=lt), "right": pa.array([1, 4, 1, 5, None, None], type=rt)})
        L, R = col("left"), col("right")
        out = E.eval_expression_list(t, [(L + R).alias("add"), (L - R).alias("sub"), (L * R).alias("mul"), (L / R).alias("div"),
                                         (L // R).alias("fdiv"), (L % R).alias("mod")])
        assert out.column("add").to_pylist() == [2, 6, 4, None, None, None], (lt, rt)
        if pa.types.is_signed_integer(lt) or pa.types.is_signed_integer(rt):
            assert out.column("sub").to_pylist() == [0, -2, 2, None, None, None], (lt, rt)
        assert out.column("mul").to_pylist() == [1, 8, 3, None, None, None]
        assert out.column("div").to_pylist() == [1.0, 0.5, 3.0, None, None, None] and out.schema.field("div").type == pa.float64()
        assert out.column("fdiv").to_pylist() == [1, 0, 3, None, None, None]
        assert out.column("mod").to_pylist() == [0, 2, 0, None, None, None]
        if pa.types.is_integer(lt) and pa.types.is_integer(rt):
            assert pa.types.is_integer(out.schema.field("mod").type), (lt, rt)
    # the widening itself: unsigned wins, and a negative operand of an unsigned `%` is null
    t = pa.table({"a": pa.array([-7, 7, None], type=pa.int8()), "b": pa.array([3, 3, 3], type=pa.uint8()), "c": pa.array([2, 2, 2], type=pa.uint64()),
                  "f": pa.array([2.0, 2.0, 2.0], type=pa.float32()), "g": pa.array([9, 9, 9], type=pa.int64())})
    out = E.eval_expression_list(t, [(col("a") % col("b")).alias("ab"), (col("g") % col("c")).alias("gc"), (col("g") % col("f")).alias("gf"), (col("a") % col("g")).alias("ag")])
    assert [out.schema.field(n).type for n in ("ab", "gc", "gf", "ag")] == [pa.uint8(), pa.uint64(), pa.float32(), pa.int64()]
    assert out.to_pydict() == {"ab": [None, 1, None], "gc": [1, 1, 1], "gf": [1.0, 1.0, 1.0], "ag": [-7, 7, None]}


def _rows(t: pa.Table):
    d = t.to_pydict()
    return sorted((tuple((k, d[k][i]) for k in d) for i in range(t.num_rows)), key=repr)


@pytest.mark.parametrize("null_safe", [True, False])
def test_multicolumn_join_literals_from_reference_tests(null_safe):
    """tests/recordbatch/test_joins.py:168-293 (multi-column keys: empty results, no cross product, cross product, all-null keys) and
    :307-358 (boolean / null key columns with a `right.y` payload name, anti joins)."""
    S = pa.large_utf8()
    ne = [null_safe] * 2 if null_safe else None
    for left in ({"a": [], "b": []}, {"a": ["apple", "banana"], "b": [3, 4]}):
        for right in ({"x": [], "y": []}, {"x": ["banana", "apple"], "y": [3, 4]}):
            lt = pa.table({"a": pa.array(left["a"], type=S), "b": pa.array(left["b"], type=pa.int32())})
            rt = pa.table({"x": pa.array(right["x"], type=S), "y": pa.array(right["y"], type=pa.int32())})
            assert O.hash_join(lt, rt, ["a", "b"], ["x", "y"], "inner", null_equals_null=ne).to_pydict() == {"a": [], "b": [], "x": [], "y": []}
    lt = pa.table({"a": pa.array(["apple", "apple", "banana", "banana", "carrot", None], type=S), "b": [1, 2, 2, 2, 3, 3], "c": [1, 2, 3, 4, 5, 5]})
    rt = pa.table({"x": pa.array(["banana", "carrot", "apple", "banana", "apple", "durian", None], type=S), "y": [1, 3, 2, 1, 3, 6, 3], "z": [1, 2, 3, 4, 5, 6, 6]})
    exp = [{"a": "apple", "b": 2, "c": 2, "x": "apple", "y": 2, "z": 3}, {"a": "carrot", "b": 3, "c": 5, "x": "carrot", "y": 3, "z": 2}]
    if null_safe:
        exp.append({"a": None, "b": 3, "c": 5, "x": None, "y": 3, "z": 6})
    assert _rows(O.hash_join(lt, rt, ["a", "b"], ["x", "y"], "inner", null_equals_null=ne)) == _rows(pa.Table.from_pylist(exp, schema=pa.schema(
        [("a", S), ("b", pa.int64()), ("c", pa.int64()), ("x", S), ("y", pa.int64()), ("z", pa.int64())])))
    lt = pa.table({"a": pa.array(["apple", "apple", "banana", "banana", "banana", None], type=S), "b": [1, 0, 1, 1, 1, 1], "c": [1, 2, 3, 4, 5, 5]})
    rt = pa.table({"x": pa.array(["apple", "apple", "banana", "banana", "banana", None], type=S), "y": [1, 0, 1, 1, 0, 1], "z": [1, 2, 3, 4, 5, 5]})
    got = O.hash_join(lt, rt, ["a", "b"], ["x", "y"], "inner", null_equals_null=ne)
    exp = [("apple", 1, 1, 1), ("apple", 0, 2, 2)] + [("banana", 1, c, z) for c in (3, 4, 5) for z in (3, 4)] + ([(None, 1, 5, 5)] if null_safe else [])
    assert sorted(zip(*[got.column(n).to_pylist() for n in ("a", "b", "c", "z")]), key=repr) == sorted(exp, key=repr)
    lt = pa.table({"a": pa.array([None, None], type=pa.int64()), "b": pa.array([None, None], type=S), "c": [1, 2]})
    rt = pa.table({"x": pa.array([None, None], type=pa.int64()), "y": pa.array([None, None], type=S), "z": [1, 2]})
    got = O.hash_join(lt, rt, ["a", "b"], ["x", "y"], "inner", null_equals_null=ne)
    assert sorted(zip(got.column("c").to_pylist(), got.column("z").to_pylist())) == ([(1, 1), (1, 2), (2, 1), (2, 2)] if null_safe else [])
    # boolean key, same key name on both sides, payload called "right.y"
    lt = pa.table({"x": pa.array([False, True, None]), "y": [0, 1, 2]})
    rt = pa.table({"x": pa.array([None, True, False, None]), "right.y": [0, 1, 2, 3]})
    got = O.hash_join(lt, rt, ["x"], ["x"], "inner", null_equals_null=[True] if null_safe else None)
    assert got.column_names == ["x", "y", "right.y"]
    got = O.sort(got, ["x", "right.y"], [False, False], [False, False])
    assert got.column("y").to_pylist() == ([0, 1, 2, 2] if null_safe else [0, 1])
    assert got.column("right.y").to_pylist() == ([2, 1, 0, 3] if null_safe else [2, 1])
    lt = pa.table({"x": pa.array([None, None, None], type=pa.int64()), "y": [0, 1, 2]})
    rt = pa.table({"x": pa.array([None, None, None, None], type=pa.int64()), "right.y": [0, 1, 2, 3]})
    got = O.hash_join(lt, rt, ["x"], ["x"], "inner")
    assert got.column_names == ["x", "y", "right.y"] and got.num_rows == 0
    # anti joins keep the left schema, whatever the right key is called
    lt = pa.table({"x": [1, 2, 3, 4], "y": [3, 4, 5, 6]})
    for rname in ("x", "z"):
        got = O.hash_join(lt, pa.table({rname: [2, 3, 5]}), ["x"], [rname], "anti")
        assert got.to_pydict() == {"x": [1, 4], "y": [3, 6]}


def test_single_column_boolean_and_string_sort_literals_from_reference_tests():
    """tests/recordbatch/test_sorting.py:49-125 (single key, every numeric / string dtype, ascending puts nulls last, descending is the exact
    reverse), :280-350 (boolean first key), :443-462 (two string keys), :465-468 (sorting by a literal keeps the order)."""
    S = pa.large_utf8()
    for typ in _NUMERIC_TYPES + [S]:
        vals = [None, 4, 2, 1, 5]
        a = pa.array([None if v is None else str(v) for v in vals], type=S) if typ == S else pa.array(vals, type=typ)
        assert O.argsort([a]).tolist() == [3, 2, 1, 4, 0], typ
        assert O.argsort([a], [True]).tolist() == [0, 4, 1, 2, 3], typ
    cases = [([None, True, False, True, False], [0, 1, 2, 3, None], False, False, [2, 4, 1, 3, 0]),
             ([None, True, False, True, False], [0, 1, 2, 3, None], True, False, [0, 1, 3, 2, 4]),
             ([True, True, True, True, True], [None, 3, 1, 2, 0], False, False, [4, 2, 3, 1, 0]),
             ([True, True, True, True, True], [None, 3, 1, 2, 0], True, False, [4, 2, 3, 1, 0]),
             ([None, None, None, None, None], [None, 3, 1, 2, 0], False, False, [4, 2, 3, 1, 0]),
             ([None, None, None, None, None], [None, 3, 1, 2, 0], True, False, [4, 2, 3, 1, 0])]
    for a, b, ad, bd, expected in cases:
        for typ in _NUMERIC_TYPES + [S]:
            bb = pa.array([None if v is None else str(v) for v in b], type=S) if typ == S else pa.array(b, type=typ)
            keys = [pa.array(a, type=pa.bool_()), bb]
            assert O.argsort(keys, [ad, bd]).tolist() == expected, (a, b, ad, typ)
            assert O.argsort(keys, [not ad, not bd]).tolist() == expected[::-1], (a, b, ad, typ)
    t = pa.table({"firstname": pa.array(["bob", "alice", "eve", None, None, "bob", "alice"], type=S), "lastname": pa.array(["a", "a", "a", "bond", None, None, "a"], type=S)})
    assert O.sort(t, ["firstname", "lastname"]).to_pydict() == {"firstname": ["alice", "alice", "bob", "bob", "eve", None, None],
                                                                  "lastname": ["a", "a", "a", None, "a", "bond", None]}
    lit_key = E.eval_expression_list(pa.table({"a": [1, 2, 3, 4, 5]}), [lit(1)]).column("literal")
    assert O.argsort([lit_key]).tolist() == [0, 1, 2, 3, 4]


def test_string_and_binary_minmax_literals_from_reference_tests():
    """tests/recordbatch/test_table_aggs.py:137-158 (strings) and :188-213 (binary): nulls are ignored, an empty or
    all-null input gives null (grouped_cmp_utf8, src/daft-core/src/array/ops/compare_agg.rs:99-134)."""
    from daft_b200 import col
    from helpers import oracle_agg
    cases = [([], None, None), ([None], None, None), ([None, None, None], None, None), (["abc"], "abc", "abc"), ([None, "abc"], "abc", "abc"),
             ([None, "aaa", None, "b", "c", "aaa", None], "aaa", "c")]
    for vals, lo, hi in cases:
        for typ, conv in ((pa.large_utf8(), lambda x: x), (pa.large_binary(), lambda x: None if x is None else x.encode())):
            tbl = pa.table({"input": pa.array([conv(v) for v in vals], type=typ)})
            got = oracle_agg(tbl, [col("input").min().alias("min"), col("input").max().alias("max")], []).to_pydict()
            assert got == {"min": [conv(lo)], "max": [conv(hi)]}, (vals, got)


def test_list_set_concat_literals_from_reference_tests():
    """The oracle's list / set / concat against the reference's literal expectations: grouped_list with null keys
    (tests/recordbatch/test_table_aggs.py:811-828), test_table_agg_groupby_cases[0] (:470-490, list + set next to sum / count), the
    empty input (:902-913) and string_agg with and without nulls (:1077-1130)."""
    groups = pa.array([None, 1, None, 1, 2, 2], type=pa.int64())
    vals = pa.array([None, 0, 1, 2, None, 4], type=pa.int64())
    got = oracle_agg(pa.table({"groups": groups, "input": vals}), [col("input").agg_list().alias("list")], [col("groups")])
    got = got.sort_by([("groups", "ascending")]).to_pydict()
    assert got == {"groups": [1, 2, None], "list": [[0, 2], [None, 4], [None, 1]]}
    assert oracle_agg(pa.table({"input": vals}), [col("input").agg_list().alias("list")], []).to_pydict() == {"list": [[None, 0, 1, 2, None, 4]]}

    values = [("Bob", None), ("Bob", None), ("Bob", 5), ("Bob", 5), (None, None), (None, 5), (None, None), (None, 2),
              ("Alice", None), ("Alice", None), ("Alice", None), ("Alice", None)]
    tbl = pa.table({"name": pa.array([v[0] for v in values], type=pa.large_utf8()), "cookies": pa.array([v[1] for v in values], type=pa.int64())})
    aggs = [col("cookies").sum().alias("sum"), col("name").count().alias("count"), col("cookies").agg_list().alias("list"), col("cookies").agg_set().alias("set")]
    got = oracle_agg(tbl, aggs, [col("name")]).sort_by([("name", "ascending")]).to_pydict()
    assert got["name"] == ["Alice", "Bob", None] and got["sum"] == [None, 10, 7] and got["count"] == [4, 4, 0]
    assert got["list"] == [[None] * 4, [None, None, 5, 5], [None, 5, None, 2]]
    assert [set(x) for x in got["set"]] == [set(), {5}, {2, 5}]

    empty = pa.table({"col_A": pa.array([], type=pa.int32()), "col_B": pa.array([], type=pa.int64())})
    assert oracle_agg(empty, [col("col_A").agg_list().alias("list")], [col("col_B")]).to_pydict() == {"col_B": [], "list": []}

    t = pa.table({"a": pa.array(["the", " quick", " brown", " fox"], type=pa.large_utf8()), "b": pa.array([1, 2, 1, 2], type=pa.int64())})
    assert oracle_agg(t, [col("a").agg_concat()], []).to_pydict() == {"a": ["the quick brown fox"]}
    assert oracle_agg(t, [col("a").agg_concat()], [col("b")]).to_pydict() == {"b": [1, 2], "a": ["the brown", " quick fox"]}
    t = pa.table({"a": pa.array(["the", " quick", None, " fox"], type=pa.large_utf8()), "b": pa.array([1, 2, 1, 2], type=pa.int64())})
    assert oracle_agg(t, [col("a").agg_concat()], []).to_pydict() == {"a": ["the quick fox"]}
    assert oracle_agg(t, [col("a").agg_concat()], [col("b")]).to_pydict() == {"b": [1, 2], "a": ["the", " quick fox"]}
    t = pa.table({"a": pa.array([None, None, None, None], type=pa.large_utf8()), "b": pa.array([1, 2, 1, 2], type=pa.int64())})
    assert oracle_agg(t, [col("a").agg_concat()], []).to_pydict() == {"a": [None]}
    # set keeps 0.0 and -0.0 apart and never merges NaNs (probe table: hash of the bytes, then Literal ==; series/mod.rs:85-100)
    f = pa.table({"x": pa.array([0.0, -0.0, float("nan"), 1.5, float("nan"), 1.5, None, 0.0])})
    s = oracle_agg(f, [col("x").agg_set().alias("s")], []).to_pydict()["s"][0]
    assert [repr(v) for v in s] == ["0.0", "-0.0", "nan", "1.5", "nan"]


def test_two_stage_list_and_concat_rewrite():
    """populate_aggregation_stages for List / Concat (src/daft-local-plan/src/agg.rs:286-301): list -> concat of the partial lists."""
    from daft_b200.aggstages import populate_aggregation_stages
    first, second, final = populate_aggregation_stages([col("v").agg_list().alias("l"), col("s").agg_concat().alias("c"), col("v").sum().alias("t")], [col("k")])
    assert [e.name() for e in first] == ["__s1_0", "__s1_1", "__s1_2"] and [repr(e) for e in final][0] == "col(k)"
    ops = lambda es: [e.args[0].op if e.kind == "alias" else e.op for e in es]
    assert ops(first) == ["list", "concat", "sum"] and ops(second) == ["concat", "concat", "sum"]
    with pytest.raises(NotImplementedError):
        populate_aggregation_stages([col("v").agg_set()], [])

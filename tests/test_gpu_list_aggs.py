"""list / set / concat aggregates (AggExpr::List / Set / Concat): grouped_list (src/daft-core/src/array/ops/list_agg.rs:30-57),
grouped_set (src/daft-core/src/series/ops/agg.rs:531-575), grouped_concat over Utf8 (src/daft-core/src/array/ops/concat_agg.rs:
138-161).  Literal cases are the reference's own (tests/recordbatch/test_table_aggs.py:763-830,902-913,470-490,1077-1130); the
randomised ones compare with the oracle's row-by-row restatement."""
import datetime

import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal, oracle_agg, random_table

pytestmark = pytest.mark.gpu


def RB(gpu, tbl):
    return gpu.RecordBatch.from_arrow_table(tbl)


TYPES = [pa.int8(), pa.int32(), pa.int64(), pa.uint16(), pa.float32(), pa.float64(), pa.bool_(), pa.date32(), pa.large_utf8()]


def _as_type(vals, typ):
    if typ == pa.date32():
        return pa.array([None if x is None else datetime.date(2020 + x, 1 + x, 1 + x) for x in vals], type=typ)
    if typ == pa.large_utf8():
        return pa.array([None if x is None else str(x) for x in vals], type=typ)
    if typ == pa.bool_():
        return pa.array([None if x is None else bool(x) for x in vals], type=typ)
    return pa.array(vals, type=typ)


@pytest.mark.parametrize("typ", TYPES, ids=str)
def test_global_list_aggs(gpu, typ):
    """test_table_aggs.py:763-771: the whole column, nulls included, as one list."""
    arr = _as_type([None, 0, 1, 2, None, 4], typ)
    got = RB(gpu, pa.table({"input": arr})).agg([gpu.col("input").agg_list().alias("list")], []).to_arrow_table()
    assert got.schema.field("list").type == pa.large_list(typ)
    assert got.to_pydict() == {"list": [arr.to_pylist()]}


@pytest.mark.parametrize("typ", TYPES, ids=str)
def test_grouped_list_aggs(gpu, typ):
    """test_table_aggs.py:811-828: groups [None, 1, None, 1, 2, 2] -> rows [[1, 3], [4, 5], [0, 2]] after sorting by the key."""
    groups = pa.array([None, 1, None, 1, 2, 2], type=pa.int64())
    arr = _as_type([None, 0, 1, 2, None, 4], typ)
    c = gpu.col
    res = RB(gpu, pa.table({"groups": groups, "input": arr})).agg([c("input").agg_list().alias("list")], [c("groups")])
    got = res.sort([c("groups")]).to_arrow_table()     # the sort gathers the list column
    vals = arr.to_pylist()
    assert got.schema.field("list").type == pa.large_list(typ)
    assert got.to_pydict() == {"groups": [1, 2, None], "list": [[vals[i] for i in g] for g in ([1, 3], [4, 5], [0, 2])]}


def test_list_aggs_empty(gpu):
    """test_table_aggs.py:902-913."""
    tbl = pa.table({"col_A": pa.array([], type=pa.int32()), "col_B": pa.array([], type=pa.int64())})
    c = gpu.col
    got = RB(gpu, tbl).agg([c("col_A").agg_list().alias("list"), c("col_A").agg_set().alias("set")], [c("col_B")]).to_arrow_table()
    assert got.schema.field("list").type == pa.large_list(pa.int32())
    assert got.to_pydict() == {"col_B": [], "list": [], "set": []}
    got = RB(gpu, tbl).agg([c("col_A").agg_list().alias("list"), c("col_A").agg_set().alias("set")], []).to_pydict()
    assert got == {"list": [[]], "set": [[]]}


def test_reference_groupby_case_with_list_and_set(gpu):
    """test_table_agg_groupby_cases[0] (test_table_aggs.py:440-490): sum, count, list, set grouped by a string column."""
    values = [("Bob", None), ("Bob", None), ("Bob", 5), ("Bob", 5), (None, None), (None, 5), (None, None), (None, 2),
              ("Alice", None), ("Alice", None), ("Alice", None), ("Alice", None)]
    tbl = pa.table({"name": pa.array([v[0] for v in values], type=pa.large_utf8()), "cookies": pa.array([v[1] for v in values], type=pa.int64())})
    c = gpu.col
    aggs = [c("cookies").sum().alias("sum"), c("name").count().alias("count"), c("cookies").agg_list().alias("list"), c("cookies").agg_set().alias("set")]
    got = RB(gpu, tbl).agg(aggs, [c("name")]).sort([c("name")]).to_pydict()
    assert got["name"] == ["Alice", "Bob", None]
    assert got["sum"] == [None, 10, 7] and got["count"] == [4, 4, 0]
    assert got["list"] == [[None] * 4, [None, None, 5, 5], [None, 5, None, 2]]
    assert got["set"] == [[], [5], [5, 2]]
    assert_tables_equal(RB(gpu, tbl).agg(aggs, [c("name")]).to_arrow_table(), oracle_agg(tbl, aggs, [c("name")]))


def test_agg_concat_on_strings(gpu):
    """test_table_aggs.py:1077-1130 (string_agg without a delimiter)."""
    c = gpu.col
    t = pa.table({"a": pa.array(["the", " quick", " brown", " fox"], type=pa.large_utf8()), "b": pa.array([1, 2, 1, 2], type=pa.int64())})
    assert RB(gpu, t).agg([c("a").agg_concat()], []).to_pydict() == {"a": ["the quick brown fox"]}
    assert RB(gpu, t).agg([c("a").agg_concat()], [c("b")]).to_pydict() == {"b": [1, 2], "a": ["the brown", " quick fox"]}
    t = pa.table({"a": pa.array(["the", " quick", None, " fox"], type=pa.large_utf8()), "b": pa.array([1, 2, 1, 2], type=pa.int64())})
    assert RB(gpu, t).agg([c("a").agg_concat()], []).to_pydict() == {"a": ["the quick fox"]}
    assert RB(gpu, t).agg([c("a").agg_concat()], [c("b")]).to_pydict() == {"b": [1, 2], "a": ["the", " quick fox"]}
    t = pa.table({"a": pa.array([None, None, None, None], type=pa.large_utf8()), "b": pa.array([1, 2, 1, 2], type=pa.int64())})
    assert RB(gpu, t).agg([c("a").agg_concat()], []).to_pydict() == {"a": [None]}
    assert RB(gpu, t).agg([c("a").agg_concat()], [c("b")]).to_pydict() == {"b": [1, 2], "a": [None, None]}
    t = pa.table({"a": pa.array([], type=pa.large_utf8())})
    assert RB(gpu, t).agg([c("a").agg_concat()], []).to_pydict() == {"a": [""]}
    with pytest.raises(Exception):
        RB(gpu, pa.table({"x": pa.array([1, 2])})).agg([c("x").agg_concat()], [])


@pytest.mark.parametrize("n", [1, 37, 5000, 150_000])
def test_list_set_concat_match_the_oracle(gpu, n):
    """Every key shape of the group-by matrix (low / high cardinality, null keys, string keys, two keys), a fused predicate, list
    aggregates of expressions, and ordinary aggregates in the same call."""
    rng = np.random.default_rng(n)
    tbl = random_table(rng, n, 0.2)
    tbl = tbl.set_column(tbl.schema.get_field_index("f64"), "f64", pa.array(np.round(rng.standard_normal(n), 1), mask=rng.random(n) < 0.2))  # duplicates for set
    c = gpu.col
    aggs = [c("i32").agg_list().alias("l_i32"), c("s").agg_list().alias("l_s"), c("u8").agg_set().alias("set_u8"), c("s").agg_set().alias("set_s"),
            c("f64").agg_set().alias("set_f"), c("b").agg_set().alias("set_b"), c("s").agg_concat().alias("cat"), (c("i16") * 2).agg_list().alias("l_e"),
            c("i64").sum().alias("tot"), c("d").agg_list().alias("l_d"), c("f32").max().alias("mx"), c("i32").count().alias("cnt")]
    for keys, pred in (([], None), (["k"], None), (["u8"], [c("i32") > 0]), (["s"], None), (["b", "i8"], None), (["i32"], None), ([], [c("i32") > 5000])):
        got = RB(gpu, tbl).agg(aggs, [c(k) for k in keys], predicate=pred).to_arrow_table()
        want = oracle_agg(tbl, aggs, [c(k) for k in keys], pred)
        assert_tables_equal(got, want)


def test_set_of_high_cardinality_groups(gpu):
    """Many groups with a handful of values each (the global-table build on both the group pass and the dedup pass)."""
    rng = np.random.default_rng(5)
    n = 400_000
    tbl = pa.table({"g": pa.array(rng.integers(0, 120_000, n).astype(np.int64)), "v": pa.array(rng.integers(0, 4, n).astype(np.int32), mask=rng.random(n) < 0.1)})
    c = gpu.col
    aggs = [c("v").agg_set().alias("set"), c("v").agg_list().alias("list"), c("v").count("all").alias("rows")]
    got = RB(gpu, tbl).agg(aggs, [c("g")]).to_arrow_table()
    assert_tables_equal(got, oracle_agg(tbl, aggs, [c("g")]))


def test_list_columns_are_refused_by_row_moving_operators(gpu):
    c = gpu.col
    res = RB(gpu, pa.table({"k": pa.array([1, 1, 2]), "v": pa.array([1, 2, 3])})).agg([c("v").agg_list().alias("l")], [c("k")])
    assert res.head(1).to_pydict() == {"k": [1], "l": [[1, 2]]}                    # slice = gather
    assert res.eval_expression_list([c("l"), c("k")]).to_pydict() == {"l": [[1, 2], [3]], "k": [1, 2]}    # a selection, no kernel
    with pytest.raises(Exception):
        res.eval_expression_list([c("l").is_null()])
    with pytest.raises(Exception):
        res.partition_by_hash([c("k")], 2)


@pytest.mark.parametrize("morsel", [900, 7000])
def test_list_and_concat_through_the_two_stage_sink(gpu, morsel):
    """AggExpr::List / Concat across morsels (src/daft-local-plan/src/agg.rs:286-301): a list per morsel, then the CONCAT of the
    partial lists (concat over a List column, src/daft-core/src/array/ops/concat_agg.rs:68-116) -- pushed in row order the result
    equals the single-batch aggregate."""
    from daft_b200.aggstages import GroupedAggregateSink
    rng = np.random.default_rng(morsel)
    tbl = random_table(rng, 20_000, 0.2)
    c = gpu.col
    aggs = [c("i32").agg_list().alias("l"), c("s").agg_list().alias("ls"), c("s").agg_concat().alias("cat"), c("i64").sum().alias("tot"),
            c("f64").mean().alias("m")]
    for gb, pred in (([c("k")], None), ([c("u8"), c("b")], [c("i32") > 0]), ([], None)):
        sink = GroupedAggregateSink(aggs, gb, pred)
        for o in range(0, tbl.num_rows, morsel):
            sink.sink(RB(gpu, tbl.slice(o, morsel)))
        got = sink.finalize().to_arrow_table()
        want = oracle_agg(tbl, aggs, gb, pred)
        assert_tables_equal(got, want, float_rtol=1e-9, sort_by=[g.name() for g in gb] or None)


def test_concat_of_list_columns(gpu):
    """grouped_concat over a List column: null lists are skipped, a group without a valid list is null, empty lists count as valid
    (concat_agg.rs:68-116 and its tests :170-230)."""
    c = gpu.col
    t = pa.table({"g": pa.array([1, 1, 2, 2, 3, 3, 3], type=pa.int64()), "v": pa.array([1, 2, 3, None, None, None, 7], type=pa.int64()),
                  "p": pa.array([0, 1, 0, 0, 0, 1, 1], type=pa.int64())})
    lists = RB(gpu, t).agg([c("v").agg_list().alias("l")], [c("g"), c("p")])       # (1,0):[1] (1,1):[2] (2,0):[3,None] (3,0):[None] (3,1):[None,7]
    got = lists.agg([c("l").agg_concat().alias("all")], [c("g")]).to_pydict()
    assert got == {"g": [1, 2, 3], "all": [[1, 2], [3, None], [None, None, 7]]}
    assert lists.agg([c("l").agg_concat().alias("all")], []).to_pydict() == {"all": [[1, 2, 3, None, None, None, 7]]}
    with pytest.raises(Exception):
        lists.agg([c("l").agg_list().alias("ll")], [c("g")])

"""populate_aggregation_stages (host mirror of src/daft-local-plan/src/agg.rs:13-416): the three derived plans, run on
the oracle over row-range pieces of a table, must reproduce the single-stage aggregate of the whole table."""
import numpy as np
import pyarrow as pa
import pytest

from daft_b200 import col
from daft_b200.aggstages import populate_aggregation_stages
from helpers import assert_tables_equal, random_table
from oracle.recordbatch import OracleRecordBatch


def two_stage(tbl, aggs, gb, pieces, predicate=None):
    first, second, final = populate_aggregation_stages(aggs, gb)
    n = tbl.num_rows
    cuts = [n * i // pieces for i in range(pieces + 1)]
    parts = [OracleRecordBatch(tbl.slice(a, b - a)).agg(first, gb, predicate).to_arrow_table() for a, b in zip(cuts, cuts[1:])]
    keys = [col(g.name()) for g in gb]
    return OracleRecordBatch(pa.concat_tables(parts)).agg(second, keys).eval_expression_list(final).to_arrow_table()


def test_stage_shapes_follow_the_reference():
    aggs = [col("a").sum().alias("s"), col("a").mean().alias("m"), col("a").count().alias("c"), col("b").min().alias("lo"),
            col("a").sum().alias("s_again")]
    first, second, final = populate_aggregation_stages(aggs, [col("k")])
    # Sum(a) appears once in each stage although three outputs need it (IndexSet, agg.rs:56-76); Mean -> Sum + Count(Valid)
    assert [repr(e.args[0]) for e in first] == ["agg:sum(col(a))", "agg:count(col(a))", "agg:min(col(b))"]
    assert [repr(e.args[0]) for e in second] == ["agg:sum(col(__s1_0))", "agg:sum(col(__s1_1))", "agg:min(col(__s1_2))"]
    assert [e.name() for e in final] == ["k", "s", "m", "c", "lo", "s_again"]


@pytest.mark.parametrize("pieces", [1, 3, 7])
def test_two_stage_equals_single_stage(pieces):
    rng = np.random.default_rng(11 + pieces)
    tbl = random_table(rng, 4000, 0.15)
    aggs = [col("i64").sum().alias("sm"), col("f64").mean().alias("m"), col("i32").count("all").alias("ca"), col("i32").count("null").alias("cn"),
            col("u16").min().alias("lo"), col("f64").max().alias("hi"), col("b").bool_and().alias("ball"), (col("f64") * 2).stddev().alias("sd"),
            col("i16").var(1).alias("v"), col("i8").mean().alias("m8")]
    for gb in ([col("k")], [col("s"), col("u8")], []):
        got = two_stage(tbl, aggs, gb, pieces)
        want = OracleRecordBatch(tbl).agg(aggs, gb).to_arrow_table()
        keys = [g.name() for g in gb]
        assert_tables_equal(got, want, float_rtol=1e-9, sort_by=keys or None)


def test_q1_through_the_generic_rewrite():
    from daft_b200 import tpch
    pred, gb, aggs = tpch.q1_exprs()
    li = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS)
    got = two_stage(li, aggs, gb, 5, pred)
    want = OracleRecordBatch(li).agg(aggs, gb, pred).to_arrow_table()
    assert_tables_equal(got, want, float_rtol=1e-12, sort_by=["L_RETURNFLAG", "L_LINESTATUS"])

"""Synthetic TPC-H tables (generated in HBM or on the host by the same counter-based generator) and the
benchmark queries expressed against the RecordBatch operator surface.

Query shapes follow the reference's DataFrame forms in benchmarking/tpch/answers.py (q1 :12-32,
q5 :124-143, q6 :146-155); the data generator is daft_b200/csrc/tpch_gen.cu.
"""
from __future__ import annotations

import ctypes
import datetime
import functools
from typing import Sequence

import pyarrow as pa

from . import _lib
from .expressions import col, lit
from .recordbatch import RecordBatch

TABLES = {"lineitem": 0, "orders": 1, "customer": 2, "supplier": 3, "nation": 4, "region": 5, "part": 6, "partsupp": 7}
DEFAULT_SEED = 20260921

Q1_COLUMNS = ["L_QUANTITY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_TAX", "L_RETURNFLAG", "L_LINESTATUS", "L_SHIPDATE"]
# bytes per row of the Arrow layout the reference holds for Q1's 7 columns:
# 2 x (8 B offset + 1 B) + i64 + 3 x f64 + date32
Q1_BYTES_PER_ROW = 54


def num_rows(table: str, sf: float) -> int:
    return int(_lib.lib().b200_tpch_num_rows(ctypes.c_int32(TABLES[table]), ctypes.c_double(sf)))


def _names(columns: Sequence[str] | None):
    if not columns:
        return None, 0
    arr = (ctypes.c_char_p * len(columns))(*[c.encode() for c in columns])
    return arr, len(columns)


def generate(table: str, sf: float, columns: Sequence[str] | None = None, row_begin: int = 0, row_end: int = -1,
             seed: int = DEFAULT_SEED) -> RecordBatch:
    """Rows [row_begin, row_end) of `table` generated straight into HBM."""
    _lib.ensure_init()
    arr, n = _names(columns)
    out = ctypes.c_void_p()
    _lib.call("b200_tpch_generate", ctypes.c_int32(TABLES[table]), ctypes.c_double(sf), ctypes.c_int64(row_begin),
              ctypes.c_int64(row_end), ctypes.c_uint64(seed), arr, ctypes.c_int32(n), ctypes.byref(out))
    return RecordBatch._wrap(out)


def generate_host(table: str, sf: float, columns: Sequence[str] | None = None, row_begin: int = 0, row_end: int = -1,
                  seed: int = DEFAULT_SEED) -> pa.Table:
    """Same rows, generated on the host cores into Arrow buffers (no GPU needed)."""
    arr, n = _names(columns)
    c_arr, c_sch = _lib.ArrowArray(), _lib.ArrowSchema()
    _lib.call("b200_tpch_generate_host", ctypes.c_int32(TABLES[table]), ctypes.c_double(sf), ctypes.c_int64(row_begin),
              ctypes.c_int64(row_end), ctypes.c_uint64(seed), arr, ctypes.c_int32(n), ctypes.byref(c_arr), ctypes.byref(c_sch))
    rb = pa.RecordBatch._import_from_c(ctypes.addressof(c_arr), ctypes.addressof(c_sch))
    return pa.Table.from_batches([rb])


# ------------------------------------------------------------------------------------------------
# Q1 (benchmarking/tpch/answers.py:12-32)
# ------------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def q1_exprs():
    disc_price = col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))
    charge = disc_price * (1 + col("L_TAX"))
    predicate = [col("L_SHIPDATE") <= datetime.date(1998, 9, 2)]
    group_by = [col("L_RETURNFLAG"), col("L_LINESTATUS")]
    aggs = [
        col("L_QUANTITY").sum().alias("sum_qty"),
        col("L_EXTENDEDPRICE").sum().alias("sum_base_price"),
        disc_price.sum().alias("sum_disc_price"),
        charge.sum().alias("sum_charge"),
        col("L_QUANTITY").mean().alias("avg_qty"),
        col("L_EXTENDEDPRICE").mean().alias("avg_price"),
        col("L_DISCOUNT").mean().alias("avg_disc"),
        col("L_QUANTITY").count().alias("count_order"),
    ]
    return predicate, group_by, aggs


@functools.lru_cache(maxsize=None)
def q1_partial_exprs():
    """First stage of the two-stage rewrite (src/daft-local-plan/src/agg.rs:107-166): Mean -> Sum + Count."""
    disc_price = col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))
    charge = disc_price * (1 + col("L_TAX"))
    predicate = [col("L_SHIPDATE") <= datetime.date(1998, 9, 2)]
    group_by = [col("L_RETURNFLAG"), col("L_LINESTATUS")]
    aggs = [
        col("L_QUANTITY").sum().alias("sum_qty"),
        col("L_EXTENDEDPRICE").sum().alias("sum_base_price"),
        disc_price.sum().alias("sum_disc_price"),
        charge.sum().alias("sum_charge"),
        col("L_DISCOUNT").sum().alias("sum_disc"),
        col("L_QUANTITY").count().alias("count_qty"),
        col("L_EXTENDEDPRICE").count().alias("count_price"),
        col("L_DISCOUNT").count().alias("count_disc"),
    ]
    return predicate, group_by, aggs


@functools.lru_cache(maxsize=None)
def q1_final_exprs():
    """Second stage + final projection: Sum of partial sums/counts, merge_mean = sum / count
    (src/daft-dsl/src/functions/agg/mod.rs:29-53)."""
    group_by = [col("L_RETURNFLAG"), col("L_LINESTATUS")]
    aggs = [col(n).sum().alias(n) for n in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "sum_disc",
                                            "count_qty", "count_price", "count_disc")]
    project = [
        col("L_RETURNFLAG"), col("L_LINESTATUS"), col("sum_qty"), col("sum_base_price"), col("sum_disc_price"),
        col("sum_charge"),
        (col("sum_qty") / col("count_qty")).alias("avg_qty"),
        (col("sum_base_price") / col("count_price")).alias("avg_price"),
        (col("sum_disc") / col("count_disc")).alias("avg_disc"),
        col("count_qty").alias("count_order"),
    ]
    return group_by, aggs, project


def q1(lineitem: RecordBatch) -> RecordBatch:
    """Filter -> Project -> HashAggregate fused into one kernel, then the 4-row sort."""
    predicate, group_by, aggs = q1_exprs()
    out = lineitem.agg(aggs, group_by, predicate=predicate)
    return out.sort([col("L_RETURNFLAG"), col("L_LINESTATUS")])


def q1_partial(lineitem: RecordBatch) -> RecordBatch:
    predicate, group_by, aggs = q1_partial_exprs()
    return lineitem.agg(aggs, group_by, predicate=predicate)


def q1_final(partials: RecordBatch) -> RecordBatch:
    group_by, aggs, project = q1_final_exprs()
    return partials.agg(aggs, group_by).eval_expression_list(project).sort([col("L_RETURNFLAG"), col("L_LINESTATUS")])


def q1_distributed(lineitem: RecordBatch, comm) -> RecordBatch:
    """Q1 over one scan partition per rank (distributed.Communicator.agg): the per-GPU group tables are all-reduced over
    NVLink and every rank finishes with the global answer (src/daft-distributed/src/pipeline_node/aggregate.rs:281-340)."""
    predicate, group_by, aggs = q1_exprs()
    return comm.agg(lineitem, aggs, group_by, predicate).sort([col("L_RETURNFLAG"), col("L_LINESTATUS")])


# ------------------------------------------------------------------------------------------------
# Q6 (answers.py:146-155): pure scan + ungrouped aggregate
# ------------------------------------------------------------------------------------------------
def q6(lineitem: RecordBatch) -> RecordBatch:
    pred = [
        col("L_SHIPDATE") >= datetime.date(1994, 1, 1), col("L_SHIPDATE") < datetime.date(1995, 1, 1),
        col("L_DISCOUNT") >= 0.05, col("L_DISCOUNT") <= 0.07, col("L_QUANTITY") < 24,
    ]
    return lineitem.agg([(col("L_EXTENDEDPRICE") * col("L_DISCOUNT")).sum().alias("revenue")], [], predicate=pred)


# ------------------------------------------------------------------------------------------------
# Q5 (answers.py:124-143): region -> nation -> supplier/customer -> orders -> lineitem, 2-key last join
# ------------------------------------------------------------------------------------------------
Q5_COLUMNS = {
    "lineitem": ["L_ORDERKEY", "L_SUPPKEY", "L_EXTENDEDPRICE", "L_DISCOUNT"],
    "orders": ["O_ORDERKEY", "O_CUSTKEY", "O_ORDERDATE"],
    "customer": ["C_CUSTKEY", "C_NATIONKEY"],
    "supplier": ["S_SUPPKEY", "S_NATIONKEY"],
    "nation": ["N_NATIONKEY", "N_REGIONKEY", "N_NAME"],
    "region": ["R_REGIONKEY", "R_NAME"],
}


def q5(t: dict[str, RecordBatch]) -> RecordBatch:
    """The join order of answers.py:124-143; every join builds on its smaller side, as the reference planner decides
    (src/daft-local-execution/src/pipeline.rs:1106-1194).  Join outputs are views (selection vectors), so the chain
    moves row ids only: L_ORDERKEY is read through the first join's selection vector by the second probe, and
    L_EXTENDEDPRICE / L_DISCOUNT are gathered for the rows that survive all five joins."""
    orders = t["orders"].filter([col("O_ORDERDATE") >= datetime.date(1994, 1, 1), col("O_ORDERDATE") < datetime.date(1995, 1, 1)])
    region = t["region"].filter([col("R_NAME") == "ASIA"])
    j = region.hash_join(t["nation"], [col("R_REGIONKEY")], [col("N_REGIONKEY")], build_on_left=True)
    j = j.hash_join(t["supplier"], [col("N_NATIONKEY")], [col("S_NATIONKEY")], build_on_left=True)
    j = j.hash_join(t["lineitem"], [col("S_SUPPKEY")], [col("L_SUPPKEY")], build_on_left=True)
    j = j.eval_expression_list([col("N_NAME"), col("L_EXTENDEDPRICE"), col("L_DISCOUNT"), col("L_ORDERKEY"), col("N_NATIONKEY")])
    j = j.hash_join(orders, [col("L_ORDERKEY")], [col("O_ORDERKEY")], build_on_left=len(j) < len(orders))
    j = j.hash_join(t["customer"], [col("O_CUSTKEY"), col("N_NATIONKEY")], [col("C_CUSTKEY"), col("C_NATIONKEY")],
                    build_on_left=len(j) < len(t["customer"]))
    out = j.agg([(col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))).sum().alias("revenue")], [col("N_NAME")])
    return out.sort([col("revenue")], descending=True)


# ------------------------------------------------------------------------------------------------
# More of benchmarking/tpch/answers.py, composed from the same operators.  Each follows the reference's DataFrame form
# (join order, projections, aggregation lists, final sort / limit); joins build on their smaller side as the planner would.
# ------------------------------------------------------------------------------------------------
def _when(pred: "Expression", then, otherwise):
    """daft.functions.when(pred, then).otherwise(x)"""
    from .expressions import _to_expr
    return pred.if_else(_to_expr(then), _to_expr(otherwise))


QUERY_COLUMNS = {
    "q3": {"customer": ["C_CUSTKEY", "C_MKTSEGMENT"], "orders": ["O_ORDERKEY", "O_CUSTKEY", "O_ORDERDATE", "O_SHIPPRIORITY"],
           "lineitem": ["L_ORDERKEY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_SHIPDATE"]},
    "q4": {"orders": ["O_ORDERKEY", "O_ORDERDATE", "O_ORDERPRIORITY"], "lineitem": ["L_ORDERKEY", "L_COMMITDATE", "L_RECEIPTDATE"]},
    "q12": {"orders": ["O_ORDERKEY", "O_ORDERPRIORITY"], "lineitem": ["L_ORDERKEY", "L_SHIPMODE", "L_COMMITDATE", "L_RECEIPTDATE", "L_SHIPDATE"]},
    "q14": {"lineitem": ["L_PARTKEY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_SHIPDATE"], "part": ["P_PARTKEY", "P_TYPE"]},
    "q18": {"customer": ["C_CUSTKEY"], "orders": ["O_ORDERKEY", "O_CUSTKEY", "O_ORDERDATE", "O_TOTALPRICE"], "lineitem": ["L_ORDERKEY", "L_QUANTITY"]},
}


def q3(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:78-103: BUILDING customers x orders before 1995-03-15 x lineitems shipped after it; revenue per order, top 10."""
    customer = t["customer"].filter([col("C_MKTSEGMENT") == "BUILDING"])
    orders = t["orders"].filter([col("O_ORDERDATE") < datetime.date(1995, 3, 15)])
    lineitem = t["lineitem"].filter([col("L_SHIPDATE") > datetime.date(1995, 3, 15)])
    j = customer.hash_join(orders, [col("C_CUSTKEY")], [col("O_CUSTKEY")], build_on_left=len(customer) <= len(orders))
    j = j.eval_expression_list([col("O_ORDERKEY"), col("O_ORDERDATE"), col("O_SHIPPRIORITY")])
    j = j.hash_join(lineitem, [col("O_ORDERKEY")], [col("L_ORDERKEY")], build_on_left=len(j) <= len(lineitem))
    j = j.eval_expression_list([col("O_ORDERKEY"), (col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))).alias("volume"), col("O_ORDERDATE"),
                                col("O_SHIPPRIORITY")])
    g = j.agg([col("volume").sum().alias("revenue")], [col("O_ORDERKEY"), col("O_ORDERDATE"), col("O_SHIPPRIORITY")])
    top = g.top_n([col("revenue"), col("O_ORDERDATE")], 10, descending=[True, False])
    return top.eval_expression_list([col("O_ORDERKEY"), col("revenue"), col("O_ORDERDATE"), col("O_SHIPPRIORITY")])


def q4(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:106-124: orders of 1993 Q3 with at least one late lineitem (semi join), counted per priority."""
    orders = t["orders"].filter([col("O_ORDERDATE") >= datetime.date(1993, 7, 1), col("O_ORDERDATE") < datetime.date(1993, 10, 1)])
    lineitems = t["lineitem"].filter([col("L_COMMITDATE") < col("L_RECEIPTDATE")])
    j = orders.hash_join(lineitems, [col("O_ORDERKEY")], [col("L_ORDERKEY")], how="semi")
    return j.agg([col("O_ORDERKEY").count().alias("order_count")], [col("O_ORDERPRIORITY")]).sort([col("O_ORDERPRIORITY")])


def q12(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:373-395: late MAIL / SHIP lineitems received in 1994, high- and low-priority line counts per ship mode.  The
    DataFrame text filters after the join; every conjunct reads lineitem only, so the reference's optimizer moves the filter below
    the join (PushDownFilter, src/daft-logical-plan/src/optimization/rules/push_down_filter.rs) -- the plan it executes is this one."""
    lineitem = t["lineitem"].filter([col("L_SHIPMODE").is_in(["MAIL", "SHIP"]), col("L_COMMITDATE") < col("L_RECEIPTDATE"),
                                     col("L_SHIPDATE") < col("L_COMMITDATE"), col("L_RECEIPTDATE") >= datetime.date(1994, 1, 1),
                                     col("L_RECEIPTDATE") < datetime.date(1995, 1, 1)])
    j = t["orders"].hash_join(lineitem, [col("O_ORDERKEY")], [col("L_ORDERKEY")], build_on_left=len(t["orders"]) <= len(lineitem))
    urgent = col("O_ORDERPRIORITY").is_in(["1-URGENT", "2-HIGH"])
    out = j.agg([_when(urgent, 1, 0).sum().alias("high_line_count"), _when(~urgent, 1, 0).sum().alias("low_line_count")], [col("L_SHIPMODE")])
    return out.sort([col("L_SHIPMODE")])


def q14(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:418-436: promotional revenue share of September 1995 (the ship-date filter below the join, as PushDownFilter puts it)."""
    lineitem = t["lineitem"].filter([col("L_SHIPDATE") >= datetime.date(1995, 9, 1), col("L_SHIPDATE") < datetime.date(1995, 10, 1)])
    j = lineitem.hash_join(t["part"], [col("L_PARTKEY")], [col("P_PARTKEY")], build_on_left=len(lineitem) <= len(t["part"]))
    rev = col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))
    a = j.agg([_when(col("P_TYPE").startswith("PROMO"), rev, 0.0).sum().alias("tmp_1"), rev.sum().alias("tmp_2")], [])
    return a.eval_expression_list([(lit(100.00) * (col("tmp_1") / col("tmp_2"))).alias("promo_revenue")])


def q18(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:510-528 (the generator has no C_NAME: the group key list is C_CUSTKEY, O_ORDERKEY, O_ORDERDATE, O_TOTALPRICE):
    orders whose lineitems sum to more than 300 units -- a high-cardinality group-by over all of lineitem -- top 100 by price."""
    res_1 = t["lineitem"].agg([col("L_QUANTITY").sum().alias("sum_qty")], [col("L_ORDERKEY")]).filter([col("sum_qty") > 300])
    j = t["orders"].hash_join(res_1, [col("O_ORDERKEY")], [col("L_ORDERKEY")], build_on_left=len(t["orders"]) <= len(res_1))
    j = j.hash_join(t["customer"], [col("O_CUSTKEY")], [col("C_CUSTKEY")], build_on_left=len(j) <= len(t["customer"]))
    j = j.eval_expression_list([col("C_CUSTKEY"), col("O_ORDERKEY"), col("O_ORDERDATE"), col("O_TOTALPRICE")])
    j = j.hash_join(t["lineitem"], [col("O_ORDERKEY")], [col("L_ORDERKEY")], build_on_left=True)
    g = j.agg([col("L_QUANTITY").sum().alias("sum")], [col("C_CUSTKEY"), col("O_ORDERKEY"), col("O_ORDERDATE"), col("O_TOTALPRICE")])
    g = g.eval_expression_list([col("C_CUSTKEY"), col("O_ORDERKEY"), col("O_ORDERDATE").alias("O_ORDERDAT"), col("O_TOTALPRICE"), col("sum")])
    return g.top_n([col("O_TOTALPRICE"), col("O_ORDERDAT")], 100, descending=[True, False])


# ------------------------------------------------------------------------------------------------
# The rest of answers.py (Q2, Q7-Q11, Q13, Q15-Q17, Q19-Q22).  DataFrame verbs map onto the RecordBatch surface:
# where -> filter, select / with_column -> eval_expression_list, distinct -> dedup, sort + limit -> top_n,
# join(on=lit) (the reference's way of writing a cross join with a one-row table) -> hash_join on the literal column.
# ------------------------------------------------------------------------------------------------
def _cols(*names: str):
    return [col(n) for n in names]


def _join(left: RecordBatch, right: RecordBatch, left_on: Sequence[str], right_on: Sequence[str], how: str = "inner") -> RecordBatch:
    """Inner joins build on the smaller side, as the planner decides (pipeline.rs:1106-1194)."""
    build_on_left = (len(left) <= len(right)) if how == "inner" else None
    return left.hash_join(right, _cols(*left_on), _cols(*right_on), how=how, build_on_left=build_on_left)


def _with_column(rb: RecordBatch, name: str, e) -> RecordBatch:
    return rb.eval_expression_list(_cols(*rb.column_names()) + [e.alias(name)])


def _volume():
    return col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))


QUERY_COLUMNS.update({
    "q2": {"region": ["R_REGIONKEY", "R_NAME"], "nation": ["N_NATIONKEY", "N_REGIONKEY", "N_NAME"],
           "supplier": ["S_SUPPKEY", "S_NATIONKEY", "S_ACCTBAL", "S_NAME", "S_ADDRESS", "S_PHONE", "S_COMMENT"],
           "partsupp": ["PS_PARTKEY", "PS_SUPPKEY", "PS_SUPPLYCOST"], "part": ["P_PARTKEY", "P_SIZE", "P_TYPE", "P_MFGR"]},
    "q7": {"nation": ["N_NATIONKEY", "N_NAME"], "supplier": ["S_SUPPKEY", "S_NATIONKEY"], "customer": ["C_CUSTKEY", "C_NATIONKEY"],
           "orders": ["O_ORDERKEY", "O_CUSTKEY"], "lineitem": ["L_ORDERKEY", "L_SUPPKEY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_SHIPDATE"]},
    "q8": {"region": ["R_REGIONKEY", "R_NAME"], "nation": ["N_NATIONKEY", "N_REGIONKEY", "N_NAME"], "supplier": ["S_SUPPKEY", "S_NATIONKEY"],
           "customer": ["C_CUSTKEY", "C_NATIONKEY"], "orders": ["O_ORDERKEY", "O_CUSTKEY", "O_ORDERDATE"], "part": ["P_PARTKEY", "P_TYPE"],
           "lineitem": ["L_ORDERKEY", "L_PARTKEY", "L_SUPPKEY", "L_EXTENDEDPRICE", "L_DISCOUNT"]},
    "q9": {"nation": ["N_NATIONKEY", "N_NAME"], "supplier": ["S_SUPPKEY", "S_NATIONKEY"], "part": ["P_PARTKEY", "P_NAME"],
           "partsupp": ["PS_PARTKEY", "PS_SUPPKEY", "PS_SUPPLYCOST"], "orders": ["O_ORDERKEY", "O_ORDERDATE"],
           "lineitem": ["L_ORDERKEY", "L_PARTKEY", "L_SUPPKEY", "L_QUANTITY", "L_EXTENDEDPRICE", "L_DISCOUNT"]},
    "q10": {"nation": ["N_NATIONKEY", "N_NAME"], "customer": ["C_CUSTKEY", "C_NATIONKEY", "C_NAME", "C_ACCTBAL", "C_ADDRESS", "C_PHONE", "C_COMMENT"],
            "orders": ["O_ORDERKEY", "O_CUSTKEY", "O_ORDERDATE"], "lineitem": ["L_ORDERKEY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_RETURNFLAG"]},
    "q11": {"nation": ["N_NATIONKEY", "N_NAME"], "supplier": ["S_SUPPKEY", "S_NATIONKEY"],
            "partsupp": ["PS_PARTKEY", "PS_SUPPKEY", "PS_AVAILQTY", "PS_SUPPLYCOST"]},
    "q13": {"customer": ["C_CUSTKEY"], "orders": ["O_ORDERKEY", "O_CUSTKEY", "O_COMMENT"]},
    "q15": {"supplier": ["S_SUPPKEY", "S_NAME", "S_ADDRESS", "S_PHONE"], "lineitem": ["L_SUPPKEY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_SHIPDATE"]},
    "q16": {"part": ["P_PARTKEY", "P_BRAND", "P_TYPE", "P_SIZE"], "partsupp": ["PS_PARTKEY", "PS_SUPPKEY"], "supplier": ["S_SUPPKEY", "S_COMMENT"]},
    "q17": {"part": ["P_PARTKEY", "P_BRAND", "P_CONTAINER"], "lineitem": ["L_PARTKEY", "L_QUANTITY", "L_EXTENDEDPRICE"]},
    "q19": {"part": ["P_PARTKEY", "P_BRAND", "P_CONTAINER", "P_SIZE"],
            "lineitem": ["L_PARTKEY", "L_QUANTITY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_SHIPMODE", "L_SHIPINSTRUCT"]},
    "q20": {"nation": ["N_NATIONKEY", "N_NAME"], "supplier": ["S_SUPPKEY", "S_NATIONKEY", "S_NAME", "S_ADDRESS"], "part": ["P_PARTKEY", "P_NAME"],
            "partsupp": ["PS_PARTKEY", "PS_SUPPKEY", "PS_AVAILQTY"], "lineitem": ["L_PARTKEY", "L_SUPPKEY", "L_QUANTITY", "L_SHIPDATE"]},
    "q21": {"nation": ["N_NATIONKEY", "N_NAME"], "supplier": ["S_SUPPKEY", "S_NATIONKEY", "S_NAME"], "orders": ["O_ORDERKEY", "O_ORDERSTATUS"],
            "lineitem": ["L_ORDERKEY", "L_SUPPKEY", "L_COMMITDATE", "L_RECEIPTDATE"]},
    "q22": {"customer": ["C_CUSTKEY", "C_PHONE", "C_ACCTBAL"], "orders": ["O_CUSTKEY"]},
})


def q2(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:35-75: for size-15 BRASS parts, the European supplier(s) with the minimum supply cost; top 100."""
    europe = t["region"].filter([col("R_NAME") == "EUROPE"])
    europe = _join(europe, t["nation"], ["R_REGIONKEY"], ["N_REGIONKEY"])
    europe = _join(europe, t["supplier"], ["N_NATIONKEY"], ["S_NATIONKEY"])
    europe = _join(europe, t["partsupp"], ["S_SUPPKEY"], ["PS_SUPPKEY"])
    brass = t["part"].filter([(col("P_SIZE") == 15) & col("P_TYPE").endswith("BRASS")])
    brass = _join(brass, europe, ["P_PARTKEY"], ["PS_PARTKEY"])
    min_cost = brass.agg([col("PS_SUPPLYCOST").min().alias("min")], _cols("P_PARTKEY"))
    j = _join(brass, min_cost, ["P_PARTKEY"], ["P_PARTKEY"]).filter([col("PS_SUPPLYCOST") == col("min")])
    j = j.eval_expression_list(_cols("S_ACCTBAL", "S_NAME", "N_NAME", "P_PARTKEY", "P_MFGR", "S_ADDRESS", "S_PHONE", "S_COMMENT"))
    return j.top_n(_cols("S_ACCTBAL", "N_NAME", "S_NAME", "P_PARTKEY"), 100, descending=[True, False, False, False])


def q7(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:158-203: revenue shipped between FRANCE and GERMANY (either direction) per year, 1995-1996."""
    lineitem = t["lineitem"].filter([(col("L_SHIPDATE") >= datetime.date(1995, 1, 1)) & (col("L_SHIPDATE") <= datetime.date(1996, 12, 31))])
    nation = t["nation"].filter([(col("N_NAME") == "FRANCE") | (col("N_NAME") == "GERMANY")])
    sup = _join(nation, t["supplier"], ["N_NATIONKEY"], ["S_NATIONKEY"])
    sup = _join(sup, lineitem, ["S_SUPPKEY"], ["L_SUPPKEY"])
    sup = sup.eval_expression_list([col("N_NAME").alias("supp_nation")] + _cols("L_ORDERKEY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_SHIPDATE"))
    cust = _join(nation, t["customer"], ["N_NATIONKEY"], ["C_NATIONKEY"])
    cust = _join(cust, t["orders"], ["C_CUSTKEY"], ["O_CUSTKEY"])
    cust = cust.eval_expression_list([col("N_NAME").alias("cust_nation"), col("O_ORDERKEY")])
    j = _join(cust, sup, ["O_ORDERKEY"], ["L_ORDERKEY"])
    j = j.filter([((col("supp_nation") == "FRANCE") & (col("cust_nation") == "GERMANY"))
                  | ((col("supp_nation") == "GERMANY") & (col("cust_nation") == "FRANCE"))])
    j = j.eval_expression_list([col("supp_nation"), col("cust_nation"), col("L_SHIPDATE").year().alias("l_year"), _volume().alias("volume")])
    g = j.agg([col("volume").sum().alias("revenue")], _cols("supp_nation", "cust_nation", "l_year"))
    return g.sort(_cols("supp_nation", "cust_nation", "l_year"))


def q8(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:206-251: BRAZIL's share of the AMERICA market for ECONOMY ANODIZED STEEL, 1995 and 1996."""
    region = t["region"].filter([col("R_NAME") == "AMERICA"])
    orders = t["orders"].filter([(col("O_ORDERDATE") <= datetime.date(1996, 12, 31)) & (col("O_ORDERDATE") >= datetime.date(1995, 1, 1))])
    part = t["part"].filter([col("P_TYPE") == "ECONOMY ANODIZED STEEL"])
    nat = _join(t["nation"], t["supplier"], ["N_NATIONKEY"], ["S_NATIONKEY"])
    line = t["lineitem"].eval_expression_list(_cols("L_PARTKEY", "L_SUPPKEY", "L_ORDERKEY") + [_volume().alias("volume")])
    line = _join(line, part, ["L_PARTKEY"], ["P_PARTKEY"])
    line = _join(line, nat, ["L_SUPPKEY"], ["S_SUPPKEY"])
    j = _join(t["nation"], region, ["N_REGIONKEY"], ["R_REGIONKEY"]).eval_expression_list(_cols("N_NATIONKEY"))
    j = _join(j, t["customer"], ["N_NATIONKEY"], ["C_NATIONKEY"]).eval_expression_list(_cols("C_CUSTKEY"))
    j = _join(j, orders, ["C_CUSTKEY"], ["O_CUSTKEY"]).eval_expression_list(_cols("O_ORDERKEY", "O_ORDERDATE"))
    j = _join(j, line, ["O_ORDERKEY"], ["L_ORDERKEY"])
    j = j.eval_expression_list([col("O_ORDERDATE").year().alias("o_year"), col("volume"),
                                _when(col("N_NAME") == "BRAZIL", col("volume"), 0.0).alias("case_volume")])
    g = j.agg([col("case_volume").sum().alias("case_volume_sum"), col("volume").sum().alias("volume_sum")], _cols("o_year"))
    return g.eval_expression_list([col("o_year"), (col("case_volume_sum") / col("volume_sum")).alias("mkt_share")]).sort(_cols("o_year"))


def q9(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:254-284: profit on green parts per supplier nation and year (the two-key partsupp join)."""
    linepart = _join(t["part"].filter([col("P_NAME").contains("green")]), t["lineitem"], ["P_PARTKEY"], ["L_PARTKEY"])
    natsup = _join(t["nation"], t["supplier"], ["N_NATIONKEY"], ["S_NATIONKEY"])
    j = _join(linepart, natsup, ["L_SUPPKEY"], ["S_SUPPKEY"])
    j = _join(j, t["partsupp"], ["L_SUPPKEY", "P_PARTKEY"], ["PS_SUPPKEY", "PS_PARTKEY"])
    j = _join(j, t["orders"], ["L_ORDERKEY"], ["O_ORDERKEY"])
    amount = _volume() - col("PS_SUPPLYCOST") * col("L_QUANTITY")
    j = j.eval_expression_list([col("N_NAME"), col("O_ORDERDATE").year().alias("o_year"), amount.alias("amount")])
    return j.agg([col("amount").sum().alias("amount")], _cols("N_NAME", "o_year")).sort(_cols("N_NAME", "o_year"), descending=[False, True])


def q10(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:287-335: customers who returned items ordered in 1993 Q4, by lost revenue; top 20."""
    lineitem = t["lineitem"].filter([col("L_RETURNFLAG") == "R"])
    orders = t["orders"].filter([(col("O_ORDERDATE") < datetime.date(1994, 1, 1)) & (col("O_ORDERDATE") >= datetime.date(1993, 10, 1))])
    j = _join(orders, t["customer"], ["O_CUSTKEY"], ["C_CUSTKEY"])
    j = _join(j, t["nation"], ["C_NATIONKEY"], ["N_NATIONKEY"])
    j = _join(j, lineitem, ["O_ORDERKEY"], ["L_ORDERKEY"])
    j = j.eval_expression_list([col("O_CUSTKEY"), col("C_NAME"), _volume().alias("volume")] + _cols("C_ACCTBAL", "N_NAME", "C_ADDRESS", "C_PHONE", "C_COMMENT"))
    g = j.agg([col("volume").sum().alias("revenue")], _cols("O_CUSTKEY", "C_NAME", "C_ACCTBAL", "C_PHONE", "N_NAME", "C_ADDRESS", "C_COMMENT"))
    top = g.top_n(_cols("revenue"), 20, descending=[True])
    return top.eval_expression_list(_cols("O_CUSTKEY", "C_NAME", "revenue", "C_ACCTBAL", "N_NAME", "C_ADDRESS", "C_PHONE", "C_COMMENT"))


def q11(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:338-370: GERMANY's stock value per part above 0.01 % of the total (the scalar subquery is a one-row join)."""
    res_1 = _join(t["partsupp"], t["supplier"], ["PS_SUPPKEY"], ["S_SUPPKEY"])
    res_1 = _join(res_1, t["nation"], ["S_NATIONKEY"], ["N_NATIONKEY"]).filter([col("N_NAME") == "GERMANY"])
    value = col("PS_SUPPLYCOST") * col("PS_AVAILQTY")
    res_2 = res_1.agg([value.sum().alias("tmp")], []).eval_expression_list([(col("tmp") * (0.0001 / 1)).alias("tmp"), lit(1).alias("lit")])
    g = _with_column(res_1.agg([value.sum().alias("value")], _cols("PS_PARTKEY")), "lit", lit(1))
    j = _join(g, res_2, ["lit"], ["lit"]).filter([col("value") > col("tmp")])
    return j.eval_expression_list([col("PS_PARTKEY"), col("value").round(2)]).sort(_cols("value"), descending=True)


def q13(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:396-415: distribution of customers by their number of orders that are not special requests (left join)."""
    orders = t["orders"].filter([~col("O_COMMENT").match(".*special.*requests.*")])
    j = t["customer"].hash_join(orders, _cols("C_CUSTKEY"), _cols("O_CUSTKEY"), how="left")
    g = j.agg([col("O_ORDERKEY").count().alias("c_count")], _cols("C_CUSTKEY")).sort(_cols("C_CUSTKEY"))
    g = g.agg([col("c_count").count().alias("custdist")], _cols("c_count"))
    return g.sort(_cols("custdist", "c_count"), descending=[True, True])


def q15(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:437-459: the supplier(s) with the top revenue of 1996 Q1 (a join on the maximum: a Float64 key)."""
    li = t["lineitem"].filter([(col("L_SHIPDATE") >= datetime.date(1996, 1, 1)) & (col("L_SHIPDATE") < datetime.date(1996, 4, 1))])
    revenue = li.agg([_volume().sum().alias("total_revenue")], _cols("L_SUPPKEY"))
    revenue = revenue.eval_expression_list([col("L_SUPPKEY").alias("supplier_no"), col("total_revenue")])
    top = revenue.agg([col("total_revenue").max().alias("total_revenue")], [])
    revenue = _join(revenue, top, ["total_revenue"], ["total_revenue"])
    j = _join(t["supplier"], revenue, ["S_SUPPKEY"], ["supplier_no"])
    return j.eval_expression_list(_cols("S_SUPPKEY", "S_NAME", "S_ADDRESS", "S_PHONE", "total_revenue")).sort(_cols("S_SUPPKEY"))


def q16(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:462-489: suppliers per (brand, type, size) excluding the ones with complaints (left join + is_null = anti join)."""
    suppkeys = t["supplier"].filter([col("S_COMMENT").match(".*Customer.*Complaints.*")])
    suppkeys = suppkeys.eval_expression_list([col("S_SUPPKEY"), col("S_SUPPKEY").alias("PS_SUPPKEY_RIGHT")])
    j = _join(t["part"], t["partsupp"], ["P_PARTKEY"], ["PS_PARTKEY"])
    j = j.filter([(col("P_BRAND") != "Brand#45") & ~col("P_TYPE").startswith("MEDIUM POLISHED")
                  & col("P_SIZE").is_in([49, 14, 23, 45, 19, 3, 36, 9])])
    j = j.hash_join(suppkeys, _cols("PS_SUPPKEY"), _cols("S_SUPPKEY"), how="left").filter([col("PS_SUPPKEY_RIGHT").is_null()])
    j = j.eval_expression_list(_cols("P_BRAND", "P_TYPE", "P_SIZE", "PS_SUPPKEY"))
    j = j.dedup(_cols("P_BRAND", "P_TYPE", "P_SIZE", "PS_SUPPKEY"))
    g = j.agg([col("PS_SUPPKEY").count().alias("supplier_cnt")], _cols("P_BRAND", "P_TYPE", "P_SIZE"))
    return g.sort(_cols("supplier_cnt", "P_BRAND", "P_TYPE", "P_SIZE"), descending=[True, False, False, False])


def q17(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:492-507: yearly revenue lost if small orders (< 20 % of the part's mean quantity) of Brand#23 MED BOX parts went away."""
    part = t["part"].filter([(col("P_BRAND") == "Brand#23") & (col("P_CONTAINER") == "MED BOX")])
    res_1 = part.hash_join(t["lineitem"], _cols("P_PARTKEY"), _cols("L_PARTKEY"), how="left")
    g = res_1.agg([(0.2 * col("L_QUANTITY")).mean().alias("avg_quantity")], _cols("P_PARTKEY"))
    g = g.eval_expression_list([col("P_PARTKEY").alias("key"), col("avg_quantity")])
    j = _join(g, res_1, ["key"], ["P_PARTKEY"]).filter([col("L_QUANTITY") < col("avg_quantity")])
    return j.agg([(col("L_EXTENDEDPRICE") / 7.0).sum().alias("avg_yearly")], [])


def q19(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:531-572: revenue of three brand / container / quantity / size brackets delivered in person by air.  The three
    OR-ed brackets share the ship-mode, ship-instruction and P_SIZE >= 1 conjuncts: the reference's simplifier factors them out of the
    OR (simplify_or, src/daft-algebra/src/simplify/boolean.rs:154-260) and PushDownFilter moves each below the join to the side it
    reads; the rest of the disjunction stays above the join."""
    lineitem = t["lineitem"].filter([col("L_SHIPMODE").is_in(["AIR", "AIR REG"]), col("L_SHIPINSTRUCT") == "DELIVER IN PERSON"])
    part = t["part"].filter([col("P_SIZE") >= 1])
    j = _join(lineitem, part, ["L_PARTKEY"], ["P_PARTKEY"])

    def bracket(brand, containers, qlo, qhi, smax):
        return ((col("P_BRAND") == brand) & col("P_CONTAINER").is_in(containers) & (col("L_QUANTITY") >= qlo) & (col("L_QUANTITY") <= qhi)
                & (col("P_SIZE") <= smax))
    pred = (bracket("Brand#12", ["SM CASE", "SM BOX", "SM PACK", "SM PKG"], 1, 11, 5)
            | bracket("Brand#23", ["MED BAG", "MED BOX", "MED PKG", "MED PACK"], 10, 20, 10)
            | bracket("Brand#34", ["LG CASE", "LG BOX", "LG PACK", "LG PKG"], 20, 30, 15))
    return j.filter([pred]).agg([_volume().sum().alias("revenue")], [])


def q20(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:575-611: CANADA's suppliers holding more than half of 1994's shipped quantity of a forest part."""
    li = t["lineitem"].filter([(col("L_SHIPDATE") >= datetime.date(1994, 1, 1)) & (col("L_SHIPDATE") < datetime.date(1995, 1, 1))])
    res_1 = li.agg([(col("L_QUANTITY") * 0.5).sum().alias("sum_quantity")], _cols("L_PARTKEY", "L_SUPPKEY"))
    res_3 = _join(t["supplier"], t["nation"].filter([col("N_NAME") == "CANADA"]), ["S_NATIONKEY"], ["N_NATIONKEY"])
    j = t["part"].filter([col("P_NAME").startswith("forest")]).eval_expression_list(_cols("P_PARTKEY")).dedup(_cols("P_PARTKEY"))
    j = _join(j, t["partsupp"], ["P_PARTKEY"], ["PS_PARTKEY"])
    j = _join(j, res_1, ["PS_SUPPKEY", "P_PARTKEY"], ["L_SUPPKEY", "L_PARTKEY"]).filter([col("PS_AVAILQTY") > col("sum_quantity")])
    j = j.eval_expression_list(_cols("PS_SUPPKEY")).dedup(_cols("PS_SUPPKEY"))
    j = _join(j, res_3, ["PS_SUPPKEY"], ["S_SUPPKEY"])
    return j.eval_expression_list(_cols("S_NAME", "S_ADDRESS")).sort(_cols("S_NAME"))


def q21(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:614-648: SAUDI ARABIA's suppliers that were the only late one on a multi-supplier order with status F; top 100.
    The DataFrame text counts an order's lines (its `.distinct()` is commented out), which equals the SQL's "another supplier
    exists" only while no order repeats a supplier; the SQL's meaning is used here: distinct suppliers per order.  "More than one
    distinct supplier" is written min(L_SUPPKEY) != max(L_SUPPKEY) and "exactly one" min == max: the same predicate as
    count_distinct > 1 / == 1, through the inline min / max accumulators of the fused aggregate kernel instead of a dedup of all
    of lineitem (at SF100 the count_distinct form spent 893 ms here, profiles/r02_tpch22.md)."""
    li = t["lineitem"]
    late = li.filter([col("L_RECEIPTDATE") > col("L_COMMITDATE")])
    span = [col("L_SUPPKEY").min().alias("lo"), col("L_SUPPKEY").max().alias("hi")]
    multi = li.eval_expression_list(_cols("L_SUPPKEY", "L_ORDERKEY")).agg(span, _cols("L_ORDERKEY")).filter([col("lo") != col("hi")])
    res_1 = _join(multi.eval_expression_list(_cols("L_ORDERKEY")), late, ["L_ORDERKEY"], ["L_ORDERKEY"])
    g = res_1.eval_expression_list(_cols("L_SUPPKEY", "L_ORDERKEY")).agg(span, _cols("L_ORDERKEY")).filter([col("lo") == col("hi")])
    j = _join(g.eval_expression_list(_cols("L_ORDERKEY")), res_1, ["L_ORDERKEY"], ["L_ORDERKEY"])
    j = _join(j, t["supplier"], ["L_SUPPKEY"], ["S_SUPPKEY"])
    j = _join(j, t["nation"], ["S_NATIONKEY"], ["N_NATIONKEY"])
    j = _join(j, t["orders"], ["L_ORDERKEY"], ["O_ORDERKEY"])
    j = j.filter([(col("N_NAME") == "SAUDI ARABIA") & (col("O_ORDERSTATUS") == "F")])
    g = j.agg([col("O_ORDERKEY").count().alias("numwait")], _cols("S_NAME"))
    return g.top_n(_cols("numwait", "S_NAME"), 100, descending=[True, False])


def q22(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:651-676: customers of seven country codes with above-average balances and no orders (anti join)."""
    res_1 = _with_column(t["customer"], "cntrycode", col("C_PHONE").left(2))
    res_1 = res_1.filter([col("cntrycode").is_in(["13", "31", "23", "29", "30", "18", "17"])]).eval_expression_list(_cols("C_ACCTBAL", "C_CUSTKEY", "cntrycode"))
    res_2 = _with_column(res_1.filter([col("C_ACCTBAL") > 0]).agg([col("C_ACCTBAL").mean().alias("avg_acctbal")], []), "lit", lit(1))
    j = res_1.hash_join(t["orders"], _cols("C_CUSTKEY"), _cols("O_CUSTKEY"), how="anti")
    j = _join(_with_column(j, "lit", lit(1)), res_2, ["lit"], ["lit"]).filter([col("C_ACCTBAL") > col("avg_acctbal")])
    g = j.agg([col("C_ACCTBAL").count().alias("numcust"), col("C_ACCTBAL").sum().alias("totacctbal")], _cols("cntrycode"))
    return g.sort(_cols("cntrycode"))

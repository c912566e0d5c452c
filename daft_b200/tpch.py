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
    """answers.py:373-395: late MAIL / SHIP lineitems received in 1994, high- and low-priority line counts per ship mode."""
    j = t["orders"].hash_join(t["lineitem"], [col("O_ORDERKEY")], [col("L_ORDERKEY")], build_on_left=True)
    j = j.filter([col("L_SHIPMODE").is_in(["MAIL", "SHIP"]), col("L_COMMITDATE") < col("L_RECEIPTDATE"), col("L_SHIPDATE") < col("L_COMMITDATE"),
                  col("L_RECEIPTDATE") >= datetime.date(1994, 1, 1), col("L_RECEIPTDATE") < datetime.date(1995, 1, 1)])
    urgent = col("O_ORDERPRIORITY").is_in(["1-URGENT", "2-HIGH"])
    out = j.agg([_when(urgent, 1, 0).sum().alias("high_line_count"), _when(~urgent, 1, 0).sum().alias("low_line_count")], [col("L_SHIPMODE")])
    return out.sort([col("L_SHIPMODE")])


def q14(t: dict[str, RecordBatch]) -> RecordBatch:
    """answers.py:418-436: promotional revenue share of September 1995."""
    j = t["lineitem"].hash_join(t["part"], [col("L_PARTKEY")], [col("P_PARTKEY")])
    pred = [col("L_SHIPDATE") >= datetime.date(1995, 9, 1), col("L_SHIPDATE") < datetime.date(1995, 10, 1)]
    rev = col("L_EXTENDEDPRICE") * (1 - col("L_DISCOUNT"))
    a = j.agg([_when(col("P_TYPE").startswith("PROMO"), rev, 0.0).sum().alias("tmp_1"), rev.sum().alias("tmp_2")], [], predicate=pred)
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

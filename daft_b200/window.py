"""Window functions and pivot on the RecordBatch surface.

The reference runs its window kernels one partition at a time (RecordBatch::window_row_number / window_rank / window_offset /
window_grouped_agg, src/daft-recordbatch/src/ops/window.rs:16-82,636-790, inside the Window*Sink operators of
src/daft-local-execution/src/sinks/window_*.rs).  Here ONE sort by (partition keys, order keys) is followed by segmented passes
over the whole table (daft_b200/csrc/sort.cu op_window_ranks / op_window_offset), and per-partition aggregates are a group-by
on the dense partition id joined back through the direct-address join table.  Rows come back sorted by (partition keys, order
keys); the reference emits them partition by partition in hash-partition order, each sorted by the order keys.

pivot follows RecordBatch::pivot (src/daft-recordbatch/src/ops/pivot.rs:70-150) and PivotSink (sinks/pivot.rs).
"""
from __future__ import annotations

from typing import Any, Sequence

from . import datatypes as dt
from .expressions import Expression, _to_expr, col
from .recordbatch import RecordBatch

RANK_FUNCS = {"row_number": 0, "rank": 1, "dense_rank": 2}


def window(batch: RecordBatch, partition_by: Sequence[Expression], order_by: Sequence[Expression] = (),
           descending: Sequence[bool] | bool | None = None, nulls_first: Sequence[bool] | bool | None = None,
           functions: Sequence[tuple] = ()) -> RecordBatch:
    """`functions`: ("row_number" | "rank" | "dense_rank", name), ("offset", name, expr, offset[, default]) -- lead for a positive
    offset, lag for a negative one -- ("agg", name, aggregate expression) for an aggregate over the whole partition, or
    ("frame", name, aggregate expression, start, end[, min_periods = 1]) for sum / count / mean / min / max over the ROWS frame
    [current row + start, current row + end] (None = unbounded on that side; RecordBatch::window_agg_dynamic_frame);
    ("range_frame", ...) with the same fields takes start / end as integer offsets on the value of the single ORDER BY key (an integer,
    a Date in days, a Timestamp in microseconds): the frame holds the rows whose key lies in [key + start, key + end].
    Returns the batch's columns plus one column per function, rows ordered by (partition keys, order keys)."""
    partition_by, order_by = list(partition_by), list(order_by)
    if not partition_by and not order_by:
        raise ValueError("window: partition_by and order_by cannot both be empty")
    n_o = len(order_by)
    desc = [descending] * n_o if isinstance(descending, bool) else list(descending or [False] * n_o)
    nf = [nulls_first] * n_o if isinstance(nulls_first, bool) else list(nulls_first if nulls_first is not None else desc)
    if len(desc) != n_o or len(nf) != n_o:
        raise ValueError("window: descending / nulls_first must have one entry per order_by expression")
    base_names = batch.column_names()
    both, idx = batch._resolve_keys(partition_by + order_by)
    names = both.column_names()
    n_p = len(partition_by)
    keys = [col(names[i]) for i in idx]
    srt = both.sort(keys, descending=[False] * n_p + desc, nulls_first=[False] * n_p + nf)
    ranks = srt._window_ranks(idx[:n_p], idx[n_p:])   # __row_number, __rank, __dense_rank, __partition_begin, __partition_id
    work = srt.union(ranks)
    result = [col(n) for n in base_names]
    seen = set(base_names)
    for f in functions:
        kind, name = f[0], f[1]
        if name in seen:
            raise ValueError(f"window: duplicate output column name {name!r}")
        seen.add(name)
        if kind in RANK_FUNCS:
            result.append(col(f"__{kind}").alias(name))
        elif kind == "offset":
            expr, offset = f[2], int(f[3])
            default = _to_expr(f[4]) if len(f) > 4 and f[4] is not None else None
            vals = work.eval_expression_list([expr.alias("__v")])
            if offset == 0:
                shifted = vals.eval_expression_list([col("__v").alias(f"__w_{name}")])
            else:
                sh = vals._window_offset(work.eval_expression_list([col("__partition_id")]), offset)   # columns __v, __src
                if default is None:
                    shifted = sh.eval_expression_list([col("__v").alias(f"__w_{name}")])
                else:
                    # rows whose shifted position leaves the partition take the default evaluated at that row, cast to the
                    # value's type (window.rs:720-745); nulls that were read from inside the partition stay null
                    tmp = work.union(sh.rename([f"__w0_{name}", f"__src_{name}"]))
                    shifted = tmp.eval_expression_list([(col(f"__src_{name}") >= 0).if_else(col(f"__w0_{name}"), default.cast(vals.dtypes()[0]))
                                                        .alias(f"__w_{name}")])
            work = work.union(shifted)
            result.append(col(f"__w_{name}").alias(name))
        elif kind == "agg":
            if not partition_by:
                raise ValueError("Partition by cannot be empty for window aggregation")
            g = work.agg([f[2].alias(f"__w_{name}")], [col("__partition_id")])
            j = work.eval_expression_list([col("__partition_id")]).hash_join(g, [col("__partition_id")], [col("__partition_id")], build_on_left=False)
            work = work.union(j.eval_expression_list([col(f"__w_{name}")]))
            result.append(col(f"__w_{name}").alias(name))
        elif kind in ("frame", "range_frame"):
            start, end = f[3], f[4]
            min_periods = int(f[5]) if len(f) > 5 else 1
            order = None
            if kind == "range_frame":
                if n_o != 1:   # validate_range_frame_order_by, window.rs:109-116
                    raise ValueError("Range frame requires exactly one ORDER BY column, multiple columns are not supported")
                order = (col(names[idx[n_p]]), desc[0])
            work = work.union(_frame_agg(work, f[2], start, end, min_periods, f"__w_{name}", order))
            result.append(col(f"__w_{name}").alias(name))
        else:
            raise ValueError(f"window: unknown function {kind!r}")
    return work.eval_expression_list(result)


FRAME_AGGS = {"sum": 0, "count": {"all": 1, "valid": 2, "null": 3}, "mean": 4, "min": 5, "max": 6}


def _frame_agg(work: RecordBatch, agg_expr: Expression, start, end, min_periods: int, out_name: str, order=None) -> RecordBatch:
    """One aggregate over a ROWS frame (or, with `order` = (key column, descending), a RANGE frame) of the sorted, rank-annotated batch
    `work` -> a one-column batch named `out_name`.
    The value is cast to the type the reference's window state accumulates in (window_states/sum.rs:137-158: signed integers as Int64,
    unsigned as UInt64, floats as they are; mean.rs:92-100: Float64); min / max run on the Int64 / UInt64 image of integer-like values
    and are cast back (the reference re-aggregates every frame with the ordinary aggregate, window.rs:441-483)."""
    from .expressions import _agg_inner
    inner = _agg_inner(agg_expr)
    if inner.kind != "agg" or inner.op not in FRAME_AGGS:
        raise NotImplementedError(f"window frame: aggregate {getattr(inner, 'op', inner.kind)!r} is not built (sum, count, mean, min, max are)")
    if order is None and start is not None and end is not None and int(start) > int(end) + 1:
        raise ValueError("Negative frame size is not allowed")   # window.rs:156-160 (raised there by the first row it happens on)
    kw = {}
    if order is not None:
        key = work.eval_expression_list([order[0].alias("__o")])
        if not dt.int_like(key.dtypes()[0]) or key.dtypes()[0] == dt.U64:
            raise NotImplementedError(f"window range frame: ORDER BY key of type {dt.NAMES[key.dtypes()[0]]} is not built (integers below UInt64, Date, Timestamp are)")
        if key.dtypes()[0] != dt.I64:
            key = key.eval_expression_list([col("__o").cast(dt.I64).alias("__o")])
        kw = dict(kind=2, order=key, descending=order[1])
    pb = work.eval_expression_list([col("__partition_begin")])
    vals = work.eval_expression_list([inner.args[0].alias("__v")])
    src = vals.dtypes()[0]
    if inner.op == "count":
        if src == dt.NULL:
            raise NotImplementedError("window frame: count over a Null column is not built")
        out = vals._window_frame(pb, FRAME_AGGS["count"][inner.value or "valid"], start, end, min_periods, dt.U64, **kw)
        return out.rename([out_name])
    if inner.op == "mean":
        if not (dt.is_integer(src) or dt.is_float(src)):
            raise TypeError(f"Cannot run Mean over type {dt.NAMES[src]}")
        acc, out_dt, back = dt.F64, dt.F64, None
    elif inner.op == "sum":
        if dt.is_signed(src): acc = dt.I64
        elif dt.is_integer(src): acc = dt.U64
        elif dt.is_float(src): acc = src
        else: raise TypeError(f"Cannot run Sum over type {dt.NAMES[src]}")
        out_dt, back = acc, None
    else:   # min / max
        if not dt.int_like(src):
            raise NotImplementedError(f"window frame: {inner.op} over {dt.NAMES[src]} is not built (integer-like values are)")
        acc = dt.U64 if src == dt.U64 else dt.I64
        out_dt, back = acc, (src if src != acc else None)
    if src != acc:
        vals = vals.eval_expression_list([col("__v").cast(acc).alias("__v")])
    out = vals._window_frame(pb, FRAME_AGGS[inner.op], start, end, min_periods, out_dt, **kw).rename([out_name])
    if back is not None:
        out = out.eval_expression_list([col(out_name).cast(back).alias(out_name)])
    return out


def _pivot_literal(name: str, dtype: int) -> Any:
    if dt.is_string(dtype):
        return name
    if dtype == dt.BOOL:
        return {"true": True, "false": False}[name.lower()]
    if dt.is_integer(dtype):
        return int(name)
    if dt.is_float(dtype):
        return float(name)
    raise ValueError(f"pivot: pivot columns of type {dt.NAMES[dtype]} are not supported")


def pivot(batch: RecordBatch, group_by: Sequence[Expression], pivot_col: Expression, values_col: Expression, names: Sequence[str]) -> RecordBatch:
    """One row per distinct group key (first-occurrence order), one column per entry of `names` holding the value of the row whose
    pivot column spells that name, null where there is none.  Like the reference this assumes (group, pivot value) pairs are
    unique -- PivotSink aggregates first (agg_pivot)."""
    group_by = list(group_by)
    if not group_by:
        raise ValueError("pivot: group_by cannot be empty")
    gnames = [g.name() for g in group_by]
    base = batch.eval_expression_list([g.alias(f"__g{i}") for i, g in enumerate(group_by)] + [pivot_col.alias("__p"), values_col.alias("__v")])
    gcols = [col(f"__g{i}") for i in range(len(group_by))]
    pdtype = base.dtypes()[len(group_by)]
    out = base.eval_expression_list(gcols).dedup(gcols)
    keys_only = out
    never = col("__p").is_null() & col("__p").not_null()   # a name the pivot column cannot spell selects no row: an all-null column
    for i, name in enumerate(names):
        try:
            pred = col("__p") == _pivot_literal(name, pdtype)
            if name == "None" and dt.is_string(pdtype):
                pred = pred | col("__p").is_null()
        except (ValueError, KeyError):
            # the reference matches names against the pivot values' string form, in which a null spells "None" (pivot.rs:24-44)
            pred = col("__p").is_null() if name == "None" else never
        rows = base.filter([pred]).eval_expression_list(gcols + [col("__v").alias(f"__n{i}")])
        j = keys_only.hash_join(rows, gcols, gcols, how="left", null_equals_null=[True] * len(gcols))
        colv = j.eval_expression_list([col(f"__n{i}")])
        out = out.union(colv)
    return out.eval_expression_list([col(f"__g{i}").alias(n) for i, n in enumerate(gnames)] + [col(f"__n{i}").alias(n) for i, n in enumerate(names)])


def agg_pivot(batch: RecordBatch, group_by: Sequence[Expression], pivot_col: Expression, value_col: Expression, agg: str,
              names: Sequence[str]) -> RecordBatch:
    """PivotSink (src/daft-local-execution/src/sinks/pivot.rs): aggregate the value per (group, pivot value), then pivot."""
    group_by = list(group_by)
    pre = batch.eval_expression_list([g.alias(f"__g{i}") for i, g in enumerate(group_by)] + [pivot_col.alias("__p"), value_col.alias("__v")])
    keys = [col(f"__g{i}") for i in range(len(group_by))] + [col("__p")]
    g = pre.agg([getattr(col("__v"), agg)().alias("__v")], keys)
    res = pivot(g, [col(f"__g{i}") for i in range(len(group_by))], col("__p"), col("__v"), names)
    return res.rename([gb.name() for gb in group_by] + list(names))

"""CPU restatement of the reference's window kernels and pivot (TEST INFRASTRUCTURE ONLY).

Per partition, as the Window*Sink operators drive them (src/daft-local-execution/src/sinks/window_partition_and_order_by.rs,
window_partition_only.rs): partition by value, sort the partition by the order keys, then
* row_number -- RecordBatch::window_row_number, src/daft-recordbatch/src/ops/window.rs:636-642: 1..n
* rank / dense_rank -- window_rank, :644-704: walk the sorted rows; rows equal to their predecessor on the order keys (null == null,
  NaN == NaN) share a rank; rank jumps to the row's position, dense_rank advances by one
* offset (lead / lag) -- window_offset, :706-789: the value `offset` rows away inside the partition, else the default (cast to the
  value's type) or null
* partition aggregate -- window_grouped_agg, :16-82: the aggregate per partition broadcast to its rows
* aggregate over a ROWS frame -- window_agg_dynamic_frame, :299-385: sum / count / mean slide a state over the partition
  (window_agg_rows_incremental :388-438, update_incremental_state :274-296, ops/window_states/{sum,count,mean}.rs), min / max
  re-aggregate every frame (window_agg_rows :441-483)
pivot -- RecordBatch::pivot, src/daft-recordbatch/src/ops/pivot.rs:70-150.
Small inputs only (pure-Python loops)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import pyarrow as pa

from . import expr_eval as E
from . import oracle as O


def _same(a, b) -> bool:
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, float) and isinstance(b, float) and a != a and b != b:
        return True
    return a == b


def window(tbl: pa.Table, partition_by: Sequence, order_by: Sequence, descending: Sequence[bool], nulls_first: Sequence[bool],
           functions: Sequence[tuple]) -> pa.Table:
    """Same arguments as daft_b200.window.window (expressions are daft_b200 Expression objects); rows come back partition by
    partition (first-occurrence order), each sorted by the order keys."""
    n = tbl.num_rows
    pcols = [E.eval_expression(tbl, e) for e in partition_by]
    ocols = [E.eval_expression(tbl, e) for e in order_by]
    if pcols and n:
        gid, first, _ = O.group_ids(pcols)
        parts = [np.nonzero(gid == g)[0] for g in range(len(first))]
    else:
        parts = [np.arange(n)] if n else []
    order, new_cols, frame_types = [], {f[1]: [] for f in functions}, {}
    ovals = list(zip(*[O.normalize(c).to_pylist() for c in ocols])) if ocols else [()] * n
    for rows in parts:
        if ocols:
            sub = [c.take(pa.array(rows)) for c in ocols]
            rows = rows[O.argsort(sub, list(descending), list(nulls_first))]
        order += list(rows)
        m = len(rows)
        rank, dense = [1] * m, [1] * m
        for i in range(1, m):
            eq = all(_same(a, b) for a, b in zip(ovals[rows[i - 1]], ovals[rows[i]]))
            rank[i] = rank[i - 1] if eq else i + 1
            dense[i] = dense[i - 1] if eq else dense[i - 1] + 1
        sub_tbl = tbl.take(pa.array(rows))
        for f in functions:
            kind, name = f[0], f[1]
            if kind == "row_number":
                new_cols[name] += list(range(1, m + 1))
            elif kind == "rank":
                new_cols[name] += rank
            elif kind == "dense_rank":
                new_cols[name] += dense
            elif kind == "offset":
                vals = O.normalize(E.eval_expression(sub_tbl, f[2])).to_pylist()
                off = int(f[3])
                dflt = None
                if len(f) > 4 and f[4] is not None:
                    from daft_b200.expressions import _to_expr
                    d = E.eval_expression(sub_tbl, _to_expr(f[4]))
                    dflt = O.normalize(d).to_pylist()
                    if len(dflt) == 1 and m != 1:
                        dflt = dflt * m
                for i in range(m):
                    j = i + off
                    if 0 <= j < m:
                        new_cols[name].append(vals[j])
                    else:
                        v = None if dflt is None else dflt[i]
                        if v is not None and vals and isinstance(next((x for x in vals if x is not None), v), float):
                            v = float(v)
                        new_cols[name].append(v)
            elif kind in ("frame", "range_frame"):
                rng_kw = {}
                if kind == "range_frame":
                    if len(order_by) != 1:
                        raise ValueError("Range frame requires exactly one ORDER BY column, multiple columns are not supported")
                    rng_kw = dict(order=E.eval_expression(sub_tbl, order_by[0]), descending=bool(list(descending)[0]))
                vals, frame_types[name] = frame_agg(sub_tbl, f[2], f[3], f[4], int(f[5]) if len(f) > 5 else 1, **rng_kw)
                new_cols[name] += vals
            elif kind == "agg":
                from daft_b200.expressions import _agg_inner
                inner = _agg_inner(f[2])
                one = O.agg(sub_tbl, [(inner.op, E.eval_expression(sub_tbl, inner.args[0]), name, inner.value)], [])
                new_cols[name] += one.column(name).to_pylist() * m
    out = tbl.take(pa.array(np.array(order, dtype=np.int64)))
    for f in functions:
        kind, name = f[0], f[1]
        if kind in ("row_number", "rank", "dense_rank"):
            typ = pa.uint64()
        elif kind == "offset":
            typ = O.normalize(E.eval_expression(tbl, f[2])).type
        elif kind in ("frame", "range_frame"):
            typ = frame_types.get(name) or frame_agg(tbl.slice(0, 0), f[2], f[3], f[4], 1)[1]
        else:   # the aggregate's output type: what the ungrouped aggregate over the whole table yields
            from daft_b200.expressions import _agg_inner
            inner = _agg_inner(f[2])
            typ = O.agg(tbl, [(inner.op, E.eval_expression(tbl, inner.args[0]), name, inner.value)], []).column(name).type
        out = out.append_column(name, pa.array(new_cols[name], type=typ))
    return out


FRAME_ACC = {"int8": "int64", "int16": "int64", "int32": "int64", "int64": "int64", "uint8": "uint64", "uint16": "uint64", "uint32": "uint64",
             "uint64": "uint64", "float": "float", "double": "double"}   # window_states/sum.rs:137-158


def frame_agg_partition(vals: list, op: str, mode, start, end, min_periods: int, acc: str, order=None, descending: bool = False) -> list:
    """One partition, already in frame order.  `vals`: Python values (None = null) in the accumulate type `acc`.  The state machine of
    window_agg_rows_incremental: a frame with fewer than min_periods rows leaves the state alone and gives null; otherwise rows
    [prev_start, start) are removed and rows [prev_end, end) added -- in that order, with `+=` / `-=` on the accumulate type."""
    m = len(vals)
    ftype = {"float": np.float32, "double": np.float64}.get(acc)

    def wrap(x):
        return (x + 2**63) % 2**64 - 2**63

    def bounds(r, ps=0, pe=0):
        if order is None:   # calculate_row_frame_bounds, window.rs:136-163
            fs = 0 if start is None else min(max(r + start, 0), m)
            fe = m if end is None else min(max(r + 1 + end, 0), m)
        else:
            # calculate_range_frame_bounds, window.rs:167-272: each boundary is searched forward from the previous frame's; comparisons with
            # a null are false; descending order with two offsets swaps them (window.rs:507-513)
            so, eo = (end, start) if descending and start is not None and end is not None else (start, end)
            cur = order[r]
            fs, fe = 0, m
            if start is not None:
                fs = ps
                while fs < m and not (cur is not None and order[fs] is not None and
                                      (order[fs] <= wrap(cur + so) if descending else order[fs] >= wrap(cur + so))):
                    fs += 1
            if end is not None:
                fe = pe
                while fe < m and not (cur is not None and order[fe] is not None and
                                      (order[fe] < wrap(cur + eo) if descending else order[fe] > wrap(cur + eo))):
                    fe += 1
        if fe < fs:
            raise ValueError("Negative frame size is not allowed")
        return fs, fe
    out = []
    if op in ("min", "max"):   # no incremental state (window_states/mod.rs:47-49): aggregate the slice
        ps = pe = 0
        for r in range(m):
            fs, fe = ps, pe = bounds(r, ps, pe)   # window_agg_range: the search resumes from every previous row's frame (window.rs:609-610)
            live = [v for v in vals[fs:fe] if v is not None] if fe - fs >= min_periods else []
            out.append((min(live) if op == "min" else max(live)) if live else None)
        return out
    total = sm = valid = valid_not_nan = nan = 0
    if ftype:
        sm = ftype(0)
    ps = pe = 0
    for r in range(m):
        fs, fe = bounds(r, ps, pe)   # ... from the last frame that was large enough (update_incremental_state is skipped otherwise)
        enough = fe - fs >= min_periods
        if enough:
            for lo, hi, sign in ((ps, fs, -1), (pe, fe, 1)):
                if hi <= lo:
                    continue
                total += sign * (hi - lo)
                for v in vals[lo:hi]:
                    if v is None:
                        continue
                    valid += sign
                    if op == "count":
                        continue
                    if ftype and v != v:
                        nan += sign
                    else:
                        valid_not_nan += sign
                        with np.errstate(all="ignore"):
                            sm = (sm + ftype(v) if sign > 0 else sm - ftype(v)) if ftype else sm + sign * v
            ps, pe = fs, fe
        if op == "count":
            out.append(({"all": total, "valid": valid, "null": total - valid}[mode]) if enough else None)
        elif not enough or valid_not_nan <= 0:
            out.append(None)
        elif op == "sum":
            if ftype:
                out.append(float("nan") if nan > 0 else float(sm))
            else:
                out.append(((sm + 2**63) % 2**64 - 2**63) if acc == "int64" else sm % 2**64)
        else:   # mean: sum series / count series (mean.rs:75-80)
            with np.errstate(all="ignore"):
                out.append(float("nan") if nan > 0 else float(np.float64(sm) / np.float64(valid)))
    return out


def frame_agg(sub_tbl: pa.Table, agg_expr, start, end, min_periods: int, order=None, descending: bool = False):
    """-> (values, arrow type) of `agg_expr` over the ROWS frame [start, end] (None = unbounded) of every row of one partition; with `order`
    (the partition's single ORDER BY column) a RANGE frame: offsets on the key's integer value (days / microseconds for temporal keys)"""
    from daft_b200.expressions import _agg_inner
    inner = _agg_inner(agg_expr)
    kw = {}
    if order is not None:
        o = O.normalize(order)
        phys = pa.int32() if pa.types.is_date32(o.type) else pa.int64()
        if not (pa.types.is_integer(o.type) or pa.types.is_date32(o.type) or pa.types.is_timestamp(o.type)) or pa.types.is_uint64(o.type):
            raise NotImplementedError("window range frame over a non-integer ORDER BY key")
        kw = dict(order=(o.cast(phys) if not pa.types.is_integer(o.type) else o).to_pylist(), descending=descending)
    if inner.op not in ("sum", "count", "mean", "min", "max"):
        raise NotImplementedError(f"window frame aggregate {inner.op}")
    child = O.normalize(E.eval_expression(sub_tbl, inner.args[0]))
    if inner.op == "count":
        return frame_agg_partition(child.to_pylist(), "count", inner.value or "valid", start, end, min_periods, "", **kw), pa.uint64()
    if inner.op == "mean":
        acc, typ = "double", pa.float64()
        vals = child.cast(pa.float64(), safe=False).to_pylist()   # i64 -> f64 rounds to nearest, as the reference cast does
    elif inner.op in ("min", "max"):
        if not (pa.types.is_integer(child.type) or pa.types.is_date32(child.type) or pa.types.is_timestamp(child.type)):
            raise NotImplementedError("window frame min / max over non-integer values")
        # plain comparisons of the values themselves (integers, dates, instants): no accumulate type involved
        return frame_agg_partition(child.to_pylist(), inner.op, None, start, end, min_periods, "", **kw), child.type
    else:
        acc = FRAME_ACC.get(str(child.type))
        if acc is None:
            raise TypeError(f"Cannot run Sum over type {child.type}")
        typ = {"int64": pa.int64(), "uint64": pa.uint64(), "float": pa.float32(), "double": pa.float64()}[acc]
        vals = child.cast(typ).to_pylist()
    return frame_agg_partition(vals, inner.op, None, start, end, min_periods, acc, **kw), typ


def pivot(tbl: pa.Table, group_by: Sequence[str], pivot_col: str, values_col: str, names: Sequence[str]) -> pa.Table:
    """pivot.rs:70-150: group keys in first-occurrence order; names match the pivot values' string form."""
    keys = list(zip(*[O.normalize(tbl.column(g)).to_pylist() for g in group_by]))
    pv, vv = O.normalize(tbl.column(pivot_col)).to_pylist(), O.normalize(tbl.column(values_col)).to_pylist()
    groups: dict = {}
    for i, k in enumerate(keys):
        groups.setdefault(k, {})
        s = "None" if pv[i] is None else (str(pv[i]).lower() if isinstance(pv[i], bool) else str(pv[i]))   # str_value of a null is "None"
        groups[k][s] = vv[i]
    cols = {g: [k[j] for k in groups] for j, g in enumerate(group_by)}
    out = pa.table({g: pa.array(cols[g], type=tbl.schema.field(g).type) for g in group_by})
    for nme in names:
        out = out.append_column(nme, pa.array([groups[k].get(nme) for k in groups], type=tbl.schema.field(values_col).type))
    return out

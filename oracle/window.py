"""CPU restatement of the reference's window kernels and pivot (TEST INFRASTRUCTURE ONLY).

Per partition, as the Window*Sink operators drive them (src/daft-local-execution/src/sinks/window_partition_and_order_by.rs,
window_partition_only.rs): partition by value, sort the partition by the order keys, then
* row_number -- RecordBatch::window_row_number, src/daft-recordbatch/src/ops/window.rs:636-642: 1..n
* rank / dense_rank -- window_rank, :644-704: walk the sorted rows; rows equal to their predecessor on the order keys (null == null,
  NaN == NaN) share a rank; rank jumps to the row's position, dense_rank advances by one
* offset (lead / lag) -- window_offset, :706-789: the value `offset` rows away inside the partition, else the default (cast to the
  value's type) or null
* partition aggregate -- window_grouped_agg, :16-82: the aggregate per partition broadcast to its rows
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
    order, new_cols = [], {f[1]: [] for f in functions}
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
        else:   # the aggregate's output type: what the ungrouped aggregate over the whole table yields
            from daft_b200.expressions import _agg_inner
            inner = _agg_inner(f[2])
            typ = O.agg(tbl, [(inner.op, E.eval_expression(tbl, inner.args[0]), name, inner.value)], []).column(name).type
        out = out.append_column(name, pa.array(new_cols[name], type=typ))
    return out


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

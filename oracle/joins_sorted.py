"""CPU restatement of the reference's joins over SORTED inputs and its cross join (TEST INFRASTRUCTURE ONLY: tests/ compare the
CUDA path with these; nothing under daft_b200/ imports this module).

* merge_inner_join   -- src/daft-recordbatch/src/ops/joins/merge_join.rs:45-330.  The reference walks both sorted key tables with two
  pointers; its run-staging state machine elides comparisons but emits exactly the pairs of a plain merge: for every run of
  equal keys, every (left row, right row) pair.  Keys compare through build_partial_compare_with_nulls
  (src/daft-core/src/kernels/search_sorted.rs:337-361): ascending, nulls last, null vs null incomparable -> never a match.
  The ORDER of the emitted pairs is an artefact of that machine; callers here compare pair multisets.
* asof_join_backward -- src/daft-recordbatch/src/ops/joins/asof_join.rs:12-160 driven per by-group as
  MicroPartition::asof_join does (src/daft-micropartition/src/ops/join.rs:127-187): partition both sides by their by-values,
  map left groups to right groups (null by-keys map to nothing), sort each pair of groups by the on-key, two-pointer walk.
* cross_join         -- src/daft-recordbatch/src/ops/joins/mod.rs:231-284.
Small inputs only (pure-Python loops)."""
from __future__ import annotations

from typing import Sequence

import numpy as np
import pyarrow as pa

from . import oracle as O


def _rows(cols: Sequence) -> list[tuple]:
    return list(zip(*[O.normalize(c).to_pylist() for c in cols])) if cols else []


def _cmp_partial(a, b):
    """one key column: None when both are null (search_sorted.rs:353-360)"""
    if a is None and b is None:
        return None
    if a is None:
        return 1
    if b is None:
        return -1
    return O._cmp_values(a, b, False)


def _cmp_rows(x: tuple, y: tuple):
    for a, b in zip(x, y):
        c = _cmp_partial(a, b)
        if c != 0:
            return c          # None (incomparable) or the first difference
    return 0


def merge_inner_join_pairs(left_keys: Sequence, right_keys: Sequence) -> list[tuple[int, int]]:
    """(left row, right row) pairs of two key tables sorted ascending, nulls last."""
    L, R = _rows(left_keys), _rows(right_keys)
    out, i, j = [], 0, 0
    while i < len(L) and j < len(R):
        c = _cmp_rows(L[i], R[j])
        if c is None:          # MergeJoinState::BothNull: both pointers move on
            i += 1
            j += 1
        elif c < 0:
            i += 1
        elif c > 0:
            j += 1
        else:
            i2 = i
            while i2 < len(L) and _cmp_rows(L[i2], R[j]) == 0:
                i2 += 1
            j2 = j
            while j2 < len(R) and _cmp_rows(L[i], R[j2]) == 0:
                j2 += 1
            out += [(a, b) for a in range(i, i2) for b in range(j, j2)]
            i, j = i2, j2
    return out


def sort_merge_join(left: pa.Table, right: pa.Table, left_on: Sequence[str], right_on: Sequence[str]) -> pa.Table:
    """RecordBatch::sort_merge_join (ops/joins/mod.rs:154-229) with the inner join's output schema."""
    n = len(left_on)
    ls = O.sort(left, list(left_on), [False] * n, [False] * n)
    rs = O.sort(right, list(right_on), [False] * n, [False] * n)
    pairs = merge_inner_join_pairs([ls.column(k) for k in left_on], [rs.column(k) for k in right_on])
    lt = ls.take(pa.array([p[0] for p in pairs], type=pa.int64()))
    rt = rs.take(pa.array([p[1] for p in pairs], type=pa.int64()))
    common = [l for l, r in zip(left_on, right_on) if l == r]
    cols, names = [lt.column(k) for k in common], list(common)
    for nme in left.column_names:
        if nme not in common:
            cols.append(lt.column(nme)); names.append(nme)
    for nme in right.column_names:
        if nme not in common:
            cols.append(rt.column(nme)); names.append(nme if nme not in left.column_names else f"right.{nme}")
    return pa.table(cols, names=names)


def _asof_backward_sorted(lon: list, ron: list) -> list:
    """asof_join.rs:100-150 without by-keys: both on-key lists sorted ascending, nulls last; index of the match per left row."""
    out, r, best = [], 0, None
    for x in lon:
        if x is None:
            out.append(None)
            continue
        while r < len(ron):
            c = _cmp_partial(x, ron[r])
            if c is not None and c >= 0:      # Greater | Equal: this right row qualifies, look further
                best = r
                r += 1
            else:                              # Less | None (a null right on-key): stop
                break
        out.append(best)
    return out


def asof_join(left: pa.Table, right: pa.Table, left_by: Sequence[str], right_by: Sequence[str], left_on: str, right_on: str) -> pa.Table:
    """Output: all left columns, then the right columns minus the right by-columns (and the right on-column when it is named
    like the left one); one row per left row.  Row order: left groups in first-occurrence order, sorted by the on-key."""
    drop = set(right_by) | ({right_on} if left_on == right_on else set())
    lby, rby = _rows([left.column(k) for k in left_by]), _rows([right.column(k) for k in right_by])
    if not left_by:
        lby, rby = [()] * left.num_rows, [()] * right.num_rows
    lgroups: dict = {}
    for i, k in enumerate(lby):
        lgroups.setdefault(k, []).append(i)
    rgroups: dict = {}
    for j, k in enumerate(rby):
        if not any(v is None for v in k):      # build_left_to_right_map: a null by-key matches nothing
            rgroups.setdefault(k, []).append(j)
    lon_all, ron_all = O.normalize(left.column(left_on)).to_pylist(), O.normalize(right.column(right_on)).to_pylist()
    lidx, ridx = [], []

    def sort_key(v):
        return (v is None, 0 if v is None else (float("inf") if isinstance(v, float) and v != v else v))
    for k, rows in lgroups.items():
        lrows = sorted(rows, key=lambda i: sort_key(lon_all[i]))
        rrows = sorted(rgroups.get(k, []) if not any(v is None for v in k) else [], key=lambda j: sort_key(ron_all[j]))
        m = _asof_backward_sorted([lon_all[i] for i in lrows], [ron_all[j] for j in rrows])
        lidx += lrows
        ridx += [-1 if x is None else rrows[x] for x in m]
    lt = left.take(pa.array(lidx, type=pa.int64()))
    rt = O._take_nullable(right, np.array(ridx, dtype=np.int64))
    cols, names = list(lt.columns), list(left.column_names)
    for nme in right.column_names:
        if nme not in drop:
            cols.append(rt.column(nme)); names.append(nme)
    return pa.table(cols, names=names)


def cross_join(left: pa.Table, right: pa.Table, outer_loop_side: str = "left") -> pa.Table:
    nl, nr = left.num_rows, right.num_rows
    if outer_loop_side == "left":
        li, ri = np.repeat(np.arange(nl), nr), np.tile(np.arange(nr), nl)
    else:
        li, ri = np.tile(np.arange(nl), nr), np.repeat(np.arange(nr), nl)
    lt, rt = left.take(pa.array(li, type=pa.int64())), right.take(pa.array(ri, type=pa.int64()))
    return pa.table(list(lt.columns) + list(rt.columns), names=left.column_names + right.column_names)

"""Hash-join operator protocol on top of the C ABI (build all -> probe batches -> finalize).

Mirrors HashJoinOperator / HashJoinBuildState / probe_* / finalize_*
(src/daft-local-execution/src/join/hash_join.rs:33-362, inner_join.rs:11-85, left_right_join.rs:15-269,
outer_join.rs:57-288, anti_semi_join.rs) and ProbeableBuilder/Probeable
(src/daft-recordbatch/src/probeable/mod.rs:33-126).
"""
from __future__ import annotations

import ctypes
from typing import Sequence

from . import _lib
from . import datatypes as dt
from .expressions import Expression, col
from .recordbatch import JOIN_TYPES, RecordBatch, _i32, _u8


class ProbeTable:
    """ProbeTableBuilder + ProbeTable: add_table()* -> build() -> probe_*()."""

    def __init__(self, nkeys: int, null_equals_null: Sequence[bool] | None = None, track_indices: bool = True):
        _lib.ensure_init()
        self.nkeys = nkeys
        self._h = ctypes.c_void_p()
        ne = _u8(null_equals_null) if null_equals_null is not None else None
        _lib.call("b200_probe_table_create", ctypes.c_int32(nkeys), ne, ctypes.c_int32(1 if track_indices else 0), ctypes.byref(self._h))
        self._keep = []

    def __del__(self):
        try:
            if self._h and self._h.value:
                _lib.lib().b200_probe_table_free(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    def add_table(self, batch: RecordBatch, key_cols: Sequence[int]) -> None:
        self._keep.append(batch)
        _lib.call("b200_probe_table_add", self._h, batch._h, _i32(key_cols))

    def build(self) -> "ProbeTable":
        _lib.call("b200_probe_table_build", self._h)
        return self

    def num_rows(self) -> int:
        return int(_lib.lib().b200_probe_table_num_rows(self._h))

    def probe_exists(self, probe: RecordBatch, key_cols: Sequence[int]) -> RecordBatch:
        out = ctypes.c_void_p()
        _lib.call("b200_probe_exists", self._h, probe._h, _i32(key_cols), ctypes.byref(out))
        return RecordBatch._wrap(out)

    def probe_indices(self, probe: RecordBatch, key_cols: Sequence[int]) -> RecordBatch:
        out = ctypes.c_void_p()
        _lib.call("b200_probe_indices", self._h, probe._h, _i32(key_cols), ctypes.byref(out))
        return RecordBatch._wrap(out)


class JoinState:
    def __init__(self, pt: ProbeTable):
        self._h = ctypes.c_void_p()
        self._pt = pt
        _lib.call("b200_join_state_create", pt._h, ctypes.byref(self._h))

    def __del__(self):
        try:
            if self._h and self._h.value:
                _lib.lib().b200_join_state_free(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


def _with_keys(batch: RecordBatch, on: Sequence[Expression], key_dtypes: Sequence[int]) -> tuple[RecordBatch, list[int]]:
    """Append the evaluated (and supertype-cast) join key expressions as trailing columns.  A key that is a plain
    column of the right dtype is passed through untouched (zero-copy; it may be a view, which the probe reads through
    its selection vector), so a chain of joins never gathers its keys."""
    from .expressions import bind
    names, dts = batch.column_names(), batch.dtypes()
    exprs = []
    for i, (e, t) in enumerate(zip(on, key_dtypes)):
        same = bind(e, names, dts).dtype == t
        exprs.append((e if same else e.cast(t)).alias(f"__key{i}"))
    keys = batch.eval_expression_list(exprs)
    base = batch.num_columns()
    return batch.union(keys), [base + i for i in range(len(on))]


def hash_join(left: RecordBatch, right: RecordBatch, left_on: Sequence[Expression], right_on: Sequence[Expression],
              how: str = "inner", null_equals_null: Sequence[bool] | None = None,
              build_on_left: bool | None = None) -> RecordBatch:
    if len(left_on) != len(right_on):
        raise ValueError(f"Mismatch of number of join keys, left_on: {len(left_on)}, right_on: {len(right_on)}")
    if len(left_on) == 0:
        raise ValueError("No columns were passed in to join on")
    if not isinstance(right, RecordBatch):
        raise TypeError(f"Expected a Table for `right` in join but got {type(right)}")
    if how not in JOIN_TYPES:
        raise ValueError(f"unknown join type {how!r}")
    from .expressions import bind
    lnames, ldt, rnames, rdt = left.column_names(), left.dtypes(), right.column_names(), right.dtypes()
    # join keys are cast to their common supertype (match_types_for_tables, ops/joins/mod.rs:22-46):
    # the same number hashes differently as Int32 and Int64
    kd = []
    for le, re_ in zip(left_on, right_on):
        kd.append(dt.supertype(bind(le, lnames, ldt).dtype, bind(re_, rnames, rdt).dtype))
    common = [le.name() for le, re_ in zip(left_on, right_on)
              if le.kind == "col" and re_.kind == "col" and le.name() == re_.name()]
    if build_on_left is None:
        # the planner builds on the smaller side (pipeline.rs:1106-1194); batch-level default: build right,
        # except right joins which must preserve the probe (right) side's rows inline
        build_on_left = how == "right"
    if how in ("semi", "anti"):
        build_on_left = False
    laug, lkeys = _with_keys(left, left_on, kd)
    raug, rkeys = _with_keys(right, right_on, kd)
    build, bkeys, probe, pkeys = (laug, lkeys, raug, rkeys) if build_on_left else (raug, rkeys, laug, lkeys)
    pt = ProbeTable(len(kd), null_equals_null, track_indices=how not in ("semi", "anti"))
    pt.add_table(build, bkeys)
    pt.build()
    jt = JOIN_TYPES[how]
    need_state = how in ("left", "right", "outer") and (
        how == "outer" or (how == "left" and build_on_left) or (how == "right" and not build_on_left))
    st = JoinState(pt) if need_state else None
    out = ctypes.c_void_p()
    _lib.call("b200_hash_join_probe", pt._h, st._h if st else None, probe._h, _i32(pkeys), ctypes.c_int32(jt),
              ctypes.c_int32(1 if build_on_left else 0), ctypes.byref(out))
    # joined columns: [left_aug..., right_aug...] (the build side comes first iff it is the left side)
    joined = RecordBatch._wrap(out, names=laug.column_names() + raug.column_names(), dtypes=laug.dtypes() + raug.dtypes()) \
        if how not in ("semi", "anti") else RecordBatch._wrap(out, like=probe)
    if how in ("semi", "anti"):
        return joined.select_columns(list(range(left.num_columns())))
    if st is not None:
        fin = ctypes.c_void_p()
        probe_like = probe.slice(0, 0)  # keep the handle alive across the call
        _lib.call("b200_hash_join_finalize", pt._h, st._h, probe_like._h, ctypes.c_int32(1 if build_on_left else 0), ctypes.byref(fin))
        joined = RecordBatch.concat([joined, RecordBatch._wrap(fin)])
    # joined columns: [left_aug..., right_aug...]
    nl, nr = laug.num_columns(), raug.num_columns()
    lbase, rbase = 0, nl
    sel: list[Expression] = []
    jn = joined.column_names()
    uniq = [f"c{i}" for i in range(len(jn))]
    joined = joined.rename(uniq)
    for k in common:
        li, ri = lnames.index(k), rnames.index(k)
        e = col(uniq[lbase + li])
        if how in ("right", "outer"):
            # rows that exist only on the right still carry their key
            e = e.fill_null(col(uniq[rbase + ri]))
        sel.append(e.alias(k))
    for i, n in enumerate(lnames):
        if n not in common:
            sel.append(col(uniq[lbase + i]).alias(n))
    for i, n in enumerate(rnames):
        if n not in common:
            name = n if n not in lnames else f"right.{n}"
            sel.append(col(uniq[rbase + i]).alias(name))
    return joined.eval_expression_list(sel)


# ------------------------------------------------------------------------------------------------------------------------
# Sort-merge, asof and cross joins: RecordBatch::sort_merge_join / asof_join / cross_join of
# src/daft-recordbatch/src/ops/joins/mod.rs:81-111,154-229,231-284, composed from sort + binary search + view gather
# (daft_b200/csrc/sort.cu op_merge_join_sorted / op_asof_join_sorted / op_cross_join).
# ------------------------------------------------------------------------------------------------------------------------
def _call_join(fn: str, laug: RecordBatch, lkeys, raug: RecordBatch, rkeys) -> RecordBatch:
    out = ctypes.c_void_p()
    _lib.call(fn, laug._h, _i32(lkeys), raug._h, _i32(rkeys), ctypes.c_int32(len(lkeys)), ctypes.byref(out))
    joined = RecordBatch._wrap(out, names=laug.column_names() + raug.column_names(), dtypes=laug.dtypes() + raug.dtypes())
    return joined.rename([f"c{i}" for i in range(joined.num_columns())])


def _key_dtypes(left: RecordBatch, right: RecordBatch, left_on, right_on) -> list[int]:
    from .expressions import bind
    lnames, ldt, rnames, rdt = left.column_names(), left.dtypes(), right.column_names(), right.dtypes()
    return [dt.supertype(bind(le, lnames, ldt).dtype, bind(re_, rnames, rdt).dtype) for le, re_ in zip(left_on, right_on)]


def sort_merge_join(left: RecordBatch, right: RecordBatch, left_on: Sequence[Expression], right_on: Sequence[Expression],
                    is_sorted: bool = False) -> RecordBatch:
    """Inner join of two batches over their sorted keys.  Both sides are sorted ascending, nulls last, on the join keys unless
    `is_sorted`; output schema as for the inner hash join (infer_join_schema): same-named keys once, left columns, right
    columns (`right.` prefix on clashes).  Rows come left-major in key order; a null in any key matches nothing."""
    if len(left_on) != len(right_on):
        raise ValueError(f"Mismatch of join on clauses: left: {len(left_on)} vs right: {len(right_on)}")
    if len(left_on) == 0:
        raise ValueError("No columns were passed in to join on")
    kd = _key_dtypes(left, right, left_on, right_on)
    lnames, rnames = left.column_names(), right.column_names()
    laug, lkeys = _with_keys(left, left_on, kd)
    raug, rkeys = _with_keys(right, right_on, kd)
    if not is_sorted:
        nk = len(kd)
        laug = laug.sort([col(f"__key{i}") for i in range(nk)], descending=[False] * nk, nulls_first=[False] * nk)
        raug = raug.sort([col(f"__key{i}") for i in range(nk)], descending=[False] * nk, nulls_first=[False] * nk)
    joined = _call_join("b200_merge_join_sorted", laug, lkeys, raug, rkeys)
    common = [le.name() for le, re_ in zip(left_on, right_on) if le.kind == "col" and re_.kind == "col" and le.name() == re_.name()]
    rbase = laug.num_columns()
    sel = [col(f"c{lnames.index(k)}").alias(k) for k in common]
    sel += [col(f"c{i}").alias(n) for i, n in enumerate(lnames) if n not in common]
    sel += [col(f"c{rbase + i}").alias(n if n not in lnames else f"right.{n}") for i, n in enumerate(rnames) if n not in common]
    return joined.eval_expression_list(sel)


def asof_join(left: RecordBatch, right: RecordBatch, left_by: Sequence[Expression], right_by: Sequence[Expression],
              left_on: Expression, right_on: Expression) -> RecordBatch:
    """Backward asof join (MicroPartition::asof_join, src/daft-micropartition/src/ops/join.rs:127-187): every left row with the
    last right row of the same by-group whose on-key is <= its own, or nulls.  Output: all left columns, then the right
    columns except the right by-columns and a right on-column named like the left one (infer_asof_join_schema,
    src/daft-dsl/src/join.rs:69-120).  Left rows keep their order (the reference emits them per by-group, sorted by the on-key)."""
    if len(left_by) != len(right_by):
        raise ValueError(f"Mismatch of asof by-keys: left: {len(left_by)} vs right: {len(right_by)}")
    lon, ron = list(left_by) + [left_on], list(right_by) + [right_on]
    kd = _key_dtypes(left, right, lon, ron)
    lnames, rnames = left.column_names(), right.column_names()
    drop = {e.name() for e in right_by if e.kind == "col"}
    if left_on.kind == "col" and right_on.kind == "col" and left_on.name() == right_on.name():
        drop.add(right_on.name())
    clash = [n for n in rnames if n not in drop and n in lnames]
    if clash:
        raise ValueError(f"asof_join: duplicate output column name(s) {clash}")
    laug, lkeys = _with_keys(left, lon, kd)
    raug, rkeys = _with_keys(right, ron, kd)
    nk = len(kd)
    raug = raug.sort([col(f"__key{i}") for i in range(nk)], descending=[False] * nk, nulls_first=[False] * nk)
    joined = _call_join("b200_asof_join_sorted", laug, lkeys, raug, rkeys)
    rbase = laug.num_columns()
    sel = [col(f"c{i}").alias(n) for i, n in enumerate(lnames)]
    sel += [col(f"c{rbase + i}").alias(n) for i, n in enumerate(rnames) if n not in drop]
    return joined.eval_expression_list(sel)


def cross_join(left: RecordBatch, right: RecordBatch, outer_loop_side: str = "left") -> RecordBatch:
    """Every pair of rows; the outer-loop side's rows stay in order, each repeated len(inner) times in a row."""
    if outer_loop_side not in ("left", "right"):
        raise ValueError(f"outer_loop_side must be 'left' or 'right', not {outer_loop_side!r}")
    clash = [n for n in right.column_names() if n in left.column_names()]
    if clash:
        raise ValueError(f"cross_join: duplicate output column name(s) {clash}")
    out = ctypes.c_void_p()
    _lib.call("b200_cross_join", left._h, right._h, ctypes.c_int32(1 if outer_loop_side == "left" else 0), ctypes.byref(out))
    return RecordBatch._wrap(out, names=left.column_names() + right.column_names(), dtypes=left.dtypes() + right.dtypes())

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

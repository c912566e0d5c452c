"""Python face of the CPU oracle (oracle/daft_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(daft_b200/) never imports this module.

Inputs and outputs are pyarrow Tables/Arrays so that parity tests can compare
the CUDA path and the oracle on literally the same Arrow data.  Semantics cited
per function (paths relative to /root/reference).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdaft_oracle.so")

# dtype codes: must match include/daft_b200.h
NULL, BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, DATE32, TIMESTAMP, LARGE_UTF8, LARGE_BINARY = range(16)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "daft_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libdaft_oracle.so"])
    return _SO


class _Col(ctypes.Structure):
    _fields_ = [
        ("dtype", ctypes.c_int32),
        ("_pad", ctypes.c_int32),
        ("length", ctypes.c_int64),
        ("validity", ctypes.c_void_p),
        ("offsets", ctypes.c_void_p),
        ("data", ctypes.c_void_p),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.orc_xxh3_64.restype = ctypes.c_uint64
        _lib.orc_xxh3_64.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_uint64]
        _lib.orc_group_ids.restype = ctypes.c_int64
        _lib.orc_join_probe.restype = ctypes.c_int64
    return _lib


def xxh3_64(b: bytes, seed: int = 0) -> int:
    return lib().orc_xxh3_64(b, len(b), seed)


_PA2CODE = {
    pa.null(): NULL, pa.bool_(): BOOL,
    pa.int8(): I8, pa.int16(): I16, pa.int32(): I32, pa.int64(): I64,
    pa.uint8(): U8, pa.uint16(): U16, pa.uint32(): U32, pa.uint64(): U64,
    pa.float32(): F32, pa.float64(): F64, pa.date32(): DATE32,
    pa.large_utf8(): LARGE_UTF8, pa.large_binary(): LARGE_BINARY,
}


def normalize(arr) -> pa.Array:
    """Daft holds strings/binary as Large (i64 offsets) and everything contiguous."""
    if isinstance(arr, pa.ChunkedArray):
        arr = arr.combine_chunks() if arr.num_chunks != 1 else arr.chunk(0)
    if pa.types.is_string(arr.type):
        arr = arr.cast(pa.large_utf8())
    elif pa.types.is_binary(arr.type):
        arr = arr.cast(pa.large_binary())
    if arr.offset != 0:
        arr = pa.concat_arrays([arr])  # re-materialise at offset 0
        if arr.offset != 0:
            arr = arr.take(pa.array(np.arange(len(arr), dtype=np.int64)))
    return arr


def dtype_code(t: pa.DataType) -> int:
    if pa.types.is_timestamp(t):
        if t.tz is not None:
            raise NotImplementedError("tz-aware timestamps")
        return TIMESTAMP
    return _PA2CODE[t]


class ColView:
    """Keeps numpy views alive while a C call uses the raw pointers."""

    def __init__(self, arr: pa.Array):
        arr = normalize(arr)
        self.arr = arr
        self.code = dtype_code(arr.type)
        bufs = arr.buffers()
        self.keep = []
        c = _Col()
        c.dtype = self.code
        c.length = len(arr)
        c.validity = c.offsets = c.data = None
        if self.code != NULL:
            if bufs[0] is not None and arr.null_count > 0:
                c.validity = bufs[0].address
            if self.code in (LARGE_UTF8, LARGE_BINARY):
                c.offsets = bufs[1].address
                c.data = bufs[2].address if bufs[2] is not None else None
            else:
                c.data = bufs[1].address if bufs[1] is not None else None
        self.c = c


def _cols(arrs: Sequence[pa.Array]):
    views = [ColView(a) for a in arrs]
    carr = (_Col * len(views))(*[v.c for v in views])
    return views, carr


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------- a1/a2
def hash_column(arr, seed=None) -> np.ndarray:
    """kernels::hashing::hash, XxHash3_64 (src/daft-core/src/kernels/hashing.rs:304)."""
    v = ColView(arr)
    out = np.empty(len(v.arr), dtype=np.uint64)
    sp = None
    if seed is not None:
        seed = np.ascontiguousarray(seed, dtype=np.uint64)
        assert len(seed) == len(v.arr)
        sp = _ptr(seed)
    rc = lib().orc_hash_column(ctypes.byref(v.c), sp, _ptr(out))
    if rc:
        raise ValueError(f"hash not implemented for {v.arr.type}")
    return out


def hash_rows(arrs: Sequence) -> np.ndarray:
    """RecordBatch::hash_rows (src/daft-recordbatch/src/ops/hash.rs:14-26)."""
    views, carr = _cols(arrs)
    out = np.empty(len(views[0].arr), dtype=np.uint64)
    rc = lib().orc_hash_rows(carr, ctypes.c_int(len(views)), _ptr(out))
    if rc:
        raise ValueError("hash_rows failed")
    return out


# --------------------------------------------------------------------------- a3
def partition_by_hash(tbl: pa.Table, keys: Sequence[str], num_partitions: int) -> list[pa.Table]:
    """RecordBatch::partition_by_hash (src/daft-recordbatch/src/ops/partition.rs:51-70):
    target = hash_rows(keys) % P; rows keep input order inside a partition."""
    h = hash_rows([tbl.column(k) for k in keys])
    tgt = (h % np.uint64(num_partitions)).astype(np.int64)
    out = []
    for p in range(num_partitions):
        idx = np.nonzero(tgt == p)[0]
        out.append(tbl.take(pa.array(idx, type=pa.int64())))
    return out


# --------------------------------------------------------------------------- a4
def group_ids(key_arrs: Sequence):
    """Phase 1 of agg_generic_hash_path (inline_agg.rs:748-824)."""
    views, carr = _cols(key_arrs)
    n = len(views[0].arr)
    gid = np.empty(max(n, 1), dtype=np.uint32)
    first = np.empty(max(n, 1), dtype=np.uint64)
    sizes = np.empty(max(n, 1), dtype=np.uint64)
    ng = lib().orc_group_ids(carr, ctypes.c_int(len(views)), _ptr(gid), _ptr(first), _ptr(sizes))
    if ng < 0:
        raise ValueError("group_ids failed")
    return gid[:n], first[:ng], sizes[:ng]


_SUM_FN = {I64: ("orc_sum_i64", np.int64), U64: ("orc_sum_u64", np.uint64), F32: ("orc_sum_f32", np.float32), F64: ("orc_sum_f64", np.float64)}
_NP = {I8: np.int8, I16: np.int16, I32: np.int32, I64: np.int64, U8: np.uint8, U16: np.uint16, U32: np.uint32,
       U64: np.uint64, F32: np.float32, F64: np.float64}
_SFX = {I8: "i8", I16: "i16", I32: "i32", I64: "i64", U8: "u8", U16: "u16", U32: "u32", U64: "u64", F32: "f32", F64: "f64"}


def _values_np(v: ColView) -> np.ndarray:
    bufs = v.arr.buffers()
    return np.frombuffer(bufs[1], dtype=_NP[v.code], count=len(v.arr)) if len(v.arr) else np.empty(0, _NP[v.code])


def _from_np(vals: np.ndarray, valid: np.ndarray, typ: pa.DataType) -> pa.Array:
    mask = None if valid.all() else ~valid.astype(bool)
    return pa.array(vals, type=typ, mask=mask)


def agg_sum(arr, gid: np.ndarray, ng: int) -> pa.Array:
    """SumAccum* (inline_agg.rs:126-198); widening per inline_agg.rs:424-460."""
    arr = normalize(arr)
    t = arr.type
    if pa.types.is_signed_integer(t):
        arr = arr.cast(pa.int64())
    elif pa.types.is_unsigned_integer(t):
        arr = arr.cast(pa.uint64())
    v = ColView(arr)
    if v.code not in _SUM_FN:
        raise ValueError(f"Expected input to sum() to be numeric, got {t}")
    fn, npt = _SUM_FN[v.code]
    out = np.zeros(max(ng, 1), dtype=npt)
    ov = np.zeros(max(ng, 1), dtype=np.uint8)
    vals = _values_np(v)
    getattr(lib(), fn)(_ptr(vals), ctypes.c_void_p(v.c.validity), _ptr(np.ascontiguousarray(gid)), ctypes.c_int64(len(v.arr)),
                       ctypes.c_int64(ng), _ptr(out), _ptr(ov))
    return _from_np(out[:ng], ov[:ng], v.arr.type)


def agg_minmax(arr, gid: np.ndarray, ng: int, which: str) -> pa.Array:
    """Min/MaxAccum* (inline_agg.rs:200-298); Boolean = bool_and / bool_or over the valid rows (array/ops/bool_agg.rs:19-60)."""
    arr0 = normalize(arr)
    if pa.types.is_boolean(arr0.type):
        return agg_minmax(arr0.cast(pa.uint8()), gid, ng, which).cast(pa.bool_())
    if pa.types.is_large_string(arr0.type) or pa.types.is_large_binary(arr0.type):
        # grouped_cmp_utf8 / binary (src/daft-core/src/array/ops/compare_agg.rs:99-134 and below): a left fold in row order
        # with bytewise comparison, nulls skipped, an all-null group -> null
        best = [None] * ng
        for g, x in zip(gid.tolist(), arr0.to_pylist()):
            if x is None:
                continue
            xb = x.encode() if isinstance(x, str) else x
            cur = best[g]
            if cur is None:
                best[g] = x
            else:
                cb = cur.encode() if isinstance(cur, str) else cur
                if (xb < cb) if which == "min" else (xb > cb):
                    best[g] = x
        return pa.array(best, type=arr0.type)
    v = ColView(arr)
    typ = v.arr.type
    code = {DATE32: I32, TIMESTAMP: I64}.get(v.code, v.code)
    npt = _NP[code]
    out = np.zeros(max(ng, 1), dtype=npt)
    ov = np.zeros(max(ng, 1), dtype=np.uint8)
    bufs = v.arr.buffers()
    vals = np.frombuffer(bufs[1], dtype=npt, count=len(v.arr)) if len(v.arr) else np.empty(0, npt)
    getattr(lib(), f"orc_{which}_{_SFX[code]}")(_ptr(vals), ctypes.c_void_p(v.c.validity), _ptr(np.ascontiguousarray(gid)),
                                                ctypes.c_int64(len(v.arr)), ctypes.c_int64(ng), _ptr(out), _ptr(ov))
    mask = None if ov[:ng].all() else ~ov[:ng].astype(bool)
    if code != v.code:
        return pa.array(out[:ng], mask=mask).cast(typ)
    return pa.array(out[:ng], type=typ, mask=mask)


def agg_count(arr, gid: np.ndarray, ng: int, mode: str = "valid") -> pa.Array:
    """CountAccum (inline_agg.rs:36-124); output UInt64."""
    v = ColView(arr)
    out = np.zeros(max(ng, 1), dtype=np.uint64)
    m = {"all": 0, "valid": 1, "null": 2}[mode]
    lib().orc_count(ctypes.c_void_p(v.c.validity), ctypes.c_int(int(v.code == NULL)), ctypes.c_int(m), _ptr(np.ascontiguousarray(gid)),
                    ctypes.c_int64(len(v.arr)), ctypes.c_int64(ng), _ptr(out))
    return pa.array(out[:ng], type=pa.uint64())


def agg(tbl: pa.Table, aggs: Sequence[tuple], group_by: Sequence) -> pa.Table:
    """RecordBatch::agg_groupby_inline (inline_agg.rs:862-928).

    aggs: (kind, values, out_name[, count_mode]) where values is a column name or a
    pyarrow array already evaluated from the child expression; group_by: column names
    or (name, array).  Output: key columns (take of first rows) ++ agg columns, groups
    in first-occurrence order.  group_by == [] -> one global group (agg.rs ungrouped)."""
    n = tbl.num_rows
    keys = [(g, tbl.column(g)) if isinstance(g, str) else g for g in group_by]
    if keys:
        gid, first, sizes = group_ids([k[1] for k in keys])
        ng = len(first)
    else:
        gid, first, ng = np.zeros(n, dtype=np.uint32), np.zeros(0, np.uint64), 1
    cols, names = [], []
    for name, arr in keys:
        cols.append(normalize(arr).take(pa.array(first.astype(np.int64))))
        names.append(name)
    for a in aggs:
        kind, values, out_name = a[0], a[1], a[2]
        arr = tbl.column(values) if isinstance(values, str) else values
        if kind == "sum":
            r = agg_sum(arr, gid, ng)
        elif kind in ("min", "max"):
            r = agg_minmax(arr, gid, ng, kind)
        elif kind == "count":
            r = agg_count(arr, gid, ng, a[3] if len(a) > 3 else "valid")
        elif kind == "mean":
            # two-stage rewrite: Sum / Count(Valid) as Float64 (src/daft-local-plan/src/agg.rs:156-165)
            s = agg_sum(arr, gid, ng).cast(pa.float64())
            c = agg_count(arr, gid, ng, "valid")
            sn = s.to_numpy(zero_copy_only=False).astype(np.float64)
            cn = c.to_numpy().astype(np.float64)
            with np.errstate(all="ignore"):
                q = sn / cn
            r = pa.array(q, mask=(cn == 0))
        elif kind == "count_distinct":
            # distinct NON-NULL values per group, UInt64 (src/daft-functions-list/src/series.rs:37-40,234-247), by definition
            vals = normalize(arr).to_pylist()
            seen = [set() for _ in range(ng)]
            for g, v in zip(gid.tolist(), vals):
                if v is not None:
                    seen[g].add(v)
            r = pa.array([len(s_) for s_ in seen], type=pa.uint64())
        elif kind == "skew":
            # calculate_skew (src/daft-core/src/utils/stats.rs:134-147): (m3 / n) / sqrt((m2 / n)^3) around the group's mean; the planner's
            # staged form (src/daft-local-plan/src/agg.rs:302-348) computes the same quantity from power sums
            vals = normalize(arr).cast(pa.float64()).to_pylist()
            per = [[] for _ in range(ng)]
            for g, v in zip(gid.tolist(), vals):
                if v is not None:
                    per[g].append(v)
            out = []
            for vs in per:
                if not vs:
                    out.append(None)
                    continue
                mean = sum(vs) / len(vs)
                m3 = sum((v - mean) ** 3 for v in vs)
                m2 = sum((v - mean) ** 2 for v in vs)
                with np.errstate(all="ignore"):
                    out.append(float(np.float64(m3 / len(vs)) / np.sqrt(np.float64(m2 / len(vs)) ** 3)))
            r = pa.array(out, type=pa.float64())
        elif kind == "percentile":
            # exact_percentile (src/daft-core/src/utils/stats.rs:31-59) over the group's non-null values cast to Float64: sort by
            # f64::total_cmp, rank = p (n - 1), (upper - lower).mul_add(rank - floor(rank), lower)
            import math
            p_ = float(a[3])
            vals = normalize(arr).cast(pa.float64()).to_pylist()
            per = [[] for _ in range(ng)]
            for g, v in zip(gid.tolist(), vals):
                if v is not None:
                    per[g].append(v)
            out = []
            for vs in per:
                if not vs:
                    out.append(None)
                    continue
                vs.sort(key=lambda v: (2, 0.0) if v != v else (1, v))   # a (positive) NaN is the greatest under total_cmp
                rank = p_ * (len(vs) - 1)
                lo, hi = int(math.floor(rank)), int(math.ceil(rank))
                out.append(vs[lo] if lo == hi else math.fma(vs[hi] - vs[lo], rank - lo, vs[lo]) if hasattr(math, "fma") else (vs[hi] - vs[lo]) * (rank - lo) + vs[lo])
            r = pa.array(out, type=pa.float64())
        elif kind == "any_value":
            # Series::any_value (src/daft-core/src/series/ops/agg.rs:353-382): the first row of each group, or with
            # ignore_nulls its first non-null row (null when there is none); ungrouped: row 0 / first valid row / null
            arr_n = normalize(arr)
            ignore = (a[3] if len(a) > 3 else "first") == "ignore_nulls"
            valid = np.ones(n, dtype=bool) if arr_n.null_count == 0 else ~np.asarray(arr_n.is_null())
            if pa.types.is_null(arr_n.type):
                idx = [None] * ng
            elif keys:
                idx = [None] * ng
                if ignore:
                    for row in np.flatnonzero(valid)[::-1]:
                        idx[int(gid[row])] = int(row)
                else:
                    idx = [int(x) for x in first]
            else:
                cand = np.flatnonzero(valid) if ignore else np.arange(n)
                idx = [int(cand[0]) if len(cand) else None]
            r = arr_n.take(pa.array(idx, type=pa.int64()))
        elif kind in ("list", "set", "concat"):
            # grouped_list (src/daft-core/src/array/ops/list_agg.rs:30-57): each group's values in row order, nulls kept;
            # grouped_set (src/daft-core/src/series/ops/agg.rs:531-575): the group's distinct NON-NULL values in first-occurrence
            # order (probe table without nulls); grouped_concat over Utf8 (src/daft-core/src/array/ops/concat_agg.rs:138-161): the
            # valid strings back to back, null when the column has nulls and the group has no valid value.  Ungrouped: one group.
            arr_n = normalize(arr)
            vals = arr_n.to_pylist()
            rows_of = [[] for _ in range(ng)]
            for row, g in enumerate(gid.tolist()):
                rows_of[g].append(row)
            if kind == "list":
                r = pa.array([[vals[i] for i in rows] for rows in rows_of], type=pa.large_list(arr_n.type))
            elif kind == "set":
                out = []
                for rows in rows_of:
                    seen, lst = set(), []
                    for i in rows:
                        v = vals[i]
                        if v is None:
                            continue
                        # the probe table matches on (XXH3 of the value's bytes, Literal ==): 0.0 and -0.0 hash apart, and a NaN
                        # equals nothing, itself included (src/daft-core/src/series/mod.rs:85-100)
                        if isinstance(v, float):
                            if v != v:
                                lst.append(v)
                                continue
                            k_ = ("f", np.float64(v).tobytes() if pa.types.is_float64(arr_n.type) else np.float32(v).tobytes())
                        else:
                            k_ = v
                        if k_ not in seen:
                            seen.add(k_)
                            lst.append(v)
                    out.append(lst)
                r = pa.array(out, type=pa.large_list(arr_n.type))
            else:
                if not (pa.types.is_large_string(arr_n.type) or pa.types.is_string(arr_n.type)):
                    raise NotImplementedError("concat over list columns")
                has_nulls = arr_n.null_count > 0
                out = []
                for rows in rows_of:
                    parts = [vals[i] for i in rows if vals[i] is not None]
                    out.append(None if (has_nulls and not parts and rows) else "".join(parts))
                r = pa.array(out, type=pa.large_utf8())
        else:
            raise NotImplementedError(kind)
        cols.append(r)
        names.append(out_name)
    return pa.table(cols, names=names)


# --------------------------------------------------------------------------- a7/a8
def join_matches(build_keys: Sequence, probe_keys: Sequence, null_equals_null=None):
    """ProbeTable::add_table + probe (probe_table.rs:56-185) -> CSR (offsets, build_rows)."""
    bv, bc = _cols(build_keys)
    pv, pc_ = _cols(probe_keys)
    nk = len(bv)
    ne = np.array([1 if x else 0 for x in (null_equals_null or [False] * nk)], dtype=np.uint8)
    npb = len(pv[0].arr)
    offs = np.empty(npb + 1, dtype=np.int64)
    cap = max(1024, 2 * (len(bv[0].arr) + npb))
    while True:
        rows = np.empty(cap, dtype=np.int64)
        total = lib().orc_join_probe(bc, pc_, ctypes.c_int(nk), _ptr(ne), _ptr(offs), _ptr(rows), ctypes.c_int64(cap))
        if total < 0:
            raise ValueError("join probe failed")
        if total <= cap:
            return offs, rows[:total]
        cap = int(total)


def _take_nullable(tbl: pa.Table, idx: np.ndarray) -> pa.Table:
    """take with -1 -> null row."""
    mask = idx < 0
    arr = pa.array(np.where(mask, 0, idx), type=pa.int64(), mask=mask if mask.any() else None)
    if tbl.num_rows == 0:
        return pa.table([pa.nulls(len(idx), f.type) for f in tbl.schema], schema=tbl.schema)
    return tbl.take(arr)


def hash_join(left: pa.Table, right: pa.Table, left_on: Sequence[str], right_on: Sequence[str], how: str = "inner",
              null_equals_null=None, build_on_left: bool | None = None) -> pa.Table:
    """HashJoinOperator semantics (src/daft-local-execution/src/join/*.rs), one probe batch.

    Output schema: common-name join keys (from left), left non-key columns, right
    non-key columns (inner_join.rs:50-78).  Row order: probe rows ascending, matches in
    build insertion order; unmatched build rows (left/right/outer) appended at the end
    (left_right_join.rs:173-269, outer_join.rs:17-55).  semi/anti output = left columns."""
    if build_on_left is None:
        build_on_left = how in ("right",)  # left/semi/anti/inner default: build on right
    lk = [left.column(k) for k in left_on]
    rk = [right.column(k) for k in right_on]
    common = [l for l, r in zip(left_on, right_on) if l == r]
    if how in ("semi", "anti"):
        offs, _ = join_matches(rk, lk, null_equals_null)
        cnt = np.diff(offs)
        keep = np.nonzero(cnt > 0 if how == "semi" else cnt == 0)[0]
        return left.take(pa.array(keep, type=pa.int64()))
    if build_on_left:
        offs, brow = join_matches(lk, rk, null_equals_null)
        n_probe, n_build = right.num_rows, left.num_rows
    else:
        offs, brow = join_matches(rk, lk, null_equals_null)
        n_probe, n_build = left.num_rows, right.num_rows
    cnt = np.diff(offs)
    probe_idx = np.repeat(np.arange(n_probe, dtype=np.int64), cnt)
    build_idx = brow.astype(np.int64)
    probe_is_left = not build_on_left
    keep_probe_unmatched = (how == "outer") or (how == "left" and probe_is_left) or (how == "right" and not probe_is_left)
    keep_build_unmatched = (how == "outer") or (how == "left" and not probe_is_left) or (how == "right" and probe_is_left)
    if keep_probe_unmatched:
        # unmatched probe rows are emitted inline, at their probe position
        cnt2 = np.maximum(cnt, 1)
        probe_idx = np.repeat(np.arange(n_probe, dtype=np.int64), cnt2)
        build_full = np.full(int(cnt2.sum()), -1, dtype=np.int64)
        starts = np.concatenate([[0], np.cumsum(cnt2)[:-1]]) if n_probe else np.zeros(0, np.int64)
        pos = np.repeat(starts, cnt) + (np.arange(len(brow)) - np.repeat(offs[:-1], cnt))
        build_full[pos.astype(np.int64)] = brow
        build_idx = build_full
    if keep_build_unmatched:
        matched = np.zeros(n_build, dtype=bool)
        matched[brow] = True
        un = np.nonzero(~matched)[0]
        probe_idx = np.concatenate([probe_idx, np.full(len(un), -1, dtype=np.int64)])
        build_idx = np.concatenate([build_idx, un])
    lidx, ridx = (build_idx, probe_idx) if build_on_left else (probe_idx, build_idx)
    lt = _take_nullable(left, lidx)
    rt = _take_nullable(right, ridx)
    cols, names = [], []
    for k in common:
        lc, rc = lt.column(k).combine_chunks(), rt.column(k).combine_chunks()
        if how in ("right", "outer"):
            # key column coalesces left/right so null-extended rows still carry the key
            import pyarrow.compute as pc
            lc = pc.coalesce(lc, rc.cast(lc.type))
        cols.append(lc)
        names.append(k)
    for name in left.column_names:
        if name not in common:
            cols.append(lt.column(name)); names.append(name)
    for name in right.column_names:
        if name not in common:
            cols.append(rt.column(name)); names.append(name)
    return pa.table(cols, names=names)


# --------------------------------------------------------------------------- a9
def argsort(key_arrs: Sequence, descending=None, nulls_first=None) -> np.ndarray:
    """RecordBatch::argsort (src/daft-recordbatch/src/ops/sort.rs:7-45); canonical
    tie-break = ascending row index (the reference sort is unstable)."""
    views, carr = _cols(key_arrs)
    nk = len(views)
    desc = np.array([1 if x else 0 for x in (descending or [False] * nk)], dtype=np.uint8)
    nf = np.array([1 if x else 0 for x in (nulls_first if nulls_first is not None else list(desc))], dtype=np.uint8)
    out = np.empty(len(views[0].arr), dtype=np.int64)
    lib().orc_argsort(carr, ctypes.c_int(nk), _ptr(desc), _ptr(nf), _ptr(out))
    return out


def sort(tbl: pa.Table, keys: Sequence[str], descending=None, nulls_first=None) -> pa.Table:
    idx = argsort([tbl.column(k) for k in keys], descending, nulls_first)
    return tbl.take(pa.array(idx, type=pa.int64()))


# --------------------------------------------------------------------------- a12 (range repartition)
def _cmp_values(a, b, descending: bool) -> int:
    """make_daft_comparator with SortOptions::new(descending, nulls_first=descending)
    (src/daft-core/src/kernels/search_sorted.rs:401-406): NaN is the greatest float (cmp_float :238-245)."""
    if a is None or b is None:
        if a is None and b is None:
            return 0
        r = -1 if a is None else 1          # nulls_first == descending: nulls are "largest" ascending, and first descending
        return r if descending else -r
    an = isinstance(a, float) and a != a
    bn = isinstance(b, float) and b != b
    if an or bn:
        r = int(an) - int(bn)
    else:
        r = -1 if a < b else (1 if a > b else 0)
    return -r if descending else r


def search_sorted(sorted_tbl: pa.Table, keys_tbl: pa.Table, descending: Sequence[bool]) -> np.ndarray:
    """RecordBatch::search_sorted (src/daft-recordbatch/src/ops/search_sorted.rs:13-36; binary search of
    src/daft-core/src/kernels/search_sorted.rs:420-433): rows of the sorted table that compare <= each key row."""
    srows = list(zip(*[normalize(c).to_pylist() for c in sorted_tbl.columns])) if sorted_tbl.num_columns else []
    krows = list(zip(*[normalize(c).to_pylist() for c in keys_tbl.columns]))
    out = np.zeros(len(krows), dtype=np.uint64)

    def cmp(sr, kr):
        for a, b, d in zip(sr, kr, descending):
            r = _cmp_values(a, b, d)
            if r:
                return r
        return 0
    for i, kr in enumerate(krows):
        left, right = 0, len(srows)
        while left < right:
            mid = left + ((right - left) >> 1)
            if cmp(srows[mid], kr) <= 0:
                left = mid + 1
            else:
                right = mid
        out[i] = left
    return out


def partition_by_range(tbl: pa.Table, keys: Sequence[str], boundaries: pa.Table, descending: Sequence[bool]) -> list[pa.Table]:
    """RecordBatch::partition_by_range (src/daft-recordbatch/src/ops/partition.rs:88-101)."""
    if boundaries.num_rows == 0:
        return [tbl]
    target = search_sorted(boundaries, tbl.select(list(keys)), descending)
    return [tbl.filter(pa.array(target == p)) for p in range(boundaries.num_rows + 1)]

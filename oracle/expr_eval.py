"""CPU oracle for Filter / Project: column-at-a-time numpy evaluation of daft_b200 Expression trees.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  Follows the reference's tree-walking interpreter
(src/daft-recordbatch/src/lib.rs:1280-1400): every AST node materialises a full column (values +
validity mask); binary operands are first cast to their supertype; IEEE f64 arithmetic is elementwise
(numpy float64 == Rust f64, no fused multiply-add), so results are bit-exact comparable.
"""
from __future__ import annotations

import datetime
import re

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc

from daft_b200 import datatypes as dt
from daft_b200.expressions import Bound, Expression, bind, T_COL, T_LIT, T_BINARY, T_UNARY, T_CAST, T_IF_ELSE, T_IS_IN, \
    T_BETWEEN, T_FILL_NULL, T_FUNCTION, BINOPS, UNOPS, FUNCS

_NP = {dt.BOOL: np.bool_, dt.I8: np.int8, dt.I16: np.int16, dt.I32: np.int32, dt.I64: np.int64, dt.U8: np.uint8,
       dt.U16: np.uint16, dt.U32: np.uint32, dt.U64: np.uint64, dt.F32: np.float32, dt.F64: np.float64,
       dt.DATE32: np.int32, dt.TIMESTAMP: np.int64, dt.NULL: np.bool_}
_BIN = {v: k for k, v in BINOPS.items()}
_UN = {v: k for k, v in UNOPS.items()}
_FN = {v: k for k, v in FUNCS.items()}


def _col_to_np(arr: pa.Array, code: int):
    n = len(arr)
    valid = np.ones(n, dtype=bool) if arr.null_count == 0 else np.asarray(arr.is_valid())
    if code == dt.NULL:
        return np.zeros(n, dtype=bool), np.zeros(n, dtype=bool)
    if dt.is_string(code):
        vals = np.array([x if x is not None else (b"" if code == dt.LARGE_BINARY else "") for x in arr.to_pylist()], dtype=object)
        if code == dt.LARGE_UTF8:
            vals = np.array([v.encode() if isinstance(v, str) else v for v in vals], dtype=object)
        return vals, valid
    if code == dt.DATE32:
        vals = np.asarray(arr.cast(pa.int32()).fill_null(0))
    elif code == dt.TIMESTAMP:
        vals = np.asarray(arr.cast(pa.int64()).fill_null(0))
    elif code == dt.BOOL:
        vals = np.asarray(arr.fill_null(False))
    else:
        vals = np.asarray(arr.fill_null(0))
    return vals.astype(_NP[code], copy=False), valid


def _lit_np(b: Bound, n: int):
    v, d = b.value, b.dtype
    if v is None or d == dt.NULL:
        if dt.is_string(d):
            return np.array([b""] * n, dtype=object), np.zeros(n, dtype=bool)
        return np.zeros(n, dtype=_NP.get(d, np.bool_)), np.zeros(n, dtype=bool)
    if dt.is_string(d):
        raw = v.encode() if isinstance(v, str) else bytes(v)
        return np.array([raw] * n, dtype=object), np.ones(n, dtype=bool)
    if d == dt.DATE32 and isinstance(v, datetime.date):
        v = (v - datetime.date(1970, 1, 1)).days
    if d == dt.TIMESTAMP and isinstance(v, datetime.datetime):
        v = int((v - datetime.datetime(1970, 1, 1)).total_seconds() * 1_000_000)
    return np.full(n, v, dtype=_NP[d]), np.ones(n, dtype=bool)


def _civil(days: np.ndarray):
    z = days.astype(np.int64) + 719468
    era = np.where(z >= 0, z, z - 146096) // 146097
    doe = z - era * 146097
    yoe = (doe - doe // 1460 + doe // 36524 - doe // 146096) // 365
    y = yoe + era * 400
    doy = doe - (365 * yoe + yoe // 4 - yoe // 100)
    mp = (5 * doy + 2) // 153
    d = doy - (153 * mp + 2) // 5 + 1
    m = np.where(mp < 10, mp + 3, mp - 9)
    return y + (m <= 2), m, d


_INT_TEXT = re.compile(rb"[+-]?[0-9]+")
# Rust `char::is_whitespace` (Unicode White_Space), which `str::trim` removes
_RUST_SPACE = "\t\n\x0b\x0c\r \x85\xa0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000"


def _parse_ints(a, ka, to):
    """Utf8 -> integer: `.trim()` (src/daft-core/src/array/ops/cast.rs:164-173), then arrow-cast 57.3.0's `parser_primitive!` over atoi 2.0
    `from_radix_10_signed_checked`: optional sign, decimal digits to the end, null when malformed or out of range (cast.rs:2088-2150)."""
    lo, hi = dt.int_range(to)
    out, ok = np.zeros(len(a), dtype=_NP[to]), np.zeros(len(a), dtype=bool)
    for i, (s, k) in enumerate(zip(a, ka)):
        if not k:
            continue
        t = bytes(s).decode("utf-8", "surrogateescape").strip(_RUST_SPACE).encode("utf-8", "surrogateescape")
        if _INT_TEXT.fullmatch(t) and lo <= int(t) <= hi:
            out[i], ok[i] = int(t), True
    return out, ok


def _total_order_key(a: np.ndarray) -> np.ndarray:
    """Floats compare by IEEE 754 totalOrder, as the arrow-rs `cmp` kernels the reference calls do
    (src/daft-core/src/array/ops/comparison.rs:541-603 -> arrow::compute::kernels::cmp, arrow-ord 57.3.0): -NaN < -inf < ... <
    -0.0 < +0.0 < ... < +inf < +NaN, and equality is equality of the bit patterns (NaN == NaN, -0.0 != 0.0).  The key is the
    sign-magnitude bit pattern mapped to a two's-complement integer with the same order (what f64::total_cmp does)."""
    if a.dtype == np.float64:
        bits = a.view(np.int64)
        return bits ^ ((bits >> 63) & np.int64(0x7FFFFFFFFFFFFFFF))
    if a.dtype == np.float32:
        bits = a.view(np.int32)
        return bits ^ ((bits >> 31) & np.int32(0x7FFFFFFF))
    return a


def _eval(b: Bound, cols, n: int):
    if b.tag == T_COL:
        return cols[b.col]
    if b.tag == T_LIT:
        return _lit_np(b, n)
    kids = [_eval(k, cols, n) for k in b.kids]
    with np.errstate(all="ignore"):
        if b.tag == T_BINARY:
            (a, ka), (c, kc) = kids
            op = _BIN[b.op]
            t = b.kids[0].dtype if b.kids[0].dtype != dt.NULL else b.kids[1].dtype
            both = ka & kc
            if op in ("add", "sub", "mul"):
                f = {"add": np.add, "sub": np.subtract, "mul": np.multiply}[op]
                return f(a, c), both
            if op == "true_div":
                return a / c, both
            if op in ("floor_div", "mod"):
                # Rust semantics (src/daft-core/src/array/ops/arithmetic.rs:245-300, series/ops/arithmetic.rs:456-466): integer `//` truncates
                # toward zero, `%` takes the sign of the dividend (floats: fmod); float `//` = floor(l / r).  x // 0, x % 0 on
                # integers: null here (the reference panics)
                with np.errstate(all="ignore"):
                    if dt.is_float(t):
                        return (np.floor(a / c) if op == "floor_div" else np.fmod(a, c)), both
                    ok = both & (c != 0)
                    cs = np.where(c == 0, 1, c)
                    rem = np.fmod(a, cs)
                    r = ((a - rem) // cs) if op == "floor_div" else rem
                return np.where(ok, r, 0).astype(a.dtype), ok
            if op in ("lt", "le", "ge", "gt", "eq", "ne"):
                f = {"lt": np.less, "le": np.less_equal, "ge": np.greater_equal, "gt": np.greater, "eq": np.equal, "ne": np.not_equal}[op]
                return np.asarray(f(_total_order_key(a), _total_order_key(c)), dtype=bool), both
            if op == "eq_null_safe":
                return np.where(both, np.asarray(_total_order_key(a) == _total_order_key(c), dtype=bool), ka == kc), np.ones(n, dtype=bool)
            if op in ("and", "or", "xor") and dt.is_integer(t):  # bitwise on integers, null if either side is null (array/ops/bitwise.rs)
                f = {"and": np.bitwise_and, "or": np.bitwise_or, "xor": np.bitwise_xor}[op]
                return f(a, c), both
            if op == "and":  # Kleene, src/daft-core/src/array/ops/comparison.rs:725-786
                va, vc = ka & a.astype(bool), kc & c.astype(bool)
                return va & vc, both | (ka & ~a.astype(bool)) | (kc & ~c.astype(bool))
            if op == "or":
                va, vc = ka & a.astype(bool), kc & c.astype(bool)
                return va | vc, both | va | vc
            if op == "xor":
                return a.astype(bool) != c.astype(bool), both
            if op == "shl":
                return np.left_shift(a, c), both
            if op == "shr":
                return np.right_shift(a, c), both
        if b.tag == T_UNARY:
            (a, ka), = kids
            op = _UN[b.op]
            if op == "not": return ~a.astype(bool), ka
            if op == "neg":  # Rust `-v` (array/ops/sign.rs:14-16): IEEE negation for floats (-(+0.0) = -0.0), wrapping for integers
                return (np.negative(a) if a.dtype.kind == "f" else (0 - a).astype(a.dtype)), ka
            if op == "abs": return np.abs(a), ka
            if op == "is_null": return ~ka, np.ones(n, dtype=bool)
            if op == "not_null": return ka.copy(), np.ones(n, dtype=bool)
        if b.tag == T_CAST:
            (a, ka), = kids
            src = b.kids[0].dtype
            if src == dt.NULL:
                return np.zeros(n, dtype=_NP[b.dtype]) if not dt.is_string(b.dtype) else np.array([b""] * n, dtype=object), np.zeros(n, dtype=bool)
            if dt.is_string(b.dtype):
                return a, ka
            if dt.is_string(src):
                return _parse_ints(a, ka, b.dtype)
            if b.dtype == dt.BOOL:
                return a != 0, ka
            if dt.int_like(b.dtype) and (dt.is_float(src) or dt.int_like(src)):
                # arrow-rs `cast` with its default CastOptions { safe: true } (src/daft-core/src/array/ops/cast.rs:194-196; arrow-cast
                # 57.3.0, Cargo.lock:298): a value the target integer type cannot hold becomes null.  Floats convert iff the
                # value truncated toward zero is in range (num-traits 0.2.19 ToPrimitive; NaN fails).
                lo, hi = dt.int_range(b.dtype)
                with np.errstate(all="ignore"):
                    if dt.is_float(src):
                        t = np.trunc(a.astype(np.float64))
                        ok = (t >= float(lo)) & (t < float(hi + 1))
                        v = np.where(ok, a, 0).astype(_NP[b.dtype])
                    else:
                        v = a.astype(_NP[b.dtype])
                        ok = (v.astype(a.dtype) == a) & ((v < 0) == (a < 0))
                return np.where(ok, v, 0).astype(_NP[b.dtype]), ka & ok
            return a.astype(_NP[b.dtype]), ka
        if b.tag == T_IF_ELSE:
            (p, kp), (x, kx), (y, ky) = kids
            s = kp & p.astype(bool)
            return np.where(s, x, y), kp & np.where(s, kx, ky)
        if b.tag == T_IS_IN:
            (a, ka) = kids[0]
            r = np.zeros(n, dtype=bool)
            for (v, kv), kb in zip(kids[1:], b.kids[1:]):
                if kb.value is None or kb.dtype == dt.NULL:
                    continue
                r |= np.asarray(a == v, dtype=bool)
                if dt.is_float(b.kids[0].dtype):
                    # float membership goes through a BTreeSet<FloatWrapper>: every NaN equals every NaN, -0.0 equals 0.0
                    # (src/daft-core/src/array/ops/is_in.rs:44-66, src/common/hashable-float-wrapper/src/lib.rs:45-57)
                    r |= np.isnan(a) & np.isnan(v)
            return r, ka
        if b.tag == T_BETWEEN:
            (a, ka), (lo, kl), (hi, kh) = kids
            # gte(lower) Kleene-and lte(upper) (src/daft-core/src/array/ops/between.rs:30-32, comparison.rs:725-786)
            a, lo, hi = _total_order_key(a), _total_order_key(lo), _total_order_key(hi)
            ge, le = np.asarray(a >= lo, dtype=bool), np.asarray(a <= hi, dtype=bool)
            kg, kle = ka & kl, ka & kh
            return ge & le, (kg & kle) | (kg & ~ge) | (kle & ~le)
        if b.tag == T_FILL_NULL:
            (a, ka), (f, kf) = kids
            return np.where(ka, a, f), ka | kf
        if b.tag == T_FUNCTION:
            fn = _FN[b.op]
            allk = np.ones(n, dtype=bool)
            for _, kk in kids:
                allk &= kk
            a = kids[0][0]
            if fn in ("year", "month", "day"):
                y, m, d = _civil(a)
                return {"year": y.astype(np.int32), "month": m.astype(np.uint32), "day": d.astype(np.uint32)}[fn], allk
            if fn in ("hour", "minute", "second"):   # time of day of a Timestamp(us), floor semantics for instants before the epoch
                us = a.astype(np.int64)
                mod, div = {"hour": (86400000000, 3600000000), "minute": (3600000000, 60000000), "second": (60000000, 1000000)}[fn]
                return (np.mod(us, mod) // div).astype(np.uint32), allk
            if fn == "ts_date":
                return np.floor_divide(a.astype(np.int64), 86400000000).astype(np.int32), allk
            if fn == "ts_trunc":
                us, unit = a.astype(np.int64), kids[1][0].astype(np.int64)
                return us - np.mod(us, unit), kids[0][1]
            if fn in ("exp", "ln", "log2", "log10", "sin", "cos", "tan", "cbrt"):
                f = {"exp": np.exp, "ln": np.log, "log2": np.log2, "log10": np.log10, "sin": np.sin, "cos": np.cos, "tan": np.tan, "cbrt": np.cbrt}[fn]
                with np.errstate(all="ignore"):
                    return f(a), allk   # `a` is already Float32 / Float64
            if fn == "sqrt":
                with np.errstate(all="ignore"):
                    return np.sqrt(a), allk   # `a` is already Float32 / Float64 (the bound plan casts to the floating representation)
            if fn == "is_nan": return np.isnan(a), allk           # array/ops/float.rs:13-110
            if fn == "not_nan": return ~np.isnan(a), allk
            if fn == "is_inf": return np.isinf(a), allk
            if fn == "ceil": return np.ceil(a), allk               # array/ops/ceil.rs
            if fn == "floor": return np.floor(a), allk
            if fn == "sign":                                       # array/ops/sign.rs: signum / sign_unsigned
                if a.dtype.kind == "f":
                    return np.where(np.isnan(a), a, np.copysign(np.ones_like(a), a)), allk
                return np.sign(a).astype(a.dtype), allk
            if fn == "round":                                      # array/ops/round.rs:14-35: (v * 10^d).round() / 10^d, half away from zero
                d = int(kids[1][0][0]) if n else 0
                x = a.astype(np.float64)
                with np.errstate(all="ignore"):
                    y = x if d == 0 else x * (10.0 ** d)
                    r = np.where(np.isfinite(y), np.copysign(np.floor(np.abs(y) + 0.5), y), y)
                    # floor(|y| + 0.5) double-rounds when |y| + 0.5 is inexact (0.49999999999999994): f64::round does not
                    r = np.where(np.abs(y) < 0.5, np.copysign(0.0, y), r)
                    r = np.where(np.abs(y) >= 2.0 ** 52, y, r)
                    out = r if d == 0 else r / (10.0 ** d)
                return out.astype(a.dtype), allk
            if fn == "match_seq":
                flags = int(kids[1][0][0]) if n else 0
                segs = [kids[j][0][0] if n else b"" for j in range(2, len(kids))]

                def hit(x: bytes) -> bool:
                    pos = 0
                    for j, sg in enumerate(segs):
                        last = j == len(segs) - 1
                        if j == 0 and (flags & 1):
                            if not x.startswith(sg):
                                return False
                            f = 0
                        elif last and (flags & 2):
                            f = len(x) - len(sg)
                            if f < pos or not x.endswith(sg):
                                return False
                        else:
                            f = x.find(sg, pos)
                            if f < 0:
                                return False
                        pos = f + len(sg)
                    if len(segs) == 1 and (flags & 3) == 3:
                        return x == segs[0]
                    if not segs:
                        return (flags & 3) != 3 or len(x) == 0
                    return True
                return np.array([hit(x) for x in a], dtype=bool), allk
            if fn == "day_of_week":
                return ((a.astype(np.int64) + 3) % 7).astype(np.uint32), allk
            if fn == "utf8_length":
                return np.array([len(x.decode("utf-8", "ignore")) for x in a], dtype=np.uint64), allk
            if fn in ("lstrip", "rstrip"):   # Rust's trim_start / trim_end: the Unicode White_Space property (not Python's isspace, which adds U+001C..001F)
                ws = set("\t\n\x0b\x0c\r \x85\xa0\u1680\u2028\u2029\u202f\u205f\u3000") | {chr(c) for c in range(0x2000, 0x200B)}
                chars = "".join(ws)
                return np.array([(x.decode().lstrip(chars) if fn == "lstrip" else x.decode().rstrip(chars)).encode() for x in a], dtype=object), allk
            pat = kids[1][0]
            if fn == "startswith": return np.array([x.startswith(p) for x, p in zip(a, pat)], dtype=bool), allk
            if fn == "endswith": return np.array([x.endswith(p) for x, p in zip(a, pat)], dtype=bool), allk
            if fn == "contains": return np.array([p in x for x, p in zip(a, pat)], dtype=bool), allk
            if fn in ("rx_match", "rx_extract", "rx_replace_group"):
                # Python's `re` as the restatement of the regex crate's leftmost-first search (daft_b200/regex.py to_python translates the
                # supported subset: `$` only at the very end, `.` without newline)
                import re
                from daft_b200 import regex as _rx
                pattern = kids[2][0][0].decode() if len(kids[2][0]) else ""
                rp = re.compile(_rx.to_python(pattern))
                group = int(b.value) if getattr(b, "value", None) is not None else 0
                texts = [x.decode("utf-8", "ignore") for x in a]
                if fn == "rx_match":
                    return np.array([rp.search(t) is not None for t in texts], dtype=bool), kids[0][1]
                out, ok = [], []
                for t in texts:
                    m = rp.search(t)
                    if fn == "rx_extract":
                        good = m is not None and m.span(group) != (-1, -1)
                        out.append(m.group(group).encode() if good else b"")
                        ok.append(good)
                    else:
                        out.append(t.encode() if m is None else (m.group(group) or "").encode())
                        ok.append(True)
                return np.array(out, dtype=object), (kids[0][1] & np.array(ok, dtype=bool)) if n else kids[0][1]
            if fn == "substr":   # substring(), src/daft-functions-utf8/src/substr.rs:269-294
                start, ln = int(kids[1][0][0]) if len(kids[1][0]) else 0, int(kids[2][0][0]) if len(kids[2][0]) else -1
                out, ok = [], []
                for x in a:
                    chars = x.decode("utf-8", "ignore")
                    good = start < len(chars) and ln != 0
                    ok.append(good)
                    out.append((chars[start:] if ln < 0 else chars[start:start + ln]).encode() if good else b"")
                return np.array(out, dtype=object), kids[0][1] & np.array(ok, dtype=bool) if n else kids[0][1]
            if fn == "right":
                return np.array([(x.decode("utf-8", "ignore")[-int(k):] if int(k) > 0 else "").encode() for x, k in zip(a, pat)], dtype=object), allk
            if fn == "find":
                return np.array([x.find(p) for x, p in zip(a, pat)], dtype=np.int64), allk
            if fn == "left":
                return np.array([x.decode("utf-8", "ignore")[:max(int(k), 0)].encode() for x, k in zip(a, pat)], dtype=object), allk
    raise NotImplementedError(b.tag)


def _to_arrow(vals, valid, code: int) -> pa.Array:
    mask = None if valid.all() else ~valid
    if code == dt.NULL:
        return pa.nulls(len(vals))
    if code == dt.LARGE_UTF8:
        return pa.array([v.decode() if k else None for v, k in zip(vals, valid)], type=pa.large_utf8())
    if code == dt.LARGE_BINARY:
        return pa.array([bytes(v) if k else None for v, k in zip(vals, valid)], type=pa.large_binary())
    if code == dt.DATE32:
        return pa.array(np.asarray(vals, dtype=np.int32), mask=mask).cast(pa.date32())
    if code == dt.TIMESTAMP:
        return pa.array(np.asarray(vals, dtype=np.int64), mask=mask).cast(pa.timestamp("us"))
    return pa.array(np.asarray(vals, dtype=_NP[code]), type=dt.to_arrow(code), mask=mask)


class _LazyCols:
    """Columns are converted to numpy only when an expression actually references them."""

    def __init__(self, tbl: pa.Table, codes):
        self.tbl, self.codes, self.cache = tbl, codes, {}

    def __getitem__(self, i):
        if i not in self.cache:
            c = self.tbl.column(i)
            c = c.combine_chunks() if c.num_chunks != 1 else c.chunk(0)
            self.cache[i] = _col_to_np(c, self.codes[i])
        return self.cache[i]


def _table_cols(tbl: pa.Table):
    names = tbl.column_names
    codes = [dt.from_arrow(f.type) for f in tbl.schema]
    return names, codes, _LazyCols(tbl, codes)


def eval_expression(tbl: pa.Table, e: Expression) -> pa.Array:
    names, codes, cols = _table_cols(tbl)
    b = bind(e, names, codes)
    if b.tag == T_COL:  # bare column reference: no copy (Expr::Column, lib.rs:1290)
        c = tbl.column(b.col)
        return c.combine_chunks() if c.num_chunks != 1 else c.chunk(0)
    vals, valid = _eval(b, cols, tbl.num_rows)
    return _to_arrow(vals, valid, b.dtype)


def eval_expression_list(tbl: pa.Table, exprs) -> pa.Table:
    return pa.table([eval_expression(tbl, e) for e in exprs], names=[e.name() for e in exprs])


def filter(tbl: pa.Table, exprs) -> pa.Table:
    """RecordBatch::filter: AND of predicates, null => drop (lib.rs:520-563)."""
    pred = exprs[0]
    for e in exprs[1:]:
        pred = pred & e
    mask = eval_expression(tbl, pred)
    if not pa.types.is_boolean(mask.type):  # mask.bool()? in mask_filter (lib.rs:543) -> "... BooleanType" ValueError
        raise ValueError(f"filter predicate must be boolean, got {mask.type}")
    return tbl.filter(pc.fill_null(mask, False))

"""Expression trees (host side) and their serialisation to the C-ABI plan encoding.

Mirrors the slice of the reference's DSL the hot path needs (daft.col / daft.lit / Expression
operators, daft/expressions/expressions.py; Expr enum src/daft-dsl/src/expr/mod.rs).  Expressions are
pure data: resolving names to column indices, inserting the supertype casts the reference decides at
plan time, and encoding to bytes happen in `bind()`; evaluation happens in CUDA (libdaftb200.so).
"""
from __future__ import annotations

import datetime
import struct
from typing import Any, Sequence

from . import datatypes as dt

# plan tags / op codes: include/daft_b200.h "expression / plan encoding"
T_COL, T_LIT, T_BINARY, T_UNARY, T_CAST, T_IF_ELSE, T_IS_IN, T_BETWEEN, T_FILL_NULL, T_FUNCTION = range(10)
BINOPS = {"add": 0, "sub": 1, "mul": 2, "true_div": 3, "floor_div": 4, "mod": 5, "lt": 6, "le": 7, "eq": 8, "ne": 9,
          "ge": 10, "gt": 11, "and": 12, "or": 13, "xor": 14, "eq_null_safe": 15, "shl": 16, "shr": 17}
UNOPS = {"not": 0, "neg": 1, "is_null": 2, "not_null": 3, "abs": 4}
FUNCS = {"year": 0, "month": 1, "day": 2, "utf8_length": 3, "startswith": 4, "endswith": 5, "contains": 6, "left": 7,
         "day_of_week": 8, "sqrt": 9, "is_nan": 10, "not_nan": 11, "is_inf": 12, "ceil": 13, "floor": 14, "sign": 15,
         "round": 16, "match_seq": 17, "hour": 18, "minute": 19, "second": 20, "ts_trunc": 21, "substr": 22, "rx_match": 23, "rx_extract": 24,
         "rx_replace_group": 25, "ts_date": 26, "exp": 27, "ln": 28, "log2": 29, "log10": 30, "sin": 31, "cos": 32, "tan": 33, "cbrt": 34, "lstrip": 35, "rstrip": 36, "right": 37, "find": 38}
_FLOAT_MATH = ("sqrt", "exp", "ln", "log2", "log10", "sin", "cos", "tan", "cbrt")
AGG_KINDS = {"count": 0, "sum": 1, "min": 2, "max": 3, "mean": 4, "any_value": 5, "list": 6, "set": 7, "concat": 8}
COUNT_MODES = {"all": 0, "valid": 1, "null": 2}
_CMP = ("lt", "le", "eq", "ne", "ge", "gt", "eq_null_safe")
_ARITH = ("add", "sub", "mul", "floor_div", "mod")

_EPOCH = datetime.date(1970, 1, 1)


class Expression:
    """A node of the expression tree.  kind in {col, lit, binary, unary, cast, if_else, is_in, between,
    fill_null, function, alias, agg}."""

    __slots__ = ("kind", "op", "args", "value", "dtype", "_name")

    def __init__(self, kind, op=None, args=(), value=None, dtype=None, name=None):
        self.kind, self.op, self.args, self.value, self.dtype, self._name = kind, op, tuple(args), value, dtype, name

    # ---- naming (Expr::name: the left-most column name, or the alias) ----
    def name(self) -> str:
        if self.kind == "alias":
            return self._name
        if self.kind == "col":
            return self.value
        if self.kind == "lit":
            return "literal"
        if self.kind == "is_in" or self.kind == "between" or self.kind == "fill_null":
            return self.args[0].name()
        if self.kind == "if_else":
            return self.args[1].name()
        return self.args[0].name() if self.args else "expr"

    def alias(self, name: str) -> "Expression":
        return Expression("alias", args=(self,), name=name)

    # ---- operators ----
    def _bin(self, op, other, reverse=False):
        o = _to_expr(other)
        return Expression("binary", op, (o, self) if reverse else (self, o))

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o): return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("true_div", o)
    def __rtruediv__(self, o): return self._bin("true_div", o, True)
    def __floordiv__(self, o): return self._bin("floor_div", o)
    def __mod__(self, o): return self._bin("mod", o)
    def __lt__(self, o): return self._bin("lt", o)
    def __le__(self, o): return self._bin("le", o)
    def __eq__(self, o): return self._bin("eq", o)  # type: ignore[override]
    def __ne__(self, o): return self._bin("ne", o)  # type: ignore[override]
    def __ge__(self, o): return self._bin("ge", o)
    def __gt__(self, o): return self._bin("gt", o)
    def __and__(self, o): return self._bin("and", o)
    def __rand__(self, o): return self._bin("and", o, True)
    def __or__(self, o): return self._bin("or", o)
    def __ror__(self, o): return self._bin("or", o, True)
    def __xor__(self, o): return self._bin("xor", o)
    def __lshift__(self, o): return self._bin("shl", o)
    def __rshift__(self, o): return self._bin("shr", o)
    def __invert__(self): return Expression("unary", "not", (self,))
    def __neg__(self): return Expression("unary", "neg", (self,))
    def __abs__(self): return Expression("unary", "abs", (self,))
    def __hash__(self): return id(self)
    def __bool__(self): raise ValueError("Expressions are lazy; use & | ~ instead of and/or/not")

    def eq_null_safe(self, o): return self._bin("eq_null_safe", o)
    def abs(self): return Expression("unary", "abs", (self,))
    def is_null(self): return Expression("unary", "is_null", (self,))
    def not_null(self): return Expression("unary", "not_null", (self,))
    def cast(self, dtype) -> "Expression":
        code = dtype if isinstance(dtype, int) else dt.from_arrow(dtype)
        return Expression("cast", args=(self,), dtype=code)
    def if_else(self, if_true, if_false): return Expression("if_else", args=(self, _to_expr(if_true), _to_expr(if_false)))
    def is_in(self, items: Sequence[Any]): return Expression("is_in", args=(self, *[_to_expr(i) for i in items]))
    def between(self, lower, upper): return Expression("between", args=(self, _to_expr(lower), _to_expr(upper)))
    def fill_null(self, fill): return Expression("fill_null", args=(self, _to_expr(fill)))

    # temporal / string namespaces, flattened
    def year(self): return Expression("function", "year", (self,))
    def month(self): return Expression("function", "month", (self,))
    def day(self): return Expression("function", "day", (self,))
    def day_of_week(self): return Expression("function", "day_of_week", (self,))
    def date(self): return Expression("function", "ts_date", (self,))    # dt.date(): the day of a Timestamp(us) as a Date
    def hour(self): return Expression("function", "hour", (self,))       # of a Timestamp; extract(hour / minute / second from ts)
    def minute(self): return Expression("function", "minute", (self,))
    def second(self): return Expression("function", "second", (self,))
    def truncate(self, unit: str):
        """dt.truncate / DATE_TRUNC to a multiple of `unit` since the epoch: 'second', 'minute', 'hour' or 'day' (calendar units that
        do not divide the epoch timeline evenly -- week, month, year -- are refused)."""
        us = {"second": 1_000_000, "minute": 60_000_000, "hour": 3_600_000_000, "day": 86_400_000_000}.get(unit)
        if us is None:
            raise ValueError(f"truncate: unsupported unit {unit!r} (second, minute, hour, day)")
        return Expression("function", "ts_trunc", (self, lit(us, dt.I64)))
    def length(self): return Expression("function", "utf8_length", (self,))
    def startswith(self, pat): return Expression("function", "startswith", (self, _to_expr(pat)))
    def endswith(self, pat): return Expression("function", "endswith", (self, _to_expr(pat)))
    def contains(self, pat): return Expression("function", "contains", (self, _to_expr(pat)))
    def left(self, n): return Expression("function", "left", (self, _to_expr(n)))
    def right(self, n): return Expression("function", "right", (self, _to_expr(n)))          # the last n characters
    def lstrip(self): return Expression("function", "lstrip", (self,))                        # Unicode White_Space, as str::trim_start
    def rstrip(self): return Expression("function", "rstrip", (self,))
    def find(self, substr): return Expression("function", "find", (self, _to_expr(substr)))   # byte index of the first occurrence, -1 = none
    def substr(self, start: int, length: int | None = None):
        """utf8 substr(start, length) in characters, 0-based (src/daft-functions-utf8/src/substr.rs:269-294): null when the value has no
        character at `start` or `length` is 0; `length=None` runs to the end.  `start` and `length` are literals."""
        if int(start) < 0 or (length is not None and int(length) < 0):
            raise ValueError("substr: start and length must not be negative")
        return Expression("function", "substr", (self, lit(int(start), dt.I64), lit(-1 if length is None else int(length), dt.I64)))

    # ---- aggregations (AggExpr, src/daft-dsl/src/expr/mod.rs) ----
    def _agg(self, kind, mode="valid"): return Expression("agg", kind, (self,), value=mode)
    def sum(self): return self._agg("sum")
    def mean(self): return self._agg("mean")
    def avg(self): return self._agg("mean")
    def min(self): return self._agg("min")
    def max(self): return self._agg("max")
    def count(self, mode: str = "valid"): return self._agg("count", mode)
    def sqrt(self): return Expression("function", "sqrt", (self,))
    def exp(self): return Expression("function", "exp", (self,))
    def ln(self): return Expression("function", "ln", (self,))
    def log2(self): return Expression("function", "log2", (self,))
    def log10(self): return Expression("function", "log10", (self,))
    def sin(self): return Expression("function", "sin", (self,))
    def cos(self): return Expression("function", "cos", (self,))
    def tan(self): return Expression("function", "tan", (self,))
    def cbrt(self): return Expression("function", "cbrt", (self,))
    def is_nan(self): return Expression("function", "is_nan", (self,))
    def not_nan(self): return Expression("function", "not_nan", (self,))
    def is_inf(self): return Expression("function", "is_inf", (self,))
    def ceil(self): return Expression("function", "ceil", (self,))
    def floor(self): return Expression("function", "floor", (self,))
    def sign(self): return Expression("function", "sign", (self,))
    def negate(self): return Expression("unary", "neg", (self,))
    def round(self, decimals: int = 0):
        """Float round to `decimals` places, half away from zero: (v * 10^d).round() / 10^d (src/daft-core/src/array/ops/round.rs:14-35)."""
        return Expression("function", "round", (self, lit(int(decimals), dt.I32)))
    def match(self, pattern: str):
        """utf8 `match` (regex search, src/daft-functions-utf8/src/match_.rs).  Literal segments separated by `.*`, optionally anchored with
        ^ / $ (".*special.*requests.*": the TPC-H shapes) run as substring searches; any other pattern of the supported regex subset
        (daft_b200/regex.py) runs in the device's Pike VM."""
        try:
            return Expression("function", "match_seq", (self, *_segments_of_regex(pattern)))
        except NotImplementedError:
            from . import regex
            return Expression("function", "rx_match", (self, lit(regex.compile_program(pattern, None), dt.LARGE_BINARY), lit(pattern)))
    def regexp(self, pattern: str):
        return self.match(pattern)
    def regexp_extract(self, pattern: str, index: int = 0):
        """utf8 `extract(pattern, index)` (src/daft-functions-utf8/src/extract.rs): the first match's capture group `index` (0 = the whole
        match); null when the pattern does not match or the group took no part in the match."""
        from . import regex
        return Expression("function", "rx_extract", (self, lit(regex.compile_program(pattern, int(index)), dt.LARGE_BINARY), lit(pattern)), value=int(index))
    def regexp_replace(self, pattern: str, replacement: str):
        """utf8 `replace(pattern, replacement, regex=True)` / SQL REGEXP_REPLACE (src/daft-functions-utf8/src/replace.rs) for the form whose
        result is a slice of the input: a pattern anchored at both ends (a match is then the whole value) replaced by ONE capture group
        (`\\1`, `$1`, `${1}`), e.g. ClickBench Q28's host extraction.  Values that do not match stay as they are.  Other forms build new
        strings and are refused."""
        from . import regex
        import re as _re
        m = _re.fullmatch(r"\\(\d)|\$(\d)|\$\{(\d)\}", replacement)
        if m is None or not regex.is_fully_anchored(pattern):
            raise NotImplementedError("regexp_replace: only a fully anchored pattern (^...$) replaced by a single group reference is supported on the device")
        group = int(next(g for g in m.groups() if g is not None))
        return Expression("function", "rx_replace_group", (self, lit(regex.compile_program(pattern, group), dt.LARGE_BINARY), lit(pattern)), value=group)
    def like(self, pattern: str):
        """SQL LIKE with `%` wildcards (src/daft-functions-utf8/src/like.rs): anchored at both ends; `_` is refused."""
        return Expression("function", "match_seq", (self, *_segments_of_like(pattern)))

    def fill_nan(self, fill_value):
        """input.if_else(fill_value, not_nan(input)) (src/daft-functions/src/float/fill_nan.rs:41-49): nulls stay null."""
        return Expression("fill_nan", args=(self, _to_expr(fill_value)))
    def stddev(self, ddof: int = 0):
        """Staged exactly like the reference's plan rewrite (src/daft-local-plan/src/agg.rs:190-226): sum(x), sum(x*x),
        count(x) over Float64, then sqrt(((sqsum/n) - (sum/n)^2) * n / (n - ddof)), null when n <= ddof."""
        return self._agg("stddev", int(ddof))
    def var(self, ddof: int = 1):
        """src/daft-local-plan/src/agg.rs:227-258: the same staging without the square root."""
        return self._agg("var", int(ddof))
    def skew(self):
        """AggExpr::Skew as the planner stages it (src/daft-local-plan/src/agg.rs:302-348): count, sum, sum of squares and sum of cubes
        of the Float64 values, then (S3 - 3 S2 S / N + 2 S^3 / N^2) / sqrt(((S2 - S^2 / N) / N)^3) / N."""
        return self._agg("skew")
    def count_distinct(self):
        """Distinct non-null values per group as UInt64 (Series::count_distinct = agg_list + list_count_distinct over a
        probe table WITHOUT nulls, src/daft-functions-list/src/series.rs:37-40,234-247)."""
        return self._agg("count_distinct")
    def approx_count_distinct(self):
        """AggExpr::ApproxCountDistinct in ONE stage (src/daft-recordbatch/src/lib.rs:749-758, array/ops/approx_count_distinct.rs:13-40):
        the number of distinct XXH3 hashes of the non-null values -- an exact distinct count unless two values collide in 64 bits
        (the HyperLogLog sketch only appears in the reference's two-stage plan).  Runs as count_distinct."""
        return self._agg("count_distinct")
    def percentile(self, percentage: float):
        """Series::percentile (src/daft-core/src/series/ops/agg.rs:269-301, utils/stats.rs:31-59): the exact percentile of the non-null
        values as Float64, linearly interpolated between the two neighbouring order statistics."""
        if not 0.0 <= float(percentage) <= 1.0:
            raise ValueError(f"Provided percentile must be between 0 and 1: {percentage}")
        return self._agg("percentile", float(percentage))
    def median(self):
        """AggExpr::Median = percentile(0.5) (src/daft-recordbatch/src/lib.rs:779)."""
        return self._agg("percentile", 0.5)
    def any_value(self, ignore_nulls: bool = False):
        """Series::any_value (src/daft-core/src/series/ops/agg.rs:353-382): the group's first row, or its first non-null row."""
        return self._agg("any_value", "ignore_nulls" if ignore_nulls else "first")
    def agg_list(self):
        """Each group's values in row order, nulls kept, as a List column (src/daft-core/src/array/ops/list_agg.rs:30-57)."""
        return self._agg("list", "all")
    def agg_set(self):
        """Each group's distinct non-null values in first-occurrence order (src/daft-core/src/series/ops/agg.rs:531-575)."""
        return self._agg("set", "all")
    def agg_concat(self):
        """Each group's non-null strings back to back (src/daft-core/src/array/ops/concat_agg.rs:138-161); List inputs are not built."""
        return self._agg("concat", "all")
    def bool_and(self): return self._agg("min")   # nulls ignored, all-null group -> null (array/ops/bool_agg.rs:19-60)
    def bool_or(self): return self._agg("max")

    def __repr__(self):
        if self.kind == "col": return f"col({self.value})"
        if self.kind == "lit": return f"lit({self.value!r})"
        if self.kind == "alias": return f"{self.args[0]!r}.alias({self._name})"
        return f"{self.kind}:{self.op}({', '.join(map(repr, self.args))})"


def _segments_of_regex(pattern: str):
    """-> (flags literal, segment literals...): flags bit 0 = anchored at the start, bit 1 = anchored at the end."""
    start = pattern.startswith("^")
    end = pattern.endswith("$") and not pattern.endswith("\\$")
    body = pattern[1 if start else 0:len(pattern) - (1 if end else 0)]
    parts = body.split(".*")
    for part in parts:
        if any(ch in part for ch in ".^$*+?()[]{}|\\"):
            raise NotImplementedError(f"match: only literal segments separated by '.*' are supported on the device, got {pattern!r}")
    # a leading / trailing '.*' cancels the anchor on that side
    if parts and parts[0] == "" and len(parts) > 1:
        start, parts = False, parts[1:]
    if parts and parts[-1] == "" and len(parts) > 1:
        end, parts = False, parts[:-1]
    segs = [p_ for p_ in parts if p_ != ""]
    return [lit((1 if start else 0) | (2 if end else 0), dt.I32)] + [lit(x) for x in segs]


def _segments_of_like(pattern: str):
    if "_" in pattern or "\\" in pattern:
        raise NotImplementedError(f"like: only '%' wildcards are supported on the device, got {pattern!r}")
    parts = pattern.split("%")
    start, end = parts[0] != "", parts[-1] != ""
    if len(parts) == 1:
        start = end = True
    segs = [p_ for p_ in parts if p_ != ""]
    return [lit((1 if start else 0) | (2 if end else 0), dt.I32)] + [lit(x) for x in segs]


def col(name: str) -> Expression:
    return Expression("col", value=name)


def lit(value: Any, dtype: int | None = None) -> Expression:
    return Expression("lit", value=value, dtype=dtype)


def _to_expr(x) -> Expression:
    return x if isinstance(x, Expression) else lit(x)


def _lit_dtype(v) -> int:
    """Literal typing of daft.lit (src/daft-dsl/src/lit): Python int -> Int32 when it fits, else Int64."""
    if v is None: return dt.NULL
    if isinstance(v, bool): return dt.BOOL
    if isinstance(v, int): return dt.I32 if -2**31 <= v < 2**31 else (dt.I64 if v < 2**63 else dt.U64)
    if isinstance(v, float): return dt.F64
    if isinstance(v, str): return dt.LARGE_UTF8
    if isinstance(v, bytes): return dt.LARGE_BINARY
    if isinstance(v, datetime.datetime): return dt.TIMESTAMP
    if isinstance(v, datetime.date): return dt.DATE32
    raise DaftTypeError(f"unsupported literal {v!r}")


# ------------------------------------------------------------------------------------------------
# binding: names -> indices, dtype inference, supertype casts, encoding
# ------------------------------------------------------------------------------------------------
class DaftTypeError(TypeError, ValueError):
    """Typing errors of the bound expression.  The reference raises DaftError::TypeError, which its Python layer surfaces as
    ValueError (e.g. tests/series/test_comparisons.py:476-500), so both exception types match."""


class Bound:
    """A typed, index-resolved node ready for encoding."""
    __slots__ = ("tag", "op", "dtype", "kids", "col", "value")

    def __init__(self, tag, dtype, op=0, kids=(), col=-1, value=None):
        self.tag, self.dtype, self.op, self.kids, self.col, self.value = tag, dtype, op, list(kids), col, value


def _cast(b: Bound, to: int) -> Bound:
    if b.dtype == to:
        return b
    if b.tag == T_LIT and b.value is not None and b.dtype != dt.NULL:
        # fold casts of literals so that `col(i64) < 5` compares against an Int64 immediate
        v = b.value
        if dt.is_float(to): v = float(v)
        elif dt.is_integer(to) and isinstance(v, float) and v == int(v): v = int(v)
        elif dt.is_integer(to) and isinstance(v, float): return Bound(T_CAST, to, kids=[b])
        if dt.is_integer(to) and isinstance(v, int) and not isinstance(v, bool):
            lo, hi = dt.int_range(to)
            if not lo <= v <= hi:
                v = None  # arrow-rs safe cast: a value the target type cannot hold becomes null (array/ops/cast.rs:194-196)
        return Bound(T_LIT, to, value=v)
    return Bound(T_CAST, to, kids=[b])


def bind(e: Expression, names: Sequence[str], dtypes: Sequence[int]) -> Bound:
    k = e.kind
    if k == "alias":
        return bind(e.args[0], names, dtypes)
    if k == "col":
        try:
            i = list(names).index(e.value)
        except ValueError:
            raise KeyError(f"column {e.value!r} not found in {list(names)}") from None
        return Bound(T_COL, dtypes[i], col=i)
    if k == "lit":
        d = e.dtype if e.dtype is not None else _lit_dtype(e.value)
        return Bound(T_LIT, d, value=e.value)
    if k == "binary":
        l, r = bind(e.args[0], names, dtypes), bind(e.args[1], names, dtypes)
        op = e.op
        if op == "true_div":
            return Bound(T_BINARY, dt.F64, BINOPS[op], [_cast(l, dt.F64), _cast(r, dt.F64)])
        if op in ("and", "or", "xor"):
            if dt.is_integer(l.dtype) or dt.is_integer(r.dtype):
                # bitwise on integers: supertype, null if either side is null (src/daft-core/src/array/ops/bitwise.rs)
                if not all(dt.is_integer(x.dtype) or x.dtype == dt.NULL for x in (l, r)):
                    raise DaftTypeError(f"Cannot perform logical / bitwise {op} on types: {dt.NAMES[l.dtype]}, {dt.NAMES[r.dtype]}")
                st = dt.supertype(l.dtype, r.dtype)
                if not dt.is_integer(st):
                    raise DaftTypeError(f"cannot apply {op} to {dt.NAMES[l.dtype]} and {dt.NAMES[r.dtype]}")
                return Bound(T_BINARY, st, BINOPS[op], [_cast(l, st), _cast(r, st)])
            if not all(x.dtype in (dt.BOOL, dt.NULL) for x in (l, r)):
                raise DaftTypeError(f"Cannot perform logical / bitwise {op} on types: {dt.NAMES[l.dtype]}, {dt.NAMES[r.dtype]}")
            return Bound(T_BINARY, dt.BOOL, BINOPS[op], [_cast(l, dt.BOOL) if l.dtype == dt.NULL else l,
                                                          _cast(r, dt.BOOL) if r.dtype == dt.NULL else r])
        if op in ("shl", "shr"):
            if not (dt.is_integer(l.dtype) and dt.is_integer(r.dtype)):  # infer_datatype.rs:494-520, test_numeric.py:773-785
                raise DaftTypeError(f"Cannot operate shift {'left' if op == 'shl' else 'right'} on types: {dt.NAMES[l.dtype]}, {dt.NAMES[r.dtype]}")
            return Bound(T_BINARY, l.dtype, BINOPS[op], [l, _cast(r, l.dtype)])
        st = dt.rem_supertype(l.dtype, r.dtype) if op == "mod" else dt.supertype(l.dtype, r.dtype)
        l, r = _cast(l, st), _cast(r, st)
        if op in _CMP:
            return Bound(T_BINARY, dt.BOOL, BINOPS[op], [l, r])
        if op in _ARITH:
            if not dt.is_numeric(st):
                raise DaftTypeError(f"cannot apply {op} to {dt.NAMES[st]}")
            if op == "floor_div" and st == dt.F32:
                # float `//` is (l / r).floor() computed as a Float64 true division, then cast to the supertype
                # (src/daft-core/src/series/ops/arithmetic.rs:466-469)
                return _cast(Bound(T_BINARY, dt.F64, BINOPS[op], [_cast(l, dt.F64), _cast(r, dt.F64)]), dt.F32)
            return Bound(T_BINARY, st, BINOPS[op], [l, r])
        raise ValueError(op)
    if k == "unary":
        c = bind(e.args[0], names, dtypes)
        if e.op == "not" and c.dtype not in (dt.BOOL, dt.NULL):
            raise ValueError(f"cannot apply not to {dt.NAMES[c.dtype]}: expected Boolean")  # tests/recordbatch/test_eval.py:47-51
        if e.op in ("is_null", "not_null", "not"):
            return Bound(T_UNARY, dt.BOOL, UNOPS[e.op], [c])
        return Bound(T_UNARY, c.dtype, UNOPS[e.op], [c])
    if k == "cast":
        c = bind(e.args[0], names, dtypes)
        if c.dtype != dt.NULL and c.dtype != e.dtype:
            if dt.is_string(c.dtype) != dt.is_string(e.dtype) and not (dt.is_string(c.dtype) and dt.is_integer(e.dtype)):
                # Utf8 -> integer parses on the device (trimmed, [+-]?digits, out of range / malformed -> null: array/ops/cast.rs:164-173)
                raise NotImplementedError(f"cast {dt.NAMES[c.dtype]} -> {dt.NAMES[e.dtype]}: of the parsing / formatting casts only "
                                          "Utf8 -> integer is built")
            if c.dtype == dt.DATE32 and e.dtype == dt.TIMESTAMP:
                # days x microseconds per day, as an Int64 product (src/daft-core/src/array/ops/cast.rs:268-279), then the physical cast
                us = Bound(T_BINARY, dt.I64, BINOPS["mul"], [_cast(c, dt.I64), Bound(T_LIT, dt.I64, value=86_400_000_000)])
                return Bound(T_CAST, dt.TIMESTAMP, kids=[us])
            if c.dtype == dt.TIMESTAMP and e.dtype == dt.DATE32:
                return Bound(T_FUNCTION, dt.DATE32, FUNCS["ts_date"], [c])   # TimestampArray::cast -> self.date() (cast.rs:347)
        return _cast(c, e.dtype)
    if k == "if_else":
        p, a, b = (bind(x, names, dtypes) for x in e.args)
        st = dt.supertype(a.dtype, b.dtype)
        return Bound(T_IF_ELSE, st, kids=[p, _cast(a, st), _cast(b, st)])
    if k == "is_in":
        c = bind(e.args[0], names, dtypes)
        items = [bind(x, names, dtypes) for x in e.args[1:]]
        st = c.dtype
        for it in items:
            st = dt.supertype(st, it.dtype)
        return Bound(T_IS_IN, dt.BOOL, kids=[_cast(c, st)] + [_cast(it, st) for it in items])
    if k == "between":
        c, lo, hi = (bind(x, names, dtypes) for x in e.args)
        st = dt.supertype(dt.supertype(c.dtype, lo.dtype), hi.dtype)
        # value.gte(lower).and(value.lte(upper)) with the Kleene `and` (src/daft-core/src/array/ops/between.rs:30-32): a null
        # bound makes the row null unless the other comparison is already false (tests/recordbatch/test_between.py:44-48)
        cv, lov, hiv = _cast(c, st), _cast(lo, st), _cast(hi, st)
        ge = Bound(T_BINARY, dt.BOOL, BINOPS["ge"], [cv, lov])
        le = Bound(T_BINARY, dt.BOOL, BINOPS["le"], [cv, hiv])
        return Bound(T_BINARY, dt.BOOL, BINOPS["and"], [ge, le])
    if k == "fill_null":
        c, f = bind(e.args[0], names, dtypes), bind(e.args[1], names, dtypes)
        st = dt.supertype(c.dtype, f.dtype)
        return Bound(T_FILL_NULL, st, kids=[_cast(c, st), _cast(f, st)])
    if k == "fill_nan":
        x, f = (bind(a, names, dtypes) for a in e.args)
        if x.dtype == dt.NULL:
            return x
        if not (dt.is_float(x.dtype) and dt.is_float(f.dtype)):
            raise DaftTypeError(f"Expects input for fill_nan to be float, but received {dt.NAMES[x.dtype]} and {dt.NAMES[f.dtype]}")
        st = dt.supertype(x.dtype, f.dtype)
        x, f = _cast(x, st), _cast(f, st)
        return Bound(T_IF_ELSE, st, kids=[Bound(T_FUNCTION, dt.BOOL, FUNCS["not_nan"], [x]), x, f])
    if k == "function" and e.op in ("is_nan", "not_nan", "is_inf", "ceil", "floor", "sign"):
        c = bind(e.args[0], names, dtypes)
        if e.op in ("is_nan", "not_nan", "is_inf"):
            if c.dtype == dt.NULL:  # float.rs:29-40: an all-null boolean column
                return Bound(T_LIT, dt.BOOL, value=None)
            if not dt.is_float(c.dtype):
                raise ValueError(f"Expected input to {e.op} to be float, got {dt.NAMES[c.dtype]}")
            return Bound(T_FUNCTION, dt.BOOL, FUNCS[e.op], [c])
        if not dt.is_numeric(c.dtype):
            raise ValueError(f"Expected input to {e.op} to be numeric, got {dt.NAMES[c.dtype]}")  # test_numeric.py:116-120,141-145,179-183
        if e.op in ("ceil", "floor") and dt.is_integer(c.dtype):
            return c  # integers are returned unchanged (src/daft-functions/src/numeric/ceil.rs:25-33, floor.rs)
        return Bound(T_FUNCTION, c.dtype, FUNCS[e.op], [c])
    if k == "function" and e.op == "round":
        c, d = bind(e.args[0], names, dtypes), bind(e.args[1], names, dtypes)
        if not dt.is_numeric(c.dtype):
            raise ValueError(f"Expected input to round to be numeric, got {dt.NAMES[c.dtype]}")
        if dt.is_integer(c.dtype):
            return c  # integers are returned unchanged (src/daft-functions/src/numeric/round.rs)
        return Bound(T_FUNCTION, c.dtype, FUNCS["round"], [c, _cast(d, dt.I32)])
    if k == "function" and e.op == "match_seq":
        kids = [bind(x, names, dtypes) for x in e.args]
        if not dt.is_string(kids[0].dtype):
            raise ValueError(f"Expects input to match to be utf8, but received {dt.NAMES[kids[0].dtype]}")
        return Bound(T_FUNCTION, dt.BOOL, FUNCS["match_seq"], [kids[0], _cast(kids[1], dt.I32)] + kids[2:])
    if k == "function":
        kids = [bind(x, names, dtypes) for x in e.args]
        out = {"year": dt.I32, "month": dt.U32, "day": dt.U32, "day_of_week": dt.U32, "utf8_length": dt.U64,
               "startswith": dt.BOOL, "endswith": dt.BOOL, "contains": dt.BOOL, "left": dt.LARGE_UTF8, "sqrt": dt.F64,
               "hour": dt.U32, "minute": dt.U32, "second": dt.U32, "ts_trunc": dt.TIMESTAMP, "substr": dt.LARGE_UTF8,
               "rx_match": dt.BOOL, "rx_extract": dt.LARGE_UTF8, "rx_replace_group": dt.LARGE_UTF8, "ts_date": dt.DATE32,
               "exp": dt.F64, "ln": dt.F64, "log2": dt.F64, "log10": dt.F64, "sin": dt.F64, "cos": dt.F64, "tan": dt.F64, "cbrt": dt.F64,
               "lstrip": dt.LARGE_UTF8, "rstrip": dt.LARGE_UTF8, "right": dt.LARGE_UTF8, "find": dt.I64}[e.op]
        if e.op in ("lstrip", "rstrip", "right", "find"):
            if not dt.is_string(kids[0].dtype) or (e.op == "find" and not dt.is_string(kids[1].dtype)):
                raise ValueError(f"Expects input to {e.op} to be utf8, but received {dt.NAMES[kids[0].dtype]}")
            if e.op == "right":
                kids[1] = _cast(kids[1], dt.I64)
            if e.op != "find":
                out = kids[0].dtype
        if e.op == "ts_date" and kids[0].dtype != dt.TIMESTAMP:
            raise ValueError(f"Expected input to date to be a timestamp, got {dt.NAMES[kids[0].dtype]}")
        if e.op in ("year", "month", "day", "day_of_week") and kids[0].dtype == dt.TIMESTAMP:
            kids[0] = Bound(T_FUNCTION, dt.DATE32, FUNCS["ts_date"], [kids[0]])   # calendar fields of a timestamp = those of its day
        if e.op in ("rx_match", "rx_extract", "rx_replace_group"):
            if not dt.is_string(kids[0].dtype):
                raise ValueError(f"Expects input to {e.op} to be utf8, but received {dt.NAMES[kids[0].dtype]}")
            if e.op != "rx_match":
                out = kids[0].dtype
        if e.op == "substr":
            if not dt.is_string(kids[0].dtype):
                raise ValueError(f"Expects input to substr to be utf8, but received {dt.NAMES[kids[0].dtype]}")
            out = kids[0].dtype
        if e.op in ("hour", "minute", "second", "ts_trunc") and kids[0].dtype != dt.TIMESTAMP:
            raise ValueError(f"Expected input to {e.op} to be a timestamp, got {dt.NAMES[kids[0].dtype]}")
        if e.op in ("utf8_length", "startswith", "endswith", "contains", "left") and not dt.is_string(kids[0].dtype):
            raise ValueError(f"Expects input to {e.op} to be utf8, but received {dt.NAMES[kids[0].dtype]}")
        if e.op in ("startswith", "endswith", "contains") and not dt.is_string(kids[1].dtype):
            raise ValueError(f"Expects pattern of {e.op} to be utf8, but received {dt.NAMES[kids[1].dtype]}")
        if e.op in ("year", "month", "day", "day_of_week") and kids[0].dtype not in (dt.DATE32,):
            raise ValueError(f"Expected input to {e.op} to be a date, got {dt.NAMES[kids[0].dtype]}")
        if e.op in _FLOAT_MATH and not dt.is_numeric(kids[0].dtype):
            raise ValueError(f"Expected input to {e.op} to be numeric, got {dt.NAMES[kids[0].dtype]}")
        if e.op == "left":
            kids[1] = _cast(kids[1], dt.I64)
        if e.op in _FLOAT_MATH:
            # computed in the operand's floating representation (src/daft-schema/src/dtype.rs:756-772, numeric/sqrt.rs:27-35)
            out = dt.F32 if kids[0].dtype in (dt.I8, dt.I16, dt.U8, dt.U16, dt.F32) else dt.F64
            kids[0] = _cast(kids[0], out)
        return Bound(T_FUNCTION, out, FUNCS[e.op], kids, value=e.value)
    if k == "agg":
        raise ValueError("aggregation expressions are only valid inside agg()")
    raise ValueError(k)


def _enc_str(s: bytes) -> bytes:
    return struct.pack("<I", len(s)) + s


def encode(b: Bound) -> bytes:
    if b.tag == T_COL:
        return struct.pack("<BI", T_COL, b.col)
    if b.tag == T_LIT:
        v, d = b.value, b.dtype
        isnull = v is None or d == dt.NULL
        head = struct.pack("<BBB", T_LIT, d, 1 if isnull else 0)
        if dt.is_string(d):
            raw = b"" if isnull else (v.encode() if isinstance(v, str) else bytes(v))
            return head + _enc_str(raw)
        if isnull:
            return head + struct.pack("<q", 0)
        if d in (dt.F32, dt.F64):
            return head + struct.pack("<d", float(v))
        if d == dt.DATE32:
            days = (v - _EPOCH).days if isinstance(v, datetime.date) else int(v)
            return head + struct.pack("<q", days)
        if d == dt.TIMESTAMP:
            if isinstance(v, datetime.datetime):
                v = int((v - datetime.datetime(1970, 1, 1)).total_seconds() * 1_000_000)
            return head + struct.pack("<q", int(v))
        if d == dt.U64:
            return head + struct.pack("<Q", int(v))
        return head + struct.pack("<q", int(v))
    if b.tag in (T_BINARY, T_UNARY):
        return struct.pack("<BB", b.tag, b.op) + b"".join(encode(k) for k in b.kids)
    if b.tag == T_CAST:
        return struct.pack("<BB", T_CAST, b.dtype) + encode(b.kids[0])
    if b.tag in (T_IF_ELSE, T_BETWEEN, T_FILL_NULL):
        return struct.pack("<B", b.tag) + b"".join(encode(k) for k in b.kids)
    if b.tag == T_IS_IN:
        return struct.pack("<B", T_IS_IN) + encode(b.kids[0]) + struct.pack("<I", len(b.kids) - 1) + b"".join(encode(k) for k in b.kids[1:])
    if b.tag == T_FUNCTION:
        return struct.pack("<BBB", T_FUNCTION, b.op, len(b.kids)) + b"".join(encode(k) for k in b.kids)
    raise ValueError(b.tag)


def encode_project(exprs: Sequence[Expression], names, dtypes) -> tuple[bytes, list[str], list[int]]:
    out = struct.pack("<I", len(exprs))
    onames, odt = [], []
    for e in exprs:
        b = bind(e, names, dtypes)
        out += _enc_str(e.name().encode()) + encode(b)
        onames.append(e.name())
        odt.append(b.dtype)
    return out, onames, odt


def encode_filter(exprs: Sequence[Expression], names, dtypes) -> bytes:
    """Conjuncts are AND-ed (RecordBatch::filter, src/daft-recordbatch/src/lib.rs:520-541)."""
    pred = exprs[0]
    for e in exprs[1:]:
        pred = pred & e
    b = bind(pred, names, dtypes)
    if b.dtype != dt.BOOL:
        raise ValueError(f"filter predicate must be boolean, got {dt.NAMES[b.dtype]}")
    return encode(b)


def encode_agg(to_agg: Sequence[Expression], group_by: Sequence[Expression], names, dtypes,
               predicate: Sequence[Expression] | None = None) -> bytes:
    out = b""
    if predicate:
        out += b"\x01" + encode_filter(predicate, names, dtypes)
    else:
        out += b"\x00"
    out += struct.pack("<I", len(group_by))
    for g in group_by:
        out += _enc_str(g.name().encode()) + encode(bind(g, names, dtypes))
    out += struct.pack("<I", len(to_agg))
    for a in to_agg:
        name = a.name()
        inner = a
        while inner.kind == "alias":
            inner = inner.args[0]
        if inner.kind != "agg":
            raise ValueError(f"expected an aggregation expression, got {a!r}")
        mode = {"first": 0, "ignore_nulls": 1}[inner.value] if inner.op == "any_value" else COUNT_MODES[inner.value]
        child = bind(inner.args[0], names, dtypes)
        if inner.op in ("sum", "mean") and not dt.is_numeric(child.dtype):
            # the reference rejects this when it types the aggregation (tests/recordbatch/test_table_aggs.py:283-286)
            raise ValueError(f"Expected input to {inner.op}() to be numeric, got {dt.NAMES[child.dtype]}")
        out += struct.pack("<BB", AGG_KINDS[inner.op], mode) + _enc_str(name.encode())
        out += encode(child)
    return out


def stage_moment_aggs(to_agg: Sequence[Expression]):
    """stddev / var / skew -> (first-stage aggregations, final projection) as populate_aggregation_stages does
    (src/daft-local-plan/src/agg.rs:190-258, 302-348); returns None when the list has none of them."""
    if not any(_agg_inner(a).op in ("stddev", "var", "skew") for a in to_agg):
        return None
    first, final = [], []
    for i, a in enumerate(to_agg):
        inner = _agg_inner(a)
        if inner.op not in ("stddev", "var", "skew"):
            first.append(a)
            final.append(col(a.name()))
            continue
        x = inner.args[0].cast(dt.F64)
        s, q, c = f"__sum{i}", f"__sqsum{i}", f"__cnt{i}"
        first += [x.sum().alias(s), (x * x).sum().alias(q), x.count().alias(c)]
        n = col(c).cast(dt.F64)
        if inner.op == "skew":
            k = f"__cbsum{i}"
            first.append((x * x * x).sum().alias(k))
            S, S2, S3 = col(s), col(q), col(k)
            denom_base = (S2 - S * S / n) / n
            denom = (denom_base * denom_base * denom_base).sqrt()
            numerator = S3 - lit(3) * S2 * S / n + lit(2) * S * S * S / n / n
            final.append((numerator / denom / n).alias(a.name()))
            continue
        mean = col(s) / n
        pop_var = col(q) / n - mean * mean
        ddof = lit(float(inner.value))
        adjusted = pop_var * n / (n - ddof)
        result = (n <= ddof).if_else(lit(None), adjusted.sqrt() if inner.op == "stddev" else adjusted)
        final.append(result.alias(a.name()))
    return first, final


def _agg_inner(a: Expression) -> Expression:
    inner = a
    while inner.kind == "alias":
        inner = inner.args[0]
    return inner

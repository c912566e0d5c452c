"""Physical dtypes of device-resident columns (codes match include/daft_b200.h b200_dtype)."""
from __future__ import annotations

import pyarrow as pa

NULL, BOOL, I8, I16, I32, I64, U8, U16, U32, U64, F32, F64, DATE32, TIMESTAMP, LARGE_UTF8, LARGE_BINARY = range(16)
LIST = 16   # output of the list / set aggregates; a schema code carries the child's code in bits 8..15 (b200_batch_dtypes)

NAMES = ["null", "bool", "int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64", "float32", "float64",
         "date32", "timestamp", "large_utf8", "large_binary", "list"]

_PA = {
    pa.null(): NULL, pa.bool_(): BOOL, pa.int8(): I8, pa.int16(): I16, pa.int32(): I32, pa.int64(): I64,
    pa.uint8(): U8, pa.uint16(): U16, pa.uint32(): U32, pa.uint64(): U64, pa.float32(): F32, pa.float64(): F64,
    pa.date32(): DATE32, pa.large_utf8(): LARGE_UTF8, pa.large_binary(): LARGE_BINARY,
    pa.utf8(): LARGE_UTF8, pa.binary(): LARGE_BINARY,
}
_TO_PA = {
    NULL: pa.null(), BOOL: pa.bool_(), I8: pa.int8(), I16: pa.int16(), I32: pa.int32(), I64: pa.int64(), U8: pa.uint8(),
    U16: pa.uint16(), U32: pa.uint32(), U64: pa.uint64(), F32: pa.float32(), F64: pa.float64(), DATE32: pa.date32(),
    TIMESTAMP: pa.timestamp("us"), LARGE_UTF8: pa.large_utf8(), LARGE_BINARY: pa.large_binary(),
}


def from_arrow(t: pa.DataType) -> int:
    if pa.types.is_timestamp(t):
        if t.tz is not None:
            raise NotImplementedError("timezone-aware timestamps are not supported on the device")
        return TIMESTAMP
    try:
        return _PA[t]
    except KeyError:
        raise NotImplementedError(f"dtype {t} is not supported on the device") from None


def to_arrow(code: int) -> pa.DataType:
    if code & 0xFF == LIST:
        return pa.large_list(_TO_PA[(code >> 8) & 0xFF])
    return _TO_PA[code]


def is_signed(c): return I8 <= c <= I64
def is_unsigned(c): return U8 <= c <= U64
def is_integer(c): return I8 <= c <= U64
def is_float(c): return c in (F32, F64)
def is_numeric(c): return is_integer(c) or is_float(c)
def is_string(c): return c in (LARGE_UTF8, LARGE_BINARY)


_WIDTH = {I8: 1, I16: 2, I32: 4, I64: 8, U8: 1, U16: 2, U32: 4, U64: 8, F32: 4, F64: 8, DATE32: 4, TIMESTAMP: 8}


def width(c): return _WIDTH.get(c, 0)


def int_like(c): return is_integer(c) or c in (DATE32, TIMESTAMP)
def int_like_signed(c): return is_signed(c) or c in (DATE32, TIMESTAMP)


def int_range(c):
    """[min, max] of an integer (or integer-backed) dtype."""
    bits = 8 * width(c)
    return (-(1 << (bits - 1)), (1 << (bits - 1)) - 1) if int_like_signed(c) else (0, (1 << bits) - 1)


def rem_supertype(a: int, b: int) -> int:
    """Operand type of `%` (try_integer_widen_for_rem, src/daft-core/src/datatypes/infer_datatype.rs:662-745): Float64 if either side
    is, else Float32 if either side is, else the integer of the larger width -- UNSIGNED as soon as one side is unsigned (a negative
    value on the other side then becomes null in the cast, as every out-of-range cast does)."""
    if is_numeric(a) and is_numeric(b):
        if F64 in (a, b):
            return F64
        if F32 in (a, b):
            return F32
        w = max(width(a), width(b))
        if is_signed(a) and is_signed(b):
            return {1: I8, 2: I16, 4: I32, 8: I64}[w]
        return {1: U8, 2: U16, 4: U32, 8: U64}[w]
    return supertype(a, b)


def supertype(a: int, b: int) -> int:
    """Common type both operands are cast to before a binary op (the reference decides this at plan
    time: src/daft-core/src/utils/supertype.rs try_get_supertype; restated for the numeric lattice)."""
    if a == b:
        return a
    if a == NULL:
        return b
    if b == NULL:
        return a
    if a == BOOL and is_numeric(b):
        return b
    if b == BOOL and is_numeric(a):
        return a
    if is_integer(a) and is_integer(b):
        wa, wb = width(a), width(b)
        if is_signed(a) == is_signed(b):
            w = max(wa, wb)
            return {1: I8, 2: I16, 4: I32, 8: I64}[w] if is_signed(a) else {1: U8, 2: U16, 4: U32, 8: U64}[w]
        ws, wu = (wa, wb) if is_signed(a) else (wb, wa)
        need = max(ws, wu * 2)
        if need > 8:
            return F64
        return {1: I8, 2: I16, 4: I32, 8: I64}[need]
    if is_float(a) and is_float(b):
        return F64
    if is_numeric(a) and is_numeric(b):
        f, i = (a, b) if is_float(a) else (b, a)
        if f == F32 and width(i) <= 2:
            return F32
        return F64
    if is_string(a) and is_string(b):
        return LARGE_BINARY if LARGE_BINARY in (a, b) else LARGE_UTF8
    # Date vs Timestamp: the reference compares at Date after a temporal conversion of the timestamp (infer_datatype.rs:132-137); that
    # conversion is not built, so mixing the two is refused instead of reinterpreting days as microseconds
    raise TypeError(f"no supertype for {NAMES[a]} and {NAMES[b]}")

"""Host-side expression logic: name resolution, literal typing, supertype casts, plan encoding."""
import datetime
import struct

import pytest

from daft_b200 import datatypes as dt
from daft_b200.expressions import T_BINARY, T_CAST, T_COL, T_LIT, bind, col, encode, encode_agg, encode_filter, lit

NAMES = ["i32", "i64", "f32", "f64", "s", "d", "b", "u8"]
DTS = [dt.I32, dt.I64, dt.F32, dt.F64, dt.LARGE_UTF8, dt.DATE32, dt.BOOL, dt.U8]


def test_literal_typing_follows_daft_lit():
    assert bind(lit(1), NAMES, DTS).dtype == dt.I32
    assert bind(lit(2**40), NAMES, DTS).dtype == dt.I64
    assert bind(lit(1.5), NAMES, DTS).dtype == dt.F64
    assert bind(lit("x"), NAMES, DTS).dtype == dt.LARGE_UTF8
    assert bind(lit(None), NAMES, DTS).dtype == dt.NULL
    assert bind(lit(datetime.date(1998, 9, 2)), NAMES, DTS).dtype == dt.DATE32


def test_supertype_casts_are_inserted():
    b = bind(col("i32") + col("i64"), NAMES, DTS)
    assert b.dtype == dt.I64 and b.kids[0].tag == T_CAST and b.kids[1].tag == T_COL
    b = bind(col("i64") < 5, NAMES, DTS)          # literal folded to the column's type
    assert b.dtype == dt.BOOL and b.kids[1].tag == T_LIT and b.kids[1].dtype == dt.I64
    b = bind(col("i32") / col("u8"), NAMES, DTS)  # true division is always Float64
    assert b.dtype == dt.F64 and all(k.dtype == dt.F64 for k in b.kids)
    assert bind(col("f32") + col("i32"), NAMES, DTS).dtype == dt.F64
    assert bind(col("f32") + col("u8"), NAMES, DTS).dtype == dt.F32
    assert bind(col("f64") * (1 - col("f64")), NAMES, DTS).dtype == dt.F64
    assert dt.supertype(dt.U64, dt.I8) == dt.F64 and dt.supertype(dt.U8, dt.I8) == dt.I16
    with pytest.raises(TypeError):
        bind(col("s") + col("i32"), NAMES, DTS)
    with pytest.raises(KeyError):
        bind(col("nope"), NAMES, DTS)


def test_encoding_layout():
    raw = encode(bind(col("i64"), NAMES, DTS))
    assert raw == struct.pack("<BI", 0, 1)
    raw = encode(bind(lit(datetime.date(1998, 9, 2)), NAMES, DTS))
    assert raw == struct.pack("<BBBq", 1, dt.DATE32, 0, 10471)
    raw = encode_filter([col("b"), col("i32") > 0], NAMES, DTS)
    assert raw[0] == T_BINARY and raw[1] == 12  # AND of the conjuncts
    plan = encode_agg([col("f64").sum().alias("s")], [col("s")], NAMES, DTS, [col("b")])
    assert plan[0] == 1
    with pytest.raises(ValueError):
        encode_agg([col("f64")], [col("s")], NAMES, DTS)
    with pytest.raises(ValueError):
        encode_filter([col("i32")], NAMES, DTS)


def test_expression_names():
    assert (col("a") + col("b")).name() == "a"
    assert (col("a") * 2).alias("z").name() == "z"
    assert col("a").sum().name() == "a"
    assert col("a").is_in([1, 2]).name() == "a"

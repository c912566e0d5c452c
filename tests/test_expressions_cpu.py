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


def test_substr_reference_literal_generated_code_and_nvrtc(tmp_path):
    """utf8 substr (tests/recordbatch/utf8/test_substr.py:7-10; substring(), src/daft-functions-utf8/src/substr.rs:269-294): oracle == the
    reference's literal, the generated code on the host == oracle (multi-byte characters, starts past the end, zero lengths), and the
    kernel compiles for sm_100a"""
    import ctypes
    import pyarrow as pa
    from daft_b200 import _lib, col
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_project
    from jit_emu import emulate_project
    from oracle import expr_eval as E
    t = pa.table({"col": pa.array(["foo", None, "barbarbar", "quux", "1", "", "héllo wörld", "日本語テキスト"], pa.large_string())})
    got = E.eval_expression_list(t, [col("col").substr(0, 5)])
    assert got.to_pydict() == {"col": ["foo", None, "barba", "quux", "1", None, "héllo", "日本語テキ"]}
    exprs = [col("col").substr(0, 5).alias("a"), col("col").substr(2).alias("b"), col("col").substr(1, 0).alias("c"), col("col").substr(3, 2).alias("d"),
             col("col").substr(4, 100).alias("e"), col("col").substr(1, 3).length().alias("n")]
    assert emulate_project(t, exprs).equals(E.eval_expression_list(t, exprs))
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    plan, _, odt = encode_project(exprs, ["col"], [dt.LARGE_UTF8])
    assert odt[:5] == [dt.LARGE_UTF8] * 5
    for nullable in (0, 1):
        _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 1)(dt.LARGE_UTF8), (ctypes.c_uint8 * 1)(nullable), ctypes.c_int32(1))
    import pytest
    with pytest.raises(ValueError, match="must not be negative"):
        col("col").substr(-1)
    with pytest.raises(ValueError, match="to be utf8"):
        encode_project([col("x").substr(0, 1)], ["x"], [dt.I64])


def test_calendar_fields_of_timestamps(tmp_path):
    """dt.date() / year / month / day / day_of_week of a Timestamp(us): oracle == pyarrow's calendar, generated code == oracle (instants on
    both sides of the epoch, day boundaries), NVRTC for sm_100a"""
    import ctypes
    import numpy as np
    import pyarrow as pa
    import pyarrow.compute as pc
    from daft_b200 import _lib, col
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_project
    from jit_emu import emulate_project
    from oracle import expr_eval as E
    rng = np.random.default_rng(0)
    us = np.concatenate([rng.integers(-2 * 10**15, 2 * 10**15, 500), np.array([0, -1, 86400000000 - 1, 86400000000, -86400000000, -86400000001])])
    t = pa.table({"ts": pa.array(us, type=pa.timestamp("us"), mask=rng.random(len(us)) < 0.1)})
    exprs = [col("ts").date().alias("d"), col("ts").year().alias("y"), col("ts").month().alias("m"), col("ts").day().alias("dd"), col("ts").day_of_week().alias("dow")]
    want = E.eval_expression_list(t, exprs)
    assert emulate_project(t, exprs).equals(want)
    for name, fn in (("y", pc.year), ("m", pc.month), ("dd", pc.day), ("dow", pc.day_of_week)):
        assert want.column(name).combine_chunks().cast(pa.int64()).equals(fn(t["ts"]).combine_chunks()), name
    assert want.column("d").combine_chunks().equals(t["ts"].combine_chunks().cast(pa.date32()))
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    plan, _, odt = encode_project(exprs, ["ts"], [dt.TIMESTAMP])
    assert odt == [dt.DATE32, dt.I32, dt.U32, dt.U32, dt.U32]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 1)(dt.TIMESTAMP), (ctypes.c_uint8 * 1)(1), ctypes.c_int32(1))


def test_unary_float_math(tmp_path):
    """exp / ln / log2 / log10 / sin / cos / tan / cbrt in the operand's floating representation (Float32 for narrow integers and Float32, as
    sqrt): the generated code on the host against numpy, to the floating-point tolerance the device's libdevice needs anyway; NVRTC"""
    import ctypes
    import numpy as np
    import pyarrow as pa
    from daft_b200 import _lib, col
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_project
    from helpers import assert_tables_equal
    from jit_emu import emulate_project
    from oracle import expr_eval as E
    rng = np.random.default_rng(0)
    t = pa.table({"x": pa.array(rng.standard_normal(1000) * 5, mask=rng.random(1000) < 0.1), "f": pa.array((rng.random(1000) * 10).astype(np.float32)),
                  "i": pa.array(rng.integers(1, 100, 1000).astype(np.int16))})
    fns = ("exp", "ln", "log2", "log10", "sin", "cos", "tan", "cbrt")
    exprs = [getattr(col(c), fn)().alias(f"{fn}_{c}") for fn in fns for c in ("x", "f", "i")]
    want = E.eval_expression_list(t, exprs)
    assert [str(f.type) for f in want.schema][:3] == ["double", "float", "float"]
    assert_tables_equal(emulate_project(t, exprs), want, float_rtol=1e-6)
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    plan, _, _ = encode_project(exprs, ["x", "f", "i"], [dt.F64, dt.F32, dt.I16])
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(dt.F64, dt.F32, dt.I16), (ctypes.c_uint8 * 3)(1, 0, 0), ctypes.c_int32(3))


def test_lstrip_rstrip_right_find(tmp_path):
    """utf8 lstrip / rstrip (Rust trim_start / trim_end: the Unicode White_Space property, so U+3000 and U+00A0 go and U+001C stays),
    right(n) in characters, find() as a BYTE index (str::find): generated code == oracle, NVRTC"""
    import ctypes
    import pyarrow as pa
    from daft_b200 import _lib, col
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_project
    from jit_emu import emulate_project
    from oracle import expr_eval as E
    t = pa.table({"s": pa.array(["  foo ", "\t\nbar\r\n", "　日本 ", "", None, "   ", "x", "\x1cabc\x1c", "\xa0nb sp\x85", "héllo wörld", "no-space"], pa.large_string()),
                  "n": pa.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 100], pa.int64())})
    exprs = [col("s").lstrip().alias("l"), col("s").rstrip().alias("r"), col("s").lstrip().rstrip().alias("lr"), col("s").right(3).alias("r3"),
             col("s").right(col("n")).alias("rn"), col("s").find("o").alias("fo"), col("s").find("ö").alias("f2"), col("s").find("").alias("fe")]
    want = E.eval_expression_list(t, exprs)
    assert want.column("lr").to_pylist() == ["foo", "bar", "日本", "", None, "", "x", "\x1cabc\x1c", "nb sp", "héllo wörld", "no-space"]
    assert want.column("fo").to_pylist() == [3, -1, -1, -1, None, -1, -1, -1, -1, 5, 1] and want.column("f2").to_pylist()[9] == 8   # byte index: é is two bytes
    assert emulate_project(t, exprs).equals(want)
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    plan, _, odt = encode_project(exprs, ["s", "n"], [dt.LARGE_UTF8, dt.I64])
    assert odt == [dt.LARGE_UTF8] * 5 + [dt.I64] * 3
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 2)(dt.LARGE_UTF8, dt.I64), (ctypes.c_uint8 * 2)(1, 0), ctypes.c_int32(2))


def test_cast_utf8_to_integers(tmp_path):
    """Utf8 -> integer casts: the reference trims (Unicode whitespace) and parses with arrow-cast's integer parser; malformed or out-of-range
    text is null.  The reference's own cases (src/daft-core/src/array/ops/cast.rs:2088-2150) as known answers, the limits of every width,
    random digit strings: generated code == oracle, NVRTC"""
    import ctypes
    import numpy as np
    import pyarrow as pa
    from daft_b200 import _lib, col
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_project
    from jit_emu import emulate_project
    from oracle import expr_eval as E
    # cast.rs:2088-2100 (Int32), 2103-2116 (Int64), 2147-2160 (invalid -> None)
    t = pa.table({"s": pa.array(["  42  ", "-1", "  100  ", None, "  -9999999999  ", "\t123\n", "not_a_number", "  "], pa.large_string())})
    got = E.eval_expression_list(t, [col("s").cast(dt.I32).alias("i32"), col("s").cast(dt.I64).alias("i64")])
    assert got.column("i32").to_pylist() == [42, -1, 100, None, None, 123, None, None]
    assert got.column("i64").to_pylist() == [42, -1, 100, None, -9999999999, 123, None, None]
    rng = np.random.default_rng(5)
    vals = ["", "+", "-", "+5", "-0", "+0", "--1", "1-", "1 2", "0x10", "1e3", "1.0", "１２", "　7　", "\xa08\x85", "000000000000000000000000000012", "1_000",
            "127", "128", "-128", "-129", "255", "256", "32767", "32768", "-32768", "-32769", "65535", "65536", "2147483647", "2147483648",
            "-2147483648", "-2147483649", "4294967295", "4294967296", "9223372036854775807", "9223372036854775808", "-9223372036854775808",
            "-9223372036854775809", "18446744073709551615", "18446744073709551616", "99999999999999999999999", "-1", None, "٣"]
    for _ in range(400):
        digits = "".join(rng.choice(list("0123456789"), size=int(rng.integers(1, 22))))
        vals.append(rng.choice(["", "-", "+", " "]) + digits + rng.choice(["", "", " ", "x", "\n"]))
    t = pa.table({"s": pa.array(vals, pa.large_string())})
    targets = [dt.I8, dt.I16, dt.I32, dt.I64, dt.U8, dt.U16, dt.U32, dt.U64]
    exprs = [col("s").cast(c).alias(dt.NAMES[c]) for c in targets] + [(col("s").cast(dt.I64) + 1).alias("plus1"), col("s").cast(dt.U8).is_null().alias("bad")]
    want = E.eval_expression_list(t, exprs)
    d = dict(zip(vals, want.column(dt.NAMES[dt.U8]).to_pylist()))
    assert d["-0"] == 0 and d["+5"] == 5 and d["255"] == 255 and d["256"] is None and d["-1"] is None and d["　7　"] == 7 and d["１２"] is None
    d = dict(zip(vals, want.column(dt.NAMES[dt.I64]).to_pylist()))
    assert d["-9223372036854775808"] == -2**63 and d["9223372036854775808"] is None and d["000000000000000000000000000012"] == 12
    assert dict(zip(vals, want.column(dt.NAMES[dt.U64]).to_pylist()))["18446744073709551615"] == 2**64 - 1
    assert emulate_project(t, exprs).equals(want)
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    plan, _, odt = encode_project(exprs, ["s"], [dt.LARGE_UTF8])
    assert odt == targets + [dt.I64, dt.BOOL]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 1)(dt.LARGE_UTF8), (ctypes.c_uint8 * 1)(1), ctypes.c_int32(1))
    # the casts that stay unbuilt say so
    import pytest
    for bad in (col("s").cast(dt.F64), col("s").cast(dt.DATE32)):
        with pytest.raises(NotImplementedError):
            encode_project([bad], ["s"], [dt.LARGE_UTF8])


def test_cast_date_timestamp_units(tmp_path):
    """Date -> Timestamp(us) multiplies the days by the microseconds of a day (src/daft-core/src/array/ops/cast.rs:268-279); Timestamp -> Date
    is the day the instant falls on, floored for instants before the epoch (cast.rs:347 -> TimestampArray::date): generated code == oracle,
    known answers from pyarrow's own casts, NVRTC"""
    import ctypes
    import datetime
    import numpy as np
    import pyarrow as pa
    from daft_b200 import _lib, col
    from daft_b200 import datatypes as dt
    from daft_b200.expressions import encode_project
    from jit_emu import emulate_project
    from oracle import expr_eval as E
    rng = np.random.default_rng(9)
    n = 500
    days = np.concatenate([[0, 1, -1, 19000, -25567], rng.integers(-40000, 40000, n - 5)]).astype(np.int32)
    us = np.concatenate([[0, -1, 1, 86_400_000_000, -86_400_000_000, -86_400_000_001], rng.integers(-3 * 10**15, 3 * 10**15, n - 6)])
    t = pa.table({"d": pa.array(days, pa.date32(), mask=rng.random(n) < 0.1), "ts": pa.array(us, pa.timestamp("us"), mask=rng.random(n) < 0.1)})
    exprs = [col("d").cast(dt.TIMESTAMP).alias("d_ts"), col("ts").cast(dt.DATE32).alias("ts_d"), col("d").cast(dt.TIMESTAMP).cast(dt.DATE32).alias("back"),
             (col("ts").cast(dt.DATE32) == col("ts").date()).alias("same"), (col("d").cast(dt.TIMESTAMP) < datetime.datetime(2000, 1, 1)).alias("cmp")]
    want = E.eval_expression_list(t, exprs)
    assert want.column("d_ts").combine_chunks().equals(t.column("d").combine_chunks().cast(pa.timestamp("us")))
    assert want.column("ts_d").combine_chunks().equals(t.column("ts").combine_chunks().cast(pa.date32()))   # pyarrow floors as chrono does
    assert want.column("back").combine_chunks().equals(t.column("d").combine_chunks())
    assert emulate_project(t, exprs).equals(want)
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    plan, _, odt = encode_project(exprs, ["d", "ts"], [dt.DATE32, dt.TIMESTAMP])
    assert odt == [dt.TIMESTAMP, dt.DATE32, dt.DATE32, dt.BOOL, dt.BOOL]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 2)(dt.DATE32, dt.TIMESTAMP), (ctypes.c_uint8 * 2)(1, 1), ctypes.c_int32(2))

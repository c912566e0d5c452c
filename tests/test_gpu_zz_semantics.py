"""Expression / operator semantics pinned after the last GPU session of round 1: totalOrder float comparisons, arrow-rs
safe-cast nulls, Rust integer division, bitwise operators, take bounds, concat errors -- CUDA path (through the C ABI) against
the CPU oracle and against the reference's literal test cases.  Kept in its own (last-sorted) file so that a surprise here
cannot hide the results of the suites that were already verified on the B200.
"""
import datetime

import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal, random_table
from oracle import expr_eval as E

pytestmark = pytest.mark.gpu


def RB(gpu, tbl):
    return gpu.RecordBatch.from_arrow_table(tbl)


@pytest.mark.parametrize("n,nulls", [(0, 0.1), (100, 0.3), (4097, 0.1), (4097, 0.0)])
def test_division_modulo_bitwise_match_oracle(gpu, n, nulls):
    """`//` and `%` with Rust semantics (truncating integers, fmod, Float32 `//` through Float64) and bitwise and / or / xor over
    the mixed-sign columns of random_table (zero divisors included)."""
    rng = np.random.default_rng(n + 17)
    tbl = random_table(rng, n, nulls)
    c = gpu.col
    exprs = [
        (c("i64") // c("i16")).alias("e3"), (c("i64") % 7).alias("e4"), (c("i64") % -7).alias("e4n"), (c("f64") % c("f32")).alias("e37"),
        (c("f64") // 7.0).alias("e38"), (c("i32") % c("i8")).alias("e39"), (c("i32") // c("i8")).alias("e39d"), (c("f32") // c("i8")).alias("e40"),
        (c("f32") % c("i8")).alias("e41"), (c("i64") & c("i32")).alias("e42"), (c("u8") | c("u16")).alias("e43"), (c("i16") ^ 255).alias("e44"),
        (c("u32") // c("u8")).alias("e45"), (c("u32") % c("u8")).alias("e46"),
    ]
    got = RB(gpu, tbl).eval_expression_list(exprs).to_arrow_table()
    want = E.eval_expression_list(tbl, exprs)
    assert_tables_equal(got, want)


def test_float_comparisons_total_order(gpu):
    """NaN / signed-zero / infinity operands: the comparison operators order floats by IEEE totalOrder (arrow-rs `cmp` kernels,
    src/daft-core/src/array/ops/comparison.rs:541-603); float is_in treats all NaNs as one element (is_in.rs:44-66)."""
    rng = np.random.default_rng(11)
    n = 5000
    special = np.array([np.nan, -np.nan, 0.0, -0.0, np.inf, -np.inf, 1.0, -1.0, 0.5])
    x = special[rng.integers(0, len(special), n)]
    y = special[rng.integers(0, len(special), n)]
    for typ in (pa.float64(), pa.float32()):
        tbl = pa.table({"x": pa.array(x, type=typ, mask=rng.random(n) < 0.1), "y": pa.array(y, type=typ, mask=rng.random(n) < 0.1)})
        c = gpu.col
        exprs = [(c("x") == c("y")).alias("eq"), (c("x") != c("y")).alias("ne"), (c("x") < c("y")).alias("lt"), (c("x") <= c("y")).alias("le"),
                 (c("x") > c("y")).alias("gt"), (c("x") >= c("y")).alias("ge"), (c("x") > 0.0).alias("gt0"), (c("x") == 0.0).alias("eq0"),
                 c("x").between(0.0, float("inf")).alias("bt"), c("x").between(c("y"), 1.0).alias("bt2"), c("x").is_in([float("nan"), 0.0, 0.5]).alias("in"),
                 c("x").eq_null_safe(c("y")).alias("ens")]
        got = RB(gpu, tbl).eval_expression_list(exprs).to_arrow_table()
        want = E.eval_expression_list(tbl, exprs)
        assert_tables_equal(got, want)
        for pred in ([c("x") > 0.0], [c("x") <= c("y")], [c("x") == c("y")], [c("x").is_in([float("nan")])]):
            g, w = RB(gpu, tbl).filter(pred).to_arrow_table(), E.filter(tbl, pred)
            assert g.num_rows == w.num_rows
            assert_tables_equal(g, w)


def test_cast_out_of_range_is_null(gpu):
    """arrow-rs safe-cast rule (src/daft-core/src/array/ops/cast.rs:194-196): what the target integer type cannot hold is null."""
    rng = np.random.default_rng(5)
    n = 3000
    edge_i = np.array([0, 127, 128, -128, -129, 255, 256, -1, 32767, 32768, -32769, 2**31 - 1, 2**31, -2**31 - 1, 2**32, 2**63 - 1, -2**63], dtype=np.int64)
    edge_f = np.array([0.9, -0.9, 127.9, 128.0, -128.9, -129.0, np.nan, np.inf, -np.inf, 255.5, 256.0, 2147483647.5, 2147483648.0,
                       -2147483648.9, 9.3e18, -9.3e18, 1.8e19, 1.9e19, 4294967295.9, 4294967296.0])
    i64 = np.concatenate([edge_i, rng.integers(-70000, 70000, n), rng.integers(-2**62, 2**62, n)])
    f64 = np.concatenate([edge_f, rng.standard_normal(n) * 300, rng.standard_normal(n // 2) * 1e10, rng.standard_normal(n) * 1e19])[:len(i64)]
    assert len(f64) == len(i64)
    tbl = pa.table({
        "i64": pa.array(i64, mask=rng.random(len(i64)) < 0.1), "u64": pa.array(i64.view(np.uint64)), "i16": pa.array((i64 & 0xffff).astype(np.uint16).view(np.int16)),
        "f64": pa.array(f64, mask=rng.random(len(i64)) < 0.1), "f32": pa.array(f64.astype(np.float32)),
    })
    c = gpu.col
    tgts = [pa.int8(), pa.int16(), pa.int32(), pa.int64(), pa.uint8(), pa.uint16(), pa.uint32(), pa.uint64(), pa.float32(), pa.float64()]
    exprs = [c(nm).cast(t).alias(f"{nm}_{t}") for nm in tbl.column_names for t in tgts]
    got = RB(gpu, tbl).eval_expression_list(exprs).to_arrow_table()
    want = E.eval_expression_list(tbl, exprs)
    assert want.column("i64_int8").null_count > n and want.column("i16_int64").null_count == 0
    assert_tables_equal(got, want)


def test_concat_literals_and_errors_from_reference_tests(gpu):
    """tests/recordbatch/test_concat.py:8-73."""
    RBc = gpu.RecordBatch
    a = RB(gpu, pa.table({"x": [1, 2, 3], "y": pa.array(["a", "b", "c"], type=pa.large_utf8())}))
    b = RB(gpu, pa.table({"x": [4, 5, 6], "y": pa.array(["d", "e", "f"], type=pa.large_utf8())}))
    assert RBc.concat([a, b]).to_pydict() == {"x": [1, 2, 3, 4, 5, 6], "y": ["a", "b", "c", "d", "e", "f"]}
    e = pa.table({"x": pa.array([], type=pa.int64()), "y": pa.array([], type=pa.large_utf8())})
    assert RBc.concat([RB(gpu, e), RB(gpu, e)]).to_pydict() == {"x": [], "y": []}
    x = RB(gpu, pa.table({"x": [1, 2, 3]}))
    with pytest.raises(TypeError, match="Expected a RecordBatch for concat"):
        RBc.concat([x, []])
    with pytest.raises(ValueError, match="Need at least 1 RecordBatch"):
        RBc.concat([])
    for other in ({"y": [1, 2, 3]}, {"x": [1.0, 2.0, 3.0]}, {"x": [1, 2, 3], "y": [2, 3, 4]}):
        with pytest.raises(ValueError, match="RecordBatch::concat requires all schemas to match"):
            RBc.concat([x, RB(gpu, pa.table(other))])


def test_take_literals_and_bounds_from_reference_tests(gpu):
    """tests/recordbatch/test_take.py:15-70: empty take, repeated / reversed indices, out-of-bounds indices are an error."""
    empty = RB(gpu, pa.table({"a": pa.array([], type=pa.int64())}))
    assert empty.take(pa.array([], type=pa.int64())).to_pydict() == {"a": []}
    with pytest.raises(Exception, match="index out of bounds"):
        empty.take(pa.array([1], type=pa.int64()))
    rb = RB(gpu, pa.table({"a": [1, 2, 3, 4]}))
    assert rb.take(pa.array([0, 1], type=pa.int64())).to_pydict() == {"a": [1, 2]}
    assert rb.take(pa.array([3, 2], type=pa.int64())).to_pydict() == {"a": [4, 3]}
    assert rb.take(pa.array([3, 2, 2, 2, 3], type=pa.int64())).to_pydict() == {"a": [4, 3, 3, 3, 4]}
    for typ in (pa.int8(), pa.int16(), pa.int32(), pa.uint8(), pa.uint16(), pa.uint32(), pa.uint64()):
        assert rb.take(pa.array([3, 0, None], type=typ)).to_pydict() == {"a": [4, 1, None]}
    for bad in ([4], [0, 1, 2, 3, 4], [-1], [2**40]):
        with pytest.raises(Exception, match="index out of bounds"):
            rb.take(pa.array(bad, type=pa.int64()))
    with pytest.raises(Exception, match="index out of bounds"):
        rb.take(pa.array([1, None, 7], type=pa.int32()))


def test_float_predicates_ceil_floor_sign_fill_nan(gpu):
    """is_nan / not_nan / is_inf, ceil / floor / sign, fill_nan against the oracle over NaN / infinity / signed-zero inputs, plus the
    reference's literal cases (tests/recordbatch/test_filter.py:223-241, test_fill.py:49-62)."""
    rng = np.random.default_rng(23)
    n = 4000
    special = np.array([np.nan, -np.nan, 0.0, -0.0, np.inf, -np.inf, 1.5, -1.5, 0.5, -0.5, 2.0, 1e300, -1e-300])
    x = np.where(rng.random(n) < 0.5, special[rng.integers(0, len(special), n)], rng.standard_normal(n) * 100)
    tbl = pa.table({"f64": pa.array(x, mask=rng.random(n) < 0.1), "f32": pa.array(x.astype(np.float32), mask=rng.random(n) < 0.1),
                    "i32": pa.array(rng.integers(-5, 5, n).astype(np.int32), mask=rng.random(n) < 0.1), "u16": pa.array(rng.integers(0, 3, n).astype(np.uint16))})
    c = gpu.col
    exprs = [e for nm in ("f64", "f32") for e in (c(nm).is_nan().alias(nm + "_nan"), c(nm).not_nan().alias(nm + "_nn"), c(nm).is_inf().alias(nm + "_inf"),
                                                   c(nm).ceil().alias(nm + "_ceil"), c(nm).floor().alias(nm + "_floor"), c(nm).sign().alias(nm + "_sign"),
                                                   c(nm).fill_nan(2.0).alias(nm + "_fn"), c(nm).fill_nan(c("f64")).alias(nm + "_fn2"))]
    exprs += [c("i32").sign().alias("i_sign"), c("u16").sign().alias("u_sign"), c("i32").ceil().alias("i_ceil"), c("u16").floor().alias("u_floor"),
              c("f32").sqrt().alias("f32_sqrt"), c("f64").sqrt().alias("f64_sqrt"), c("u16").sqrt().alias("u16_sqrt"), c("i32").sqrt().alias("i32_sqrt")]
    got = RB(gpu, tbl).eval_expression_list(exprs).to_arrow_table()
    want = E.eval_expression_list(tbl, exprs)
    assert_tables_equal(got, want)
    nan = float("nan")
    rb = RB(gpu, pa.table({"a": pa.array([1.0, nan, 3.0, None, nan])}))
    assert rb.eval_expression_list([c("a").is_nan()]).to_pydict() == {"a": [False, True, False, None, True]}
    assert rb.eval_expression_list([c("a").not_nan()]).to_pydict() == {"a": [True, False, True, None, False]}
    assert rb.filter([c("a").not_nan()]).to_pydict() == {"a": [1.0, 3.0]}
    for typ in (pa.float32(), pa.float64()):
        rb = RB(gpu, pa.table({"input": pa.array([1.0, None, 3.0, nan], type=typ)}))
        assert rb.eval_expression_list([c("input").fill_nan(2.0)]).to_pydict() == {"input": [1.0, None, 3.0, 2.0]}

"""Shared comparison helpers for the parity tests (oracle vs CUDA path on the same Arrow data)."""
from __future__ import annotations

import math

import numpy as np
import pyarrow as pa

from oracle import expr_eval as E
from oracle import oracle as O


def norm(tbl: pa.Table) -> pa.Table:
    """Normalise logical types both sides agree on: strings/binary Large, timestamps us."""
    cols = []
    for c in tbl.columns:
        c = c.combine_chunks() if isinstance(c, pa.ChunkedArray) else c
        if pa.types.is_string(c.type):
            c = c.cast(pa.large_utf8())
        elif pa.types.is_binary(c.type):
            c = c.cast(pa.large_binary())
        cols.append(c)
    return pa.table(cols, names=tbl.column_names)


def assert_tables_equal(got: pa.Table, want: pa.Table, float_rtol: float = 0.0, check_order: bool = True, sort_by=None):
    got, want = norm(got), norm(want)
    assert got.column_names == want.column_names, (got.column_names, want.column_names)
    assert got.num_rows == want.num_rows, (got.num_rows, want.num_rows)
    if sort_by is not None:
        got = got.sort_by([(k, "ascending") for k in sort_by])
        want = want.sort_by([(k, "ascending") for k in sort_by])
    for name in want.column_names:
        g, w = got.column(name).combine_chunks(), want.column(name).combine_chunks()
        assert g.type == w.type, f"{name}: {g.type} vs {w.type}"
        gl, wl = g.to_pylist(), w.to_pylist()
        if pa.types.is_floating(w.type):
            for i, (a, b) in enumerate(zip(gl, wl)):
                if a is None or b is None:
                    assert a is None and b is None, f"{name}[{i}]: {a} vs {b}"
                elif math.isnan(b):
                    assert math.isnan(a), f"{name}[{i}]: {a} vs nan"
                elif float_rtol == 0.0:
                    assert a == b or (a == 0 and b == 0), f"{name}[{i}]: {a!r} vs {b!r}"
                elif a == b:
                    pass  # also covers equal infinities, where the relative test below would compute inf - inf
                else:
                    # float32 accumulators: the summation order differs from the sequential CPU fold
                    tol = max(float_rtol, 2e-4) if pa.types.is_float32(w.type) else float_rtol
                    assert abs(a - b) <= tol * max(abs(a), abs(b), 1e-30) + (1e-30 if tol else 0), f"{name}[{i}]: {a!r} vs {b!r}"
        else:
            assert gl == wl, f"{name}: first diff at {next(i for i, (a, b) in enumerate(zip(gl, wl)) if a != b)}"


def oracle_agg(tbl: pa.Table, aggs, group_by, predicate=None) -> pa.Table:
    """Filter -> Project -> agg_groupby_inline the way the reference chains them."""
    if predicate:
        tbl = E.filter(tbl, predicate)
    specs = []
    for a in aggs:
        inner = a
        while inner.kind == "alias":
            inner = inner.args[0]
        specs.append((inner.op, E.eval_expression(tbl, inner.args[0]), a.name(), inner.value))
    keys = [(g.name(), E.eval_expression(tbl, g)) for g in group_by]
    return O.agg(tbl, specs, keys)


def random_table(rng: np.random.Generator, n: int, null_frac: float = 0.1) -> pa.Table:
    def maybe_null(arr, typ=None):
        mask = rng.random(n) < null_frac if null_frac > 0 else None
        return pa.array(arr, type=typ, mask=mask)
    words = ["", "a", "bb", "ccc", "dddd", "hello world", "TPC-H lineitem comment " * 3, "x" * 130, "y" * 300, "é✓ unicode"]
    return pa.table({
        "i8": maybe_null(rng.integers(-5, 5, n).astype(np.int8)),
        "i16": maybe_null(rng.integers(-300, 300, n).astype(np.int16)),
        "i32": maybe_null(rng.integers(-1000, 1000, n).astype(np.int32)),
        "i64": maybe_null(rng.integers(-10**12, 10**12, n).astype(np.int64)),
        "u8": maybe_null(rng.integers(0, 7, n).astype(np.uint8)),
        "u16": maybe_null(rng.integers(0, 70, n).astype(np.uint16)),
        "u32": maybe_null(rng.integers(0, 7000, n).astype(np.uint32)),
        "u64": maybe_null(rng.integers(0, 2**63, n).astype(np.uint64)),
        "f32": maybe_null(rng.standard_normal(n).astype(np.float32)),
        "f64": maybe_null(rng.standard_normal(n) * 1e3),
        "b": maybe_null(rng.random(n) < 0.5),
        "d": maybe_null(rng.integers(8000, 11000, n).astype(np.int32)).cast(pa.date32()),
        "s": maybe_null([words[i] for i in rng.integers(0, len(words), n)], pa.large_utf8()),
        "k": pa.array(rng.integers(0, 13, n).astype(np.int64)),
    })

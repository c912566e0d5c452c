"""The JIT-generated projection code, executed on the host (tests/jit_emu.py) against the CPU oracle.

No GPU involved: the library only generates the source (`b200_jit_describe`); the same text is compiled with g++ and run warp
by warp.  This is the CPU-side evidence that the expression semantics the device tests check (tests/test_gpu_parity.py,
tests/test_gpu_zz_semantics.py) are what the code generator emits -- the device tests then only add the launch itself.
"""
import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import daft_b200
import jit_emu
import test_gpu_parity as parity
import test_gpu_zz_semantics as semantics
from helpers import assert_tables_equal, random_table
from oracle import expr_eval as E


class _Result:
    def __init__(self, tbl):
        self._t = tbl

    def to_arrow_table(self):
        return self._t

    def to_pydict(self):
        return self._t.to_pydict()


class _EmulatedBatch:
    """Stands in for daft_b200.RecordBatch: projections (and filter predicates) run through the generated code on the host."""

    def __init__(self, tbl):
        self._t = tbl

    @classmethod
    def from_arrow_table(cls, tbl):
        return cls(tbl)

    def eval_expression_list(self, exprs):
        return _Result(jit_emu.emulate_project(self._t, exprs))

    def filter(self, preds):
        pred = preds[0]
        for p in preds[1:]:
            pred = pred & p
        mask = jit_emu.emulate_project(self._t, [pred.alias("__mask")]).column("__mask")
        return _Result(self._t.filter(pc.fill_null(mask, False)))


class _EmulatedModule:
    col = staticmethod(daft_b200.col)
    lit = staticmethod(daft_b200.lit)
    RecordBatch = _EmulatedBatch


EMU = _EmulatedModule()

@pytest.mark.parametrize("n,nulls", [(0, 0.1), (1, 0.0), (100, 0.3), (4097, 0.1), (4097, 0.0)])
def test_generated_projection_code_matches_oracle(n, nulls):
    rng = np.random.default_rng(n + 7)
    tbl = random_table(rng, n, nulls)
    exprs = parity.exprs_under_test(EMU.col, EMU.lit)   # string outputs included (two-pass length / copy protocol)
    got = jit_emu.emulate_project(tbl, exprs)
    want = E.eval_expression_list(tbl, exprs)
    assert_tables_equal(got, want)


@pytest.mark.parametrize("n", [5, 1000, 10007])
def test_generated_filter_predicates_match_oracle(n):
    parity.test_filter_matches_oracle(EMU, n)


@pytest.mark.parametrize("n,nulls", [(0, 0.1), (100, 0.3), (4097, 0.1), (4097, 0.0)])
def test_generated_division_modulo_bitwise(n, nulls):
    semantics.test_division_modulo_bitwise_match_oracle(EMU, n, nulls)


def test_generated_float_comparisons_total_order():
    semantics.test_float_comparisons_total_order(EMU)


def test_generated_casts_out_of_range_null():
    semantics.test_cast_out_of_range_is_null(EMU)


def test_generated_float_predicates_ceil_floor_sign_fill_nan():
    semantics.test_float_predicates_ceil_floor_sign_fill_nan(EMU)


@pytest.mark.parametrize("seed", [300, 301])
def test_random_expression_trees_match_oracle(seed):
    """Two seeded batches of the expression fuzzer (scripts/fuzz_exprs.py: arithmetic, division, casts, comparisons, Kleene logic,
    bitwise, if_else, fill_null, between, is_in, float predicates, string and date functions, depth <= 3): one projection of 40
    random trees each through the generated code, against the oracle.  Longer runs: `python scripts/fuzz_exprs.py 60 40 1000`."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import fuzz_exprs
    bad = fuzz_exprs.run_batch(seed, 40)
    assert not bad, bad

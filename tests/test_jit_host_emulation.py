"""The JIT-generated projection code, executed on the host (tests/jit_emu.py) against the CPU oracle.

No GPU involved: the library only generates the source (`b200_jit_describe`); the same text is compiled with g++ and run warp
by warp.  This is the CPU-side evidence that the expression semantics the device tests check (tests/test_gpu_parity.py,
tests/test_gpu_zz_semantics.py) are what the code generator emits -- the device tests then only add the launch itself.
"""
import datetime

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

import daft_b200
from daft_b200 import col
import jit_emu
from jit_emu import emulate_project
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

    AGG_MODE: dict = {}   # which generated accumulate path agg() drives: {} = global_update, {"tier1": True} = lane cells + flush

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


    def agg(self, aggs, group_by=None, predicate=None):
        """Aggregate columns through the generated aggregate code (jit_emu.emulate_agg); the key columns are the group-by
        expressions at each group's first row, which is how op_agg_first_rows builds them (a projection at the first rows)."""
        group_by = list(group_by or [])
        aggs = list(aggs)

        def inner(a):
            while a.kind == "alias":
                a = a.args[0]
            return a
        # let the library's code generator see the plan as submitted (it refuses what the device does not offer)
        names = self._t.column_names
        dts = [daft_b200.datatypes.from_arrow(self._t.schema.field(n).type) for n in names]
        from daft_b200.expressions import encode_agg
        jit_emu.describe_agg(encode_agg(aggs, group_by, names, dts, predicate), dts, [1 if self._t.column(n).null_count else 0 for n in names])
        anyv = [a for a in aggs if inner(a).op == "any_value"]
        ign = [a for a in anyv if inner(a).value == "ignore_nulls"]   # these keep a "first valid row" cell in the aggregate kernel
        plain = [a for a in aggs if inner(a).op != "any_value"] or [col(self._t.column_names[0]).count("all").alias("__rows")]
        takes = {}
        cols, first = jit_emu.emulate_agg(self._t, plain + ign, group_by, predicate, takes_out=takes, **self.AGG_MODE)
        out = {n: cols.column(n) for n in cols.column_names}
        for a in anyv:   # any_value = the child expression at the group's first (valid) row: a gathering projection (op_agg_first_rows)
            take = takes.get(a.name(), first)
            at = jit_emu.emulate_project(self._t, [inner(a).args[0].alias(a.name())], take=take) if self._t.num_rows else \
                pa.table({a.name(): pa.nulls(len(first))})
            out[a.name()] = at.column(a.name())
        ordered = [out[a.name()] for a in aggs]
        names = [a.name() for a in aggs]
        if not group_by:
            return _Result(pa.table(ordered, names=names))
        keys = jit_emu.emulate_project(self._t, group_by, take=first)   # group keys: the same gathering projection
        return _Result(pa.table(keys.columns + ordered, names=keys.column_names + names))


class _EmulatedModule:
    col = staticmethod(daft_b200.col)
    lit = staticmethod(daft_b200.lit)
    RecordBatch = _EmulatedBatch
    B200Error = daft_b200._lib.B200Error


EMU = _EmulatedModule()


class _EmulatedBatchTier1(_EmulatedBatch):
    AGG_MODE = {"tier1": True}


class _EmulatedModuleTier1(_EmulatedModule):
    RecordBatch = _EmulatedBatchTier1


EMU_T1 = _EmulatedModuleTier1()

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
    """Two seeded batches of the expression fuzzer (tests/fuzz_exprs.py: arithmetic, division, casts, comparisons, Kleene logic,
    bitwise, if_else, fill_null, between, is_in, float predicates, string and date functions, depth <= 3): one projection of 40
    random trees each through the generated code, against the oracle.  Longer runs: `python tests/fuzz_exprs.py 60 40 1000`."""
    import fuzz_exprs
    bad = fuzz_exprs.run_batch(seed, 40)
    assert not bad, bad


@pytest.mark.parametrize("n,nulls", [(0, 0.1), (1, 0.0), (1000, 0.2), (5000, 0.0), (5000, 0.3)])
def test_generated_aggregate_code_matches_oracle(n, nulls):
    """The group-by matrix of the device parity test (16 aggregates x 6 key shapes) through the generated load / eval / keys_equal /
    code / global_update functions, read back the way op_agg_first_rows finalises the cells."""
    parity.test_groupby_inline_aggs(EMU, n, nulls)


def test_generated_aggregate_ungrouped_predicate_null_and_nan_keys():
    parity.test_ungrouped_and_fused_predicate(EMU)
    parity.test_groupby_null_and_nan_keys(EMU)


def test_generated_tpch_q1_q6_match_golden_answers():
    """TPC-H Q1 (the headline plan) and Q6 at SF 0.01 through the generated aggregate code, against the committed answers."""
    import json
    import os
    from daft_b200 import tpch
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tpch_q1_q6_sf001.json")))
    li = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS)
    predicate, group_by, aggs = tpch.q1_exprs()
    got = _EmulatedBatch(li).agg(aggs, group_by, predicate).to_arrow_table().sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")]).to_pydict()
    want = golden["q1"]
    assert set(got) == set(want)
    for name, vals in want.items():
        for a, b in zip(got[name], vals):
            assert a == b or abs(a - b) <= 1e-9 * abs(b), (name, a, b)
    # the two-stage form the multi-GPU path runs: partial aggregate -> final aggregate of the partials -> merge_mean projection
    predicate, group_by, partial = tpch.q1_partial_exprs()
    parts = _EmulatedBatch(li.slice(0, 30000)).agg(partial, group_by, predicate).to_arrow_table(), \
        _EmulatedBatch(li.slice(30000)).agg(partial, group_by, predicate).to_arrow_table()
    group_by, final, project = tpch.q1_final_exprs()
    merged = _EmulatedBatch(pa.concat_tables(parts).combine_chunks()).agg(final, group_by).to_arrow_table()
    two = _EmulatedBatch(merged).eval_expression_list(project).to_arrow_table().sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")]).to_pydict()
    for name, vals in want.items():
        for a, b in zip(two[name], vals):
            assert a == b or abs(a - b) <= 1e-9 * abs(b), (name, a, b)
    pred = [col("L_SHIPDATE") >= datetime.date(1994, 1, 1), col("L_SHIPDATE") < datetime.date(1995, 1, 1), col("L_DISCOUNT") >= 0.05,
            col("L_DISCOUNT") <= 0.07, col("L_QUANTITY") < 24]
    got6 = _EmulatedBatch(li).agg([(col("L_EXTENDEDPRICE") * col("L_DISCOUNT")).sum().alias("revenue")], [], pred).to_pydict()["revenue"][0]
    want6 = golden["q6"]["revenue"][0]
    assert abs(got6 - want6) <= 1e-9 * abs(want6)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_random_aggregate_plans_match_oracle(seed):
    """Seeded plans of the aggregate fuzzer (tests/fuzz_aggs.py: random predicate, 0-3 group keys, 1-6 aggregates over random
    expression trees) through the generated aggregate code.  Longer runs: `python tests/fuzz_aggs.py 60 100`."""
    import fuzz_aggs
    assert fuzz_aggs.run(seed) is None


@pytest.mark.parametrize("nulls", [0.0, 0.2])
def test_generated_staged_aggregate_build(nulls):
    """The TMA-staged build of ungrouped plans (what Q6-shaped scans run): stage_issue lays the column slices out in the stage,
    load_staged reads them back; full 768-row tiles go through the stage, the tail through the direct loads."""
    import json
    import os
    from daft_b200 import tpch
    rng = np.random.default_rng(41)
    tbl = random_table(rng, 5000, nulls).drop_columns(["s"])   # string bytes are not staged; keep the fixed-width columns
    c = col
    aggs = [c("i64").sum().alias("s"), c("f64").mean().alias("m"), c("u8").max().alias("mx"), c("i32").count().alias("cnt"), (c("f64") * c("f32")).sum().alias("e")]
    pred = [c("i32") > -500, c("f64") <= 900.0]
    got, _ = jit_emu.emulate_agg(tbl, aggs, [], pred, staged=True)
    from helpers import oracle_agg
    assert_tables_equal(got, oracle_agg(tbl, aggs, [], pred), float_rtol=1e-12)
    li = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS)
    pred = [col("L_SHIPDATE") >= datetime.date(1994, 1, 1), col("L_SHIPDATE") < datetime.date(1995, 1, 1), col("L_DISCOUNT") >= 0.05,
            col("L_DISCOUNT") <= 0.07, col("L_QUANTITY") < 24]
    got, _ = jit_emu.emulate_agg(li, [(col("L_EXTENDEDPRICE") * col("L_DISCOUNT")).sum().alias("revenue")], [], pred, staged=True)
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tpch_q1_q6_sf001.json")))
    assert abs(got.column("revenue")[0].as_py() - golden["q6"]["revenue"][0]) <= 1e-9 * golden["q6"]["revenue"][0]


@pytest.mark.parametrize("n,nulls", [(0, 0.1), (5, 0.1), (1000, 0.1), (10007, 0.0), (10007, 0.3)])
def test_generated_filter_keep_and_copy(n, nulls):
    """Filter plans proper: the generated keep() and the column-major copy() of the fixed-width never-null columns (the other
    columns are gathered by the host on the device, with pyarrow here)."""
    rng = np.random.default_rng(n)
    tbl = random_table(rng, n, nulls)
    c = col
    for pred in ([c("i32") > 0], [c("b")], [c("f64") < 0.0, c("s") != "a"], [c("i64").is_null() | (c("u8") == 3)], [c("i32") > 5000],
                 [c("d") <= datetime.date(1998, 9, 2)], [c("k") % 3 == 1, c("f32").not_nan()]):
        got, copied = jit_emu.emulate_filter(tbl, pred)
        assert_tables_equal(got, E.filter(tbl, pred))
        assert 13 in copied   # "k" never has nulls: always copied by the generated code


def test_generated_global_build_tier2_functions():
    """The global-table build (mid / high cardinality): groups below T2 accumulate through t2_init / t2_update and are merged with
    t2_flush, the rest go straight to global_update -- same results as the oracle either way."""
    from helpers import oracle_agg
    rng = np.random.default_rng(3)
    tbl = random_table(rng, 4000, 0.2)
    c = col
    aggs = [c("f64").sum().alias("sf"), c("i32").max().alias("mx"), c("i32").min().alias("mn"), c("i64").count().alias("n"), c("u8").mean().alias("avg"),
            c("f32").sum().alias("s32"), c("b").max().alias("bmax"), c("d").min().alias("dmin"), c("u64").sum().alias("su")]
    for gb in ([c("k"), c("s")], [c("i32")], [c("u32"), c("b")]):
        got, first = jit_emu.emulate_agg(tbl, aggs, gb, [c("f64") > -500.0], tier2=True)
        want = oracle_agg(tbl, aggs, gb, [c("f64") > -500.0])
        assert_tables_equal(got, want.select([a.name() for a in aggs]), float_rtol=1e-12)


@pytest.mark.parametrize("n", [1, 700, 5000])
def test_generated_any_value_and_bool_aggs(n):
    """any_value (a gathering projection at the groups' first rows) and the boolean min / max aggregates."""
    parity.test_any_value_and_bool_aggs(EMU, n)


def test_generated_gathering_projection_null_rows():
    """The projection build that evaluates at gathered rows (group keys, any_value, join outputs): a negative index is a row of
    nulls, so `is_null` turns true there and literals stay valid -- compared with the oracle on the taken table."""
    rng = np.random.default_rng(5)
    tbl = random_table(rng, 1000, 0.2)
    idx = np.where(rng.random(700) < 0.15, -1, rng.integers(0, 1000, 700))
    exprs = parity.exprs_under_test(EMU.col, EMU.lit)
    got = jit_emu.emulate_project(tbl, exprs, take=idx)
    want = E.eval_expression_list(tbl.take(pa.array([None if i < 0 else int(i) for i in idx], type=pa.int64())), exprs)
    assert_tables_equal(got, want)


@pytest.mark.parametrize("n,nulls", [(0, 0.1), (1, 0.0), (1000, 0.2), (5000, 0.3)])
def test_generated_tier1_accumulate_path_matches_oracle(n, nulls):
    """The low-cardinality build's own accumulate path -- what the Q1 kernel runs per row: per-lane private cells (lane_init /
    lane_update), ballot count cells (warp_cells) and the warp-wide butterfly flush into the global cells (shuffles replayed to
    their fixed point) -- over the group-by matrix of the device parity test, plus the ungrouped / predicate / null-key cases."""
    parity.test_groupby_inline_aggs(EMU_T1, n, nulls)
    if n == 1000:
        parity.test_ungrouped_and_fused_predicate(EMU_T1)
        parity.test_groupby_null_and_nan_keys(EMU_T1)


def test_generated_tier1_tpch_q1_matches_golden_answers():
    import json
    import os
    from daft_b200 import tpch
    golden = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tpch_q1_q6_sf001.json")))
    li = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS)
    predicate, group_by, aggs = tpch.q1_exprs()
    got = _EmulatedBatchTier1(li).agg(aggs, group_by, predicate).to_arrow_table().sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")]).to_pydict()
    for name, vals in golden["q1"].items():
        for a, b in zip(got[name], vals):
            assert a == b or abs(a - b) <= 1e-9 * abs(b), (name, a, b)


def test_generated_round_and_match_patterns():
    """round (src/daft-core/src/array/ops/round.rs:14-35) and utf8 match / like over literal-segment patterns (the shapes of
    benchmarking/tpch/answers.py:365,402 and the LIKE predicates of TPC-H Q13 / Q16 / ClickBench) through the generated code."""
    rng = np.random.default_rng(21)
    n = 600
    x = np.concatenate([rng.standard_normal(n - 8) * 1000, [0.5, -0.5, 2.5, 0.125, -0.125, 1e300, np.nan, 0.49999999999999994]])
    words = ["special requests", "the special packages requests", "requests special", "", "Customer Complaints", "xCustomeryComplaintsz", "special",
             "specialrequests", "PROMO BRUSHED", "MEDIUM POLISHED TIN", "abc", "ab", "abcabc"]
    tbl = pa.table({"f64": pa.array(x, mask=rng.random(n) < 0.1), "f32": pa.array(x.astype(np.float32), mask=rng.random(n) < 0.1),
                    "s": pa.array([words[i] for i in rng.integers(0, len(words), n)], type=pa.large_utf8(), mask=rng.random(n) < 0.1)})
    exprs = [col("f64").round(2).alias("r2"), col("f64").round(0).alias("r0"), col("f32").round(1).alias("r1f"), col("f64").round(-2).alias("rm2"),
             col("s").match(".*special.*requests.*").alias("m1"), col("s").match("special").alias("m2"), col("s").match("^special.*requests$").alias("m3"),
             col("s").like("%Customer%Complaints%").alias("l1"), col("s").like("PROMO%").alias("l2"), col("s").like("%TIN").alias("l3"),
             col("s").like("abc").alias("l4"), col("s").like("ab%c").alias("l5"), col("s").like("%").alias("l6"), (~col("s").like("MEDIUM POLISHED%")).alias("l7")]
    got = emulate_project(tbl, exprs)
    assert_tables_equal(got, E.eval_expression_list(tbl, exprs))
    d = got.to_pydict()
    for i, w in enumerate(tbl.column("s").to_pylist()):
        if w is not None:
            assert d["m1"][i] == (w.find("special") >= 0 and w.find("requests", w.find("special") + 7) >= 0)
            assert d["l5"][i] == (w.startswith("ab") and w.endswith("c") and len(w) >= 3)
    assert col("s").match("a+b").op == "rx_match"          # beyond literal segments: the regex VM (tests/test_regex_cpu.py)
    with pytest.raises(ValueError):
        col("s").match("(?=a)b")                            # outside the supported regex subset
    with pytest.raises(NotImplementedError):
        col("s").like("a_c")

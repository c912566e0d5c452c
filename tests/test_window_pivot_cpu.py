"""Window functions and pivot: the host composition of daft_b200/window.py run on the CPU (generated expression / aggregate
code through the host emulation, joins / sorts / dedup by the oracle, the two native window calls restated in numpy) against
the oracle's per-partition restatement of the reference kernels (oracle/window.py), which is pinned on the reference's
literal expectations (tests/window/test_partition_only.py etc., tests/recordbatch/test_pivot.py)."""
import numpy as np
import pyarrow as pa
import pytest

from daft_b200 import col, window as W
from helpers import assert_tables_equal
from oracle import window as OW
from test_tpch_all_emulated_cpu import EmulatedQueryBatch


def _sorted(tbl, keys):
    return tbl.sort_by([(k, "ascending") for k in keys])


def window_table(rng, n, nulls=0.1):
    return pa.table({
        "id": np.arange(n),
        "g": pa.array(np.array([f"p{v}" for v in rng.integers(0, 5, n)], dtype=object), type=pa.large_string(), mask=rng.random(n) < nulls),
        "h": pa.array(rng.integers(0, 3, n)),
        "o": pa.array(rng.integers(0, 8, n), mask=rng.random(n) < nulls),
        "x": pa.array(rng.standard_normal(n).round(3), mask=rng.random(n) < nulls),
        "s": pa.array(np.array([f"v{v}" for v in rng.integers(0, 50, n)], dtype=object), type=pa.large_string(), mask=rng.random(n) < nulls),
    })


FUNCS = [("row_number", "rn"), ("rank", "rk"), ("dense_rank", "dr"), ("offset", "lead1", col("x"), 1), ("offset", "lag2", col("x"), -2, 0.5),
         ("offset", "lead_s", col("s"), 3), ("offset", "self", col("h") + 1, 0), ("agg", "psum", col("x").sum()), ("agg", "pcnt", col("o").count()),
         ("agg", "pmax", col("h").max())]


def test_oracle_window_reference_literals():
    """tests/window/test_order_by_only.py-style expectations: rank ties, dense rank, lead / lag with defaults"""
    t = pa.table({"k": ["a", "a", "a", "b", "b"], "v": [10, 10, 20, 5, 7]})
    got = OW.window(t, [col("k")], [col("v")], [False], [False],
                    [("row_number", "rn"), ("rank", "rk"), ("dense_rank", "dr"), ("offset", "nx", col("v"), 1), ("offset", "pv", col("v"), -1, -1),
                     ("agg", "tot", col("v").sum())]).to_pydict()
    assert got["rn"] == [1, 2, 3, 1, 2] and got["rk"] == [1, 1, 3, 1, 2] and got["dr"] == [1, 1, 2, 1, 2]
    assert got["nx"] == [10, 20, None, 7, None] and got["pv"] == [-1, 10, 10, -1, 5] and got["tot"] == [40, 40, 40, 12, 12]


@pytest.mark.parametrize("n", [0, 1, 37, 400])
@pytest.mark.parametrize("shape", ["partition+order", "order only", "partition only", "two keys desc"])
def test_window_composition_matches_oracle(n, shape):
    rng = np.random.default_rng(n + len(shape))
    tbl = window_table(rng, n)
    pb, ob, desc, nf = {"partition+order": ([col("g")], [col("o")], [False], [False]), "order only": ([], [col("o"), col("h")], [False, True], [False, True]),
                        "partition only": ([col("g"), col("h")], [], [], []),
                        "two keys desc": ([col("h")], [col("o"), col("x")], [True, False], [True, False])}[shape]
    funcs = [f for f in FUNCS if pb or f[0] != "agg"]
    got = W.window(EmulatedQueryBatch(tbl), pb, ob, desc, nf, funcs).to_arrow_table()
    want = OW.window(tbl, pb, ob, desc, nf, funcs)
    assert got.column_names == want.column_names
    if ob and n:
        # ties in the order keys: row_number / lead / lag depend on the (admissible) order of the tied rows, which differs between a
        # global stable sort and the per-partition sort -- compare those through a tie-free order (the id column appended)
        ob2, d2, n2 = ob + [col("id")], desc + [False], nf + [False]
        got = W.window(EmulatedQueryBatch(tbl), pb, ob2, d2, n2, funcs).to_arrow_table()
        want = OW.window(tbl, pb, ob2, d2, n2, funcs)
        got_r = W.window(EmulatedQueryBatch(tbl), pb, ob, desc, nf, [("rank", "rk"), ("dense_rank", "dr")]).to_arrow_table()
        want_r = OW.window(tbl, pb, ob, desc, nf, [("rank", "rk"), ("dense_rank", "dr")])
        assert_tables_equal(_sorted(got_r, ["id"]), _sorted(want_r, ["id"]))
        assert_tables_equal(_sorted(got, ["id"]), _sorted(want, ["id"]), float_rtol=1e-12)
    elif not ob:
        keep = [c for c in got.column_names if c not in ("rn", "lead1", "lag2", "lead_s")]   # no order: positions inside a partition are arbitrary
        assert_tables_equal(_sorted(got.select(keep), ["id"]), _sorted(want.select(keep), ["id"]), float_rtol=1e-12)


def test_oracle_pivot_reference_literals():
    """tests/recordbatch/test_pivot.py: missing combinations are null, unknown names give null columns"""
    t = pa.table({"group": ["A", "A", "B"], "pivot": ["1", "2", "1"], "value": [1, 2, 3]})
    assert OW.pivot(t, ["group"], "pivot", "value", ["1", "2"]).to_pydict() == {"group": ["A", "B"], "1": [1, 3], "2": [2, None]}
    assert OW.pivot(t, ["group"], "pivot", "value", ["1", "3"]).to_pydict() == {"group": ["A", "B"], "1": [1, 3], "3": [None, None]}


@pytest.mark.parametrize("ptype", ["str", "int", "bool"])
def test_pivot_composition_matches_oracle(ptype):
    rng = np.random.default_rng(5)
    g1, g2 = rng.integers(0, 6, 300), rng.integers(0, 2, 300)
    pv = {"str": np.array([f"c{v}" for v in rng.integers(0, 4, 300)], dtype=object), "int": rng.integers(0, 4, 300), "bool": rng.integers(0, 2, 300).astype(bool)}[ptype]
    tbl = pa.table({"g1": pa.array(g1, mask=rng.random(300) < 0.1), "g2": g2, "p": pa.array(pv, type=pa.large_string()) if ptype == "str" else pv,
                    "v": rng.standard_normal(300)})
    names = {"str": ["c0", "c3", "zz", "c1"], "int": ["0", "3", "9", "x"], "bool": ["true", "false"]}[ptype]
    # unique (group, pivot) pairs: aggregate first, as PivotSink does
    got = W.agg_pivot(EmulatedQueryBatch(tbl), [col("g1"), col("g2")], col("p"), col("v"), "sum", names).to_arrow_table()
    from oracle import oracle as O
    pre = O.agg(tbl, [("sum", tbl.column("v"), "v", "valid")], [("g1", tbl.column("g1")), ("g2", tbl.column("g2")), ("p", tbl.column("p"))])
    want = OW.pivot(pre, ["g1", "g2"], "p", "v", names)
    assert_tables_equal(got, want, float_rtol=1e-12)


# ---- percentile / median: the window-based composition of RecordBatch._agg_with_percentile against exact_percentile's restatement ----
def test_oracle_percentile_reference_literals():
    """tests/recordbatch/test_table_aggs.py percentile cases: interpolation between order statistics, nulls ignored, all-null -> null"""
    from helpers import oracle_agg
    t = pa.table({"g": [1, 1, 1, 1, 2, 2, 3], "x": [1.0, 2.0, 3.0, 4.0, 10.0, None, None]})
    got = oracle_agg(t, [col("x").median().alias("m"), col("x").percentile(0.25).alias("q1"), col("x").percentile(1.0).alias("mx")], [col("g")]).to_pydict()
    assert got == {"g": [1, 2, 3], "m": [2.5, 10.0, None], "q1": [1.75, 10.0, None], "mx": [4.0, 10.0, None]}
    got = oracle_agg(t, [col("x").median().alias("m")], []).to_pydict()
    assert got == {"m": [3.0]}


@pytest.mark.parametrize("n", [0, 1, 50, 600])
def test_percentile_composition_matches_oracle(n):
    from helpers import oracle_agg
    rng = np.random.default_rng(n)
    tbl = window_table(rng, n, nulls=0.15)
    aggs = [col("x").median().alias("med"), col("o").percentile(0.9).alias("p90"), col("x").sum().alias("sx"), col("h").percentile(0.0).alias("p0"),
            (col("x") * 2 + col("h")).percentile(0.33).alias("pe")]
    for gb in ([col("g")], [col("g"), col("h")], []):
        got = EmulatedQueryBatch(tbl).agg(aggs, gb).to_arrow_table()
        want = oracle_agg(tbl, aggs, gb)
        keys = [g.name() for g in gb]
        if keys:
            assert_tables_equal(_sorted(got, keys), _sorted(want, keys), float_rtol=1e-12)
        else:
            assert_tables_equal(got, want, float_rtol=1e-12)
    got = EmulatedQueryBatch(tbl).agg([col("x").median().alias("med")], [col("g")], predicate=[col("h") > 0]).to_arrow_table()
    assert_tables_equal(_sorted(got, ["g"]), _sorted(oracle_agg(tbl, [col("x").median().alias("med")], [col("g")], [col("h") > 0]), ["g"]), float_rtol=1e-12)


# ---- the reference's own pivot cases (tests/recordbatch/test_pivot.py:9-125), through the oracle and through the host composition ----
PIVOT_LITERALS = [
    ({"group": ["B", "A", "A", "B"], "pivot": [1, 2, 1, 2], "value": [1, 2, 3, 4]}, ["1", "2"], {"group": ["A", "B"], "1": [3, 1], "2": [2, 4]}),              # :27-54
    ({"group": ["A", "A", "B"], "pivot": [1, 2, 1], "value": [1, 2, 3]}, ["1"], {"group": ["A", "B"], "1": [1, 3]}),                                          # :57-71
    ({"group": ["A", "A", "B"], "pivot": [1, 2, 1], "value": [1, 2, 3]}, ["1", "2", "3"], {"group": ["A", "B"], "1": [1, 3], "2": [2, None], "3": [None, None]}),  # :74-90
    ({"group": ["A", None, "B", None], "pivot": [1, 1, 2, 2], "value": [1, 2, 3, 4]}, ["1", "2"],
     {"group": ["A", "B", None], "1": [1, None, 2], "2": [None, 3, 4]}),                                                                                        # :93-108
    ({"group": ["A", "A", "B", "B"], "pivot": [1, None, 1, None], "value": [1, 2, 3, 4]}, ["1", "None"], {"group": ["A", "B"], "1": [1, 3], "None": [2, 4]}),    # :111-125
]


@pytest.mark.parametrize("case", range(len(PIVOT_LITERALS)))
def test_pivot_reference_literals(case):
    data, names, want = PIVOT_LITERALS[case]
    tbl = pa.table({"group": pa.array(data["group"], pa.large_string()), "pivot": pa.array(data["pivot"], pa.int64()), "value": pa.array(data["value"], pa.int64())})
    for got in (OW.pivot(tbl, ["group"], "pivot", "value", names),
                W.pivot(EmulatedQueryBatch(tbl), [col("group")], col("pivot"), col("value"), names).to_arrow_table()):
        assert got.sort_by([("group", "ascending")]).to_pydict() == want


# ---- the reference's window cases (tests/window/*.py), through the oracle and through the host composition over the emulation twin ----
def _both(tbl, pb, ob, desc, nf, funcs):
    a = OW.window(tbl, pb, ob, desc, nf, funcs)
    b = W.window(EmulatedQueryBatch(tbl), pb, ob, desc, nf, funcs).to_arrow_table()
    return a, b


def test_window_reference_row_number_literals():
    """tests/window/test_partition_order_by_rank.py:13-50"""
    tbl = pa.table({"category": pa.array(["A", "A", "A", "B", "B", "B", "C", "C"], pa.large_string()), "sales": [100, 200, 50, 500, 100, 300, 250, 150]})
    for desc, want in ((False, [2, 3, 1, 3, 1, 2, 2, 1]), (True, [2, 1, 3, 1, 3, 2, 1, 2])):
        for got in _both(tbl, [col("category")], [col("sales")], [desc], [desc], [("row_number", "rn")]):
            got = got.sort_by([("category", "ascending"), ("sales", "ascending")]).to_pydict()
            ref = sorted(zip(tbl.column("category").to_pylist(), tbl.column("sales").to_pylist(), want))
            assert list(zip(got["category"], got["sales"], got["rn"])) == ref


@pytest.mark.parametrize("desc,nulls_first,expected_rank", [(False, False, [1, 2, 1, 2, 1, 2]), (False, True, [2, 1, 2, 1, 2, 1]),
                                                             (True, False, [1, 2, 1, 2, 1, 2]), (True, True, [2, 1, 2, 1, 2, 1])])
def test_window_reference_rank_nulls_first_or_last(desc, nulls_first, expected_rank):
    """tests/window/test_partition_order_by_rank.py:240-277"""
    tbl = pa.table({"id": [1, 2, 3, 4, 5, 6], "category": pa.array(["A", "A", "B", "B", "C", "C"], pa.large_string()), "value": [10, None, 20, None, 30, None]})
    for got in _both(tbl, [col("category")], [col("value")], [desc], [nulls_first], [("rank", "rank"), ("dense_rank", "drank")]):
        got = got.sort_by([("id", "ascending")]).to_pydict()
        assert got["rank"] == expected_rank and got["drank"] == expected_rank


@pytest.mark.parametrize("desc,nulls_first,expected_lead", [
    (False, False, [2, 3, 4, None, 6, 7, 8, None, 10, 11, 12, None]), (False, True, [2, 3, None, 1, 6, 7, None, 5, 10, 11, None, 9]),
    (True, False, [4, 1, 2, None, 8, 5, 6, None, 12, 9, 10, None]), (True, True, [None, 1, 2, 3, None, 5, 6, 7, None, 9, 10, 11])])
def test_window_reference_lead_nulls_first_or_last(desc, nulls_first, expected_lead):
    """tests/window/test_partition_order_by_lag_lead.py:65-104"""
    tbl = pa.table({"id": list(range(1, 13)), "category": pa.array(["A"] * 4 + ["B"] * 4 + ["C"] * 4, pa.large_string()), "ts": [1, 2, 3, None] * 3})
    for got in _both(tbl, [col("category")], [col("ts")], [desc], [nulls_first], [("offset", "lead_value", col("id"), 1)]):
        assert got.sort_by([("id", "ascending")]).column("lead_value").to_pylist() == expected_lead


def test_window_reference_lag_lead_with_column_default():
    """tests/window/test_partition_order_by_lag_lead.py:153-198: a default that is an expression of the row, offsets beyond the partition"""
    import random
    random.seed(45)
    rows, want = [], []
    for category in ["A", "B"]:
        defaults, values = random.sample(range(10, 100), 50), random.sample(range(10, 100), 50)
        for ts in range(50):
            rows.append((category, ts, defaults[ts], values[ts]))
            want.append((category, ts, 2 * defaults[ts] if ts < 10 else values[ts - 10], 2 * defaults[ts] if ts >= 40 else values[ts + 10],
                         2 * defaults[ts], 2 * defaults[ts], 99 if ts < 15 else values[ts - 15], 99 if ts >= 35 else values[ts + 15]))
    tbl = pa.table({"category": pa.array([r[0] for r in rows], pa.large_string()), "ts": [r[1] for r in rows], "default": [r[2] for r in rows],
                    "value": [r[3] for r in rows]})
    funcs = [("offset", "lag_10", col("value"), -10, 2 * col("default")), ("offset", "lead_10", col("value"), 10, 2 * col("default")),
             ("offset", "lag_50", col("value"), -50, 2 * col("default")), ("offset", "lead_50", col("value"), 50, 2 * col("default")),
             ("offset", "lag_15", col("value"), -15, 99), ("offset", "lead_15", col("value"), 15, 99)]
    for got in _both(tbl, [col("category")], [col("ts")], [False], [False], funcs):
        got = got.sort_by([("category", "ascending"), ("ts", "ascending")]).to_pydict()
        assert list(zip(got["category"], got["ts"], got["lag_10"], got["lead_10"], got["lag_50"], got["lead_50"], got["lag_15"], got["lead_15"])) == want


def test_window_reference_partition_only_aggregates():
    """tests/window/test_partition_only.py:14-155,211-245: sum / min / max / mean / count over the partition, null partition keys form a group"""
    tbl = pa.table({"category": pa.array(["A", "A", "B", None, "B", None, "C"], pa.large_string()), "value": [1, 2, 3, 4, 5, 6, 7]})
    funcs = [("agg", "s", col("value").sum()), ("agg", "mn", col("value").min()), ("agg", "mx", col("value").max()), ("agg", "m", col("value").mean()),
             ("agg", "c", col("value").count())]
    for got in _both(tbl, [col("category")], [], [], [], funcs):
        got = got.sort_by([("value", "ascending")]).to_pydict()
        assert got["s"] == [3, 3, 8, 10, 8, 10, 7] and got["mn"] == [1, 1, 3, 4, 3, 4, 7] and got["mx"] == [2, 2, 5, 6, 5, 6, 7]
        assert got["m"] == [1.5, 1.5, 4.0, 5.0, 4.0, 5.0, 7.0] and got["c"] == [2, 2, 2, 2, 2, 2, 1]


# ---- skew: the planner's staged power-sum form (src/daft-local-plan/src/agg.rs:302-348) executed through the generated code,
# against calculate_skew's definition (src/daft-core/src/utils/stats.rs:134-147) ----
def test_skew_staged_form_matches_definition():
    from helpers import oracle_agg
    from daft_b200.aggstages import populate_aggregation_stages
    rng = np.random.default_rng(11)
    n = 2000
    tbl = pa.table({"g": pa.array(rng.integers(0, 6, n)), "x": pa.array(rng.gamma(2.0, 3.0, n), mask=rng.random(n) < 0.1),
                    "i": pa.array(rng.integers(-50, 50, n))})
    aggs = [col("x").skew().alias("sk"), col("i").skew().alias("ski"), col("x").sum().alias("sx")]
    for gb in ([col("g")], []):
        want = oracle_agg(tbl, aggs, gb)
        got = EmulatedQueryBatch(tbl).agg(aggs, gb).to_arrow_table()
        keys = [g.name() for g in gb]
        a, b = (_sorted(got, keys), _sorted(want, keys)) if keys else (got, want)
        assert_tables_equal(a, b, float_rtol=1e-9)
        # the two-stage rewrite over two morsels: partial power sums, summed, same projection
        first, second, final = populate_aggregation_stages(aggs, gb)
        halves = [EmulatedQueryBatch(tbl.slice(0, 900)).agg(first, gb).to_arrow_table(), EmulatedQueryBatch(tbl.slice(900)).agg(first, gb).to_arrow_table()]
        merged = EmulatedQueryBatch(pa.concat_tables(halves)).agg(second, [col(g.name()) for g in gb]).eval_expression_list(final).to_arrow_table()
        a = _sorted(merged, keys) if keys else merged
        assert_tables_equal(a, b, float_rtol=1e-9)
    # reference literal (tests/recordbatch/test_table_aggs.py skew cases): symmetric data has zero skew, a single value gives NaN, no value null
    t = pa.table({"g": [1, 1, 1, 2, 3], "x": [1.0, 2.0, 3.0, 5.0, None]})
    got = EmulatedQueryBatch(t).agg([col("x").skew().alias("sk")], [col("g")]).to_arrow_table().sort_by([("g", "ascending")]).column("sk").to_pylist()
    assert got[0] == 0.0 and got[1] != got[1] and got[2] is None


def test_composite_aggregates_nest_in_one_call():
    """count_distinct + string min / max + percentile in ONE aggregation list: each is a host-side rewrite, and they nest.  Found by
    ClickBench Q22 (count_distinct next to MIN(URL)): the string rewrite sorts its groups by key, the count_distinct column was placed
    side by side in first-occurrence order -> misaligned rows.  The column is now aligned on the keys."""
    from helpers import oracle_agg
    rng = np.random.default_rng(3)
    n = 1500
    tbl = pa.table({"g": pa.array(rng.integers(0, 40, n)[::-1].copy()), "h": pa.array(rng.integers(0, 3, n)),
                    "s": pa.array(np.array([f"v{v}" for v in rng.integers(0, 60, n)], dtype=object), type=pa.large_string(), mask=rng.random(n) < 0.1),
                    "x": pa.array(rng.standard_normal(n), mask=rng.random(n) < 0.1), "u": pa.array(rng.integers(0, 25, n))})
    aggs = [col("u").count_distinct().alias("cd"), col("s").min().alias("smin"), col("s").max().alias("smax"), col("x").median().alias("med"),
            col("x").sum().alias("sx"), col("s").count_distinct().alias("cds")]
    for gb in ([col("g")], [col("g"), col("h")], []):
        for subset in (aggs, aggs[:2], [aggs[0], aggs[3]], [aggs[1], aggs[3]]):
            got = EmulatedQueryBatch(tbl).agg(subset, gb).to_arrow_table()
            want = oracle_agg(tbl, subset, gb)
            keys = [g.name() for g in gb]
            a, b = (_sorted(got, keys), _sorted(want, keys)) if keys else (got, want)
            assert_tables_equal(a, b, float_rtol=1e-12)


# ---- aggregates over ROWS frames (RecordBatch::window_agg_dynamic_frame, src/daft-recordbatch/src/ops/window.rs:299-483) ----
# three parties: a naive "aggregate the slice" expectation (what the reference's own tests compute, tests/window/test_running_agg.py),
# the oracle's restatement of the sliding state machine, and daft_b200.window's composition running the host build of the device's
# state machine (tests/wf_host.cpp <- daft_b200/csrc/window_frame_core.h)

def _naive_frames(values, start, end, min_periods, fn):
    out, m = [], len(values)
    for r in range(m):
        fs = 0 if start is None else min(max(r + start, 0), m)
        fe = m if end is None else min(max(r + 1 + end, 0), m)
        live = [v for v in values[fs:fe] if v is not None]
        out.append(fn(live, fe - fs) if fe - fs >= min_periods else None)
    return out


@pytest.mark.parametrize("start,end,min_periods", [(None, 0, 1), (-2, 0, 1), (-2, 0, 3), (-10, -3, 3), (2, 9, 3), (None, 2, 1), (-3, None, 1), (-2, 2, 1),
                                                   (-1, 1, 1), (-3, 0, 4), (0, 0, 1), (1, 0, 1), (None, None, 1), (5, 7, 1), (-7, -5, 2), (-(2**62), 2**62, 1), (-(2**63), 2**63 - 1, 2)])
def test_window_frames_reference_running_agg_cases(start, end, min_periods):
    """tests/window/test_running_agg.py:25-130 (running sum), :270-297 (mean over -2..0 with min_periods 3), :300-395 (negative / positive
    bounds with min_periods), :398-465 (one unbounded side), :468-560 (centred frames), :563-620 (min_periods 1-4): integer data, so the
    slice aggregate the reference's tests expect is exact"""
    import random
    random.seed(444)
    cats, ts, vals = [], [], []
    per = {}
    for c in ["X", "Y", "Z"]:
        v = [random.randint(1, 100) if random.random() > 0.1 else None for _ in range({"X": 100, "Y": 37, "Z": 1}[c])]
        per[c] = v
        cats += [c] * len(v); ts += list(range(len(v))); vals += v
    order = list(range(len(vals)))
    random.shuffle(order)
    tbl = pa.table({"category": pa.array([cats[i] for i in order], pa.large_string()), "ts": [ts[i] for i in order], "value": pa.array([vals[i] for i in order], pa.int32())})
    funcs = [("frame", "s", col("value").sum(), start, end, min_periods), ("frame", "m", col("value").mean(), start, end, min_periods),
             ("frame", "c", col("value").count(), start, end, min_periods), ("frame", "ca", col("value").count("all"), start, end, min_periods),
             ("frame", "cn", col("value").count("null"), start, end, min_periods), ("frame", "mn", col("value").min(), start, end, min_periods),
             ("frame", "mx", col("value").max(), start, end, min_periods)]
    want = {k: [] for k in ("s", "m", "c", "ca", "cn", "mn", "mx")}
    for c in ["X", "Y", "Z"]:
        v = per[c]
        want["s"] += _naive_frames(v, start, end, min_periods, lambda live, size: sum(live) if live else None)
        want["m"] += _naive_frames(v, start, end, min_periods, lambda live, size: sum(live) / len(live) if live else None)
        want["c"] += _naive_frames(v, start, end, min_periods, lambda live, size: len(live))
        want["ca"] += _naive_frames(v, start, end, min_periods, lambda live, size: size)
        want["cn"] += _naive_frames(v, start, end, min_periods, lambda live, size: size - len(live))
        want["mn"] += _naive_frames(v, start, end, min_periods, lambda live, size: min(live) if live else None)
        want["mx"] += _naive_frames(v, start, end, min_periods, lambda live, size: max(live) if live else None)
    for got in _both(tbl, [col("category")], [col("ts")], [False], [False], funcs):
        assert [str(got.schema.field(k).type) for k in ("s", "m", "c", "ca", "cn", "mn", "mx")] == ["int64", "double", "uint64", "uint64", "uint64", "int32", "int32"]
        got = got.sort_by([("category", "ascending"), ("ts", "ascending")]).to_pydict()
        for k in want:
            if k == "m":   # the sliding sum of integers held in Float64 is exact here (small values), the division is one rounding
                assert got[k] == pytest.approx(want[k], rel=1e-15, abs=0) if None not in want[k] else all(
                    (a is None and b is None) or (a is not None and b is not None and abs(a - b) <= 1e-12 * abs(b)) for a, b in zip(got[k], want[k]))
            else:
                assert got[k] == want[k], k


@pytest.mark.parametrize("seed", range(6))
def test_window_frames_float_state_machine_is_bit_exact(seed):
    """floats (NaN, +-inf, nulls, magnitudes 1e-8 .. 1e12): the sliding `+=` / `-=` sequence of the reference's SumWindowState is order
    dependent -- the host build of the device's state machine must agree with the oracle's restatement bit for bit; every dtype the
    reference accumulates differently (Int64 / UInt64 wrap, Float32 in Float32); descending order, several partition keys"""
    rng = np.random.default_rng(seed)
    n = 300
    x = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 12, n)
    x[rng.random(n) < 0.03] = np.nan
    x[rng.random(n) < 0.02] = np.inf
    x[rng.random(n) < 0.02] = -np.inf
    tbl = pa.table({"id": np.arange(n), "g": pa.array(rng.integers(0, 6, n)), "h": pa.array(rng.integers(0, 2, n).astype(np.int8), mask=rng.random(n) < 0.1),
                    "o": pa.array(rng.integers(0, 40, n)),
                    "f64": pa.array(x, mask=rng.random(n) < 0.15), "f32": pa.array(x.astype(np.float32), mask=rng.random(n) < 0.15),
                    "i64": pa.array(rng.integers(-2**62, 2**62, n), mask=rng.random(n) < 0.1), "u64": pa.array(rng.integers(0, 2**64, n, dtype=np.uint64)),
                    "i8": pa.array(rng.integers(-128, 128, n).astype(np.int8), mask=rng.random(n) < 0.1), "u16": pa.array(rng.integers(0, 65536, n).astype(np.uint16)),
                    "d": pa.array(rng.integers(0, 20000, n).astype(np.int32), pa.date32(), mask=rng.random(n) < 0.1)})
    frames = [(None, 0, 1), (-3, 0, 1), (-1, 4, 2), (2, 5, 1), (-6, -2, 3), (None, None, 1), (-2, None, 1), (0, 0, 1)]
    start, end, mp = frames[seed % len(frames)]
    start2, end2, mp2 = frames[(seed + 3) % len(frames)]
    funcs = []
    for c in ("f64", "f32", "i64", "u64", "i8", "u16"):
        funcs += [("frame", f"s_{c}", col(c).sum(), start, end, mp), ("frame", f"m_{c}", col(c).mean(), start2, end2, mp2),
                  ("frame", f"c_{c}", col(c).count(), start, end, mp)]
    for c in ("i64", "u64", "i8", "u16", "d"):
        funcs += [("frame", f"mn_{c}", col(c).min(), start2, end2, mp2), ("frame", f"mx_{c}", col(c).max(), start, end, mp)]
    funcs += [("frame", "expr", (col("f64") * 2.0 + col("i8")).sum(), start, end, mp), ("row_number", "rn")]
    desc = bool(seed % 2)
    a, b = _both(tbl, [col("g"), col("h")], [col("o"), col("id")], [desc, False], [desc, False], funcs)
    assert a.schema.field("s_f32").type == pa.float32() and a.schema.field("s_i8").type == pa.int64() and a.schema.field("s_u16").type == pa.uint64()
    assert a.schema.field("mn_d").type == pa.date32() and a.schema.field("m_i8").type == pa.float64()
    assert_tables_equal(_sorted(b, ["id"]), _sorted(a, ["id"]), float_rtol=0.0)


def test_window_frames_errors():
    tbl = pa.table({"g": [1, 1, 2], "o": [1, 2, 3], "x": [1.0, 2.0, 3.0], "s": pa.array(["a", "b", "c"], pa.large_string())})
    for make in (lambda t: EmulatedQueryBatch(t),):
        with pytest.raises(ValueError, match="Negative frame size"):
            W.window(make(tbl), [col("g")], [col("o")], [False], [False], [("frame", "f", col("x").sum(), 3, 1)])
        with pytest.raises(NotImplementedError):
            W.window(make(tbl), [col("g")], [col("o")], [False], [False], [("frame", "f", col("x").min(), -1, 0)])   # float min / max over frames
        with pytest.raises(NotImplementedError):
            W.window(make(tbl), [col("g")], [col("o")], [False], [False], [("frame", "f", col("x").stddev(), -1, 0)])
        with pytest.raises(TypeError):
            W.window(make(tbl), [col("g")], [col("o")], [False], [False], [("frame", "f", col("s").sum(), -1, 0)])


# ---- RANGE frames (calculate_range_frame_bounds, window.rs:167-272) ----

@pytest.mark.parametrize("start,end,desc,min_periods", [(-2, 2, False, 1), (-2, 2, True, 1), (-5, -1, False, 1), (-5, -1, True, 1), (1, 5, False, 1), (1, 5, True, 1),
                                                        (None, 1, False, 1), (-1, None, False, 1), (-10000, 10000, False, 1), (-3, 1, False, 3), (0, 0, True, 1),
                                                        (None, None, True, 1), (-3, 1, True, 2)])
def test_window_range_frames_reference_cases(start, end, desc, min_periods):
    """tests/window/test_running_agg.py:826-1370: distinct integer timestamps with gaps; the frame of a row holds the rows whose ts lies in
    [ts + start, ts + end], in either sort direction (calculate_expected_range, :978-1010)"""
    import random
    random.seed(50)
    cats, tss, vals, want = [], [], [], {k: {} for k in ("s", "m", "mn", "mx", "c")}
    for c in ["X", "Y"]:
        stamps = sorted(random.sample(range(100), 30))
        v = [random.randint(1, 100) for _ in stamps]
        for t, x in zip(stamps, v):
            cats.append(c); tss.append(t); vals.append(x)
            lo = t + start if start is not None else -10**9
            hi = t + end if end is not None else 10**9
            live = [y for u, y in zip(stamps, v) if lo <= u <= hi]
            ok = len(live) >= min_periods
            want["s"][(c, t)] = sum(live) if ok and live else None
            want["m"][(c, t)] = sum(live) / len(live) if ok and live else None
            want["mn"][(c, t)] = min(live) if ok and live else None
            want["mx"][(c, t)] = max(live) if ok and live else None
            want["c"][(c, t)] = len(live) if ok else None
    order = list(range(len(vals)))
    random.shuffle(order)
    tbl = pa.table({"category": pa.array([cats[i] for i in order], pa.large_string()), "ts": pa.array([tss[i] for i in order], pa.int32()),
                    "value": [vals[i] for i in order]})
    funcs = [("range_frame", "s", col("value").sum(), start, end, min_periods), ("range_frame", "m", col("value").mean(), start, end, min_periods),
             ("range_frame", "mn", col("value").min(), start, end, min_periods), ("range_frame", "mx", col("value").max(), start, end, min_periods),
             ("range_frame", "c", col("value").count(), start, end, min_periods)]
    for got in _both(tbl, [col("category")], [col("ts")], [desc], [desc], funcs):
        got = got.to_pydict()
        for k in want:
            for c, t, x in zip(got["category"], got["ts"], got[k]):
                w = want[k][(c, t)]
                assert (x is None and w is None) or (x is not None and w is not None and abs(x - w) <= 1e-12 * abs(w)), (k, c, t, x, w)


@pytest.mark.parametrize("seed", range(6))
def test_window_range_frames_match_oracle_with_nulls_and_ties(seed):
    """duplicate and null order keys, nulls first or last, temporal keys (offsets in days / microseconds), float sums: the host build of the
    device's state machine == the oracle's restatement (which follows the reference's forward searches, including where they resume)"""
    rng = np.random.default_rng(100 + seed)
    n = 250
    tbl = pa.table({"id": np.arange(n), "g": pa.array(rng.integers(0, 5, n)),
                    "k": pa.array(rng.integers(0, 60, n), mask=rng.random(n) < 0.08),
                    "d": pa.array(rng.integers(10000, 10100, n).astype(np.int32), pa.date32(), mask=rng.random(n) < 0.05),
                    "ts": pa.array(rng.integers(0, 50, n) * 1_000_000, pa.timestamp("us")),
                    "x": pa.array(rng.standard_normal(n) * 10.0 ** rng.integers(-3, 6, n), mask=rng.random(n) < 0.1),
                    "v": pa.array(rng.integers(-50, 50, n).astype(np.int16), mask=rng.random(n) < 0.1)})
    key, scale = [("k", 1), ("d", 1), ("ts", 1_000_000)][seed % 3]
    desc = bool(seed % 2)
    nulls_first = desc if seed < 4 else not desc   # nulls-first order keys: the reference's min / max path then resumes its searches at the end
    frames = [(-3, 3, 1), (-5, -1, 2), (1, 4, 1), (None, 2, 1), (-2, None, 1), (0, 0, 1)]
    a0, b0, mp = frames[seed]
    sc = lambda o: None if o is None else o * scale
    funcs = []
    for name, e in (("x", col("x")), ("v", col("v"))):
        funcs += [("range_frame", f"s_{name}", e.sum(), sc(a0), sc(b0), mp), ("range_frame", f"m_{name}", e.mean(), sc(a0), sc(b0), mp),
                  ("range_frame", f"c_{name}", e.count("all"), sc(a0), sc(b0), mp)]
    funcs += [("range_frame", "mn", col("v").min(), sc(a0), sc(b0), mp), ("range_frame", "mx", col("v").max(), sc(a0), sc(b0), mp), ("frame", "rows", col("v").sum(), -1, 1)]
    # ties in the key: the frame of a row does not depend on the order of its peers, the sliding float sum does -- break ties by id on both sides
    pb, ob = [col("g")], [col(key)]
    a = OW.window(tbl.sort_by([("id", "ascending")]), pb, ob, [desc], [nulls_first], funcs)
    b = W.window(EmulatedQueryBatch(tbl), pb, ob, [desc], [nulls_first], funcs).to_arrow_table()
    exact = [c for c in a.column_names if not c.endswith("_x") or c.startswith("c_")]
    assert_tables_equal(_sorted(b.select(exact), ["id"]), _sorted(a.select(exact), ["id"]), float_rtol=0.0)
    loose = [c for c in a.column_names if c not in exact] + ["id"]
    assert_tables_equal(_sorted(b.select(loose), ["id"]), _sorted(a.select(loose), ["id"]), float_rtol=1e-9)


def test_window_range_frame_errors():
    tbl = pa.table({"g": [1, 1, 1], "o": [1, 2, 3], "p": [1.0, 2.0, 3.0], "x": [1.0, 2.0, 3.0]})
    with pytest.raises(ValueError, match="exactly one ORDER BY"):
        W.window(EmulatedQueryBatch(tbl), [col("g")], [col("o"), col("x")], [False, False], [False, False], [("range_frame", "f", col("x").sum(), -1, 1)])
    with pytest.raises(NotImplementedError):
        W.window(EmulatedQueryBatch(tbl), [col("g")], [col("p")], [False], [False], [("range_frame", "f", col("x").sum(), -1, 1)])
    for run in (lambda: W.window(EmulatedQueryBatch(tbl), [col("g")], [col("o")], [False], [False], [("range_frame", "f", col("x").sum(), 2, -2)]),
                lambda: OW.window(tbl, [col("g")], [col("o")], [False], [False], [("range_frame", "f", col("x").sum(), 2, -2)])):
        with pytest.raises(ValueError, match="Negative frame size"):
            run()

"""Sort-merge, asof and cross joins on the CUDA path against the oracle (oracle/joins_sorted.py) and the
reference's literal expectations (tests/dataframe/test_asof_join.py, restated in tests/test_joins_sorted_cpu.py)."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal
from oracle import joins_sorted as J
from test_joins_sorted_cpu import ASOF_LITERALS, _sorted, random_join_tables

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", range(len(ASOF_LITERALS)))
def test_asof_join_reference_literals(gpu, case):
    left, right, lby, rby, lon, ron, keys, want = ASOF_LITERALS[case]
    c = gpu.col
    got = gpu.RecordBatch.from_arrow_table(pa.table(left)).asof_join(
        gpu.RecordBatch.from_arrow_table(pa.table(right)), [c(k) for k in lby], [c(k) for k in rby], c(lon), c(ron)).to_arrow_table()
    assert got.column_names == list(want.keys())
    assert _sorted(got, keys).to_pydict() == want


@pytest.mark.parametrize("nl,nr", [(0, 5), (5, 0), (1, 1), (40, 60), (3000, 2000), (50000, 20000)])
def test_sort_merge_join_matches_oracle(gpu, nl, nr):
    rng = np.random.default_rng(nl * 7 + nr)
    left, right = random_join_tables(rng, nl, nr, dom=12 if nl < 100 else max(nl // 3, 1))
    c = gpu.col
    for keys in (["k"], ["k", "s"]):
        r = right if len(keys) == 2 else right.rename_columns(["k", "rs", "rv"])
        got = gpu.RecordBatch.from_arrow_table(left).sort_merge_join(gpu.RecordBatch.from_arrow_table(r), [c(k) for k in keys],
                                                                      [c(k) for k in keys]).to_arrow_table()
        if nl <= 3000:
            want = J.sort_merge_join(left, r, keys, keys)
        else:   # the pure-Python merge is too slow here: the oracle's probe-table join yields the same rows (tests/test_joins_sorted_cpu.py)
            from oracle import oracle as O
            want = O.hash_join(left, r, keys, keys, "inner")
        assert_tables_equal(_sorted(got, ["lv", "rv"]), _sorted(want, ["lv", "rv"]))
    # is_sorted = True skips the sorts: same answer on pre-sorted inputs, and the output is left-major in key order
    from oracle import oracle as O
    ls, rs = left.sort_by([("k", "ascending")]), right.rename_columns(["k", "rs", "rv"]).sort_by([("k", "ascending")])
    got = gpu.RecordBatch.from_arrow_table(ls).sort_merge_join(gpu.RecordBatch.from_arrow_table(rs), [c("k")], [c("k")], is_sorted=True).to_arrow_table()
    assert_tables_equal(_sorted(got, ["lv", "rv"]), _sorted(O.hash_join(ls, rs, ["k"], ["k"], "inner"), ["lv", "rv"]))
    kk = got.column("k").to_pylist()
    assert kk == sorted(kk)


def test_sort_merge_join_mixed_key_types_and_errors(gpu):
    c = gpu.col
    left = gpu.RecordBatch.from_arrow_table(pa.table({"a": pa.array([1, 2, 3], pa.int32()), "x": [1.0, 2.0, 3.0]}))
    right = gpu.RecordBatch.from_arrow_table(pa.table({"b": pa.array([3, 2, 2], pa.int64()), "y": ["p", "q", "r"]}))
    got = left.sort_merge_join(right, [c("a")], [c("b")]).to_arrow_table()      # keys cast to their supertype (ops/joins/mod.rs:22-46)
    assert _sorted(got, ["a", "y"]).to_pydict() == {"a": [2, 2, 3], "x": [2.0, 2.0, 3.0], "b": [2, 2, 3], "y": ["q", "r", "p"]}
    with pytest.raises(ValueError, match="No columns were passed in to join on"):
        left.sort_merge_join(right, [], [])
    with pytest.raises(ValueError, match="Mismatch of join on clauses"):
        left.sort_merge_join(right, [c("a")], [])


@pytest.mark.parametrize("nl,nr,nby", [(0, 10, 1), (10, 0, 1), (300, 200, 0), (300, 200, 1), (4000, 3000, 2)])
def test_asof_join_matches_oracle(gpu, nl, nr, nby):
    rng = np.random.default_rng(nl + nr + nby)
    by = ["g1", "g2"][:nby]

    def tbl(n, p, unique_on):
        d = {"g1": pa.array(np.array([f"e{v}" for v in rng.integers(0, 5, n)], dtype=object), mask=rng.random(n) < 0.05),
             "g2": pa.array(rng.integers(0, 3, n), mask=rng.random(n) < 0.05)}
        on = rng.permutation(10 * max(n, 1))[:n] if unique_on else rng.integers(0, 10 * max(n, 1), n)   # unique right on-keys: one admissible match
        d["ts"] = pa.array(on.astype(np.float64) / 4, mask=rng.random(n) < 0.05)
        d[f"{p}v"] = pa.array(np.arange(n))
        return pa.table(d)
    left, right = tbl(nl, "l", False), tbl(nr, "r", True)
    right = right.drop_columns([g for g in ("g1", "g2") if g not in by])   # a right column named like a left one is refused (duplicate names)
    c = gpu.col
    got = gpu.RecordBatch.from_arrow_table(left).asof_join(gpu.RecordBatch.from_arrow_table(right), [c(k) for k in by], [c(k) for k in by],
                                                           c("ts"), c("ts")).to_arrow_table()
    want = J.asof_join(left, right, by, by, "ts", "ts")
    assert got.column_names == want.column_names
    assert got.column("lv").to_pylist() == list(range(nl))          # left rows keep their order on the device path
    assert_tables_equal(_sorted(got, ["lv"]), _sorted(want, ["lv"]))


@pytest.mark.parametrize("nl,nr", [(0, 3), (3, 0), (3, 2), (257, 129)])
def test_cross_join_matches_oracle(gpu, nl, nr):
    left = pa.table({"a": np.arange(nl), "s": pa.array([f"v{i % 7}" if i % 5 else None for i in range(nl)], pa.large_string())})
    right = pa.table({"b": np.arange(nr) * 0.5, "t": pa.array([bool(i % 2) for i in range(nr)])})
    for side in ("left", "right"):
        got = gpu.RecordBatch.from_arrow_table(left).cross_join(gpu.RecordBatch.from_arrow_table(right), side).to_arrow_table()
        assert_tables_equal(got, J.cross_join(left, right, side))

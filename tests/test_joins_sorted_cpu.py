"""The oracle's sort-merge / asof / cross joins pinned on the reference's own test literals
(tests/dataframe/test_asof_join.py:89-330, tests/recordbatch/test_joins.py sort-merge cases) and cross-checked against the
oracle's hash join, whose pair tables are pinned on tests/recordbatch/test_joins.py:59-140."""
import numpy as np
import pyarrow as pa
import pytest

from helpers import assert_tables_equal
from oracle import joins_sorted as J
from oracle import oracle as O


def _sorted(tbl, keys=None):
    return tbl.sort_by([(k, "ascending") for k in (keys or tbl.column_names)])


# (left, right, left_by, right_by, left_on, right_on, sort keys, expected) -- tests/dataframe/test_asof_join.py
ASOF_LITERALS = [
    ({"ts": [10, 20, 30], "v": [1, 2, 3]}, {"ts": [10, 20, 30], "w": [11, 22, 33]}, [], [], "ts", "ts", ["ts"],
     {"ts": [10, 20, 30], "v": [1, 2, 3], "w": [11, 22, 33]}),                                                        # :89-94
    ({"entity": ["X", "X", "Y"], "ts": [100, 110, 100], "v": [7, 8, 9]},
     {"entity": ["X", "X", "X", "Y", "Y"], "ts": [90, 100, 110, 90, 110], "w": [1.1, 2.2, 3.3, 4.4, 5.5]}, ["entity"], ["entity"], "ts", "ts",
     ["entity", "ts"], {"entity": ["X", "X", "Y"], "ts": [100, 110, 100], "v": [7, 8, 9], "w": [2.2, 3.3, 4.4]}),      # :96-108
    ({"g1": ["P", "P"], "g2": [3, 3], "ts": [50, 80], "v": [11, 22]}, {"g1": ["P", "P", "P"], "g2": [3, 3, 7], "ts": [40, 70, 70], "w": [5, 9, 15]},
     ["g1", "g2"], ["g1", "g2"], "ts", "ts", ["ts"], {"g1": ["P", "P"], "g2": [3, 3], "ts": [50, 80], "v": [11, 22], "w": [5, 9]}),  # :110-121
    ({"ts_left": [5, 10, 15], "v": [1, 2, 3]}, {"ts_right": [3, 8, 12], "w": [100, 200, 300]}, [], [], "ts_left", "ts_right", ["ts_left"],
     {"ts_left": [5, 10, 15], "v": [1, 2, 3], "ts_right": [3, 8, 12], "w": [100, 200, 300]}),                         # :123-133
    ({"left_date": [20230101, 20230102, 20230103, 20230104, 20230106], "left_tag": ["foo", "bar", "foo", "bar", "foo"], "key": ["a", "b", "b", "a", "b"]},
     {"right_date": [20230102, 20230105, 20230108], "right_tag": ["x", "y", "z"], "key": ["a", "b", "b"]}, ["key"], ["key"], "left_date", "right_date",
     ["left_date"], {"left_date": [20230101, 20230102, 20230103, 20230104, 20230106], "left_tag": ["foo", "bar", "foo", "bar", "foo"],
                     "key": ["a", "b", "b", "a", "b"], "right_date": [None, None, None, 20230102, 20230105],
                     "right_tag": [None, None, None, "x", "y"]}),                                                     # :135-158
    ({"grp_l": ["M", "N"], "ts_l": [50, 50], "v": [7, 8]}, {"grp_r": ["M", "N"], "ts_r": [30, 40], "w": [70, 80]}, ["grp_l"], ["grp_r"], "ts_l", "ts_r",
     ["grp_l"], {"grp_l": ["M", "N"], "ts_l": [50, 50], "v": [7, 8], "ts_r": [30, 40], "w": [70, 80]}),                # :160-171
    ({"grp_l": ["M", "N"], "ts": [50, 50], "v": [7, 8]}, {"grp_r": ["M", "N"], "ts": [30, 40], "w": [70, 80]}, ["grp_l"], ["grp_r"], "ts", "ts",
     ["grp_l"], {"grp_l": ["M", "N"], "ts": [50, 50], "v": [7, 8], "w": [70, 80]}),                                   # :173-184
    ({"a": [1.0, 2.0, 3.0], "b": ["p", "q", "r"]}, {"a": [0.4, 1.6, 2.7], "c": ["s1", "s2", "s3"]}, [], [], "a", "a", ["a"],
     {"a": [1.0, 2.0, 3.0], "b": ["p", "q", "r"], "c": ["s1", "s2", "s3"]}),                                          # :205-215
    ({"ts": [2, 6, 12], "v": [10, 20, 30]}, {"ts": [4, 9], "w": [40, 90]}, [], [], "ts", "ts", ["ts"],
     {"ts": [2, 6, 12], "v": [10, 20, 30], "w": [None, 40, 90]}),                                                     # :217-224
    ({"ts": [1, 2, 3], "v": [10, 20, 30]}, {"ts": [100, 200], "w": [11, 22]}, [], [], "ts", "ts", ["ts"],
     {"ts": [1, 2, 3], "v": [10, 20, 30], "w": [None, None, None]}),                                                  # :226-231
    ({"ts": [100, 200, 300], "v": [1, 2, 3]}, {"ts": [10, 50], "w": [11, 55]}, [], [], "ts", "ts", ["ts"],
     {"ts": [100, 200, 300], "v": [1, 2, 3], "w": [55, 55, 55]}),                                                     # :233-238
    ({"ts": [6, 6, 11], "v": [1, 2, 3]}, {"ts": [4, 8], "w": [40, 80]}, [], [], "ts", "ts", ["ts", "v"],
     {"ts": [6, 6, 11], "v": [1, 2, 3], "w": [40, 40, 80]}),                                                          # :240-245
    ({"time": [2, 5, 5, 8], "ticker": ["AAPL", "GOOG", "AAPL", "GOOG"], "price": [150, 2800, 155, 2850]},
     {"time": [1, 3, 6, 9], "ticker": ["AAPL", "GOOG", "AAPL", "GOOG"], "bid": [149, 2790, 153, 2860]}, ["ticker"], ["ticker"], "time", "time",
     ["ticker", "time"], {"time": [2, 5, 5, 8], "ticker": ["AAPL", "AAPL", "GOOG", "GOOG"], "price": [150, 155, 2800, 2850],
                          "bid": [149, 149, 2790, 2790]}),                                                            # :256-283
    ({"b": [0.0, 1.25, 2.5, 3.75, 5.0, 6.25, 7.5], "c": ["x", "x", "x", "x", "y", "y", "y"]},
     {"val": [0.0, 3.0, 3.1, 3.2, 4.5, 5.5, 7.5], "c": ["x", "x", "x", "y", "y", "y", "y"], "b": [0.0, 3.0, 3.1, 3.2, 4.5, 5.5, 7.5]}, ["c"], ["c"], "b", "b",
     ["c", "b"], {"b": [0.0, 1.25, 2.5, 3.75, 5.0, 6.25, 7.5], "c": ["x", "x", "x", "x", "y", "y", "y"],
                  "val": [0.0, 0.0, 0.0, 3.1, 4.5, 5.5, 7.5]}),                                                       # :285-307
    ({"ts": [None, 2, None], "v": [1, 2, 3]}, {"ts": [1, 2, 3], "w": [10, 20, 30]}, [], [], "ts", "ts", ["v"],
     {"ts": [None, 2, None], "v": [1, 2, 3], "w": [None, 20, None]}),                                                 # :309-318
    ({"g": ["X", None], "ts": [5, 5], "v": [1, 2]}, {"g": ["X", None], "ts": [3, 3], "w": [10, 20]}, ["g"], ["g"], "ts", "ts", ["v"],
     {"g": ["X", None], "ts": [5, 5], "v": [1, 2], "w": [10, None]}),                                                 # :320-326
]


@pytest.mark.parametrize("case", range(len(ASOF_LITERALS)))
def test_oracle_asof_join_reference_literals(case):
    left, right, lby, rby, lon, ron, keys, want = ASOF_LITERALS[case]
    got = J.asof_join(pa.table(left), pa.table(right), lby, rby, lon, ron)
    assert got.column_names == list(want.keys())
    assert _sorted(got, keys).to_pydict() == want


def random_join_tables(rng, nl, nr, nulls=0.1, dom=12):
    def tbl(n, p):
        k = rng.integers(0, dom, n)
        s = np.array([f"s{v}" for v in rng.integers(0, 4, n)], dtype=object)
        return pa.table({"k": pa.array(k, mask=rng.random(n) < nulls), "s": pa.array(s, mask=rng.random(n) < nulls),
                         f"{p}v": pa.array(np.arange(n))})
    return tbl(nl, "l"), tbl(nr, "r")


@pytest.mark.parametrize("nl,nr", [(0, 5), (5, 0), (1, 1), (40, 60), (200, 150)])
def test_oracle_sort_merge_join_equals_oracle_hash_join(nl, nr):
    """two restatements of the same relation: the merge over sorted keys and the probe-table join (pinned on the reference's
    pair tables, tests/test_oracle_cpu.py) must produce the same rows"""
    rng = np.random.default_rng(nl * 7 + nr)
    left, right = random_join_tables(rng, nl, nr)
    for keys in (["k"], ["k", "s"]):
        r = right if len(keys) == 2 else right.rename_columns(["k", "rs", "rv"])   # O.hash_join leaves name clashes to its caller
        got = J.sort_merge_join(left, r, keys, keys)
        want = O.hash_join(left, r, keys, keys, "inner")
        assert_tables_equal(_sorted(got, ["lv", "rv"]), _sorted(want, ["lv", "rv"]))


def test_oracle_merge_join_reference_literals():
    """tests/recordbatch/test_joins.py sort-merge cases: duplicate runs on both sides, nulls never match"""
    pairs = J.merge_inner_join_pairs([pa.array([1, 2, 2, 3, None])], [pa.array([2, 2, 3, 4, None])])
    assert sorted(pairs) == [(1, 0), (1, 1), (2, 0), (2, 1), (3, 2)]
    assert J.merge_inner_join_pairs([pa.array([1, 2])], [pa.array([3, 4])]) == []      # range-wise disjoint (merge_join.rs:104-117)
    assert J.merge_inner_join_pairs([pa.array([], type=pa.int64())], [pa.array([1])]) == []


def test_oracle_cross_join():
    left, right = pa.table({"a": [1, 2, 3]}), pa.table({"b": ["x", "y"]})
    assert J.cross_join(left, right, "left").to_pydict() == {"a": [1, 1, 2, 2, 3, 3], "b": ["x", "y", "x", "y", "x", "y"]}
    assert J.cross_join(left, right, "right").to_pydict() == {"a": [1, 2, 3, 1, 2, 3], "b": ["x", "x", "x", "y", "y", "y"]}


# ---- the reference runs its join pair tables through BOTH strategies (tests/recordbatch/test_joins.py:30-140, join_impl in
# ("hash_join", "sort_merge_join"), null-safe equality skipped for the sort-merge join): the same literals through the merge restatement ----
@pytest.mark.parametrize("typ", [pa.int8(), pa.int64(), pa.uint64(), pa.float32(), pa.float64(), pa.large_utf8()])
@pytest.mark.parametrize("case", range(6))
def test_sort_merge_join_pairs_from_reference_tests(typ, case):
    from test_oracle_cpu import JOIN_PAIR_CASES   # imported here: tests/test_gpu_zzz2_joins_sorted.py imports this module for its tables
    assert len(JOIN_PAIR_CASES) == 6
    left, right, pairs, _ = JOIN_PAIR_CASES[case]

    def column(vals):
        if pa.types.is_large_string(typ):
            return pa.array([None if v is None else str(v) for v in vals], type=typ)
        return pa.array(vals, type=typ)
    lt = pa.table({"x": column(left), "x_ind": list(range(len(left)))})
    rt = pa.table({"y": column(right), "y_ind": list(range(len(right)))})
    res = J.sort_merge_join(lt, rt, ["x"], ["y"])
    assert res.column_names == ["x", "x_ind", "y", "y_ind"]
    assert sorted(zip(res.column("x_ind").to_pylist(), res.column("y_ind").to_pylist())) == sorted(pairs)
    res = J.sort_merge_join(rt, lt, ["y"], ["x"])   # tables swapped (test_joins.py:121-140)
    assert sorted(zip(res.column("x_ind").to_pylist(), res.column("y_ind").to_pylist())) == sorted(pairs)

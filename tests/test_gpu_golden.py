"""The CUDA path (through the C ABI) against the committed golden vectors of tests/golden/ -- no oracle involved."""
import json
import os

import numpy as np
import pyarrow as pa
import pytest

from test_golden_cpu import PA_TYPES, golden_column, load

pytestmark = pytest.mark.gpu


def RB(gpu, tbl):
    return gpu.RecordBatch.from_arrow_table(tbl)


def test_reference_literal_hashes(gpu):
    for case in load("reference_literals.json")["hash"]:
        rb = RB(gpu, pa.table({"x": golden_column(case["type"], case["values"])}))
        seed = None if case["seed"] is None else RB(gpu, pa.table({"s": pa.array(case["seed"], type=pa.uint64())}))
        assert rb.hash_rows([gpu.col("x")], seed=seed).to_pydict()["hash"] == case["expected"], case


@pytest.mark.parametrize("typ", sorted(PA_TYPES))
def test_hash_vectors(gpu, typ):
    g = load("hash_vectors.json")
    col = g["columns"][typ]
    rb = RB(gpu, pa.table({"x": golden_column(typ, col["values"])}))
    assert rb.hash_rows([gpu.col("x")]).to_pydict()["hash"] == col["unseeded"]
    seed = RB(gpu, pa.table({"s": pa.array(g["seeds"], type=pa.uint64())}))
    assert rb.hash_rows([gpu.col("x")], seed=seed).to_pydict()["hash"] == col["seeded"]


def test_hash_rows_vector(gpu):
    g = load("hash_vectors.json")
    names = g["hash_rows"]["columns"]
    rb = RB(gpu, pa.table({t: golden_column(t, g["columns"][t]["values"]) for t in names}))
    assert rb.hash_rows([gpu.col(t) for t in names]).to_pydict()["hash"] == g["hash_rows"]["expected"]


def test_tpch_q1_q6_goldens(gpu):
    from daft_b200 import tpch
    g = load("tpch_q1_q6_sf001.json")
    li = tpch.generate("lineitem", g["sf"], sorted(set(tpch.Q1_COLUMNS)))
    assert len(li) == g["lineitem_rows"]
    q1 = tpch.q1(li).to_pydict()
    for k, v in g["q1"].items():
        if isinstance(v[0], float):
            assert q1[k] == pytest.approx(v, rel=1e-9), k   # north_star tolerance is 1e-6; the summation order differs
        else:
            assert q1[k] == v, k
    assert tpch.q6(li).to_pydict()["revenue"] == pytest.approx(g["q6"]["revenue"], rel=1e-9)

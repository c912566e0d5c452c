"""The oracle against the committed golden vectors (tests/golden/, see make_goldens.py for provenance)."""
import json
import os

import numpy as np
import pyarrow as pa
import pytest

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PA_TYPES = {"int8": pa.int8(), "int16": pa.int16(), "int32": pa.int32(), "int64": pa.int64(), "uint8": pa.uint8(), "uint16": pa.uint16(),
            "uint32": pa.uint32(), "uint64": pa.uint64(), "float32": pa.float32(), "float64": pa.float64(), "bool": pa.bool_(),
            "date32": pa.date32(), "large_utf8": pa.large_utf8()}


def load(name):
    return json.load(open(os.path.join(GOLD, name)))


def golden_column(typ, values):
    if typ == "date32":
        return pa.array(values, type=pa.int32()).cast(pa.date32())
    return pa.array(values, type=PA_TYPES[typ])


def test_reference_literal_hashes():
    for case in load("reference_literals.json")["hash"]:
        arr = golden_column(case["type"], case["values"])
        seed = None if case["seed"] is None else np.array(case["seed"], dtype=np.uint64)
        assert O.hash_column(arr, seed).tolist() == case["expected"], case


@pytest.mark.parametrize("typ", sorted(PA_TYPES))
def test_hash_vectors(typ):
    g = load("hash_vectors.json")
    col = g["columns"][typ]
    arr = golden_column(typ, col["values"])
    assert O.hash_column(arr).tolist() == col["unseeded"]
    assert O.hash_column(arr, np.array(g["seeds"], dtype=np.uint64)).tolist() == col["seeded"]


def test_hash_rows_vector():
    g = load("hash_vectors.json")
    cols = [golden_column(t, g["columns"][t]["values"]) for t in g["hash_rows"]["columns"]]
    assert O.hash_rows(cols).tolist() == g["hash_rows"]["expected"]


def test_tpch_goldens_match_the_oracle_pipeline():
    from daft_b200 import tpch
    from oracle.recordbatch import OracleRecordBatch
    g = load("tpch_q1_q6_sf001.json")
    li = tpch.generate_host("lineitem", g["sf"], sorted(set(tpch.Q1_COLUMNS)))
    assert li.num_rows == g["lineitem_rows"]
    q1 = tpch.q1(OracleRecordBatch(li)).to_arrow_table().to_pydict()
    for k, v in g["q1"].items():
        assert q1[k] == pytest.approx(v, rel=1e-12) if isinstance(v[0], float) else q1[k] == v, k
    q6 = tpch.q6(OracleRecordBatch(li)).to_arrow_table().to_pydict()
    assert q6["revenue"] == pytest.approx(g["q6"]["revenue"], rel=1e-12)

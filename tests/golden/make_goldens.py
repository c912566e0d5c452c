"""Regenerates the GENERATED golden vectors of this directory (run from the repo root: python tests/golden/make_goldens.py).

* reference_literals.json is NOT generated: it restates literal expectations of the reference's own tests.
* hash_vectors.json: XXH3-64 of seeded random values per dtype, computed with the independent Python `xxhash` package
  (the package the reference's tests/series/test_hash.py compares against), using the reference's byte rules
  (src/daft-core/src/kernels/hashing.rs:304-445: native-width little-endian ints/floats, bool as ASCII "1"/"0", raw string bytes,
  primitive null = xxh3(b"") ignoring the seed, string null = xxh3(b"", seed)); plus the chained multi-column row hash
  (src/daft-recordbatch/src/ops/hash.rs:14-26).
* tpch_q1_q6_sf001.json: TPC-H Q1 / Q6 answers over this repo's deterministic generator at SF 0.01, computed by the CPU
  oracle (oracle/).  Pins generator + oracle + CUDA path against silent drift; the reference itself cannot run here
  (no Rust toolchain), so these are oracle-derived, not reference-derived.
"""
import json
import os
import struct
import sys

import numpy as np
import xxhash

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

PRIM_NULL = xxhash.xxh3_64_intdigest(b"")


def value_bytes(v, typ):
    if typ in ("int8", "int16", "int32", "int64", "uint8", "uint16", "uint32", "uint64"):
        width = int("".join(c for c in typ if c.isdigit())) // 8
        return int(v).to_bytes(width, "little", signed=typ.startswith("int"))
    if typ == "float32":
        return struct.pack("<f", v)
    if typ == "float64":
        return struct.pack("<d", v)
    if typ == "bool":
        return b"1" if v else b"0"
    if typ == "date32":
        return int(v).to_bytes(4, "little", signed=True)
    return v.encode() if isinstance(v, str) else bytes(v)


def hash_value(v, typ, seed):
    if v is None:
        return xxhash.xxh3_64_intdigest(b"", seed) if typ == "large_utf8" else PRIM_NULL
    return xxhash.xxh3_64_intdigest(value_bytes(v, typ), seed)


def make_hash_vectors():
    rng = np.random.default_rng(12345)
    n = 64
    cols = {
        "int8": [int(x) for x in rng.integers(-128, 128, n)], "int16": [int(x) for x in rng.integers(-2**15, 2**15, n)],
        "int32": [int(x) for x in rng.integers(-2**31, 2**31, n)], "int64": [int(x) for x in rng.integers(-2**62, 2**62, n)],
        "uint8": [int(x) for x in rng.integers(0, 256, n)], "uint16": [int(x) for x in rng.integers(0, 2**16, n)],
        "uint32": [int(x) for x in rng.integers(0, 2**32, n)], "uint64": [int(x) for x in rng.integers(0, 2**63, n)],
        "float32": [float(np.float32(x)) for x in rng.standard_normal(n)], "float64": [float(x) for x in rng.standard_normal(n)],
        "bool": [bool(x) for x in rng.integers(0, 2, n)], "date32": [int(x) for x in rng.integers(0, 20000, n)],
        "large_utf8": ["".join(chr(97 + int(c)) for c in rng.integers(0, 26, int(l))) for l in rng.integers(0, 40, n)],
    }
    for typ, vals in cols.items():  # ~10% nulls
        for i in rng.choice(n, 6, replace=False):
            vals[int(i)] = None
    seeds = [int(x) for x in rng.integers(0, 2**63, n)]
    out = {"n": n, "seeds": seeds, "columns": {}}
    for typ, vals in cols.items():
        out["columns"][typ] = {"values": vals, "unseeded": [hash_value(v, typ, 0) for v in vals],
                               "seeded": [hash_value(v, typ, s) for v, s in zip(vals, seeds)]}
    order = ["int64", "large_utf8", "bool", "float64", "date32", "uint8"]
    h = [hash_value(v, order[0], 0) for v in cols[order[0]]]
    for typ in order[1:]:
        h = [hash_value(v, typ, s) for v, s in zip(cols[typ], h)]
    out["hash_rows"] = {"columns": order, "expected": h}
    json.dump(out, open(os.path.join(HERE, "hash_vectors.json"), "w"), indent=1)


def make_tpch():
    from daft_b200 import tpch
    from oracle.recordbatch import OracleRecordBatch
    sf = 0.01
    li = tpch.generate_host("lineitem", sf, sorted(set(tpch.Q1_COLUMNS) | {"L_SHIPDATE", "L_DISCOUNT", "L_QUANTITY", "L_EXTENDEDPRICE"}))
    q1 = tpch.q1(OracleRecordBatch(li)).to_arrow_table()
    q6 = tpch.q6(OracleRecordBatch(li)).to_arrow_table()
    json.dump({"sf": sf, "lineitem_rows": li.num_rows, "q1": q1.to_pydict(), "q6": q6.to_pydict()},
              open(os.path.join(HERE, "tpch_q1_q6_sf001.json"), "w"), indent=1, default=str)


if __name__ == "__main__":
    make_hash_vectors()
    make_tpch()
    print("wrote", sorted(os.listdir(HERE)))

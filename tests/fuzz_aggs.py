"""Random aggregate plans (predicate + group keys + aggregates over random expressions) through the generated aggregate code on
the host (tests/jit_emu.py emulate_agg) against the oracle.

    python tests/fuzz_aggs.py [plans] [seed]
"""
import os
import sys

import numpy as np
import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fuzz_exprs as F  # noqa: E402
import jit_emu  # noqa: E402
from daft_b200 import col  # noqa: E402
from helpers import assert_tables_equal, oracle_agg, random_table  # noqa: E402

KEYS = ["k", "s", "b", "u8", "i8", "d", "i16", "u16"]


def plan(rng):
    nk = int(rng.integers(0, 4))
    keys = []
    for i in range(nk):
        if rng.random() < 0.75:
            name = KEYS[rng.integers(0, len(KEYS))]
            if name not in [k.name() for k in keys]:
                keys.append(col(name))
        else:
            keys.append((col(["i32", "k", "u16"][rng.integers(0, 3)]) % int(rng.integers(2, 6))).alias(f"g{i}"))
    aggs = []
    for i in range(int(rng.integers(1, 7))):
        x = F.gen_num(rng, int(rng.integers(0, 3)))
        kind = rng.integers(0, 7)
        a = [x.sum(), x.min(), x.max(), x.mean(), x.count(), x.count("all"), x.count("null")][kind]
        aggs.append(a.alias(f"a{i}"))
    pred = [F.gen_bool(rng, 2)] if rng.random() < 0.6 else None
    return aggs, keys, pred


def run(seed):
    rng = np.random.default_rng(seed)
    tbl = random_table(rng, int(rng.integers(0, 3000)), float(rng.choice([0.0, 0.15, 0.4])))
    while True:
        aggs, keys, pred = plan(rng)
        try:
            want = oracle_agg(tbl, aggs, keys, pred)
            break
        except (TypeError, ValueError, NotImplementedError, KeyError):
            continue
    got, first = jit_emu.emulate_agg(tbl, aggs, keys, pred)
    try:
        assert_tables_equal(got, want.select([a.name() for a in aggs]), float_rtol=1e-9)
        return None
    except AssertionError as err:
        return (aggs, keys, pred, str(err)[:300])


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    bad = 0
    for s in range(s0, s0 + n):
        r = run(s)
        print(f"seed {s}:", "ok" if r is None else "MISMATCH")
        if r is not None:
            bad += 1
            print("   aggs", r[0], "\n   keys", r[1], "\n   pred", r[2], "\n  ", r[3])
    sys.exit(1 if bad else 0)

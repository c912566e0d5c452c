"""Random expression trees through the generated projection code on the host (tests/jit_emu.py) against the oracle.

    python tests/fuzz_exprs.py [batches] [exprs_per_batch] [seed]

Every batch is one projection (one g++ compile).  A mismatch prints the expression and the first differing row.
"""
import datetime
import os
import sys

import numpy as np
import pyarrow as pa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import jit_emu  # noqa: E402
from daft_b200 import col, lit  # noqa: E402
from helpers import assert_tables_equal, random_table  # noqa: E402
from oracle import expr_eval as E  # noqa: E402

NUM = ["i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f32", "f64"]
INTS = NUM[:8]
CASTS = [pa.int8(), pa.int16(), pa.int32(), pa.int64(), pa.uint8(), pa.uint16(), pa.uint32(), pa.uint64(), pa.float32(), pa.float64()]


def gen_num(rng, depth):
    r = rng.random()
    if depth == 0 or r < 0.25:
        if rng.random() < 0.2:
            return lit([0, 1, -1, 7, 2.5, -0.5, 1000, None][rng.integers(0, 8)])
        return col(NUM[rng.integers(0, len(NUM))])
    if r < 0.32:   # leaves derived from the date / string columns
        return [col("d").year(), col("d").month(), col("d").day(), col("d").day_of_week(), col("s").length(), gen_str(rng, 1).length(),
                col("ts").hour(), col("ts").minute(), col("ts").second(), col("ts").truncate(["second", "minute", "hour", "day"][rng.integers(0, 4)]).minute()][rng.integers(0, 10)]
    a, b = gen_num(rng, depth - 1), gen_num(rng, depth - 1)
    k = rng.integers(0, 20)
    if k == 18: return a << lit(int(rng.integers(0, 8)))
    if k == 19: return a >> lit(int(rng.integers(0, 8)))
    if k == 13: return a & b
    if k == 14: return a | b
    if k == 15: return a ^ b
    if k == 16: return a.sqrt()
    if k == 17: return a.fill_nan(b)
    if k == 0: return a + b
    if k == 1: return a - b
    if k == 2: return a * b
    if k == 3: return a / b
    if k == 4: return a // b
    if k == 5: return a % b
    if k == 6: return a.cast(CASTS[rng.integers(0, len(CASTS))])
    if k == 7: return gen_bool(rng, depth - 1).if_else(a, b)
    if k == 8: return a.fill_null(b)
    if k == 9: return abs(a)
    if k == 10: return -a
    if k == 11: return a.sign()
    return a.floor() if rng.random() < 0.5 else a.ceil()


def gen_bool(rng, depth):
    r = rng.random()
    if depth == 0 or r < 0.15:
        return col("b")
    k = rng.integers(0, 20)
    if k == 12: return gen_num(rng, depth - 1).is_nan()
    if k == 13: return gen_num(rng, depth - 1).not_nan()
    if k == 14: return gen_num(rng, depth - 1).is_inf()
    if k == 15: return gen_num(rng, depth - 1).eq_null_safe(gen_num(rng, depth - 1))
    if k == 16: return gen_bool(rng, depth - 1) ^ gen_bool(rng, depth - 1)
    if k == 17:
        x, pat = gen_str(rng, depth - 1), ["", "a", "TPC", "d", "lo w", "é"][rng.integers(0, 6)]
        if rng.random() < 0.3:   # the regex VM
            return x.match([r"a+b?", r"^[a-d]+$", r"\w+ \w+", r"l{2}|c{3}", r"[^a-z ]", r"(TPC|x)-?H", r"^.{0,3}$", r"é|✓"][rng.integers(0, 8)])
        return [x.startswith(pat), x.endswith(pat), x.contains(pat)][rng.integers(0, 3)]
    if k == 18:
        x, y = gen_str(rng, depth - 1), gen_str(rng, depth - 1)
        return [x < y, x <= y, x == y, x != y, x >= y, x > y, x.is_in(["a", "ccc", "zz"]), x.between("a", "d")][rng.integers(0, 8)]
    if k == 19:
        d = datetime.date(1992 + int(rng.integers(0, 8)), int(rng.integers(1, 13)), int(rng.integers(1, 28)))
        return [col("d") < d, col("d") >= d, col("d") == d, col("d").is_null(), col("d").between(d, datetime.date(1999, 1, 1))][rng.integers(0, 5)]
    if k < 6:
        a, b = gen_num(rng, depth - 1), gen_num(rng, depth - 1)
        return [a < b, a <= b, a == b, a != b, a >= b, a > b][k]
    if k == 6: return gen_bool(rng, depth - 1) & gen_bool(rng, depth - 1)
    if k == 7: return gen_bool(rng, depth - 1) | gen_bool(rng, depth - 1)
    if k == 8: return ~gen_bool(rng, depth - 1)
    if k == 9: return gen_num(rng, depth - 1).is_null()
    if k == 10: return gen_num(rng, depth - 1).between(gen_num(rng, 0), gen_num(rng, 0))
    return gen_num(rng, depth - 1).is_in([0, 1, 3, -2])


def gen_str(rng, depth):
    r = rng.random()
    if depth == 0 or r < 0.4:
        return col("s") if rng.random() < 0.8 else lit(["a", "bb", "", "TPC-H", None][rng.integers(0, 5)])
    k = rng.integers(0, 4)
    if k == 3 and rng.random() < 0.4:
        pat, ngroups = [(r"(\w+) (\w+)", 2), (r"[a-d]+", 0), (r"(x+)|(y+)", 2), (r"^(.)(.*)$", 2)][rng.integers(0, 4)]
        return gen_str(rng, depth - 1).regexp_extract(pat, int(rng.integers(0, ngroups + 1)))
    if k == 3: return gen_str(rng, depth - 1).substr(int(rng.integers(0, 5)), [None, 0, 1, 3, 40][rng.integers(0, 5)])
    if k == 0: return gen_str(rng, depth - 1).left(int(rng.integers(0, 6)))
    if k == 1: return gen_bool(rng, depth - 1).if_else(gen_str(rng, depth - 1), gen_str(rng, depth - 1))
    return gen_str(rng, depth - 1).fill_null(gen_str(rng, depth - 1))


def batch(seed, nexpr, n=600, nulls=0.15):
    rng = np.random.default_rng(seed)
    tbl = random_table(rng, n, nulls)
    # a timestamp column of its own (random_table is shared with the device tests and stays as it is): instants on both sides of the epoch
    ts = pa.array(rng.integers(-10**15, 2 * 10**15, n), type=pa.timestamp("us"), mask=rng.random(n) < nulls)
    tbl = tbl.append_column("ts", ts)
    exprs = []
    while len(exprs) < nexpr:
        u = rng.random()
        e = gen_bool(rng, 3) if u < 0.35 else gen_str(rng, 3) if u < 0.45 else gen_num(rng, 3)
        try:
            E.eval_expression_list(tbl.slice(0, 4), [e.alias("x")])   # typing errors (bitwise on floats, ...) are not interesting
        except (TypeError, ValueError, NotImplementedError):
            continue
        exprs.append(e.alias(f"x{len(exprs)}"))
    return tbl, exprs


def run_batch(seed, nexpr):
    tbl, exprs = batch(seed, nexpr)
    got = jit_emu.emulate_project(tbl, exprs)
    want = E.eval_expression_list(tbl, exprs)
    bad = []
    for e in exprs:
        try:
            assert_tables_equal(got.select([e.name()]), want.select([e.name()]))
        except AssertionError as err:
            bad.append((e, str(err)[:300]))
    return bad


if __name__ == "__main__":
    nb = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    ne = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    s0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    total = 0
    for s in range(s0, s0 + nb):
        bad = run_batch(s, ne)
        total += len(bad)
        print(f"seed {s}: {ne - len(bad)}/{ne} ok")
        for e, msg in bad:
            print("   MISMATCH", e, "\n     ", msg)
    sys.exit(1 if total else 0)

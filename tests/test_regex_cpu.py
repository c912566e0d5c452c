"""Regular expressions on the device path: the host compiler (daft_b200/regex.py) and the Pike VM of the code generator's prelude
(`rx_exec`, executed here through the host emulation of the GENERATED code) against Python's `re` as the restatement of the regex crate's
leftmost-first semantics (src/daft-functions-utf8/src/{match_.rs,extract.rs,replace.rs}); NVRTC compiles the same kernels for sm_100a."""
import ctypes
import random
import re

import numpy as np
import pyarrow as pa
import pytest

from daft_b200 import _lib, col
from daft_b200 import datatypes as dt
from daft_b200 import regex as R
from daft_b200.expressions import encode_project
from jit_emu import emulate_project
from oracle import expr_eval as E

PATTERNS = [r"^https?://(?:www\.)?([^/]+)/.*$", r"a+b", r"(\d+)-(\d+)", r"[a-c]x|y+?z", r"^(foo|bar)*baz$", r"\w+@\w+\.com", r"x{2,3}y", r"(a|ab)(c|bcd)(d*)",
            r"^$", r"a*?b", r"[^a-z]+", r"\S+\s\S+", r"(?:ab){2,}", r"é+", r"(.)(.)", r"\W\w", r"[^\W\d]+", r"h.p", r"(\D+)\d", r"a|b|cd", r"^.{3}$", r"q[u]?e?"]
FIXED = ["http://www.google.com/a/b", "https://x.org/", "foo bar", "12-345", "a@b.com", "xxy", "abcd", "special requests", "", "日本é", "foobarbaz", "ababab", "que"]


def _strings(rng, n):
    alphabet = "abcxyz019 -/.@\nhtps:wé日"
    out = []
    for _ in range(n):
        out.append(rng.choice(FIXED) if rng.random() < 0.25 else "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 14))))
    return out


def test_reference_literals():
    """tests/recordbatch/utf8/test_extract.py / test_match.py shapes: the whole match, a group, no match -> null"""
    t = pa.table({"col": pa.array(["123-456", "789-012", "345-678", None, "abc"], pa.large_string())})
    got = E.eval_expression_list(t, [col("col").regexp_extract(r"\d+").alias("a"), col("col").regexp_extract(r"(\d+)-(\d+)", 2).alias("b"),
                                     col("col").match(r"^\d{3}-\d{3}$").alias("m")]).to_pydict()
    assert got == {"a": ["123", "789", "345", None, None], "b": ["456", "012", "678", None, None], "m": [True, True, True, None, False]}


def test_compiler_and_vm_model_vs_python_re():
    """the program bytes the host compiles, interpreted by a Python transcription of the VM, over every pattern x capture group"""
    rng = random.Random(1)
    for p in PATTERNS:
        _, ng = R.parse(p)
        rp = re.compile(R.to_python(p))
        for g in [None] + list(range(ng + 1)):
            prog = R.compile_program(p, g)
            assert prog[0] <= R.MAX_INST
            for t in _strings(rng, 150):
                m = rp.search(t)
                got = _vm(prog, t.encode())
                assert got[0] == (m is not None), (p, t)
                if m and g is not None:
                    sp = m.span(g)
                    assert got[1] == ((-1, -1) if sp == (-1, -1) else (len(t[:sp[0]].encode()), len(t[:sp[1]].encode()))), (p, g, t)


def _vm(prog: bytes, s: bytes):
    n = prog[0]
    inst = [prog[1 + 4 * i:5 + 4 * i] for i in range(n)]
    L = len(s)

    def cont(b):
        return (b & 0xC0) == 0x80

    def add(lst, added, pc, pos, c0, c1):
        stack = [(pc, c0, c1)]
        while stack:
            pc, a, b = stack.pop()
            while True:
                if pc in added:
                    break
                added.add(pc)
                op, x, y, z = inst[pc]
                if op == R.OP_JMP:
                    pc = x
                elif op == R.OP_SPLIT:
                    stack.append((y, a, b))
                    pc = x
                elif op == R.OP_SAVE:
                    a, b = (pos, b) if x == 0 else (a, pos)
                    pc += 1
                elif op == R.OP_BOL and pos == 0 or op == R.OP_EOL and pos == L or op == R.OP_TAIL and not (pos < L and cont(s[pos])):
                    pc += 1
                elif op in (R.OP_BOL, R.OP_EOL):
                    break
                else:
                    lst.append((pc, a, b))
                    break
    clist, cadded, matched, res = [], set(), False, (-1, -1)
    for pos in range(L + 1):
        if not matched:
            add(clist, cadded, 0, pos, -1, -1)
        nlist, nadded = [], set()
        for pc, a, b in clist:
            op, x, y, z = inst[pc]
            if op == R.OP_MATCH:
                matched, res = True, (a, b)
                break
            if pos >= L:
                continue
            ch = s[pos]
            if op == R.OP_CHAR:
                ok, nxt = ch == x, pc + 1
            elif op == R.OP_ANY:
                ok, nxt = ch != 10 and not cont(ch), pc + 1
            elif op == R.OP_TAIL:
                ok, nxt = True, pc
            else:
                off = x | (z << 8)
                ok = (any(prog[off + 1 + 2 * k] <= ch <= prog[off + 2 + 2 * k] for k in range(prog[off])) != bool(y & 1)) if ch < 128 else (bool(y & 2) and not cont(ch))
                nxt = pc + 1
            if ok:
                add(nlist, nadded, nxt, pos + 1, a, b)
        clist, cadded = nlist, nadded
        if not clist and matched:
            break
    return matched, res


@pytest.mark.parametrize("batch", range(4))
def test_generated_code_vs_oracle(batch):
    """the prelude's rx_exec itself (the generated kernel compiled for the host): match / extract of every group / anchored replace"""
    rng = random.Random(100 + batch)
    vals = _strings(rng, 400)
    t = pa.table({"s": pa.array(vals, pa.large_string(), mask=np.array([rng.random() < 0.1 for _ in vals]))})
    exprs = []
    for i, p in enumerate(PATTERNS[batch::4]):
        _, ng = R.parse(p)
        exprs.append(col("s").match(p).alias(f"m{i}"))
        for g in range(ng + 1):
            exprs.append(col("s").regexp_extract(p, g).alias(f"e{i}_{g}"))
        if R.is_fully_anchored(p) and ng >= 1:
            exprs.append(col("s").regexp_replace(p, r"\1").alias(f"r{i}"))
    got, want = emulate_project(t, exprs), E.eval_expression_list(t, exprs)
    for name in want.column_names:
        assert got.column(name).equals(want.column(name)), name


def test_kernels_compile_for_sm100a_and_bad_patterns_are_refused(tmp_path):
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    exprs = [col("s").regexp_replace(r"^https?://(?:www\.)?([^/]+)/.*$", r"\1").alias("host"), col("s").match(r"\w+@\w+\.com").alias("m"),
             col("s").regexp_extract(r"(\d+)-(\d+)", 2).alias("g")]
    plan, _, odt = encode_project(exprs, ["s"], [dt.LARGE_UTF8])
    assert odt == [dt.LARGE_UTF8, dt.BOOL, dt.LARGE_UTF8]
    for nullable in (0, 1):
        _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 1)(dt.LARGE_UTF8), (ctypes.c_uint8 * 1)(nullable), ctypes.c_int32(1))
    for bad in (r"(a", r"a)", r"\1", r"(?=a)", r"(?P<n>a)", r"[é]", r"*a", r"a{40}", "a" * 70, r"a{2,1}"):
        with pytest.raises(ValueError):
            R.compile_program(bad)
    with pytest.raises(NotImplementedError):
        col("s").regexp_replace(r"a+", "b")          # builds a new string: not a slice of the input
    with pytest.raises(NotImplementedError):
        col("s").regexp_replace(r"^(a)|b$", r"\1")   # not anchored on both sides of every alternative
    # a corrupted program is refused by the plan parser, not run
    good = encode_project([col("s").match(r"a+b|c")], ["s"], [dt.LARGE_UTF8])[0]
    prog = R.compile_program(r"a+b|c")
    i = good.index(prog)
    broken = bytearray(good)
    split = next(k for k in range(prog[0]) if prog[1 + 4 * k] == R.OP_SPLIT)
    broken[i + 1 + 4 * split + 1] = 200              # a SPLIT target outside the program
    out, err = ctypes.c_void_p(), ctypes.c_void_p()
    rc = _lib.lib().b200_jit_describe(ctypes.c_int32(0), bytes(broken), ctypes.c_size_t(len(broken)), (ctypes.c_int32 * 1)(dt.LARGE_UTF8), (ctypes.c_uint8 * 1)(0),
                                      ctypes.c_int32(1), ctypes.byref(out), ctypes.byref(err))
    assert rc != 0 and b"regex" in ctypes.string_at(err.value)
    _lib.lib().b200_free_string(err)

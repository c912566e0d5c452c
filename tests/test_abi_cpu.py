"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, refuses to run without a GPU (no CPU fallback), and the JIT code generator + NVRTC work offline."""
import ctypes
import os
import re

import pytest

from daft_b200 import _lib, datatypes as dt
from daft_b200.expressions import col, encode_agg, encode_project, lit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "daft_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = header_symbols()
    assert len(declared) >= 50
    for s in declared:
        assert hasattr(L, s), f"libdaftb200.so does not export {s}"
    assert sorted(_lib.SYMBOLS) == declared
    assert L.b200_abi_version() == 1


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.B200Error, match="no CUDA device|CUDA"):
        _lib.init(0)
    # every operator refuses as well
    out = ctypes.c_void_p()
    with pytest.raises(_lib.B200Error, match="no CPU fallback|CUDA"):
        _lib.call("b200_tpch_generate", ctypes.c_int32(0), ctypes.c_double(0.001), ctypes.c_int64(0), ctypes.c_int64(-1),
                  ctypes.c_uint64(1), None, ctypes.c_int32(0), ctypes.byref(out))


def test_error_convention_message_and_free():
    # bad plan bytes -> nonzero status + malloc'd message (daft-ext convention, abi/mod.rs:84-107)
    err = ctypes.c_void_p()
    rc = _lib.lib().b200_precompile(ctypes.c_int32(0), b"\xff\xff", ctypes.c_size_t(2), (ctypes.c_int32 * 1)(dt.I64),
                                    (ctypes.c_uint8 * 1)(0), ctypes.c_int32(1), ctypes.byref(err))
    assert rc != 0 and err.value
    assert b"plan" in ctypes.string_at(err.value)
    _lib.lib().b200_free_string(err)


def test_codegen_and_nvrtc_offline(tmp_path):
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    names, dts = ["a", "b", "s"], [dt.I64, dt.F64, dt.LARGE_UTF8]
    plan, onames, odt = encode_project([(col("a") + 1).alias("x"), (col("b") * 2.0 > 1).alias("y"), col("s").left(2).alias("z")], names, dts)
    assert onames == ["x", "y", "z"] and odt == [dt.I64, dt.BOOL, dt.LARGE_UTF8]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(*dts), (ctypes.c_uint8 * 3)(1, 0, 1), ctypes.c_int32(3))
    # `//` and `%` (Rust semantics: trunc_div_s / trunc_rem_s helpers, fmod for floats) must compile for every arithmetic type
    plan, _, odt = encode_project([(col("a") // 3).alias("q"), (col("a") % -3).alias("r"), (col("b") // 2.0).alias("fq"), (col("b") % 2.0).alias("fr")], names, dts)
    assert odt == [dt.I64, dt.I64, dt.F64, dt.F64]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(*dts), (ctypes.c_uint8 * 3)(1, 1, 1), ctypes.c_int32(3))
    assert any("trunc_div_s<" in open(os.path.join(tmp_path, f)).read() and "fmod(" in open(os.path.join(tmp_path, f)).read()
               for f in os.listdir(tmp_path) if f.endswith(".cu"))
    # Float32 `//` binds as floor(f64(l) / f64(r)) cast back to Float32 (series/ops/arithmetic.rs:466-469)
    plan, _, odt = encode_project([(col("x") // col("y")).alias("q"), (col("x") % col("y")).alias("r")], ["x", "y"], [dt.F32, dt.F32])
    assert odt == [dt.F32, dt.F32]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 2)(dt.F32, dt.F32), (ctypes.c_uint8 * 2)(1, 0), ctypes.c_int32(2))
    # narrowing / sign-changing / float -> int casts carry the arrow-rs safe-cast null rule: every source x target pair compiles
    import pyarrow as pa
    srcs = {"i8": dt.I8, "i16": dt.I16, "i32": dt.I32, "i64": dt.I64, "u8": dt.U8, "u16": dt.U16, "u32": dt.U32, "u64": dt.U64,
            "f32": dt.F32, "f64": dt.F64, "d": dt.DATE32, "bo": dt.BOOL}
    tgts = [pa.int8(), pa.int16(), pa.int32(), pa.int64(), pa.uint8(), pa.uint16(), pa.uint32(), pa.uint64(), pa.float32(), pa.float64()]
    exprs = [col(n).cast(t).alias(f"{n}_{i}") for n in srcs for i, t in enumerate(tgts)]
    plan, _, _ = encode_project(exprs, list(srcs), list(srcs.values()))
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * len(srcs))(*srcs.values()),
              (ctypes.c_uint8 * len(srcs))(*([0] * len(srcs))), ctypes.c_int32(len(srcs)))
    # float comparisons / between / is_in (totalOrder keys, NaN membership) for both float widths
    fl = [e for nm in ("x", "y") for e in ((col(nm) <= col("y")).alias(nm + "le"), (col(nm) == 0.0).alias(nm + "eq"), col(nm).between(0.0, 1.0).alias(nm + "bt"),
                                           col(nm).is_in([float("nan"), 0.5]).alias(nm + "in"), col(nm).eq_null_safe(col("y")).alias(nm + "ens"))]
    plan, _, _ = encode_project(fl, ["x", "y"], [dt.F32, dt.F64])
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 2)(dt.F32, dt.F64), (ctypes.c_uint8 * 2)(1, 1), ctypes.c_int32(2))
    # bitwise and / or / xor on integers (mixed widths through the supertype)
    plan, _, odt = encode_project([(col("p") & col("q")).alias("a"), (col("p") | 3).alias("o"), (col("q") ^ col("q")).alias("x")], ["p", "q"], [dt.I8, dt.U16])
    assert odt == [dt.I32, dt.I32, dt.U16]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 2)(dt.I8, dt.U16), (ctypes.c_uint8 * 2)(1, 0), ctypes.c_int32(2))
    # float predicates, ceil / floor / sign, fill_nan (lowered to if_else(not_nan(x), x, fill))
    fx = [e for nm in ("x", "y") for e in (col(nm).is_nan().alias(nm + "n"), col(nm).not_nan().alias(nm + "nn"), col(nm).is_inf().alias(nm + "i"),
                                           col(nm).ceil().alias(nm + "c"), col(nm).floor().alias(nm + "f"), col(nm).sign().alias(nm + "s"),
                                           col(nm).fill_nan(2.0).alias(nm + "fn"))]
    fx += [col("p").sign().alias("ps"), col("q").sign().alias("qs"), col("p").ceil().alias("pc"), col("x").sqrt().alias("xsq"),
           col("y").sqrt().alias("ysq"), col("p").sqrt().alias("psq")]
    plan, _, odt = encode_project(fx, ["x", "y", "p", "q"], [dt.F32, dt.F64, dt.I8, dt.U16])
    assert odt[-6:] == [dt.I8, dt.U16, dt.I8, dt.F32, dt.F64, dt.F32] and odt[3:7] == [dt.F32, dt.F32, dt.F32, dt.F64]
    _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 4)(dt.F32, dt.F64, dt.I8, dt.U16),
              (ctypes.c_uint8 * 4)(1, 1, 1, 0), ctypes.c_int32(4))
    plan = encode_agg([col("b").sum().alias("sb"), col("a").max().alias("ma"), col("a").count().alias("n")], [col("s"), col("a") % 3], names, dts,
                      [col("b") > lit(0.5)])
    _lib.call("b200_precompile", ctypes.c_int32(2), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(*dts), (ctypes.c_uint8 * 3)(1, 1, 1), ctypes.c_int32(3))
    # list / set / concat aggregates: the row -> group pass (variant 4) and the projection of their children
    plan = encode_agg([col("a").agg_list().alias("l"), (col("b") * 2).agg_set().alias("st"), col("s").agg_concat().alias("cat"), col("a").sum().alias("t")],
                      [col("s"), col("a") % 3], names, dts, [col("b") > lit(0.5)])
    _lib.call("b200_precompile", ctypes.c_int32(2), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(*dts), (ctypes.c_uint8 * 3)(1, 1, 1), ctypes.c_int32(3))
    plan = encode_agg([col("a").agg_list().alias("l")], [], names, dts)
    _lib.call("b200_precompile", ctypes.c_int32(2), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(*dts), (ctypes.c_uint8 * 3)(1, 1, 1), ctypes.c_int32(3))
    files = os.listdir(tmp_path)
    assert any(f.endswith(".cubin") for f in files) and any(f.endswith(".cu") for f in files)
    src = open(os.path.join(tmp_path, next(f for f in files if f.endswith(".cu")))).read()
    assert "b200_project_k" in src or "b200_agg_k" in src and "xxh" in src
    _lib.call("b200_set_jit_cache_dir", os.path.join(os.path.dirname(_lib.LIB_PATH), "_jitcache").encode())


def test_host_generator_is_deterministic_and_shaped():
    from daft_b200 import tpch
    a = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS)
    b = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS, 100, 200)
    assert a.slice(100, 100).equals(b)
    assert set(a.column("L_RETURNFLAG").to_pylist()) <= {"A", "N", "R"} and set(a.column("L_LINESTATUS").to_pylist()) == {"O", "F"}
    q = a.column("L_QUANTITY").to_pylist()
    assert min(q) >= 1 and max(q) <= 50
    assert tpch.num_rows("lineitem", 10) == 59986052 and tpch.num_rows("lineitem", 100) == 600037902
    n = tpch.generate_host("nation", 1)
    assert n.num_rows == 25 and n.column("N_NAME").to_pylist()[8] == "INDIA"


def test_timestamp_parts_and_string_if_else_compile_for_sm100a(tmp_path):
    """hour / minute / second / truncate of a Timestamp and a CASE that yields strings (ClickBench Q18 / Q39 / Q42): NVRTC -> sm_100a"""
    _lib.call("b200_set_jit_cache_dir", str(tmp_path).encode())
    names, dts = ["ts", "s", "k"], [dt.TIMESTAMP, dt.LARGE_UTF8, dt.I16]
    from daft_b200 import lit
    exprs = [col("ts").hour().alias("h"), col("ts").minute().alias("m"), col("ts").second().alias("sec"), col("ts").truncate("minute").alias("tm"),
             (col("k") == 0).if_else(col("s"), lit("")).alias("src")]
    plan, onames, odt = encode_project(exprs, names, dts)
    assert odt == [dt.U32, dt.U32, dt.U32, dt.TIMESTAMP, dt.LARGE_UTF8]
    for nullable in ((0, 0, 0), (1, 1, 1)):
        _lib.call("b200_precompile", ctypes.c_int32(0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(*dts), (ctypes.c_uint8 * 3)(*nullable), ctypes.c_int32(3))
    assert any("ts_minute(" in open(os.path.join(tmp_path, f)).read() for f in os.listdir(tmp_path) if f.endswith(".cu"))
    # as group-by keys of an aggregate (Q18: GROUP BY UserID, minute, SearchPhrase)
    from daft_b200.expressions import encode_agg
    plan = encode_agg([col("k").count("all").alias("c")], [col("ts").minute().alias("m"), col("s")], names, dts)
    _lib.call("b200_precompile", ctypes.c_int32(2), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * 3)(*dts), (ctypes.c_uint8 * 3)(0, 0, 0), ctypes.c_int32(3))
    with pytest.raises(ValueError, match="to be a timestamp"):
        encode_project([col("k").minute()], names, dts)


def test_generated_rust_bindings_cover_every_symbol():
    """integration/daft_b200_sys.rs (scripts/gen_rust_ffi.py) is the `extern "C"` block the reference's Rust host binds: it must be in step
    with the header and declare every exported symbol (there is no Rust toolchain here to compile it)"""
    import re
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fresh = subprocess.check_output([sys.executable, os.path.join(root, "scripts", "gen_rust_ffi.py")]).decode()
    committed = open(os.path.join(root, "integration", "daft_b200_sys.rs")).read()
    assert fresh == committed, "run: python scripts/gen_rust_ffi.py > integration/daft_b200_sys.rs"
    declared = set(re.findall(r"pub fn (b200_[a-z0-9_]+)\(", committed))
    assert declared == set(_lib.SYMBOLS)
    assert "pub fn b200_concat(batches: *const *const B200Batch, n: i32, out: *mut *mut B200Batch, errmsg: *mut *mut c_char) -> c_int;" in committed
    assert "columns: *const *const c_char" in committed and "seed: *const B200Batch" in committed


def test_malformed_plans_are_refused_not_crashed_on():
    """Plan bytes come from the host: flipped, truncated, padded and spliced plans must come back as an error (or as a different valid
    plan), never take the process down.  600 such mutations ran in subprocesses without a crash; a deterministic sample stays here."""
    import random
    from daft_b200 import tpch
    from daft_b200.expressions import encode_agg, encode_filter
    names, dts = tpch.Q1_COLUMNS, [dt.I64, dt.F64, dt.F64, dt.F64, dt.LARGE_UTF8, dt.LARGE_UTF8, dt.DATE32]
    pred, gb, aggs = tpch.q1_exprs()
    plans = [(2, encode_agg(aggs, gb, names, dts, pred)),
             (0, encode_project([(col("L_QUANTITY") * 2 + 1).alias("x"), col("L_RETURNFLAG").left(1).alias("y"), col("L_SHIPDATE").year().alias("z"),
                                 col("L_RETURNFLAG").like("%A%").alias("w")], names, dts)[0]),
             (1, encode_filter(pred + [col("L_DISCOUNT").between(0.01, 0.05), col("L_RETURNFLAG").is_in(["A", "R"])], names, dts))]
    random.seed(3)
    refused = 0
    for it in range(150):
        op, plan = plans[it % 3]
        b = bytearray(plan)
        k = it % 4
        if k == 0:
            for _ in range(random.randint(1, 3)):
                b[random.randrange(len(b))] = random.randrange(256)
        elif k == 1:
            b = b[: random.randrange(0, len(b))]
        elif k == 2:
            b += bytes(random.randrange(256) for _ in range(random.randint(1, 20)))
        else:
            i = random.randrange(len(b))
            b[i:i] = bytes(random.randrange(256) for _ in range(random.randint(1, 8)))
        out, err = ctypes.c_void_p(), ctypes.c_void_p()
        rc = _lib.lib().b200_jit_describe(ctypes.c_int32(op), bytes(b), ctypes.c_size_t(len(b)), (ctypes.c_int32 * 7)(*dts), (ctypes.c_uint8 * 7)(*[0] * 7),
                                          ctypes.c_int32(7), ctypes.byref(out), ctypes.byref(err))
        if rc != 0:
            refused += 1
            assert err.value
            _lib.lib().b200_free_string(err)
        else:
            _lib.lib().b200_free_string(out)
    assert refused > 100


def test_plain_c_host_compiles_links_and_fails_loudly_without_a_device(tmp_path):
    """examples/c_host.c: the header is C99, every symbol it uses links against the in-tree library, and without a GPU the first
    call reports that there is no CPU fallback"""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "c_host")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", f"-I{root}/include", f"{root}/examples/c_host.c", f"-L{root}/daft_b200", "-ldaftb200",
                           f"-Wl,-rpath,{root}/daft_b200", "-o", exe])
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([exe, "missing.parquet", "a"], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr

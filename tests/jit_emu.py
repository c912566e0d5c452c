"""Host emulation of the JIT-generated PROJECT code (test infrastructure).

The projection kernel the library generates for a plan is straight-line, fully typed C++ (`Prog::row`) over a small set
of helpers in `daft_b200/csrc/jit_prelude.cuh`.  This module asks the library for that source and its parameter layout
(`b200_jit_describe`, no CUDA involved), compiles the very same text for the host with a shim for the handful of CUDA
intrinsics it uses, and runs it over Arrow inputs one emulated warp at a time.  Warp ballots -- the only cross-lane
operation of `row` -- are replayed in two sweeps per warp: sweep 0 records every lane's vote per call site, sweep 1 re-runs
the (pure, idempotent) row code with the recorded masks, so the bitmaps it stores are the ones a real warp would store.

Aggregate plans (emulate_agg: load / eval / keys_equal / code / global_update, the staged build's stage_issue / load_staged) and
filter plans (emulate_filter: keep / copy) are driven the same way further down.

What this proves: the generated expression code (casts, division, comparisons, null rules, functions) computes what the
oracle computes.  What it does not: anything about the launch, the memory system or the skeletons around `row`.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

import numpy as np
import pyarrow as pa

from daft_b200 import _lib
from daft_b200 import datatypes as dt
from daft_b200.expressions import encode_project

P_COL_DATA, P_COL_VALID, P_COL_OFFSETS, P_OUT_DATA, P_OUT_VALID, P_OUT_LEN, P_OUT_OFFSETS, P_TAKE = 0, 1, 2, 3, 4, 5, 6, 7
I_N = 0

_NP = {dt.I8: np.int8, dt.I16: np.int16, dt.I32: np.int32, dt.I64: np.int64, dt.U8: np.uint8, dt.U16: np.uint16, dt.U32: np.uint32,
       dt.U64: np.uint64, dt.F32: np.float32, dt.F64: np.float64, dt.DATE32: np.int32, dt.TIMESTAMP: np.int64}

_SHIM = r"""
#include <cstdint>
#include <cstring>
#include <math.h>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __global__
#define __launch_bounds__(...)
#define __grid_constant__
struct EmuDim3 { unsigned x, y, z; };
static EmuDim3 threadIdx, blockIdx, blockDim, gridDim;
static unsigned g_ballots[1 << 14];
static int g_site, g_phase, g_lane;
static inline unsigned __ballot_sync(unsigned, bool pred) {
    const int s = g_site++;
    if (g_phase == 0) { if (pred) g_ballots[s] |= 1u << g_lane; return 0u; }
    return g_ballots[s];
}
static inline long long __double_as_longlong(double x) { long long b; memcpy(&b, &x, 8); return b; }
static inline double __longlong_as_double(long long b) { double x; memcpy(&x, &b, 8); return x; }
static inline int __float_as_int(float x) { int b; memcpy(&b, &x, 4); return b; }
static inline float __int_as_float(int b) { float x; memcpy(&x, &b, 4); return x; }
static inline int __double2hiint(double x) { return (int)(__double_as_longlong(x) >> 32); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> static inline T max(T a, T b) { return a < b ? b : a; }
"""

_DRIVER = r"""
extern "C" void emu_run(void **ptrs, int nptr, long long *ints, int nint, long long n, int pass) {
    const long long *take = Prog::PTR_TAKE >= 0 ? (const long long *)ptrs[Prog::PTR_TAKE < 0 ? 0 : Prog::PTR_TAKE] : nullptr;
    Prog::P p;
    for (int i = 0; i < nptr; i++) p.ptr[i] = ptrs[i];
    for (int i = 0; i < nint; i++) p.ival[i] = ints[i];
    const long long n_round = (n + 31) & ~31LL;
    for (long long base = 0; base < n_round; base += 32) {
        memset(g_ballots, 0, sizeof(g_ballots));
        for (g_phase = 0; g_phase < 2; g_phase++)
            for (g_lane = 0; g_lane < 32; g_lane++) {
                g_site = 0;
                threadIdx.x = (unsigned)g_lane;
                const long long i = base + g_lane;
                Prog::row(p, i, (take && i < n) ? take[i] : i, i < n, pass);   // project_body: src = take[i] when gathering
            }
    }
}
"""


def _prelude_for_host() -> str:
    xxh = open(os.path.join(_lib.CSRC, "xxh3.cuh")).read().replace("#pragma once", "")   # prepended to every generated kernel
    text = xxh + open(os.path.join(_lib.CSRC, "jit_prelude.cuh")).read()
    cut = text.index("// Filter with in-kernel stable compaction")
    cut = text.rindex("// ====", 0, cut)
    text = text[:cut]
    text = re.sub(r"asm volatile\([^;]*\);", ";", text)   # PTX-only helpers (prefetch) are never reached from Prog::row
    return text.replace("#pragma once", "") + "\n}  // namespace jit\n}  // namespace b200\n"


_NVRTC_SEEN = set()


def _nvrtc_too(op: int, plan: bytes, dtypes, nullable) -> None:
    """B200_EMU_NVRTC=<dir>: every plan the emulation runs is ALSO compiled by NVRTC for sm_100a (b200_precompile: all the build variants of
    an aggregate), cubins into <dir> -- a sweep over whole query suites that no kernel the queries need fails to compile for the device."""
    d = os.environ.get("B200_EMU_NVRTC")
    key = (op, bytes(plan), tuple(dtypes), tuple(nullable))
    if not d or key in _NVRTC_SEEN:
        return
    _NVRTC_SEEN.add(key)
    os.makedirs(d, exist_ok=True)
    _lib.call("b200_set_jit_cache_dir", d.encode())
    n = len(dtypes)
    _lib.call("b200_precompile", ctypes.c_int32(op), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * n)(*dtypes), (ctypes.c_uint8 * n)(*nullable), ctypes.c_int32(n))


def describe_project(exprs, names, dtypes, nullable, with_take: bool = False):
    plan, onames, odt = encode_project(exprs, names, dtypes)
    _nvrtc_too(0, plan, dtypes, nullable)
    out, err = ctypes.c_void_p(), ctypes.c_void_p()
    n = len(dtypes)
    rc = _lib.lib().b200_jit_describe(ctypes.c_int32((1 << 16) if with_take else 0), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * n)(*dtypes),
                                      (ctypes.c_uint8 * n)(*nullable), ctypes.c_int32(n), ctypes.byref(out), ctypes.byref(err))
    _lib.check(rc, err)
    text = ctypes.string_at(out.value).decode()
    _lib.lib().b200_free_string(out)
    head, src = text.split("\n---\n", 1)
    fields = dict(line.split("=", 1) for line in head.split("\n"))
    ptrs = [tuple(map(int, s.split(":"))) for s in fields["ptrs"].split(",") if s]
    ints = [int(s) for s in fields["ints"].split(",") if s]
    outs = [tuple(map(int, s.split(":"))) for s in fields["outs"].split(",") if s]
    return src, ptrs, ints, outs, onames, odt



def _build_so(so: str, cpp: str, source: str) -> None:
    """g++ the generated source into `so`.  Several processes may want the same kernel at the same moment (the world-2 gloo tests): each
    writes under a name of its own and renames into place, so that nobody ever loads a half-written library."""
    uniq = f".{os.getpid()}.tmp"
    with open(cpp + uniq, "w") as f:
        f.write(source)
    # -ffp-contract=off: the device code is compiled with --fmad=false.  -fno-builtin-fmod*: gcc otherwise reasons that
    # fabs(fmod(x, y)) == fmod(x, y) for non-negative x and drops the fabs, which is visible when libm returns the
    # negative x86 default NaN for fmod(x, 0).
    try:
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fno-builtin-fmod", "-fno-builtin-fmodf", "-fPIC", "-shared", "-w", "-x", "c++",
                               *os.environ.get("B200_EMU_FLAGS", "").split(), "-o", so + uniq, cpp + uniq])
        os.replace(so + uniq, so)
    finally:
        for leftover in (cpp + uniq, so + uniq):
            if os.path.exists(leftover):
                os.remove(leftover)


def _compile(src: str) -> ctypes.CDLL:
    body = "\n".join(l for l in src.split("\n") if "__global__" not in l)
    full = _SHIM + _prelude_for_host() + body + _DRIVER
    tag = hashlib.sha1(full.encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), os.environ.get("B200_EMU_DIR", "daft_b200_jit_emu"))
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"emu_{tag}.so")
    if not os.path.exists(so):
        _build_so(so, os.path.join(d, f"emu_{tag}.cpp"), full)
    return ctypes.CDLL(so)


def _pack_bits(mask: np.ndarray, nwords32: int) -> np.ndarray:
    out = np.zeros(nwords32 * 4, dtype=np.uint8)
    packed = np.packbits(mask.astype(np.uint8), bitorder="little")
    out[:len(packed)] = packed
    return out


def emulate_project(tbl: pa.Table, exprs, take=None) -> pa.Table:
    """eval_expression_list through the generated `Prog::row`, on the host.  take: int64 row indices to evaluate at (the
    gathering build op_project uses for group keys, any_value and join outputs); a negative index yields a null row."""
    tbl = tbl.combine_chunks()
    n_in = tbl.num_rows
    n = n_in if take is None else len(take)
    names = tbl.column_names
    cols = [tbl.column(i).chunk(0) if tbl.column(i).num_chunks else pa.array([], type=tbl.column(i).type) for i in range(len(names))]
    dtypes = [dt.from_arrow(c.type) for c in cols]
    nullable = [1 if c.null_count > 0 else 0 for c in cols]
    src, ptrs, ints, outs, onames, odt = describe_project(exprs, names, dtypes, nullable, take is not None)
    lib = _compile(src)
    nw = max((n + 31) // 32, 1)
    nw_in = max((n_in + 31) // 32, 1)
    keep, slots, out_bufs = [], [], {}
    for kind, idx in ptrs:
        if kind == P_TAKE:
            buf = np.concatenate([np.asarray(take, dtype=np.int64), np.zeros(32, np.int64)])
        elif kind in (P_COL_DATA, P_COL_VALID, P_COL_OFFSETS):
            c, code = cols[idx], dtypes[idx]
            if kind == P_COL_VALID:
                buf = _pack_bits(np.asarray(c.is_valid()), nw_in + 1)
            elif kind == P_COL_OFFSETS:
                buf = np.frombuffer(c.buffers()[1], dtype=np.int64, count=n_in + 1 + c.offset)[c.offset:].copy() if n_in else np.zeros(1, np.int64)
            elif dt.is_string(code):
                raw = c.buffers()[2]
                buf = np.concatenate([np.frombuffer(raw, dtype=np.uint8) if raw is not None else np.zeros(0, np.uint8), np.zeros(16, np.uint8)])
            elif code == dt.BOOL:
                buf = _pack_bits(np.asarray(c.fill_null(False)), nw_in + 1)
            elif code == dt.NULL:
                buf = np.zeros(8, np.uint8)
            else:
                phys = c.cast(pa.int32()) if code == dt.DATE32 else c.cast(pa.int64()) if code == dt.TIMESTAMP else c
                buf = np.ascontiguousarray(np.concatenate([np.asarray(phys.fill_null(0)).astype(_NP[code]), np.zeros(4, _NP[code])]))
        elif kind == P_OUT_VALID:
            buf = np.zeros(nw, dtype=np.uint32)
            out_bufs[("valid", idx)] = buf
        elif kind == P_OUT_DATA:
            code = outs[idx][0]
            if code in (dt.BOOL, dt.NULL):
                buf = np.zeros(nw, dtype=np.uint32)
            elif dt.is_string(code):
                buf = None   # allocated after pass 0, once the lengths are known
            else:
                buf = np.zeros(nw * 32, dtype=_NP[code])
            out_bufs[("data", idx)] = buf
        elif kind == P_OUT_LEN:
            buf = np.zeros(nw * 32, dtype=np.int64)
            out_bufs[("len", idx)] = buf
        elif kind == P_OUT_OFFSETS:
            buf = np.zeros(nw * 32 + 1, dtype=np.int64)
            out_bufs[("offsets", idx)] = buf
        else:
            raise NotImplementedError(f"host emulation: parameter kind {kind}")
        keep.append(buf)
        slots.append(buf.ctypes.data if buf is not None else 0)
    ivals = [n if k == I_N else 0 for k in ints]

    def run(pass_):
        lib.emu_run((ctypes.c_void_p * max(len(slots), 1))(*slots), ctypes.c_int(len(slots)),
                    (ctypes.c_longlong * max(len(ivals), 1))(*ivals), ctypes.c_int(len(ivals)), ctypes.c_longlong(n), ctypes.c_int(pass_))
    run(0)
    string_outs = [i for i, (code, _) in enumerate(outs) if dt.is_string(code)]
    if string_outs:
        # the two-pass protocol of op_project: pass 0 wrote one length per row, the host scans them into offsets and
        # allocates the bytes, pass 1 copies
        for i in string_outs:
            lens = out_bufs[("len", i)][:n]
            out_bufs[("offsets", i)][1:n + 1] = np.cumsum(lens)
            data = np.zeros(int(lens.sum()) + 16, dtype=np.uint8)
            out_bufs[("data", i)] = data
            slots[ptrs.index((P_OUT_DATA, i))] = data.ctypes.data
        run(1)
    arrays = []
    for i, (code, nullable_out) in enumerate(outs):
        valid = np.ones(n, dtype=bool)
        if ("valid", i) in out_bufs:
            valid = np.unpackbits(out_bufs[("valid", i)].view(np.uint8), bitorder="little")[:n].astype(bool)
        mask = None if valid.all() else ~valid
        if code == dt.NULL:
            arrays.append(pa.nulls(n))
        elif code == dt.BOOL:
            vals = np.unpackbits(out_bufs[("data", i)].view(np.uint8), bitorder="little")[:n].astype(bool)
            arrays.append(pa.array(vals, type=pa.bool_(), mask=mask))
        elif dt.is_string(code):
            off = out_bufs[("offsets", i)][:n + 1]
            raw = out_bufs[("data", i)].tobytes()
            items = [raw[off[j]:off[j + 1]] if valid[j] else None for j in range(n)]
            arrays.append(pa.array([x.decode() if x is not None else None for x in items], type=pa.large_utf8()) if code == dt.LARGE_UTF8
                          else pa.array(items, type=pa.large_binary()))
        else:
            vals = out_bufs[("data", i)][:n]
            arr = pa.array(vals, mask=mask)
            arrays.append(arr.cast(dt.to_arrow(code)) if arr.type != dt.to_arrow(code) else arr)
    return pa.table(arrays, names=list(onames))


# ------------------------------------------------------------------------------------------------------------------
# Aggregate plans: the generated load / eval (filter + key + value expressions) / keys_equal / code / global_update
# functions, driven row by row on the host.  Group lookup is a plain hash map keyed by Prog::code and confirmed with
# Prog::keys_equal; the cells are read back the way ops_jit.cu op_agg_first_rows finalises them.
# ------------------------------------------------------------------------------------------------------------------
P_SLOT_HASH, P_ACC, P_OVERFLOW = 8, 10, 11
I_MASK, I_RSTRIDE = 1, 2
C_ADD_F64, C_ADD_F32, C_ADD_I64, C_MIN, C_MAX = 0, 1, 2, 3, 4
AGG_COUNT, AGG_SUM, AGG_MIN, AGG_MAX, AGG_MEAN, AGG_ANY_VALUE = 0, 1, 2, 3, 4, 5

_AGG_SHIM = r"""
#include <unordered_map>
#include <vector>
static bool g_commit = true;   // false while a warp-wide reduction is being replayed to its fixed point (see emu_agg_tier1)
static inline double atomicAdd(double *p, double v) { const double o = *p; if (g_commit) *p = o + v; return o; }
static inline float atomicAdd(float *p, float v) { const float o = *p; if (g_commit) *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (g_commit) *p = o + v; return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (g_commit && v > o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (g_commit && v < o) *p = v; return o; }
// Butterfly shuffles: lanes run one after the other, so a lane cannot see its partner's value of the same step.  The flush is
// pure, so it is replayed: every replay records each lane's value per shuffle call site and hands out the partner's value of the
// PREVIOUS replay; after as many replays as there are dependent steps every value is the true one.
static unsigned long long g_shfl_prev[512][32], g_shfl_cur[512][32];
static int g_shfl_site;
template <typename T> static inline T emu_shfl_xor(T v, int d) {
    const int s = g_shfl_site++;
    unsigned long long bits = 0;
    memcpy(&bits, &v, sizeof(T));
    g_shfl_cur[s][g_lane] = bits;
    T out;
    memcpy(&out, &g_shfl_prev[s][g_lane ^ d], sizeof(T));
    return out;
}
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int d) { return emu_shfl_xor(v, d); }
static inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long v, int d) { return emu_shfl_xor(v, d); }
static inline double __shfl_xor_sync(unsigned, double v, int d) { return emu_shfl_xor(v, d); }
static inline float __shfl_xor_sync(unsigned, float v, int d) { return emu_shfl_xor(v, d); }
static inline bool __all_sync(unsigned, bool p) { return p; }
struct PfLaneEmu;
"""

_AGG_DRIVER = r"""
extern "C" long long emu_agg(void **ptrs, int nptr, long long *ints, int nint, long long n, long long cap, int tier2) {
    std::vector<unsigned long long> t2c((size_t)Prog::T2 * Prog::NCELLS + 8);
    std::vector<char> t2_used((size_t)Prog::T2 + 1, 0);
    Prog::P p;
    for (int i = 0; i < nptr; i++) p.ptr[i] = ptrs[i];
    for (int i = 0; i < nint; i++) p.ival[i] = ints[i];
    std::unordered_multimap<uint64_t, long long> table;
    std::vector<Prog::Keys> group_keys;
    long long ngroups = 0;
    for (long long row = 0; row < n; row++) {
        Prog::Raw r;
        Prog::load(p, row, r);
        Prog::load2(p, row, r);
        Prog::Keys k;
        Prog::Vals v;
        if (!Prog::eval(p, r, k, v)) continue;
        long long slot = -1;
        if constexpr (Prog::NKEYS > 0) {
            const uint64_t code = Prog::code(k);
            auto range = table.equal_range(code);
            for (auto it = range.first; it != range.second; ++it)
                if (Prog::keys_equal(group_keys[it->second], k)) { slot = it->second; break; }
            if (slot < 0) {
                if (ngroups >= cap) return -1;
                slot = ngroups++;
                group_keys.push_back(k);
                table.emplace(code, slot);
            }
        } else {
            slot = 0;
            ngroups = 1;
        }
        if (tier2) {
            // the global build's shared-memory tier: the first T2 groups accumulate in a CTA-local table (t2_update) that is
            // flushed into the global cells at the end (t2_flush); later groups go straight to the global cells
            if (slot < Prog::T2) {
                if (!t2_used[slot]) { Prog::t2_init(t2c.data(), (int)slot); t2_used[slot] = 1; }
                Prog::t2_update(t2c.data(), (int)slot, v, row);
                continue;
            }
        }
        Prog::global_update(p, slot, v, row);
    }
    if (tier2)
        for (long long t = 0; t < Prog::T2 && t < ngroups; t++)
            if (t2_used[t]) Prog::t2_flush(p, t, t2c.data(), (int)t);
    return ngroups;
}
"""


_TIER1_DRIVER = r"""
// The low-cardinality build's accumulate path: per-lane private cells (lane_init / lane_update), ballot count cells (warp_cells) and
// the warp-wide flush into the global cells.  One emulated warp processes the rows 32 at a time; the other warps of the CTA keep
// their freshly initialised cells, which the flush folds in as it would on the device.
extern "C" long long emu_agg_tier1(void **ptrs, int nptr, long long *ints, int nint, long long n, long long cap) {
    Prog::P p;
    for (int i = 0; i < nptr; i++) p.ptr[i] = ptrs[i];
    for (int i = 0; i < nint; i++) p.ival[i] = ints[i];
    constexpr int NL = Prog::NLANE, NW = Prog::NWARP > 0 ? Prog::NWARP : 1;
    const size_t lstride = (size_t)NL * 32, wstride = NW;
    std::unordered_multimap<uint64_t, long long> table;
    std::vector<Prog::Keys> group_keys;
    std::vector<std::vector<unsigned long long>> lcells, wcells;
    long long ngroups = 0;
    auto new_group = [&]() {
        lcells.emplace_back((size_t)AGG_WARPS * lstride);
        wcells.emplace_back((size_t)AGG_WARPS * wstride, 0ULL);
        for (int w = 0; w < AGG_WARPS; w++)
            for (int l = 0; l < 32; l++) Prog::lane_init(lcells.back().data() + (size_t)w * lstride + l);
        return ngroups++;
    };
    if (Prog::NKEYS == 0) new_group();
    for (long long base = 0; base < n; base += 32) {
        Prog::Vals vals[32];
        long long gid[32];
        for (int l = 0; l < 32; l++) {
            gid[l] = -1;
            const long long row = base + l;
            if (row >= n) continue;
            Prog::Raw r;
            Prog::load(p, row, r);
            Prog::load2(p, row, r);
            Prog::Keys k;
            if (!Prog::eval(p, r, k, vals[l])) continue;
            long long slot = -1;
            if constexpr (Prog::NKEYS > 0) {
                const uint64_t code = Prog::code(k);
                auto range = table.equal_range(code);
                for (auto it = range.first; it != range.second; ++it)
                    if (Prog::keys_equal(group_keys[it->second], k)) { slot = it->second; break; }
                if (slot < 0) {
                    if (ngroups >= cap) return -1;
                    slot = new_group();
                    group_keys.push_back(k);
                    table.emplace(code, slot);
                }
            } else {
                slot = 0;
            }
            gid[l] = slot;
            Prog::lane_update(lcells[slot].data() + l, vals[l], row);   // warp 0, lane l
        }
        for (int l = 0; l < 32; l++) {
            const long long g = gid[l];
            if (g < 0) continue;
            bool first = true;
            for (int m = 0; m < l; m++) first = first && gid[m] != g;
            if (!first) continue;
            memset(g_ballots, 0, sizeof(g_ballots));
            for (g_phase = 0; g_phase < 2; g_phase++)
                for (g_lane = 0; g_lane < 32; g_lane++) {
                    g_site = 0;
                    Prog::warp_cells(vals[g_lane], gid[g_lane] == g, wcells[g].data(), g_lane);
                }
        }
    }
    for (long long g = 0; g < ngroups; g++) {
        memset(g_shfl_prev, 0, sizeof(g_shfl_prev));
        for (int rep = 0; rep < 8; rep++) {
            g_commit = rep == 7;
            for (g_lane = 0; g_lane < 32; g_lane++) {
                g_shfl_site = 0;
                Prog::flush(p, g, lcells[g].data(), lstride, wcells[g].data(), wstride, g_lane);
            }
            memcpy(g_shfl_prev, g_shfl_cur, sizeof(g_shfl_prev));
        }
        g_commit = true;
    }
    return ngroups;
}
"""


def describe_agg(plan: bytes, dtypes, nullable, variant: int = 0):
    _nvrtc_too(2, plan, dtypes, nullable)
    out, err = ctypes.c_void_p(), ctypes.c_void_p()
    n = len(dtypes)
    rc = _lib.lib().b200_jit_describe(ctypes.c_int32(2 | (variant << 8)), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * n)(*dtypes),
                                      (ctypes.c_uint8 * n)(*nullable), ctypes.c_int32(n), ctypes.byref(out), ctypes.byref(err))
    _lib.check(rc, err)
    text = ctypes.string_at(out.value).decode()
    _lib.lib().b200_free_string(out)
    head, src = text.split("\n---\n", 1)
    f = dict(line.split("=", 1) for line in head.split("\n"))
    aggs = []
    for s in f["aggs"].split(","):
        kind, mode, vdt, odt, lane, cnt, cn, cz, name = s.split(":", 8)
        aggs.append(dict(kind=int(kind), mode=int(mode), val_dtype=int(vdt), out_dtype=int(odt), lane_cell=int(lane), cnt_cell=int(cnt),
                         child_nullable=cn == "1", const_zero=cz == "1", name=name))
    for a, pair in zip(aggs, f.get("nancells", "").split(",")):   # float min / max: the NaN-forgetting fold (jit.h AggOut::nan_cell)
        a["nan_cell"], a["last_cell"] = (int(x) for x in pair.split(":")) if pair else (-1, -1)
    for a, fv in zip(aggs, f.get("firstvalid", "").split(",")):   # any_value(ignore_nulls): first valid row (jit.h AggOut::first_valid_cell)
        a["first_valid_cell"] = int(fv) if fv else -1
    return dict(src=src, ptrs=[tuple(map(int, s.split(":"))) for s in f["ptrs"].split(",") if s], ints=[int(s) for s in f["ints"].split(",") if s],
                nkeys=int(f["nkeys"]), nlane=int(f["nlane"]), nwarp=int(f["nwarp"]), aggs=aggs, cellops=[int(s) for s in f["cellops"].split(",") if s])


def _compile_agg(src: str, staged: bool = False, tier1: bool = False) -> ctypes.CDLL:
    lines = src.split("\n")
    defines = "\n".join(l for l in lines if l.startswith("#define"))     # hoisted in front of the prelude by the library too
    body = "\n".join(l for l in lines if "__global__" not in l and not l.startswith("#define"))
    text = open(os.path.join(_lib.CSRC, "jit_prelude.cuh")).read()
    # the aggregate skeletons are CUDA-only; the generated Prog needs the helpers above them plus PfLane / AGG_WARPS
    head = _prelude_for_host().rsplit("\n}  // namespace jit", 1)[0]
    m = re.search(r"struct PfLane \{.*?\};", text, re.S)
    extra = "#ifndef B200_AGG_WARPS\n#define B200_AGG_WARPS 8\n#endif\nconstexpr int AGG_WARPS = B200_AGG_WARPS;\n" + m.group(0) + "\n"
    full = _SHIM + _AGG_SHIM + (_STAGED_SHIM if staged else "") + defines + "\n" + head + extra + "\n}  // namespace jit\n}  // namespace b200\n" + body + \
        (_STAGED_DRIVER if staged else _TIER1_DRIVER if tier1 else _AGG_DRIVER)
    tag = hashlib.sha1(full.encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), os.environ.get("B200_EMU_DIR", "daft_b200_jit_emu"))
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"emuagg_{tag}.so")
    if not os.path.exists(so):
        _build_so(so, os.path.join(d, f"emuagg_{tag}.cpp"), full)
    return ctypes.CDLL(so)


def _input_buffer(kind, c, code, n, nw):
    if kind == P_COL_VALID:
        return _pack_bits(np.asarray(c.is_valid()), nw + 1)
    if kind == P_COL_OFFSETS:
        return np.frombuffer(c.buffers()[1], dtype=np.int64, count=n + 1 + c.offset)[c.offset:].copy() if n else np.zeros(1, np.int64)
    if dt.is_string(code):
        raw = c.buffers()[2]
        return np.concatenate([np.frombuffer(raw, dtype=np.uint8) if raw is not None else np.zeros(0, np.uint8), np.zeros(16, np.uint8)])
    if code == dt.BOOL:
        return _pack_bits(np.asarray(c.fill_null(False)), nw + 1)
    if code == dt.NULL:
        return np.zeros(8, np.uint8)
    phys = c.cast(pa.int32()) if code == dt.DATE32 else c.cast(pa.int64()) if code == dt.TIMESTAMP else c
    return np.ascontiguousarray(np.concatenate([np.asarray(phys.fill_null(0)).astype(_NP[code]), np.zeros(4, _NP[code])]))


def _decode_ord(cells: np.ndarray, code: int) -> np.ndarray:
    """Inverse of the order-preserving min / max keys (jit_prelude.cuh ord_*; ops_jit.cu finalize_one)."""
    if code in (dt.F32, dt.F64):
        nan = cells == np.uint64(0xFFFFFFFFFFFFFFFF)
        bits = np.where(cells >> np.uint64(63) != 0, cells & np.uint64(0x7FFFFFFFFFFFFFFF), ~cells)
        vals = bits.view(np.float64).copy()
        vals[nan] = np.nan
        return vals.astype(_NP[code])
    if dt.int_like_signed(code):
        return (cells ^ np.uint64(0x8000000000000000)).view(np.int64).astype(_NP[code])
    if code == dt.BOOL:
        return cells != 0
    return cells.astype(_NP[code])


def emulate_agg(tbl: pa.Table, aggs, group_by, predicate=None, staged: bool = False, tier2: bool = False, tier1: bool = False,
                takes_out: dict | None = None):
    """(aggregate columns in first-occurrence group order, first row of every group) through the generated aggregate code.
    staged=True runs the TMA-staged build of an ungrouped plan: full tiles go through stage_issue / load_staged."""
    from daft_b200.expressions import encode_agg
    tbl = tbl.combine_chunks()
    n = tbl.num_rows
    names = tbl.column_names
    cols = [tbl.column(i).chunk(0) if tbl.column(i).num_chunks else pa.array([], type=tbl.column(i).type) for i in range(len(names))]
    dtypes = [dt.from_arrow(c.type) for c in cols]
    nullable = [1 if c.null_count > 0 else 0 for c in cols]
    # tier2=True: the global-table build, its first T2 groups going through the shared-memory tier (t2_init / t2_update / t2_flush)
    d = describe_agg(encode_agg(aggs, group_by, names, dtypes, predicate), dtypes, nullable, 2 if staged else 1 if tier2 else 0)
    # tier1=True: the low-cardinality build's own accumulate path (lane_init / lane_update / warp_cells / flush)
    lib = _compile_agg(d["src"], staged, tier1)
    nw = max((n + 31) // 32, 1)
    cap = max(n, 1)
    ncell = d["nlane"] + d["nwarp"]
    acc = np.zeros((ncell, cap), dtype=np.uint64)
    for i, op in enumerate(d["cellops"]):
        if op == C_MIN:
            acc[i, :] = np.uint64(0xFFFFFFFFFFFFFFFF)
    keep, slots = [acc], []
    for kind, idx in d["ptrs"]:
        if kind in (P_COL_DATA, P_COL_VALID, P_COL_OFFSETS):
            buf = _input_buffer(kind, cols[idx], dtypes[idx], n, nw)
            keep.append(buf)
            slots.append(buf.ctypes.data)
        elif kind == P_ACC:
            slots.append(acc[idx].ctypes.data)
        elif kind in (P_SLOT_HASH, P_OVERFLOW):
            buf = np.zeros(cap + 8, dtype=np.uint64)
            keep.append(buf)
            slots.append(buf.ctypes.data)
        else:
            raise NotImplementedError(f"host emulation: aggregate parameter kind {kind}")
    ivals = [{I_N: n, I_MASK: cap - 1, I_RSTRIDE: 1}[k] for k in d["ints"]]
    if staged:
        lib.emu_agg_staged.restype = ctypes.c_longlong
        staged_rows = lib.emu_agg_staged((ctypes.c_void_p * len(slots))(*slots), ctypes.c_int(len(slots)), (ctypes.c_longlong * len(ivals))(*ivals),
                                         ctypes.c_int(len(ivals)), ctypes.c_longlong(n))
        assert staged_rows > 0 or n < 4096, "no tile went through the stage"
        ng = 1
    elif tier1:
        lib.emu_agg_tier1.restype = ctypes.c_longlong
        ng = lib.emu_agg_tier1((ctypes.c_void_p * len(slots))(*slots), ctypes.c_int(len(slots)), (ctypes.c_longlong * len(ivals))(*ivals),
                               ctypes.c_int(len(ivals)), ctypes.c_longlong(n), ctypes.c_longlong(cap))
    else:
        lib.emu_agg.restype = ctypes.c_longlong
        ng = lib.emu_agg((ctypes.c_void_p * len(slots))(*slots), ctypes.c_int(len(slots)), (ctypes.c_longlong * len(ivals))(*ivals),
                         ctypes.c_int(len(ivals)), ctypes.c_longlong(n), ctypes.c_longlong(cap), ctypes.c_int(1 if tier2 else 0))
    assert ng >= 0
    if d["nkeys"] == 0:
        ng = 1   # an ungrouped aggregate always yields one row (sums null, counts 0 over an empty / fully filtered input)
    nl = d["nlane"]
    nrows, minrow = acc[nl - 2, :ng], acc[nl - 1, :ng]
    order = np.argsort(minrow, kind="stable") if d["nkeys"] > 0 else np.arange(ng)
    arrays, out_names = [], []
    for a in d["aggs"]:
        cnt = (acc[nl + a["cnt_cell"], :ng] if a["cnt_cell"] >= 0 else nrows)[order]
        cell = acc[a["lane_cell"], :ng][order] if a["lane_cell"] >= 0 else None
        typ = dt.to_arrow(a["out_dtype"])
        if a["kind"] == AGG_ANY_VALUE:
            # no output column here: the caller gathers the child at these rows (ops_jit.cu cell_to_take; all ones = no valid row)
            if takes_out is not None:
                fv = a.get("first_valid_cell", -1)
                r = acc[fv, :ng][order] if fv >= 0 else minrow[order]
                takes_out[a["name"]] = np.where(r == np.uint64(0xFFFFFFFFFFFFFFFF), -1, r.astype(np.int64)).astype(np.int64)
            continue
        if a["kind"] == AGG_COUNT:
            arr = pa.array(np.zeros(ng, np.uint64) if a["const_zero"] else cnt, type=pa.uint64())
        else:
            valid = cnt > 0
            may_null = a["child_nullable"] or d["nkeys"] == 0
            mask = ~valid if may_null and not valid.all() else None
            if a["kind"] in (AGG_SUM, AGG_MEAN):
                v = a["val_dtype"]
                s = cell.view(np.float64) if v == dt.F64 else (cell & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32) if v == dt.F32 \
                    else cell if v == dt.U64 else cell.view(np.int64)
                if a["kind"] == AGG_MEAN:
                    with np.errstate(all="ignore"):
                        s = np.where(valid, s.astype(np.float64) / np.maximum(cnt, 1).astype(np.float64), 0.0)
                arr = pa.array(np.where(valid, s, 0).astype(_NP.get(a["out_dtype"], np.float64)), type=typ, mask=mask)
            else:
                vals = _decode_ord(cell, a["out_dtype"])
                if a.get("nan_cell", -1) >= 0:
                    # first pass + finalize only: NaN iff the group's last valid value is one (ops_jit.cu finalize_one); a NaN
                    # in the MIDDLE of a group needs the device's second pass, which this emulation does not run
                    nanrow, lastrow = acc[a["nan_cell"], :ng][order], acc[a["last_cell"], :ng][order]
                    assert not ((nanrow != 0) & (nanrow != lastrow)).any(), "emulation input has a NaN followed by more values in its group"
                    vals = np.where((nanrow != 0) & (nanrow == lastrow), np.nan, vals).astype(vals.dtype)
                if a["out_dtype"] == dt.BOOL:
                    arr = pa.array(vals & valid, type=pa.bool_(), mask=mask)
                else:
                    phys = pa.array(np.where(valid, vals, 0).astype(vals.dtype), mask=mask)
                    arr = phys.cast(typ) if phys.type != typ else phys
        arrays.append(arr)
        out_names.append(a["name"])
    return pa.table(arrays, names=out_names), minrow[order].astype(np.int64)


# ---- the TMA-staged aggregate build (variant 2): stage_issue copies column slices into a stage, load_staged reads the stage ----
_STAGED_SHIM = r"""
static unsigned char g_stage[1 << 20];
static inline void mbar_expect_tx(uint32_t, uint32_t) {}
static inline void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t) { memcpy(g_stage + dst, src, bytes); }
"""

_STAGED_DRIVER = r"""
extern "C" long long emu_agg_staged(void **ptrs, int nptr, long long *ints, int nint, long long n) {
    static_assert(Prog::NKEYS == 0, "the staged emulation covers ungrouped plans (what the library stages by default)");
    Prog::P p;
    for (int i = 0; i < nptr; i++) p.ptr[i] = ptrs[i];
    for (int i = 0; i < nint; i++) p.ival[i] = ints[i];
    long long staged_rows = 0;
    for (long long base = 0; base < n; base += Prog::STAGE_ROWS) {
        const bool full = base + Prog::STAGE_ROWS <= n;   // the skeleton stages full tiles and loads the tail directly
        if (full) { Prog::stage_issue(p, 0u, 0u, base); staged_rows += Prog::STAGE_ROWS; }
        for (int rr = 0; rr < Prog::STAGE_ROWS && base + rr < n; rr++) {
            Prog::Raw r;
            if (full) Prog::load_staged(p, g_stage, base + rr, rr, r);
            else { Prog::load(p, base + rr, r); Prog::load2(p, base + rr, r); }
            Prog::Keys k;
            Prog::Vals v;
            if (!Prog::eval(p, r, k, v)) continue;
            Prog::global_update(p, 0, v, base + rr);
        }
    }
    return staged_rows;
}
"""


# ---- Filter plans: the generated keep() (predicate) and copy() (column-major compaction of the fixed-width never-null columns) ----
_FILTER_DRIVER = r"""
extern "C" long long emu_filter(void **ptrs, int nptr, long long *ints, int nint, long long n, long long *kept_rows) {
    Prog::P p;
    for (int i = 0; i < nptr; i++) p.ptr[i] = ptrs[i];
    for (int i = 0; i < nint; i++) p.ival[i] = ints[i];
    long long nk = 0;
    for (long long row = 0; row < n; row++)
        if (Prog::keep(p, row)) kept_rows[nk++] = row;
    // the skeleton hands copy() FILTER_SUB (source row, destination row) pairs per thread, destination -1 for dropped rows
    for (long long base = 0; base < n; base += FILTER_SUB) {
        int64_t src[FILTER_SUB], dst[FILTER_SUB];
        for (int j = 0; j < FILTER_SUB; j++) {
            src[j] = base + j < n ? base + j : 0;
            dst[j] = -1;
            if (base + j < n) {
                // destination = rank of the row among the kept rows (binary search in kept_rows)
                long long lo = 0, hi = nk;
                while (lo < hi) { const long long mid = (lo + hi) / 2; if (kept_rows[mid] < base + j) lo = mid + 1; else hi = mid; }
                if (lo < nk && kept_rows[lo] == base + j) dst[j] = lo;
            }
        }
        Prog::copy(p, src, dst);
    }
    return nk;
}
"""


def emulate_filter(tbl: pa.Table, predicates) -> pa.Table:
    """RecordBatch::filter through the generated keep() and copy(); columns the kernel does not copy itself (nullable / string /
    boolean ones, gathered by the host through the row-id list) are taken with pyarrow at the kept rows."""
    from daft_b200.expressions import encode_filter
    tbl = tbl.combine_chunks()
    n = tbl.num_rows
    names = tbl.column_names
    cols = [tbl.column(i).chunk(0) if tbl.column(i).num_chunks else pa.array([], type=tbl.column(i).type) for i in range(len(names))]
    dtypes = [dt.from_arrow(c.type) for c in cols]
    nullable = [1 if c.null_count > 0 else 0 for c in cols]
    plan = encode_filter(list(predicates), names, dtypes)
    _nvrtc_too(1, plan, dtypes, nullable)
    out, err = ctypes.c_void_p(), ctypes.c_void_p()
    rc = _lib.lib().b200_jit_describe(ctypes.c_int32(1), plan, ctypes.c_size_t(len(plan)), (ctypes.c_int32 * len(dtypes))(*dtypes),
                                      (ctypes.c_uint8 * len(dtypes))(*nullable), ctypes.c_int32(len(dtypes)), ctypes.byref(out), ctypes.byref(err))
    _lib.check(rc, err)
    text = ctypes.string_at(out.value).decode()
    _lib.lib().b200_free_string(out)
    head, src = text.split("\n---\n", 1)
    f = dict(line.split("=", 1) for line in head.split("\n"))
    ptrs = [tuple(map(int, s.split(":"))) for s in f["ptrs"].split(",") if s]
    ints = [int(s) for s in f["ints"].split(",") if s]
    body = "\n".join(l for l in src.split("\n") if "__global__" not in l)
    full = _SHIM + "constexpr int FILTER_SUB = 8;\n" + _prelude_for_host() + body + _FILTER_DRIVER
    tag = hashlib.sha1(full.encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), os.environ.get("B200_EMU_DIR", "daft_b200_jit_emu"))
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"emufilter_{tag}.so")
    if not os.path.exists(so):
        _build_so(so, os.path.join(d, f"emufilter_{tag}.cpp"), full)
    lib = ctypes.CDLL(so)
    nw = max((n + 31) // 32, 1)
    keep, slots, copied = [], [], {}
    for kind, idx in ptrs:
        if kind in (P_COL_DATA, P_COL_VALID, P_COL_OFFSETS):
            buf = _input_buffer(kind, cols[idx], dtypes[idx], n, nw)
        elif kind == P_OUT_DATA:
            buf = np.zeros(max(n, 1) + 8, dtype=_NP[dtypes[idx]])
            copied[idx] = buf
        else:   # tile counts / row-id list: written by the skeleton, not by the generated code
            buf = np.zeros(max(n, 1) + 8, dtype=np.int64)
        keep.append(buf)
        slots.append(buf.ctypes.data)
    ivals = [n if k == I_N else 0 for k in ints]
    kept = np.zeros(max(n, 1), dtype=np.int64)
    lib.emu_filter.restype = ctypes.c_longlong
    nk = lib.emu_filter((ctypes.c_void_p * max(len(slots), 1))(*slots), ctypes.c_int(len(slots)), (ctypes.c_longlong * max(len(ivals), 1))(*ivals),
                        ctypes.c_int(len(ivals)), ctypes.c_longlong(n), kept.ctypes.data_as(ctypes.POINTER(ctypes.c_longlong)))
    rows = pa.array(kept[:nk], type=pa.int64())
    arrays = []
    for i, c in enumerate(cols):
        if i in copied:
            arr = pa.array(copied[i][:nk])
            arrays.append(arr.cast(c.type) if arr.type != c.type else arr)
        else:
            arrays.append(c.take(rows))
    return pa.table(arrays, names=names), sorted(copied)

// Device-side skeletons of the JIT-specialised pipeline kernels (sm_100a).
//
// The host (jit.cpp) turns a serialised plan (include/daft_b200.h "plan encoding") into a `Prog`
// struct -- straight-line, fully typed code for the filter / key / aggregate-input expressions --
// and instantiates one of the hand-written skeletons below with it:
//
//   project_body         Project (a11): one pass, outputs written at HBM speed, bitmaps built with
//                        warp ballots, strings in a length pass + copy pass.
//   filter_body          Filter (a10): count pass + stable compaction pass over 2048-row tiles.
//   agg_body             fused Filter -> Project -> HashAggregate (a4/a6/a10/a11), low cardinality,
//                        direct loads: every input column is read exactly once, nothing is materialised.
//   agg_body_staged      the same kernel fed by cp.async.bulk (TMA) stages + mbarriers.
//   agg_body_global      mid / high cardinality: shared-memory tier 2 + global open-addressing table.
//
// Group-by design, replacing the hashbrown raw-entry probe + per-accumulator scatter passes of
// src/daft-recordbatch/src/ops/inline_agg.rs:748-824,556-565 (details at each skeleton and in DESIGN.md 3.1):
//   * a group is identified by a 64-bit key CODE -- an injective packing of the key values when they fit 62 bits
//     (equal codes ARE equal keys), otherwise a fingerprint confirmed against the key values of the group's first
//     row (nulls equal, floats by bit pattern == the reference's "hash equal AND comparator" rule);
//   * tier 1: the first G1 groups a CTA meets live in per-lane private shared-memory cells (plain
//     read-modify-write, zero atomics, zero bank conflicts); a tier-1 kernel that keeps missing its groups raises
//     the status flag and the host reruns the plan with the global kernel;
//   * global kernel: tier 2 = shared-memory table with shared atomics, then red.global atomics on the group's
//     cells in a table whose layout (records vs one array per cell) the host picks by size;
//   * first-occurrence group order (inline_agg.rs:804-807) is recovered by tracking min(row) per group.
#pragma once

#define FULL_MASK 0xffffffffu

namespace b200 {
namespace jit {

struct Str {
    const uint8_t *p;
    int64_t len;
};

__device__ __forceinline__ bool bit_at(const uint8_t *bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }
__device__ __forceinline__ void prefetch_l2(const void *q) { asm volatile("prefetch.global.L2 [%0];" ::"l"(q)); }

// ---- order-preserving keys for min/max on any numeric type ---------------------------------------
// floats: all NaN collapse to the largest key (NaN sorts greatest, search_sorted.rs:238-245)
__device__ __forceinline__ uint64_t ord_f64(double v) {
    if (v != v) return 0xFFFFFFFFFFFFFFFFULL;
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double unord_f64(uint64_t k) {
    if (k == 0xFFFFFFFFFFFFFFFFULL) return __longlong_as_double(0x7FF8000000000000LL);
    uint64_t b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFULL) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ uint64_t ord_f32(float v) { return ord_f64((double)v); }
__device__ __forceinline__ uint64_t ord_i64(int64_t v) { return (uint64_t)v ^ 0x8000000000000000ULL; }
__device__ __forceinline__ uint64_t ord_u64(uint64_t v) { return v; }

// ---- string helpers ---------------------------------------------------------------------------------
// first min(len, 8) bytes of a string packed little-endian into a register (zero padded): short keys
// (TPC-H flags, codes) are then hashed and compared without touching memory again.
// Both packers are branch-free (selects + predicated loads of aligned words that contain bytes of the value, so
// they never leave the allocation) so that the loads of several rows/columns are all in flight before the
// first one is consumed.
__device__ __forceinline__ uint64_t str_pack8(Str a) {
    const uintptr_t addr = (uintptr_t)a.p;
    const bool any = a.len > 0;
    const int sh = (int)(addr & 7) * 8;
    const int nbits = (a.len < 8 ? (int)a.len : 8) * 8;
    const unsigned long long *w = (const unsigned long long *)(addr & ~(uintptr_t)7);
    unsigned long long w0 = 0ULL, w1 = 0ULL;
    if (any) w0 = __ldg(w);
    if (any && sh + nbits > 64) w1 = __ldg(w + 1);
    unsigned long long v = w0 >> sh;
    v |= sh ? (w1 << (64 - sh)) : 0ULL;
    v &= nbits >= 64 ? ~0ULL : ((1ULL << nbits) - 1ULL);
    return v;
}
// first min(len, 4) bytes with 32-bit arithmetic only (short key columns on the exact-code path need <= 3 bytes)
// first L (<= 4) bytes at a shared-memory address, little-endian
__device__ __forceinline__ uint32_t smem_pack4(const unsigned char *q, const int L) {
    uint32_t v = q[0];
    if (L > 1) v |= (uint32_t)q[1] << 8;
    if (L > 2) v |= (uint32_t)q[2] << 16;
    if (L > 3) v |= (uint32_t)q[3] << 24;
    return v;
}
__device__ __forceinline__ uint64_t str_pack4(Str a) {
    const uintptr_t addr = (uintptr_t)a.p;
    const bool any = a.len > 0;
    const uint32_t sh = (uint32_t)(addr & 3) * 8u;
    const uint32_t nb = a.len < 4 ? (uint32_t)a.len : 4u;
    const unsigned int *w = (const unsigned int *)(addr & ~(uintptr_t)3);
    unsigned int w0 = 0u, w1 = 0u;
    if (any) w0 = __ldg(w);
    if (any && sh + nb * 8u > 32u) w1 = __ldg(w + 1);
    unsigned int v = __funnelshift_r(w0, w1, sh);
    v &= nb >= 4u ? 0xFFFFFFFFu : ((1u << (nb * 8u)) - 1u);
    return (uint64_t)v;
}
__device__ __forceinline__ uint64_t hash_str_packed(Str a, uint64_t pk, uint64_t seed) {
    const int64_t len = a.len;
    if (len <= 8) {
        if (len == 0) return xxh::hash_empty(seed);
        if (len <= 3)
            return xxh::hash_1to3((uint32_t)(pk & 0xff), (uint32_t)((pk >> (8 * (len >> 1))) & 0xff),
                                  (uint32_t)((pk >> (8 * (len - 1))) & 0xff), (uint32_t)len, seed);
        return xxh::hash_4to8((uint32_t)pk, (uint32_t)(pk >> (8 * (len - 4))), (uint32_t)len, seed);
    }
    return xxh::hash_bytes(a.p, len, seed);
}
__device__ __forceinline__ bool str_eq_packed(Str a, uint64_t pa, Str b, uint64_t pb) {
    if (a.len != b.len || pa != pb) return false;
    for (int64_t i = 8; i < a.len; i++)
        if (a.p[i] != b.p[i]) return false;
    return true;
}
__device__ __forceinline__ bool str_eq(Str a, Str b) {
    if (a.len != b.len) return false;
    for (int64_t i = 0; i < a.len; i++)
        if (a.p[i] != b.p[i]) return false;
    return true;
}
__device__ __forceinline__ int str_cmp(Str a, Str b) {
    int64_t n = a.len < b.len ? a.len : b.len;
    for (int64_t i = 0; i < n; i++) {
        int d = (int)a.p[i] - (int)b.p[i];
        if (d) return d;
    }
    return a.len < b.len ? -1 : (a.len > b.len ? 1 : 0);
}
__device__ __forceinline__ bool str_startswith(Str a, Str pat) {
    if (pat.len > a.len) return false;
    for (int64_t i = 0; i < pat.len; i++)
        if (a.p[i] != pat.p[i]) return false;
    return true;
}
__device__ __forceinline__ bool str_endswith(Str a, Str pat) {
    if (pat.len > a.len) return false;
    const uint8_t *q = a.p + (a.len - pat.len);
    for (int64_t i = 0; i < pat.len; i++)
        if (q[i] != pat.p[i]) return false;
    return true;
}
__device__ __forceinline__ bool str_contains(Str a, Str pat) {
    if (pat.len == 0) return true;
    for (int64_t s = 0; s + pat.len <= a.len; s++) {
        int64_t i = 0;
        while (i < pat.len && a.p[s + i] == pat.p[i]) i++;
        if (i == pat.len) return true;
    }
    return false;
}
// first position >= start where `pat` occurs in `a`, -1 if none
__device__ __forceinline__ int64_t str_find_from(Str a, Str pat, int64_t start) {
    if (pat.len == 0) return start <= a.len ? start : -1;
    for (int64_t s = start; s + pat.len <= a.len; s++) {
        int64_t i = 0;
        while (i < pat.len && a.p[s + i] == pat.p[i]) i++;
        if (i == pat.len) return s;
    }
    return -1;
}
__device__ __forceinline__ int64_t str_char_len(Str a) {
    int64_t c = 0;
    for (int64_t i = 0; i < a.len; i++) c += (a.p[i] & 0xC0) != 0x80;
    return c;
}
__device__ __forceinline__ Str str_left(Str a, int64_t nchars) {
    if (nchars <= 0) return Str{a.p, 0};
    int64_t c = 0, i = 0;
    for (; i < a.len; i++) {
        if ((a.p[i] & 0xC0) != 0x80) {
            if (c == nchars) break;
            c++;
        }
    }
    return Str{a.p, i};
}

// ---- lstrip / rstrip / right / find (src/daft-functions-utf8/src/{lstrip,rstrip,right,find}.rs): views into the value ----
// Rust's trim_start / trim_end remove characters with the Unicode White_Space property
__device__ __forceinline__ bool uni_is_space(uint32_t c) {
    return (c >= 9 && c <= 13) || c == 32 || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) || c == 0x2028 || c == 0x2029 ||
           c == 0x202F || c == 0x205F || c == 0x3000;
}
__device__ __forceinline__ uint32_t utf8_decode(const uint8_t *p, int64_t avail, int &w) {
    const uint32_t b0 = p[0];
    if (b0 < 0x80 || avail < 2) { w = 1; return b0; }
    if (b0 < 0xE0) { w = 2; return ((b0 & 0x1F) << 6) | (p[1] & 0x3F); }
    if (b0 < 0xF0 && avail >= 3) { w = 3; return ((b0 & 0x0F) << 12) | ((uint32_t)(p[1] & 0x3F) << 6) | (p[2] & 0x3F); }
    if (avail >= 4) { w = 4; return ((b0 & 0x07) << 18) | ((uint32_t)(p[1] & 0x3F) << 12) | ((uint32_t)(p[2] & 0x3F) << 6) | (p[3] & 0x3F); }
    w = 1;
    return b0;
}
__device__ __forceinline__ Str str_lstrip(Str a) {
    int64_t i = 0;
    while (i < a.len) {
        int w;
        if (!uni_is_space(utf8_decode(a.p + i, a.len - i, w))) break;
        i += w;
    }
    return Str{a.p + i, a.len - i};
}
__device__ __forceinline__ Str str_rstrip(Str a) {
    int64_t j = a.len;
    while (j > 0) {
        int64_t k = j - 1;
        while (k > 0 && (a.p[k] & 0xC0) == 0x80) k--;
        int w;
        if (!uni_is_space(utf8_decode(a.p + k, j - k, w))) break;
        j = k;
    }
    return Str{a.p, j};
}
// Utf8 -> integer cast: the reference trims Unicode whitespace (src/daft-core/src/array/ops/cast.rs:164-173) and hands the value to
// arrow-cast 57.3.0's string parser (`parser_primitive!`, atoi 2.0 `from_radix_10_signed_checked`): an optional '+' or '-', then decimal
// digits only, to the end; the last byte must be a digit; a value the target type cannot hold -- or anything else -- is null
// (tests: cast.rs:2088-2150).  `bits` / `sg`: the target's width and signedness.  An unsigned target accepts "-0".
__device__ __forceinline__ bool str_parse_int(Str a, int bits, bool sg, unsigned long long &out) {
    a = str_rstrip(str_lstrip(a));
    out = 0;
    if (a.len == 0) return false;
    int64_t i = 0;
    bool neg = false;
    if (a.p[0] == '+' || a.p[0] == '-') { neg = a.p[0] == '-'; i = 1; }
    if (i == a.len) return false;
    unsigned long long acc = 0;
    for (; i < a.len; i++) {
        const unsigned d = (unsigned)a.p[i] - 48u;
        if (d > 9u) return false;
        if (acc > (0xFFFFFFFFFFFFFFFFULL - d) / 10ULL) return false;
        acc = acc * 10ULL + d;
    }
    const unsigned long long pos_max = sg ? ((1ULL << (bits - 1)) - 1ULL) : (bits == 64 ? 0xFFFFFFFFFFFFFFFFULL : ((1ULL << bits) - 1ULL));
    const unsigned long long neg_max = sg ? (1ULL << (bits - 1)) : 0ULL;
    if (acc > (neg ? neg_max : pos_max)) return false;
    out = neg ? (0ULL - acc) : acc;
    return true;
}
__device__ __forceinline__ Str str_right(Str a, int64_t nchars) {
    int64_t j = a.len, c = 0;
    while (j > 0 && c < nchars) {
        j--;
        while (j > 0 && (a.p[j] & 0xC0) == 0x80) j--;
        c++;
    }
    return Str{a.p + j, a.len - j};
}

// utf8 substr in CHARACTERS (src/daft-functions-utf8/src/substr.rs:269-294): the value has no character at `start`, or `nchars` == 0 -> null
// (ok = false); nchars < 0 = to the end of the value
__device__ __forceinline__ Str str_substr(Str a, int64_t start, int64_t nchars, bool &ok) {
    int64_t c = 0, i = 0;
    for (; i < a.len; i++) {
        if ((a.p[i] & 0xC0) != 0x80) {
            if (c == start) break;
            c++;
        }
    }
    if (i >= a.len || nchars == 0) {
        ok = false;
        return Str{a.p, 0};
    }
    ok = true;
    if (nchars < 0) return Str{a.p + i, a.len - i};
    int64_t j = i, cc = 0;
    for (; j < a.len; j++) {
        if ((a.p[j] & 0xC0) != 0x80) {
            if (cc == nchars) break;
            cc++;
        }
    }
    return Str{a.p + i, j - i};
}

// ---- regular expressions: a Pike VM over the program daft_b200/regex.py compiles on the host ------------------------
// Leftmost-first semantics (the `regex` crate's, src/daft-functions-utf8/src/{match_.rs,extract.rs,replace.rs}); O(len x instructions) per
// value, no backtracking.  Program: [ninst <= 64] [ninst x (op, a, b, c)] [class tables].  Steps over BYTES: `.` and classes that can
// match a non-ASCII character take a character's first byte, RX_TAIL its continuation bytes; threads never start inside a character.
enum { RX_MATCH = 0, RX_CHAR, RX_ANY, RX_CLASS, RX_SPLIT, RX_JMP, RX_SAVE, RX_BOL, RX_EOL, RX_TAIL };
struct RxResult {
    bool matched;
    int32_t c0, c1;   // byte span of the selected capture group (-1: it took no part in the match)
};
struct RxList {
    uint64_t added;
    int n;
    uint8_t pc[64];
    int32_t c0[64], c1[64];
};
__device__ __forceinline__ void rx_add(RxList &l, const uint8_t *prog, Str s, int pc0, int64_t pos, int32_t a0, int32_t b0) {
    uint8_t spc[64];
    int32_t sa[64], sb[64];
    int sp = 0;
    spc[0] = (uint8_t)pc0;
    sa[0] = a0;
    sb[0] = b0;
    sp = 1;
    while (sp > 0) {
        sp--;
        int pc = spc[sp];
        int32_t a = sa[sp], b = sb[sp];
        for (;;) {
            if ((l.added >> pc) & 1ULL) break;
            l.added |= 1ULL << pc;
            const uint8_t *in = prog + 1 + 4 * pc;
            const int op = in[0];
            if (op == RX_JMP) {
                pc = in[1];
                continue;
            }
            if (op == RX_SPLIT) {
                spc[sp] = in[2];   // the second branch waits: everything the first one reaches comes before it in the list
                sa[sp] = a;
                sb[sp] = b;
                sp++;
                pc = in[1];
                continue;
            }
            if (op == RX_SAVE) {
                if (in[1] == 0) a = (int32_t)pos;
                else b = (int32_t)pos;
                pc++;
                continue;
            }
            if (op == RX_BOL) {
                if (pos == 0) {
                    pc++;
                    continue;
                }
                break;
            }
            if (op == RX_EOL) {
                if (pos == s.len) {
                    pc++;
                    continue;
                }
                break;
            }
            if (op == RX_TAIL && !(pos < s.len && (s.p[pos] & 0xC0) == 0x80)) {
                pc++;
                continue;
            }
            l.pc[l.n] = (uint8_t)pc;
            l.c0[l.n] = a;
            l.c1[l.n] = b;
            l.n++;
            break;
        }
    }
}
__device__ __forceinline__ RxResult rx_exec(Str s, Str program) {
    const uint8_t *prog = program.p;
    RxList lists[2];
    int cur = 0;
    lists[0].added = 0;
    lists[0].n = 0;
    RxResult res{false, -1, -1};
    for (int64_t pos = 0; pos <= s.len; pos++) {
        RxList &cl = lists[cur], &nl = lists[cur ^ 1];
        if (!res.matched) rx_add(cl, prog, s, 0, pos, -1, -1);   // unanchored search: a fresh thread, behind every running one
        nl.added = 0;
        nl.n = 0;
        for (int t = 0; t < cl.n; t++) {
            const int pc = cl.pc[t];
            const uint8_t *in = prog + 1 + 4 * pc;
            const int op = in[0];
            if (op == RX_MATCH) {
                res.matched = true;
                res.c0 = cl.c0[t];
                res.c1 = cl.c1[t];
                break;   // threads behind this one have lower priority
            }
            if (pos >= s.len) continue;
            const uint8_t ch = s.p[pos];
            const bool cont = (ch & 0xC0) == 0x80;
            bool ok = false;
            int next = pc + 1;
            if (op == RX_CHAR) ok = ch == in[1];
            else if (op == RX_ANY) ok = ch != 10 && !cont;
            else if (op == RX_TAIL) {
                ok = true;   // parked on a continuation byte by rx_add
                next = pc;
            } else if (op == RX_CLASS) {
                if (ch < 128) {
                    const uint8_t *tab = prog + ((int)in[1] | ((int)in[3] << 8));
                    bool in_set = false;
                    for (int k = 0; k < tab[0]; k++) in_set = in_set || (tab[1 + 2 * k] <= ch && ch <= tab[2 + 2 * k]);
                    ok = in_set != ((in[2] & 1) != 0);
                } else {
                    ok = (in[2] & 2) != 0 && !cont;
                }
            }
            if (ok) rx_add(nl, prog, s, next, pos + 1, cl.c0[t], cl.c1[t]);
        }
        cur ^= 1;
        if (lists[cur].n == 0 && res.matched) break;
    }
    return res;
}

// ---- civil calendar from days since 1970-01-01 (proleptic Gregorian) --------------------------------
__device__ __forceinline__ void civil_from_days(int64_t z, int32_t &y, int32_t &m, int32_t &d) {
    z += 719468;
    const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    const uint32_t doe = (uint32_t)(z - era * 146097);
    const uint32_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const int64_t yy = (int64_t)yoe + era * 400;
    const uint32_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const uint32_t mp = (5 * doy + 2) / 153;
    d = (int32_t)(doy - (153 * mp + 2) / 5 + 1);
    m = (int32_t)(mp < 10 ? mp + 3 : mp - 9);
    y = (int32_t)(yy + (m <= 2));
}
__device__ __forceinline__ int32_t date_year(int32_t days) { int32_t y, m, d; civil_from_days(days, y, m, d); return y; }
__device__ __forceinline__ uint32_t date_month(int32_t days) { int32_t y, m, d; civil_from_days(days, y, m, d); return (uint32_t)m; }
__device__ __forceinline__ uint32_t date_day(int32_t days) { int32_t y, m, d; civil_from_days(days, y, m, d); return (uint32_t)d; }
__device__ __forceinline__ uint32_t date_dow(int32_t days) {  // Monday = 0
    int64_t r = ((int64_t)days + 3) % 7;
    return (uint32_t)(r < 0 ? r + 7 : r);
}

// ---- time of day of a Timestamp(us); negative instants count back from the epoch (floor semantics, as chrono does) ----
__device__ __forceinline__ int64_t ts_floor_mod(int64_t a, int64_t m) {
    const int64_t r = a % m;
    return r < 0 ? r + m : r;
}
__device__ __forceinline__ uint32_t ts_hour(int64_t us) { return (uint32_t)(ts_floor_mod(us, 86400000000LL) / 3600000000LL); }
__device__ __forceinline__ uint32_t ts_minute(int64_t us) { return (uint32_t)(ts_floor_mod(us, 3600000000LL) / 60000000LL); }
__device__ __forceinline__ uint32_t ts_second(int64_t us) { return (uint32_t)(ts_floor_mod(us, 60000000LL) / 1000000LL); }
__device__ __forceinline__ int64_t ts_trunc(int64_t us, int64_t unit) { return us - ts_floor_mod(us, unit); }

// ---- integer division helpers ------------------------------------------------------------------------
// The reference's `//` and `%` on integers are Rust's `/` and `%` (src/daft-core/src/array/ops/arithmetic.rs:245-300,
// series/ops/arithmetic.rs:456-466): quotient truncated toward zero, remainder with the sign of the dividend
// (its `pmod` function is the positive modulo).  Divisor 0 is handled by the caller as null (the reference panics);
// MIN / -1 wraps instead of trapping.
template <typename T>
__device__ __forceinline__ T trunc_div_s(T a, T b) {
    return b == (T)-1 ? (T)(0ULL - (unsigned long long)a) : (T)(a / b);
}
template <typename T>
__device__ __forceinline__ T trunc_rem_s(T a, T b) {
    return b == (T)-1 ? (T)0 : (T)(a % b);
}

// ---- float comparison keys ---------------------------------------------------------------------------
// The reference compares with the arrow-rs `cmp` kernels (src/daft-core/src/array/ops/comparison.rs:541-603), which order floats
// by IEEE 754 totalOrder: -NaN < -inf < ... < -0.0 < +0.0 < ... < +inf < +NaN, equality = equal bit patterns.  f_key maps the
// sign-magnitude pattern to a two's-complement integer with that order (f64::total_cmp); literals fold at compile time.
__device__ __forceinline__ long long f_key(double x) {
    const long long b = __double_as_longlong(x);
    return b ^ (long long)((unsigned long long)(b >> 63) >> 1);
}
__device__ __forceinline__ int f_key(float x) {
    const int b = __float_as_int(x);
    return b ^ (int)((unsigned)(b >> 31) >> 1);
}

__device__ __forceinline__ bool f_posnan(double x) { return x != x && __double2hiint(x) >= 0; }
__device__ __forceinline__ bool f_negnan(double x) { return x != x && __double2hiint(x) < 0; }
__device__ __forceinline__ bool f_posnan(float x) { return x != x && __float_as_int(x) >= 0; }
__device__ __forceinline__ bool f_negnan(float x) { return x != x && __float_as_int(x) < 0; }

// ---- kernel parameter block: 8-byte slots the host fills in the order codegen allocated them ---------
template <int NPTR, int NINT>
struct Params {
    void *ptr[NPTR > 0 ? NPTR : 1];
    long long ival[NINT > 0 ? NINT : 1];
};

// ======================================================================================================
// Filter / Project
// ======================================================================================================
// Prog provides:
//   using P = Params<..>;  static constexpr int IV_N (ival slot of the row count), PTR_TAKE (ptr slot of
//   the optional int64 gather index, -1 if none)
//   __device__ static void row(const P&, int64_t out_row, int64_t src_row, bool active, int pass)
// `row` is called by ALL lanes (inactive ones too) because bitmap outputs are built with warp ballots.
template <class Prog>
__device__ __forceinline__ void project_body(const typename Prog::P &p, int pass) {
    const int64_t n = p.ival[Prog::IV_N];
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool active = i < n;
        int64_t src = i;
        if (Prog::PTR_TAKE >= 0 && active) src = ((const int64_t *)p.ptr[Prog::PTR_TAKE < 0 ? 0 : Prog::PTR_TAKE])[i];
        Prog::row(p, i, src, active, pass);
    }
}

// ======================================================================================================
// Filter with in-kernel stable compaction (RecordBatch::filter, src/daft-recordbatch/src/lib.rs:520-570)
// ======================================================================================================
// pass 0: kept rows per 2048-row tile -> counts[tile];  (host: exclusive scan)  pass 1: predicate again, every
// surviving row's fixed-width columns are written straight to their final position (row order preserved: warp
// ballots give the in-tile rank).  Traffic: predicate columns twice, payload once in + once out -- instead of a
// mask column, an index column and one index-driven gather per column.
// Prog: IV_N, PTR_COUNTS (int64 per tile: counts in pass 0, exclusive offsets in pass 1), PTR_IDX (-1 or slot of an
// int64 row-id output for the columns the host gathers), keep(p,row), copy(p,src,dst)
constexpr int FILTER_THREADS = 256, FILTER_SUB = 8, FILTER_TILE = FILTER_THREADS * FILTER_SUB;

template <class Prog>
__device__ __forceinline__ void filter_body(const typename Prog::P &p, int pass) {
    // One tile = FILTER_SUB rows per thread, all evaluated (loads in flight together) before any synchronisation;
    // the (sub-row, warp) survivor counts are scanned by warp 0, so a tile costs two barriers, not four per 256 rows.
    constexpr int NWF = FILTER_THREADS / 32, E = FILTER_SUB * NWF;
    static_assert(E % 32 == 0, "scan layout");
    __shared__ int s_w[E];
    __shared__ int s_tot;
    const int64_t n = p.ival[Prog::IV_N];
    const int64_t ntiles = (n + FILTER_TILE - 1) / FILTER_TILE;
    long long *counts = (long long *)p.ptr[Prog::PTR_COUNTS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t tbase = tile * FILTER_TILE;
        bool keep[FILTER_SUB];
        unsigned m[FILTER_SUB];
#pragma unroll
        for (int j = 0; j < FILTER_SUB; j++) {
            const int64_t row = tbase + (int64_t)j * FILTER_THREADS + threadIdx.x;
            keep[j] = row < n && Prog::keep(p, row);
        }
#pragma unroll
        for (int j = 0; j < FILTER_SUB; j++) {
            m[j] = __ballot_sync(FULL_MASK, keep[j]);
            if (lane == 0) s_w[j * NWF + warp] = __popc(m[j]);
        }
        __syncthreads();
        if (warp == 0) {  // exclusive scan of the E counts, E / 32 consecutive entries per lane
            int v[E / 32], sum = 0;
#pragma unroll
            for (int k = 0; k < E / 32; k++) {
                v[k] = s_w[lane * (E / 32) + k];
                sum += v[k];
            }
            int x = sum;
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(FULL_MASK, x, d);
                if (lane >= d) x += y;
            }
            int run = x - sum;
#pragma unroll
            for (int k = 0; k < E / 32; k++) {
                s_w[lane * (E / 32) + k] = run;
                run += v[k];
            }
            if (lane == 31) s_tot = x;
        }
        __syncthreads();
        if (!pass) {
            if (threadIdx.x == 0) counts[tile] = s_tot;
        } else {
            const int64_t obase = counts[tile];
            int64_t src[FILTER_SUB], dst[FILTER_SUB];  // dst < 0: row dropped
#pragma unroll
            for (int j = 0; j < FILTER_SUB; j++) {
                src[j] = tbase + (int64_t)j * FILTER_THREADS + threadIdx.x;
                dst[j] = keep[j] ? obase + s_w[j * NWF + warp] + __popc(m[j] & ((1u << lane) - 1u)) : -1;
            }
            Prog::copy(p, src, dst);
            if (Prog::PTR_IDX >= 0) {
#pragma unroll
                for (int j = 0; j < FILTER_SUB; j++)
                    if (dst[j] >= 0) ((int64_t *)p.ptr[Prog::PTR_IDX < 0 ? 0 : Prog::PTR_IDX])[dst[j]] = src[j];
            }
        }
        __syncthreads();
    }
}

// ======================================================================================================
// fused Filter -> Project -> HashAggregate
// ======================================================================================================
// Prog provides (all generated):
//   using P;  IV_N, IV_MASK (table capacity - 1), PTR_TAG, PTR_MINROW (first row per slot = representative row),
//   PTR_FLAG (int: 0 ok, 1 table overflow, 2 low-cardinality kernel gave up);
//   IV_RSTRIDE (words between table records), ASTRIDE (record size);  NKEYS, G1, NLANE, NWARP, U;
//   MODE (0: row-dependent exact 62-bit codes, 2: never exact);  struct Raw / Keys / Vals;
//   load / load2 / eval / eval_keys / keys_equal / code / lane_init / lane_update / global_update / warp_cells / flush;
//   t2_init / t2_update / t2_flush (global build);  STAGE_ROWS / STAGES / STAGE_BYTES / RPL / LATE, stage_issue / load_staged
//   (TMA build)
//
// Group identity is carried by a 64-bit key CODE (see jit.cpp "Exact key code"): when it is exact (injective in
// the key values) equal codes ARE equal keys; otherwise the code is a fingerprint and a match is confirmed against
// the key values of the slot's representative row (nulls equal, floats by bit pattern -- the reference's
// "hash equal AND comparator" rule, src/daft-recordbatch/src/ops/inline_agg.rs:760-784).
#ifndef B200_AGG_WARPS
#define B200_AGG_WARPS 8  // warps per CTA of the aggregate kernels (the staged build may ask for 16)
#endif
constexpr int AGG_WARPS = B200_AGG_WARPS;
constexpr unsigned long long AGG_EMPTY = 0xFFFFFFFFFFFFFFFFULL;     // min-row cell of an unclaimed slot
constexpr int AGG_PROBE_LIMIT = 1024;

__device__ __forceinline__ uint64_t agg_mix(uint64_t x) {
    x ^= x >> 32;
    x *= 0xD6E8FEB86659FD93ULL;
    x ^= x >> 32;
    return x;
}

// Two layouts of the global table, chosen by the host per launch (ival[IV_RSTRIDE] = words between records):
//  * records (stride = Prog::ASTRIDE): [code, cell 0, cell 1, ...] padded to whole 32-byte sectors -- everything one
//    row touches sits in one or two adjacent sectors: for tables larger than L2 (DRAM-bound random access);
//  * one array per cell (stride = 1): a row's atomics land on different L2 slices and proceed in parallel: for tables
//    that live in L2 (atomic-throughput bound).  Measured on B200, 50 M rows: 100 k groups 1.6 ms vs 2.3 ms with
//    records; 10 M groups 9.9 ms vs 6.9 ms with records.
// All "slot" values handed around are ELEMENT indices (record number * stride): a cell is `cell_base[slot]`.
// Find or claim the record of the group with code `tag`.  Returns its element index, or -1 after flagging overflow.
template <class Prog>
__device__ __forceinline__ long long agg_find_slot(const typename Prog::P &p, unsigned long long tag, int64_t row) {
    unsigned long long *slot_tag = (unsigned long long *)p.ptr[Prog::PTR_TAG];
    unsigned long long *slot_min = (unsigned long long *)p.ptr[Prog::PTR_MINROW];
    const uint64_t mask = (uint64_t)p.ival[Prog::IV_MASK];
    const bool exact = Prog::MODE == 0 && !(tag >> 63);
    const uint64_t rstride = (uint64_t)p.ival[Prog::IV_RSTRIDE];
    uint64_t rec = agg_mix(tag) & mask;
    for (int probes = 0; probes < AGG_PROBE_LIMIT; probes++) {
        const uint64_t s = rec * rstride;
        // tags are never 0 (bit 62 or bit 63 is set): one CAS on the tag word claims the record
        unsigned long long cur = slot_tag[s];  // a stale 0 from L1 is corrected by the CAS below
        if (cur == 0ULL) {
            const unsigned long long prev = atomicCAS(&slot_tag[s], 0ULL, tag);
            if (prev == 0ULL) return (long long)s;  // claimed; the caller's update publishes the first row
            cur = prev;
        }
        if (cur == tag) {
            if (exact) return (long long)s;
            unsigned long long r = *((volatile unsigned long long *)&slot_min[s]);
            while (r == AGG_EMPTY) r = *((volatile unsigned long long *)&slot_min[s]);  // claimer is about to publish
            if ((int64_t)r == row) return (long long)s;
            typename Prog::Keys mine, other;
            Prog::eval_keys(p, row, mine);
            Prog::eval_keys(p, (int64_t)r, other);
            if (Prog::keys_equal(mine, other)) return (long long)s;
        }
        rec = (rec + 1) & mask;
    }
    *((volatile int *)p.ptr[Prog::PTR_FLAG]) = 1;  // table too full: the host retries with a larger one
    return -1;
}

// ---- global kernel: mid / high cardinality -------------------------------------------------------------
// Tier 2: a per-CTA shared-memory table of (exact code -> cells) absorbs the rows of up to T2 groups with
// shared-memory atomics and is merged into the global table once per CTA, so a few thousand hot groups do not
// hammer a handful of L2 addresses.  Rows whose group does not fit (or whose code is inexact) go straight to the
// global table: one CAS to find/claim the slot, then red.global atomics on its cells.
template <class Prog>
__device__ __forceinline__ void agg_body_global(const typename Prog::P &p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int T2 = Prog::T2;
    unsigned long long *t2_tag = (unsigned long long *)smem_raw;  // 0 = empty (codes are never 0)
    unsigned long long *t2c = t2_tag + T2;                        // cell c of slot t: t2c[c * T2 + t]
    for (int t = threadIdx.x; t < T2; t += blockDim.x) {
        t2_tag[t] = 0ULL;
        Prog::t2_init(t2c, t);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = p.ival[Prog::IV_N];
    constexpr int U = Prog::U;
    const int64_t chunk = 32 * U;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    const int64_t total_warps = (int64_t)gridDim.x * AGG_WARPS;
    volatile int *flag = (volatile int *)p.ptr[Prog::PTR_FLAG];
    int t2_hits = 0, t2_miss = 0;
    bool use_t2 = T2 >= 64 && Prog::NKEYS > 0;
    int abandoned = 0;  // status flag as of the previous chunk: its load overlaps this chunk's column loads
    for (int64_t c = (int64_t)blockIdx.x * AGG_WARPS + warp; c < nchunks; c += total_warps) {
        if (__any_sync(FULL_MASK, abandoned != 0)) break;  // someone hit a full table: stop early, the host will retry
        const int64_t base = c * chunk;
        typename Prog::Raw raw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t row = base + u * 32 + lane;
            if (row < n) Prog::load(p, row, raw[u]);
        }
        abandoned = *flag;
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t row = base + u * 32 + lane;
            if (row < n) Prog::load2(p, row, raw[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t row = base + u * 32 + lane;
            typename Prog::Vals vals;
            bool live = row < n;
            unsigned long long tag = 1ULL;
            {
                typename Prog::Keys keys;
                if (live) live = Prog::eval(p, raw[u], keys, vals);
                if (live && Prog::NKEYS > 0) tag = Prog::code(keys);
            }
            if (!live) continue;
            bool done = false;
            if (use_t2 && !(tag >> 63)) {  // exact codes only: equal code == equal key, no verification needed
                unsigned int t = (unsigned int)(agg_mix(tag) >> 40) & (T2 - 1);
#pragma unroll 1
                for (int probes = 0; probes < 8 && !done; probes++) {
                    unsigned long long cur = ((volatile unsigned long long *)t2_tag)[t];
                    if (cur == 0ULL) {
                        const unsigned long long prev = atomicCAS(&t2_tag[t], 0ULL, tag);
                        cur = prev == 0ULL ? tag : prev;
                    }
                    if (cur == tag) {
                        Prog::t2_update(t2c, (int)t, vals, row);
                        done = true;
                    }
                    t = (t + 1) & (T2 - 1);
                }
                if (done) t2_hits++;
                else if (++t2_miss > 256 && t2_miss > 4 * t2_hits) use_t2 = false;  // high cardinality: stop probing tier 2
            }
            if (!done) {
                const long long slot = Prog::NKEYS > 0 ? agg_find_slot<Prog>(p, tag, row) : 0;
                if (slot >= 0) Prog::global_update(p, slot, vals, row);
            }
        }
    }
    __syncthreads();
    // merge tier 2 into the global table (the slot's first row doubles as its representative row)
    if (!*flag) {
        for (int t = threadIdx.x; t < T2; t += blockDim.x) {
            const unsigned long long tag = t2_tag[t];
            if (tag == 0ULL) continue;
            const long long slot = agg_find_slot<Prog>(p, tag, (int64_t)t2c[(Prog::NLANE - 1) * T2 + t]);
            if (slot >= 0) Prog::t2_flush(p, slot, t2c, t);
        }
    }
}

// ---- NaN-fold fix-up pass (variant 3) --------------------------------------------------------------------
// Float min / max follow the reference's left fold, which forgets everything before a NaN (jit.h AggOut::nan_cell).
// The aggregate kernels fold the non-NaN candidates of ALL rows and record each group's last NaN row; when some group
// has a NaN in the middle of its values the host clears those candidate cells and runs this pass over the same input
// and the same (finished) table: every row finds its group's record again and re-enters its candidates only if it
// lies after the group's last NaN.
template <class Prog>
__device__ __forceinline__ void agg_body_nanfix(const typename Prog::P &p) {
    const int64_t n = p.ival[Prog::IV_N];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
        typename Prog::Raw raw;
        Prog::load(p, row, raw);
        Prog::load2(p, row, raw);
        typename Prog::Keys keys;
        typename Prog::Vals vals;
        if (!Prog::eval(p, raw, keys, vals)) continue;
        long long slot = 0;
        if (Prog::NKEYS > 0) slot = agg_find_slot<Prog>(p, Prog::code(keys), row);
        if (slot >= 0) Prog::nanfix_update(p, slot, vals, row);
    }
}

// ---- row -> group pass (variant 4) ------------------------------------------------------------------------
// The list / set / concat aggregates need every row's group.  The table is finished when this runs: each row finds its
// record again (no claims happen) and its element index is written out; rows the predicate drops get -1.
template <class Prog>
__device__ __forceinline__ void agg_body_rowslot(const typename Prog::P &p) {
    const int64_t n = p.ival[Prog::IV_N];
    long long *out = (long long *)p.ptr[Prog::PTR_ROWSLOT];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += stride) {
        typename Prog::Raw raw;
        Prog::load(p, row, raw);
        Prog::load2(p, row, raw);
        typename Prog::Keys keys;
        typename Prog::Vals vals;
        long long slot = -1;
        if (Prog::eval(p, raw, keys, vals)) slot = Prog::NKEYS > 0 ? agg_find_slot<Prog>(p, Prog::code(keys), row) : 0;
        out[row] = slot;
    }
}

// Rare path of the low-cardinality kernel, kept out of line so that its registers do not burden the streaming loop:
// find/claim the group's global slot, try to adopt it into the CTA's tier 1, otherwise update it globally.
// Returns the tier-1 index (>= 0), -2 when the row was accumulated globally, -1 when the table is full.
template <class Prog>
__device__ __noinline__ long long agg_slow_path(const typename Prog::P &p, const typename Prog::Vals &vals, int64_t row,
                                                unsigned long long tag, long long *s_slot, unsigned long long *s_fp,
                                                typename Prog::Keys *s_keys) {
    const long long slot = agg_find_slot<Prog>(p, tag, row);
    if (slot < 0) return -1;
    for (int g = 0; g < Prog::G1; g++) {
        const long long old = (long long)atomicCAS((unsigned long long *)&s_slot[g], (unsigned long long)-1LL, (unsigned long long)slot);
        if (old == -1) {  // this lane owns entry g: publish the key values (for inexact codes), then the code
            typename Prog::Keys keys;
            Prog::eval_keys(p, row, keys);
            s_keys[g] = keys;
            __threadfence_block();
            ((volatile unsigned long long *)s_fp)[g] = tag;
            // make sure the slot has a representative row even if every later row of this group stays in tier 1
            atomicMin(((unsigned long long *)p.ptr[Prog::PTR_MINROW]) + slot, (unsigned long long)row);
            return g;
        }
        if (old == slot) return g;
    }
    Prog::global_update(p, slot, vals, row);
    return -2;
}

// ---- low-cardinality kernels: tier 1 in shared memory, give up (flag 2) when most rows miss it ------------------
template <class Prog>
struct AggTier1 {
    static constexpr int G1 = Prog::G1, NL = Prog::NLANE, NW = Prog::NWARP, NWS = NW > 0 ? NW : 1;
    unsigned long long *s_lcell, *s_wcell, *s_fp;
    long long *s_slot;
    typename Prog::Keys *s_keys;
    unsigned long long *my_lcell, *my_wcell;
    unsigned long long tfp[G1];  // register copy of the tier-1 codes (entries only go 0 -> final value: a stale copy
                                 // can only miss an entry -> slow path -> refresh)
    int slow_rows;
    bool gave_up;  // this lane saw (or raised) the status flag on its slow path: the kernel's result will be discarded
    volatile int *flag;

    static __host__ __device__ constexpr size_t smem_bytes() {
        return 8 * ((size_t)AGG_WARPS * G1 * NL * 32 + (size_t)AGG_WARPS * G1 * NWS + 2 * (size_t)G1) + (size_t)G1 * sizeof(typename Prog::Keys);
    }
    // layout: lane cells [W][G1][NL][32] u64 | warp cells [W][G1][NW] u64 | s_fp[G1] | s_slot[G1] | s_keys[G1]
    __device__ __forceinline__ void init(unsigned char *smem, const typename Prog::P &p) {
        s_lcell = (unsigned long long *)smem;
        s_wcell = s_lcell + (size_t)AGG_WARPS * G1 * NL * 32;
        s_fp = s_wcell + AGG_WARPS * G1 * NWS;
        s_slot = (long long *)(s_fp + G1);
        s_keys = (typename Prog::Keys *)(s_slot + G1);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (threadIdx.x < G1) {
            s_slot[threadIdx.x] = Prog::NKEYS == 0 ? 0 : -1;
            s_fp[threadIdx.x] = Prog::NKEYS == 0 ? 1ULL : 0ULL;
        }
        for (int i = threadIdx.x; i < AGG_WARPS * G1 * NWS; i += blockDim.x) s_wcell[i] = 0ULL;
        my_lcell = s_lcell + (size_t)warp * G1 * NL * 32 + lane;  // cell c of group g: [(g*NL+c)*32]
        for (int g = 0; g < G1; g++) Prog::lane_init(my_lcell + (size_t)g * NL * 32);
        my_wcell = s_wcell + (size_t)warp * G1 * NWS;
        slow_rows = 0;
        gave_up = false;
        flag = (volatile int *)p.ptr[Prog::PTR_FLAG];
    }
    __device__ __forceinline__ void refresh() {
#pragma unroll
        for (int g = 0; g < G1; g++) tfp[g] = ((volatile unsigned long long *)s_fp)[g];
    }
    // one row: filter + projections + key code -> tier-1 cell update (or the rare slow path)
    __device__ __forceinline__ void row(const typename Prog::P &p, const typename Prog::Raw &raw, int64_t r, bool live, int lane) {
        typename Prog::Vals vals;
        unsigned long long fp = 0ULL;
        {
            typename Prog::Keys keys;  // dead after the code is formed: the rare paths re-read the keys by row
            if (live) live = Prog::eval(p, raw, keys, vals);
            if (live) fp = Prog::NKEYS > 0 ? Prog::code(keys) : 1ULL;
        }
        int lid = -1;
#pragma unroll
        for (int g = 0; g < G1; g++) lid = (tfp[g] == fp) ? g : lid;
        if (Prog::NKEYS > 0 && live && lid >= 0 && (Prog::MODE == 2 || (fp >> 63))) {
            typename Prog::Keys kv;  // inexact code: confirm against the published key values
            Prog::eval_keys(p, r, kv);
            if (!Prog::keys_equal(kv, s_keys[lid])) lid = -1;
        }
        if (live && lid < 0) {  // rare here: first rows of a CTA, or a group beyond tier 1
            if (*flag) {  // already abandoned by someone: do not spend a table probe on a result that is discarded
                gave_up = true;
                live = false;
            } else {
                const typename Prog::Vals vcopy = vals;
                const long long q = agg_slow_path<Prog>(p, vcopy, r, fp, s_slot, s_fp, s_keys);
                if (q >= 0) lid = (int)q;
                else live = false;  // accumulated globally (or table overflow)
                refresh();
                // most rows missing tier 1 means the input is not low-cardinality: hand over to the global kernel
                if (++slow_rows > 64 + G1) {
                    *flag = 2;
                    gave_up = true;
                }
            }
        }
        if (live) Prog::lane_update(my_lcell + (size_t)lid * NL * 32, vals, r);
        if (NW > 0) {  // valid / null counts of nullable inputs: one ballot per tier-1 group
#pragma unroll
            for (int g = 0; g < G1; g++) {
                const bool in_g = live && lid == g;
                if (__any_sync(FULL_MASK, in_g)) Prog::warp_cells(vals, in_g, my_wcell + g * NW, lane);
            }
        }
    }
    // flush: warp w reduces groups w, w+W, ... across all warps' private copies (call after __syncthreads)
    __device__ __forceinline__ void flush(const typename Prog::P &p, int lane, int warp) {
        for (int g = warp; g < G1; g += AGG_WARPS) {
            const long long slot = s_slot[g];
            if (slot < 0) continue;
            Prog::flush(p, slot, s_lcell + (size_t)g * NL * 32, (size_t)G1 * NL * 32, s_wcell + g * NWS, (size_t)G1 * NWS, lane);
        }
    }
};

// the line(s) of a chunk's footprint a lane prefetches: address = b + first_row_of_chunk * w
struct PfLane {
    const char *b0, *b1;
    int w0, w1;
};

// one chunk of 32*U rows of a warp; FULL: every row of the chunk exists (no per-row bounds checks or guards)
template <class Prog, bool FULL>
__device__ __forceinline__ void agg_chunk(AggTier1<Prog> &t1, const typename Prog::P &p, const int64_t base, const int64_t n, const int lane,
                                          int &abandoned, const int64_t nextstep, const PfLane &pf) {
    constexpr int U = Prog::U;
    typename Prog::Raw raw[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t row = base + u * 32 + lane;
        if (FULL || row < n) Prog::load(p, row, raw[u]);
    }
    abandoned = *t1.flag;  // consumed at the next chunk: the load overlaps this chunk's column loads
    if (FULL && Prog::PREFETCH) {
        // pull the cache lines of the chunk this warp handles NEXT into L2 (no scoreboard entry): when the warp gets there
        // its loads pay L2 latency instead of HBM latency.  Every lane owns one (or two) fixed lines of the chunk's
        // footprint -- base pointer and bytes-per-row are set up once per kernel -- so a chunk costs two instructions.
        const int64_t next = base + nextstep * Prog::PFDIST;
        if (next + 32 * U <= n) {
            if (pf.b0) prefetch_l2(pf.b0 + next * pf.w0);
            if (Prog::PF_LINES > 32 && pf.b1) prefetch_l2(pf.b1 + next * pf.w1);
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {  // dependent loads (string prefixes) of all U rows, issued back to back
        const int64_t row = base + u * 32 + lane;
        if (FULL || row < n) Prog::load2(p, row, raw[u]);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int64_t row = base + u * 32 + lane;
        t1.row(p, raw[u], row, FULL || row < n, lane);
    }
}

template <class Prog>
__device__ __forceinline__ void agg_body(const typename Prog::P &p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AggTier1<Prog> t1;
    t1.init(smem_raw, p);
    __syncthreads();
    t1.refresh();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = p.ival[Prog::IV_N];
    const int64_t chunk = 32 * Prog::U;
    const int64_t nchunks = (n + chunk - 1) / chunk, nfull = n / chunk;
    const int64_t total_warps = (int64_t)gridDim.x * AGG_WARPS;
    int abandoned = 0;  // status flag as of the previous chunk
    PfLane pf;
    Prog::prefetch_setup(p, lane, pf);
    int64_t c = (int64_t)blockIdx.x * AGG_WARPS + warp;
    for (; c < nfull; c += total_warps) {
        if (__any_sync(FULL_MASK, abandoned != 0 || t1.gave_up)) break;
        agg_chunk<Prog, true>(t1, p, c * chunk, n, lane, abandoned, total_warps * chunk, pf);
    }
    if (c < nchunks && c >= nfull && !__any_sync(FULL_MASK, abandoned != 0))  // the one partial chunk at the end of the input
        agg_chunk<Prog, false>(t1, p, c * chunk, n, lane, abandoned, 0, pf);
    __syncthreads();
    t1.flush(p, lane, warp);
}

// ---- the same kernel fed by TMA bulk copies ---------------------------------------------------------------------
// The CTA owns a ring of S stages in shared memory.  A stage holds R = blockDim.x * RPL consecutive rows of every
// column the plan reads: R values of a fixed-width column, R+2 offsets of a string column, R*L value bytes of a
// fixed-length string column.  Thread 0 issues one cp.async.bulk (UBLKCP) per column slice, completing on the
// stage's "full" mbarrier; every warp copies its rows from the stage into registers, arrives on the stage's "empty"
// mbarrier and goes on to aggregate them while the copies of the next S-1 stages are in flight -- HBM latency is
// covered by shared-memory capacity instead of registers and resident warps.  Copies are kilobytes each: the TMA
// unit retires roughly one bulk copy per ~46 cycles per SM, so per-warp sized copies (256-512 B) are issue bound.
// Tail chunks that cannot be staged (the offsets slice needs row base+R+1) use the direct-load path.
__device__ __forceinline__ uint32_t smem_u32(const void *q) { return (uint32_t)__cvta_generic_to_shared(q); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// wait for a phase; gives up (false) when the kernel has been abandoned (status flag set by any CTA)
__device__ __forceinline__ bool mbar_wait_or_abort(uint32_t bar, uint32_t parity, volatile int *flag) {
    for (int spins = 1; !mbar_try_wait(bar, parity); spins++)
        if ((spins & 15) == 0 && *flag) return false;  // the flag lives in global memory: poll it rarely
    return true;
}

template <class Prog>
__device__ __forceinline__ void agg_body_staged(const typename Prog::P &p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int R = Prog::STAGE_ROWS, S = Prog::STAGES, SB = Prog::STAGE_BYTES, RPL = Prog::RPL, T = AGG_WARPS * 32;
    static_assert(R == T * RPL, "stage rows");
    // layout: stages [S][SB] (16 B aligned) | full barriers [S] | empty barriers [S] | tier 1
    unsigned char *stages = smem_raw;
    unsigned long long *full = (unsigned long long *)(stages + (size_t)S * SB);
    unsigned long long *empty = full + S;
    AggTier1<Prog> t1;
    t1.init((unsigned char *)(empty + S), p);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) {
            mbar_init(smem_u32(&full[s]), 1);
            mbar_init(smem_u32(&empty[s]), AGG_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    t1.refresh();
    const int64_t n = p.ival[Prog::IV_N];
    const int64_t nchunks = (n + R - 1) / R;
    const int64_t stride = gridDim.x;
    auto stageable = [&](int64_t c) { return c < nchunks && (c + 1) * R < n; };  // offsets slice reads row (c+1)*R + 1
    int64_t issued = 0;  // (thread 0) iterations whose copies have been issued
    if (threadIdx.x == 0) {
#pragma unroll 1
        for (int s = 0; s < S; s++) {
            const int64_t c = blockIdx.x + (int64_t)s * stride;
            if (stageable(c)) {
                Prog::stage_issue(p, smem_u32(stages + (size_t)s * SB), smem_u32(&full[s]), c * R);
                issued = s + 1;
            }
        }
    }
    bool aborted = false;
    int abandoned = 0;  // status flag as loaded 16 iterations ago
    int64_t i = 0;
    for (int64_t c = blockIdx.x; c < nchunks; c += stride, i++) {
        if (__any_sync(FULL_MASK, t1.gave_up)) {  // (a register: set on the slow path only)
            aborted = true;
            break;
        }
        if ((i & 15) == 0) {  // iterations are shorter than a global-memory round trip: look at the flag loaded 16 ago
            if (__any_sync(FULL_MASK, abandoned != 0)) {
                aborted = true;
                break;
            }
            abandoned = *t1.flag;
        }
        if (threadIdx.x == 0 && i > 0) {  // refill the stage every warp finished reading one iteration ago
            const int64_t j = i - 1 + S, cj = blockIdx.x + j * stride;
            if (stageable(cj)) {
                const int sp = (int)((i - 1) % S);
                if (!mbar_wait_or_abort(smem_u32(&empty[sp]), (uint32_t)(((i - 1) / S) & 1), t1.flag)) {
                    aborted = true;
                } else {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    Prog::stage_issue(p, smem_u32(stages + (size_t)sp * SB), smem_u32(&full[sp]), cj * R);
                    issued = j + 1;
                }
            }
        }
        if (warp == 0) aborted = __any_sync(FULL_MASK, aborted);  // keep warp 0 convergent with its producer lane
        if (aborted) break;
        const int s = (int)(i % S);
        const int64_t base = c * R;
        if (stageable(c)) {
            const bool landed = mbar_wait_or_abort(smem_u32(&full[s]), (uint32_t)((i / S) & 1), t1.flag);
            if (!__all_sync(FULL_MASK, landed)) {
                aborted = true;
                break;
            }
            const unsigned char *st = stages + (size_t)s * SB;
            if (Prog::LATE) {
                // one row at a time straight from the stage (lowest register pressure); the stage is released after
                // the chunk has been aggregated
#pragma unroll
                for (int u = 0; u < RPL; u++) {
                    typename Prog::Raw r1;
                    const int rr = u * T + threadIdx.x;
                    Prog::load_staged(p, st, base + rr, rr, r1);
                    t1.row(p, r1, base + rr, true, lane);  // staged chunks are full
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
            } else {
                // copy the warp's rows to registers, release the stage at once, then aggregate
                typename Prog::Raw raw[RPL];
#pragma unroll
                for (int u = 0; u < RPL; u++) {
                    const int rr = u * T + threadIdx.x;
                    Prog::load_staged(p, st, base + rr, rr, raw[u]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&empty[s]));
#pragma unroll
                for (int u = 0; u < RPL; u++) t1.row(p, raw[u], base + u * T + threadIdx.x, true, lane);
            }
        } else {  // the unstaged tail of the input: one row at a time, direct loads
#pragma unroll 1
            for (int u = 0; u < RPL; u++) {
                typename Prog::Raw r1;
                const int64_t row = base + u * T + threadIdx.x;
                if (row < n) {
                    Prog::load(p, row, r1);
                    Prog::load2(p, row, r1);
                }
                t1.row(p, r1, row, row < n, lane);
            }
        }
    }
    if (threadIdx.x == 0 && aborted) {
        // abandoned: copies already issued must land before the CTA's shared memory is released
        for (int64_t j = i; j < issued; j++)
            while (!mbar_try_wait(smem_u32(&full[(int)(j % S)]), (uint32_t)((j / S) & 1))) {
            }
    }
    __syncthreads();
    t1.flush(p, lane, warp);
}

}  // namespace jit
}  // namespace b200

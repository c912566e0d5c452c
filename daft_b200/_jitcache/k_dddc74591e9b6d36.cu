// XXH3-64 (seeded, default secret) for sm_100a device code.
//
// Replaces xxhash_rust::xxh3::xxh3_64_with_seed as called by the reference at
// src/daft-core/src/kernels/hashing.rs:164-170.  Bit-exact with the canonical xxHash v0.8
// XXH3_64bits_withSeed for every input length class; the hot classes for this path (1-3 B flag
// strings, 4 B dates, 8 B keys, 9-16 B names) are straight-line register code with the secret
// folded into immediates.  This header is compiled both ahead of time (hash.cu, join.cu, ...)
// and textually prepended to every JIT-generated kernel.


#ifdef __CUDACC_RTC__
typedef signed char int8_t;
typedef short int16_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
typedef unsigned long size_t;
#else
#include <cstdint>
#include <cstddef>
#endif

namespace b200 {
namespace xxh {

#define B200_HD __host__ __device__ __forceinline__

constexpr uint8_t kSecret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};

// secret words as compile-time constants (offset must be a constant expression)
constexpr uint64_t sec64c(int off) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | kSecret[off + i];
    return v;
}
constexpr uint32_t sec32c(int off) {
    uint32_t v = 0;
    for (int i = 3; i >= 0; --i) v = (v << 8) | kSecret[off + i];
    return v;
}
template <int OFF>
struct Sec {
    static constexpr uint64_t v64 = sec64c(OFF);
};
constexpr uint32_t kBitflip1to3 = sec32c(0) ^ sec32c(4);

constexpr uint32_t P32_1 = 0x9E3779B1U, P32_2 = 0x85EBCA77U, P32_3 = 0xC2B2AE3DU;
constexpr uint64_t P64_1 = 0x9E3779B185EBCA87ULL, P64_2 = 0xC2B2AE3D27D4EB4FULL, P64_3 = 0x165667B19E3779F9ULL,
                   P64_4 = 0x85EBCA77C2B2AE63ULL, P64_5 = 0x27D4EB2F165667C5ULL, PMX1 = 0x165667919E3779F9ULL,
                   PMX2 = 0x9FB21C651E98DF25ULL;

B200_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
B200_HD uint32_t bswap32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __byte_perm(x, 0, 0x0123);
#else
    return __builtin_bswap32(x);
#endif
}
B200_HD uint64_t bswap64(uint64_t x) {
    return ((uint64_t)bswap32((uint32_t)x) << 32) | bswap32((uint32_t)(x >> 32));
}
B200_HD uint64_t mul128_fold64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return (a * b) ^ __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128)a * b;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
#endif
}
B200_HD uint64_t xxh64_avalanche(uint64_t h) {
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32; return h;
}
B200_HD uint64_t xxh3_avalanche(uint64_t h) {
    h ^= h >> 37; h *= PMX1; h ^= h >> 32; return h;
}
B200_HD uint64_t rrmxmx(uint64_t h, uint64_t len) {
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= PMX2;
    h ^= (h >> 35) + len;
    h *= PMX2;
    return h ^ (h >> 28);
}

// ---- fixed-width fast paths (value already in a register) ------------------------------------
B200_HD uint64_t hash_empty(uint64_t seed) {
    return xxh64_avalanche(seed ^ (Sec<56>::v64 ^ Sec<64>::v64));
}
// len in {1,2,3}: c1 = first byte, c2 = byte[len>>1], c3 = last byte
B200_HD uint64_t hash_1to3(uint32_t c1, uint32_t c2, uint32_t c3, uint32_t len, uint64_t seed) {
    uint32_t comb = (c1 << 16) | (c2 << 24) | c3 | (len << 8);
    uint64_t bitflip = (uint64_t)kBitflip1to3 + seed;
    return xxh64_avalanche((uint64_t)comb ^ bitflip);
}
B200_HD uint64_t hash_u8(uint8_t v, uint64_t seed) { return hash_1to3(v, v, v, 1, seed); }
B200_HD uint64_t hash_u16(uint16_t v, uint64_t seed) { return hash_1to3(v & 0xff, v >> 8, v >> 8, 2, seed); }
// len in [4,8]: lo32 = first 4 bytes, hi32 = last 4 bytes
B200_HD uint64_t hash_4to8(uint32_t first4, uint32_t last4, uint32_t len, uint64_t seed) {
    seed ^= (uint64_t)bswap32((uint32_t)seed) << 32;
    uint64_t bitflip = (Sec<8>::v64 ^ Sec<16>::v64) - seed;
    uint64_t in64 = (uint64_t)last4 + ((uint64_t)first4 << 32);
    return rrmxmx(in64 ^ bitflip, len);
}
B200_HD uint64_t hash_u32(uint32_t v, uint64_t seed) { return hash_4to8(v, v, 4, seed); }
B200_HD uint64_t hash_u64(uint64_t v, uint64_t seed) { return hash_4to8((uint32_t)v, (uint32_t)(v >> 32), 8, seed); }
// len in [9,16]: lo = first 8 bytes, hi = last 8 bytes
B200_HD uint64_t hash_9to16(uint64_t first8, uint64_t last8, uint64_t len, uint64_t seed) {
    uint64_t bf1 = (Sec<24>::v64 ^ Sec<32>::v64) + seed;
    uint64_t bf2 = (Sec<40>::v64 ^ Sec<48>::v64) - seed;
    uint64_t lo = first8 ^ bf1, hi = last8 ^ bf2;
    uint64_t acc = len + bswap64(lo) + hi + mul128_fold64(lo, hi);
    return xxh3_avalanche(acc);
}

// ---- byte-stream path --------------------------------------------------------------------------
static __device__ const uint8_t d_secret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c,
    0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb, 0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f,
    0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c,
    0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb, 0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3,
    0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d,
    0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31, 0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64,
    0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e,
    0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc, 0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce,
    0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};

// Unaligned little-endian loads from global memory, byte granular (never reads past p+width).
B200_HD uint32_t ld32u(const uint8_t *p) {
    if ((((uintptr_t)p) & 3) == 0) return *reinterpret_cast<const uint32_t *>(p);
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
B200_HD uint64_t ld64u(const uint8_t *p) {
    if ((((uintptr_t)p) & 7) == 0) return *reinterpret_cast<const uint64_t *>(p);
    if ((((uintptr_t)p) & 3) == 0) {
        const uint32_t *q = reinterpret_cast<const uint32_t *>(p);
        return (uint64_t)q[0] | ((uint64_t)q[1] << 32);
    }
    uint64_t v = 0;
#pragma unroll
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}
B200_HD uint64_t sec64(int off) {
#ifdef __CUDA_ARCH__
    return ld64u(d_secret + off);
#else
    return ld64u(kSecret + off);
#endif
}

B200_HD uint64_t mix16(const uint8_t *in, int soff, uint64_t seed) {
    return mul128_fold64(ld64u(in) ^ (sec64(soff) + seed), ld64u(in + 8) ^ (sec64(soff + 8) - seed));
}

__host__ __device__ __noinline__ uint64_t hash_long(const uint8_t *in, int64_t len, uint64_t seed) {
    // secret word i (8-byte granular at byte offset `off`) with the seed folded in:
    // custom[16j..16j+8) = k + seed, custom[16j+8..16j+16) = k - seed; unaligned offsets straddle words.
    auto csec = [&](int off) -> uint64_t {
        if (seed == 0) return sec64(off);
        uint64_t v = 0;
        // rebuild byte by byte from the two (or one) derived words it touches
        int w0 = off >> 3, sh = (off & 7) * 8;
        uint64_t a = sec64(w0 * 8);
        a = (w0 & 1) ? a - seed : a + seed;
        if (sh == 0) return a;
        uint64_t b = sec64(w0 * 8 + 8);
        b = ((w0 + 1) & 1) ? b - seed : b + seed;
        v = (a >> sh) | (b << (64 - sh));
        return v;
    };
    uint64_t acc[8] = {P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1};
    auto accumulate = [&](const uint8_t *p, int soff) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t dv = ld64u(p + 8 * i);
            uint64_t dk = dv ^ csec(soff + 8 * i);
            acc[i ^ 1] += dv;
            acc[i] += (uint64_t)(uint32_t)dk * (dk >> 32);
        }
    };
    const int64_t block_len = 1024;
    const int64_t nb_blocks = (len - 1) / block_len;
    for (int64_t n = 0; n < nb_blocks; n++) {
        for (int s = 0; s < 16; s++) accumulate(in + n * block_len + s * 64, s * 8);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint64_t a = acc[i];
            a ^= a >> 47;
            a ^= csec(128 + 8 * i);
            a *= P32_1;
            acc[i] = a;
        }
    }
    const int nb_stripes = (int)(((len - 1) - block_len * nb_blocks) / 64);
    for (int s = 0; s < nb_stripes; s++) accumulate(in + nb_blocks * block_len + s * 64, s * 8);
    accumulate(in + len - 64, 192 - 64 - 7);
    uint64_t r = (uint64_t)len * P64_1;
#pragma unroll
    for (int i = 0; i < 4; i++) r += mul128_fold64(acc[2 * i] ^ csec(11 + 16 * i), acc[2 * i + 1] ^ csec(11 + 16 * i + 8));
    return xxh3_avalanche(r);
}

__host__ __device__ __noinline__ uint64_t hash_17plus(const uint8_t *in, int64_t len, uint64_t seed) {
    if (len <= 128) {
        uint64_t acc = (uint64_t)len * P64_1;
        if (len > 32) {
            if (len > 64) {
                if (len > 96) {
                    acc += mix16(in + 48, 96, seed);
                    acc += mix16(in + len - 64, 112, seed);
                }
                acc += mix16(in + 32, 64, seed);
                acc += mix16(in + len - 48, 80, seed);
            }
            acc += mix16(in + 16, 32, seed);
            acc += mix16(in + len - 32, 48, seed);
        }
        acc += mix16(in, 0, seed);
        acc += mix16(in + len - 16, 16, seed);
        return xxh3_avalanche(acc);
    }
    if (len <= 240) {
        uint64_t acc = (uint64_t)len * P64_1;
        int nb = (int)(len / 16);
        for (int i = 0; i < 8; i++) acc += mix16(in + 16 * i, 16 * i, seed);
        acc = xxh3_avalanche(acc);
        for (int i = 8; i < nb; i++) acc += mix16(in + 16 * i, 16 * (i - 8) + 3, seed);
        acc += mix16(in + len - 16, 136 - 17, seed);
        return xxh3_avalanche(acc);
    }
    return hash_long(in, len, seed);
}

// any length, any alignment
B200_HD uint64_t hash_bytes(const uint8_t *in, int64_t len, uint64_t seed) {
    if (len <= 16) {
        if (len == 0) return hash_empty(seed);
        if (len <= 3) return hash_1to3(in[0], in[len >> 1], in[len - 1], (uint32_t)len, seed);
        if (len <= 8) return hash_4to8(ld32u(in), ld32u(in + len - 4), (uint32_t)len, seed);
        return hash_9to16(ld64u(in), ld64u(in + len - 8), (uint64_t)len, seed);
    }
    return hash_17plus(in, len, seed);
}

}  // namespace xxh
}  // namespace b200

// Device-side skeletons of the JIT-specialised pipeline kernels (sm_100a).
//
// The host (jit.cpp) turns a serialised plan (include/daft_b200.h "plan encoding") into a `Prog`
// struct -- straight-line, fully typed code for the filter / key / aggregate-input expressions --
// and instantiates one of the hand-written skeletons below with it:
//
//   project_body         Filter/Project (a10/a11): one pass, outputs written at HBM speed,
//                        bitmaps built with warp ballots, strings in a length pass + copy pass.
//   agg_body             fused Filter -> Project -> HashAggregate (a4/a6/a10/a11): every input
//                        column is read exactly once; nothing is materialised.
//
// Group-by design (agg_body), replacing the hashbrown raw-entry probe + per-accumulator scatter
// passes of src/daft-recordbatch/src/ops/inline_agg.rs:748-824,556-565:
//   * global open-addressing table keyed by the XXH3 row hash (same hash_rows as the reference), slot
//     claimed with one atomicCAS; equal hashes are verified against the claiming row's key values
//     (nulls equal, floats by bit pattern == the reference's "hash equal AND comparator" rule);
//   * accumulators are indexed by table slot, three tiers by where the group lives:
//       tier 1  the first G1 groups a CTA meets: per-lane private shared-memory cells, plain
//               read-modify-write, zero atomics, zero bank conflicts (cell address = lane);
//               counts and first-row tracking are done per warp with ballots;
//       tier 3  every other group: red.global atomics on the slot's cells.
//     Low-cardinality aggregates (TPC-H Q1: 4 groups) never leave tier 1; high-cardinality ones
//     (Q18: one group per order) spread their atomics over millions of addresses.
//   * first-occurrence group order (inline_agg.rs:804-807) is recovered by tracking min(row) per group.


#define FULL_MASK 0xffffffffu

namespace b200 {
namespace jit {

struct Str {
    const uint8_t *p;
    int64_t len;
};

__device__ __forceinline__ bool bit_at(const uint8_t *bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

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

// ---- integer division helpers (floor semantics; divisor 0 handled by the caller as null) ------------
template <typename T>
__device__ __forceinline__ T floor_div_s(T a, T b) {
    T q = a / b;
    if ((a % b != 0) && ((a < 0) != (b < 0))) q--;
    return q;
}
template <typename T>
__device__ __forceinline__ T floor_mod_s(T a, T b) {
    T r = a % b;
    if (r != 0 && ((r < 0) != (b < 0))) r += b;
    return r;
}

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
// fused Filter -> Project -> HashAggregate
// ======================================================================================================
// Prog provides (all generated):
//   using P;  IV_N, IV_MASK (table capacity - 1), PTR_SLOT_HASH, PTR_SLOT_ROW, PTR_MINROW, PTR_NROWS,
//   PTR_OVERFLOW;  NKEYS, G1, NLANE (lane cells), NWARP (warp count cells), U (rows per lane per step)
//   struct Raw;  struct Keys;  struct Vals;
//   load(p,row,Raw&) ; bool eval(p,Raw&,Keys&,Vals&) (false = filtered out) ; eval_keys(p,row,Keys&)
//   uint64 hash(Keys&) ; bool keys_equal(Keys&,Keys&)
//   lane_init(cell*) ; lane_update(cell* base, Vals&) ; global_update(p, slot, Vals&)
//   warp_cells(p, Vals&, bool in_group, unsigned long long* wc)  (ballot-based count cells)
//   flush(...)  (generated reduction of the lane cells)
constexpr int AGG_WARPS = 8;

template <class Prog>
__device__ __forceinline__ long long agg_find_slot(const typename Prog::P &p, const typename Prog::Keys &k, uint64_t h,
                                                  int64_t row) {
    unsigned long long *slot_hash = (unsigned long long *)p.ptr[Prog::PTR_SLOT_HASH];
    long long *slot_row = (long long *)p.ptr[Prog::PTR_SLOT_ROW];
    const uint64_t mask = (uint64_t)p.ival[Prog::IV_MASK];
    if (h == 0) h = 1;  // 0 marks an empty slot; equality is decided by the key comparison anyway
    uint64_t s = h & mask;
    for (uint64_t probes = 0; probes <= mask; probes++) {
        unsigned long long cur = slot_hash[s];  // may be a stale 0 from L1: the CAS below re-reads at L2
        bool mine = false;
        if (cur == 0) {
            unsigned long long prev = atomicCAS(&slot_hash[s], 0ULL, (unsigned long long)h);
            if (prev == 0) {
                // claimed: publish the representative row (written before anyone can need it below)
                *((volatile long long *)&slot_row[s]) = (long long)row;
                __threadfence();
                mine = true;
            }
            cur = prev == 0 ? h : prev;
        }
        if (mine) return (long long)s;
        if (cur == h) {
            long long r = slot_row[s];
            while (r < 0) r = *((volatile long long *)&slot_row[s]);
            if (r == row) return (long long)s;
            typename Prog::Keys other;
            Prog::eval_keys(p, r, other);
            if (Prog::keys_equal(k, other)) return (long long)s;
        }
        s = (s + 1) & mask;
    }
    *((int *)p.ptr[Prog::PTR_OVERFLOW]) = 1;  // table full: host retries with a larger table
    return -1;
}

template <class Prog>
__device__ __forceinline__ void agg_body(const typename Prog::P &p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: int s_small[G1] | pad | warp cells [W][G1][NWARP+1] u64 | lane cells [W][G1][NLANE][32] u64
    constexpr int G1 = Prog::G1, NL = Prog::NLANE, NW = Prog::NWARP;
    long long *s_small = (long long *)smem_raw;                                  // G1 (slot or -1)
    unsigned long long *s_wcell = (unsigned long long *)(s_small + G1);          // W*G1*(NW+1); [NW] = min row
    unsigned long long *s_lcell = s_wcell + AGG_WARPS * G1 * (NW + 1);           // W*G1*NL*32

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < G1) s_small[threadIdx.x] = -1;
    for (int i = threadIdx.x; i < AGG_WARPS * G1 * (NW + 1); i += blockDim.x)
        s_wcell[i] = ((i % (NW + 1)) == NW) ? 0xFFFFFFFFFFFFFFFFULL : 0ULL;
    unsigned long long *my_lcell = s_lcell + (size_t)warp * G1 * NL * 32 + lane;  // cell c of group g: [(g*NL+c)*32]
    if (NL > 0)
        for (int g = 0; g < G1; g++) Prog::lane_init(my_lcell + (size_t)g * NL * 32);
    unsigned long long *my_wcell = s_wcell + (size_t)warp * G1 * (NW + 1);
    __syncthreads();

    const int64_t n = p.ival[Prog::IV_N];
    constexpr int U = Prog::U;
    const int64_t chunk = 32 * U;
    const int64_t nchunks = (n + chunk - 1) / chunk;
    const int64_t total_warps = (int64_t)gridDim.x * AGG_WARPS;
    for (int64_t c = (int64_t)blockIdx.x * AGG_WARPS + warp; c < nchunks; c += total_warps) {
        const int64_t base = c * chunk;
        typename Prog::Raw raw[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t row = base + u * 32 + lane;
            if (row < n) Prog::load(p, row, raw[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int64_t row = base + u * 32 + lane;
            typename Prog::Keys keys;
            typename Prog::Vals vals;
            bool live = row < n;
            if (live) live = Prog::eval(p, raw[u], keys, vals);
            long long slot = -1;
            if (live) {
                if (Prog::NKEYS == 0) slot = 0;
                else slot = agg_find_slot<Prog>(p, keys, Prog::hash(keys), row);
                if (slot < 0) live = false;
            }
            // tier 1 lookup / insert
            int lid = -1;
            if (live) {
#pragma unroll
                for (int g = 0; g < G1; g++)
                    if (((volatile long long *)s_small)[g] == slot) lid = g;
                if (lid < 0) {
                    for (int g = 0; g < G1; g++) {
                        long long old = (long long)atomicCAS((unsigned long long *)&s_small[g], (unsigned long long)-1LL,
                                                             (unsigned long long)slot);
                        if (old == -1 || old == slot) { lid = g; break; }
                    }
                }
            }
            if (live && lid >= 0) {
                if (NL > 0) Prog::lane_update(my_lcell + (size_t)lid * NL * 32, vals);
            } else if (live) {
                Prog::global_update(p, slot, vals, row);
            }
            // warp-level cells: row counts / valid counts by ballot, first row by ffs
#pragma unroll
            for (int g = 0; g < G1; g++) {
                const bool in_g = live && lid == g;
                const unsigned m = __ballot_sync(FULL_MASK, in_g);
                if (m) {
                    Prog::warp_cells(vals, in_g, my_wcell + g * (NW + 1), lane);
                    if (lane == 0) {
                        unsigned long long first = (unsigned long long)(base + u * 32 + (__ffs(m) - 1));
                        if (first < my_wcell[g * (NW + 1) + NW]) my_wcell[g * (NW + 1) + NW] = first;
                    }
                }
            }
        }
    }
    __syncthreads();
    // flush tier 1: warp w reduces groups w, w+W, ... across all warps' private copies
    for (int g = warp; g < G1; g += AGG_WARPS) {
        const long long slot = s_small[g];
        if (slot < 0) continue;
        Prog::flush(p, slot, s_lcell + (size_t)g * NL * 32, (size_t)G1 * NL * 32, s_wcell + g * (NW + 1), G1 * (NW + 1), lane);
    }
}

}  // namespace jit
}  // namespace b200

using namespace b200;
using namespace b200::jit;
struct Prog {
  typedef Params<49, 2> P;
  static constexpr int IV_N = 0, IV_MASK = 1, PTR_SLOT_HASH = 0, PTR_SLOT_ROW = 1, PTR_OVERFLOW = 2;
  static constexpr int NKEYS = 1, G1 = 2, NLANE = 12, NWARP = 13, U = 4;
  struct Raw {
    bool k13;
    int64_t c13;
    bool k2;
    int32_t c2;
    bool k5;
    uint16_t c5;
    bool k9;
    double c9;
    bool k8;
    float c8;
    bool k0;
    int8_t c0;
    bool k3;
    int64_t c3;
    bool k7;
    uint64_t c7;
    bool k11;
    int32_t c11;
    bool k1;
    int16_t c1;
    int pad_;
  };
  struct Keys {
    int64_t v0; bool k0;
    int pad_;
  };
  struct Vals {
    int64_t a0; bool k0;
    uint64_t a1; bool k1;
    double a2; bool k2;
    float a3; bool k3;
    int8_t a4; bool k4;
    int64_t a5; bool k5;
    uint64_t a6; bool k6;
    int32_t a7; bool k7;
    int32_t a8; bool k8;
    int64_t a9; bool k9;
    double a10; bool k10;
    int pad_;
  };
  __device__ __forceinline__ static void load(const P& p, const int64_t row, Raw& r) {
    r.k13 = bit_at((const uint8_t*)p.ptr[29], row);
    r.c13 = ((const int64_t*)p.ptr[30])[row];
    r.k2 = bit_at((const uint8_t*)p.ptr[31], row);
    r.c2 = ((const int32_t*)p.ptr[32])[row];
    r.k5 = bit_at((const uint8_t*)p.ptr[33], row);
    r.c5 = ((const uint16_t*)p.ptr[34])[row];
    r.k9 = bit_at((const uint8_t*)p.ptr[35], row);
    r.c9 = ((const double*)p.ptr[36])[row];
    r.k8 = bit_at((const uint8_t*)p.ptr[37], row);
    r.c8 = ((const float*)p.ptr[38])[row];
    r.k0 = bit_at((const uint8_t*)p.ptr[39], row);
    r.c0 = ((const int8_t*)p.ptr[40])[row];
    r.k3 = bit_at((const uint8_t*)p.ptr[41], row);
    r.c3 = ((const int64_t*)p.ptr[42])[row];
    r.k7 = bit_at((const uint8_t*)p.ptr[43], row);
    r.c7 = ((const uint64_t*)p.ptr[44])[row];
    r.k11 = bit_at((const uint8_t*)p.ptr[45], row);
    r.c11 = ((const int32_t*)p.ptr[46])[row];
    r.k1 = bit_at((const uint8_t*)p.ptr[47], row);
    r.c1 = ((const int16_t*)p.ptr[48])[row];
  }
  __device__ __forceinline__ static bool eval(const P& p, const Raw& r, Keys& keys, Vals& vals) {
    const int64_t v0 = r.c13; const bool k0 = r.k13;
    const int64_t v1 = (int64_t)(3LL); const bool k1 = true;
    const bool k2 = k0 && k1 && (v1 != 0);
    const int64_t v2 = k2 ? floor_mod_s<int64_t>(v0, v1) : (int64_t)0;
    keys.v0 = v2; keys.k0 = k2;
    const int32_t v3 = r.c2; const bool k3 = r.k2;
    vals.a0 = (int64_t)v3; vals.k0 = k3;
    const uint16_t v4 = r.c5; const bool k4 = r.k5;
    vals.a1 = (uint64_t)v4; vals.k1 = k4;
    const double v5 = r.c9; const bool k5 = r.k9;
    vals.a2 = (double)v5; vals.k2 = k5;
    const float v6 = r.c8; const bool k6 = r.k8;
    vals.a3 = (float)v6; vals.k3 = k6;
    const int8_t v7 = r.c0; const bool k7 = r.k0;
    vals.a4 = (int8_t)v7; vals.k4 = k7;
    const int64_t v8 = r.c3; const bool k8 = r.k3;
    vals.a5 = (int64_t)v8; vals.k5 = k8;
    const uint64_t v9 = r.c7; const bool k9 = r.k7;
    vals.a6 = (uint64_t)v9; vals.k6 = k9;
    const int32_t v12 = r.c11; const bool k12 = r.k11;
    vals.a7 = (int32_t)v12; vals.k7 = k12;
    vals.a8 = (int32_t)v3; vals.k8 = k3;
    const int16_t v17 = r.c1; const bool k17 = r.k1;
    vals.a9 = (int64_t)v17; vals.k9 = k17;
    const double v19 = __longlong_as_double(4607182418800017408LL); const bool k19 = true;
    const double v21 = (double)v6; const bool k21 = k6;
    const double v22 = v19 - v21;
    const bool k22 = k19 && k21;
    const double v23 = v5 * v22;
    const bool k23 = k5 && k22;
    vals.a10 = (double)v23; vals.k10 = k23;
    return true;
  }
  __device__ __forceinline__ static void eval_keys(const P& p, const int64_t row, Keys& keys) {
    const bool rowok = true;
    const bool k0 = rowok && bit_at((const uint8_t*)p.ptr[29], row);
    const int64_t v0 = rowok ? ((const int64_t*)p.ptr[30])[row] : (int64_t)0;
    const int64_t v1 = (int64_t)(3LL); const bool k1 = true;
    const bool k2 = k0 && k1 && (v1 != 0);
    const int64_t v2 = k2 ? floor_mod_s<int64_t>(v0, v1) : (int64_t)0;
    keys.v0 = v2; keys.k0 = k2;
  }
  __device__ __forceinline__ static uint64_t hash(const Keys& k) {
    uint64_t h = 0;
    h = (k.k0 ? xxh::hash_u64((uint64_t)k.v0, 0ULL) : 0x2D06800538D394C2ULL);
    return h;
  }
  __device__ __forceinline__ static bool keys_equal(const Keys& a, const Keys& b) {
    return true && (a.k0 == b.k0) && (!a.k0 || (a.v0 == b.v0));
  }
  __device__ __forceinline__ static void lane_init(unsigned long long* cells) {
    cells[0] = 0ULL;
    cells[32] = 0ULL;
    cells[64] = 0ULL;
    cells[96] = 0ULL;
    cells[128] = 0xFFFFFFFFFFFFFFFFULL;
    cells[160] = 0ULL;
    cells[192] = 0xFFFFFFFFFFFFFFFFULL;
    cells[224] = 0xFFFFFFFFFFFFFFFFULL;
    cells[256] = 0ULL;
    cells[288] = 0ULL;
    cells[320] = 0ULL;
    cells[352] = 0ULL;
  }
  __device__ __forceinline__ static void lane_update(unsigned long long* cells, const Vals& v) {
    if (v.k0) cells[0] += (unsigned long long)v.a0;
    if (v.k1) cells[32] += (unsigned long long)v.a1;
    if (v.k2) cells[64] = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)cells[64]) + v.a2);
    if (v.k3) cells[96] = (unsigned long long)(unsigned)__float_as_int(__int_as_float((int)cells[96]) + v.a3);
    if (v.k4) { const unsigned long long o_ = ord_i64((int64_t)v.a4); if (o_ < cells[128]) cells[128] = o_; }
    if (v.k5) { const unsigned long long o_ = ord_i64((int64_t)v.a5); if (o_ > cells[160]) cells[160] = o_; }
    if (v.k6) { const unsigned long long o_ = ord_u64((uint64_t)v.a6); if (o_ < cells[192]) cells[192] = o_; }
    if (v.k2) { const unsigned long long o_ = ord_f64(v.a2); if (o_ < cells[224]) cells[224] = o_; }
    if (v.k2) { const unsigned long long o_ = ord_f64(v.a2); if (o_ > cells[256]) cells[256] = o_; }
    if (v.k7) { const unsigned long long o_ = ord_i64((int64_t)v.a7); if (o_ > cells[288]) cells[288] = o_; }
    if (v.k9) cells[320] += (unsigned long long)v.a9;
    if (v.k10) cells[352] = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)cells[352]) + v.a10);
  }
  __device__ __forceinline__ static void global_update(const P& p, const long long slot, const Vals& v, const int64_t row) {
    if (v.k0) atomicAdd(((unsigned long long*)p.ptr[3]) + slot, (unsigned long long)v.a0);
    if (v.k1) atomicAdd(((unsigned long long*)p.ptr[4]) + slot, (unsigned long long)v.a1);
    if (v.k2) atomicAdd((double*)(((unsigned long long*)p.ptr[5]) + slot), v.a2);
    if (v.k3) atomicAdd((float*)(((unsigned long long*)p.ptr[6]) + slot), v.a3);
    if (v.k4) atomicMin(((unsigned long long*)p.ptr[7]) + slot, (unsigned long long)ord_i64((int64_t)v.a4));
    if (v.k5) atomicMax(((unsigned long long*)p.ptr[8]) + slot, (unsigned long long)ord_i64((int64_t)v.a5));
    if (v.k6) atomicMin(((unsigned long long*)p.ptr[9]) + slot, (unsigned long long)ord_u64((uint64_t)v.a6));
    if (v.k2) atomicMin(((unsigned long long*)p.ptr[10]) + slot, (unsigned long long)ord_f64(v.a2));
    if (v.k2) atomicMax(((unsigned long long*)p.ptr[11]) + slot, (unsigned long long)ord_f64(v.a2));
    if (v.k7) atomicMax(((unsigned long long*)p.ptr[12]) + slot, (unsigned long long)ord_i64((int64_t)v.a7));
    if (v.k9) atomicAdd(((unsigned long long*)p.ptr[13]) + slot, (unsigned long long)v.a9);
    if (v.k10) atomicAdd((double*)(((unsigned long long*)p.ptr[14]) + slot), v.a10);
    if (v.k0) atomicAdd(((unsigned long long*)p.ptr[15]) + slot, 1ULL);
    if (v.k1) atomicAdd(((unsigned long long*)p.ptr[16]) + slot, 1ULL);
    if (v.k2) atomicAdd(((unsigned long long*)p.ptr[17]) + slot, 1ULL);
    if (v.k3) atomicAdd(((unsigned long long*)p.ptr[18]) + slot, 1ULL);
    if (v.k4) atomicAdd(((unsigned long long*)p.ptr[19]) + slot, 1ULL);
    if (v.k5) atomicAdd(((unsigned long long*)p.ptr[20]) + slot, 1ULL);
    if (v.k6) atomicAdd(((unsigned long long*)p.ptr[21]) + slot, 1ULL);
    if (v.k7) atomicAdd(((unsigned long long*)p.ptr[22]) + slot, 1ULL);
    if (v.k8) atomicAdd(((unsigned long long*)p.ptr[23]) + slot, 1ULL);
    if (!v.k8) atomicAdd(((unsigned long long*)p.ptr[24]) + slot, 1ULL);
    if (v.k9) atomicAdd(((unsigned long long*)p.ptr[25]) + slot, 1ULL);
    if (v.k10) atomicAdd(((unsigned long long*)p.ptr[26]) + slot, 1ULL);
    atomicAdd(((unsigned long long*)p.ptr[27]) + slot, 1ULL);
    atomicMin(((unsigned long long*)p.ptr[28]) + slot, (unsigned long long)row);
  }
  __device__ __forceinline__ static void warp_cells(const Vals& v, const bool in_g, unsigned long long* wcell, const int lane) {
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k0); if (lane == 0) wcell[0] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k1); if (lane == 0) wcell[1] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k2); if (lane == 0) wcell[2] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k3); if (lane == 0) wcell[3] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k4); if (lane == 0) wcell[4] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k5); if (lane == 0) wcell[5] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k6); if (lane == 0) wcell[6] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k7); if (lane == 0) wcell[7] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k8); if (lane == 0) wcell[8] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && !v.k8); if (lane == 0) wcell[9] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k9); if (lane == 0) wcell[10] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && v.k10); if (lane == 0) wcell[11] += __popc(m_); }
    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g); if (lane == 0) wcell[12] += __popc(m_); }
  }
  __device__ __forceinline__ static void flush(const P& p, const long long slot, const unsigned long long* lcell, const size_t lstride,
                                               const unsigned long long* wcell, const size_t wstride, const int lane) {
    {
      unsigned long long a_ = 0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += lcell[(size_t)w_ * lstride + 0 + lane];
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0) atomicAdd(((unsigned long long*)p.ptr[3]) + slot, a_);
    }
    {
      unsigned long long a_ = 0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += lcell[(size_t)w_ * lstride + 32 + lane];
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0) atomicAdd(((unsigned long long*)p.ptr[4]) + slot, a_);
    }
    {
      double a_ = 0.0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += __longlong_as_double((long long)lcell[(size_t)w_ * lstride + 64 + lane]);
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0) atomicAdd((double*)(((unsigned long long*)p.ptr[5]) + slot), a_);
    }
    {
      float a_ = 0.f; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += __int_as_float((int)lcell[(size_t)w_ * lstride + 96 + lane]);
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0) atomicAdd((float*)(((unsigned long long*)p.ptr[6]) + slot), a_);
    }
    {
      unsigned long long a_ = 0xFFFFFFFFFFFFFFFFULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = lcell[(size_t)w_ * lstride + 128 + lane]; a_ = t_ < a_ ? t_ : a_; }
      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ < a_ ? t_ : a_; }
      if (lane == 0) atomicMin(((unsigned long long*)p.ptr[7]) + slot, a_);
    }
    {
      unsigned long long a_ = 0ULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = lcell[(size_t)w_ * lstride + 160 + lane]; a_ = t_ > a_ ? t_ : a_; }
      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ > a_ ? t_ : a_; }
      if (lane == 0) atomicMax(((unsigned long long*)p.ptr[8]) + slot, a_);
    }
    {
      unsigned long long a_ = 0xFFFFFFFFFFFFFFFFULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = lcell[(size_t)w_ * lstride + 192 + lane]; a_ = t_ < a_ ? t_ : a_; }
      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ < a_ ? t_ : a_; }
      if (lane == 0) atomicMin(((unsigned long long*)p.ptr[9]) + slot, a_);
    }
    {
      unsigned long long a_ = 0xFFFFFFFFFFFFFFFFULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = lcell[(size_t)w_ * lstride + 224 + lane]; a_ = t_ < a_ ? t_ : a_; }
      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ < a_ ? t_ : a_; }
      if (lane == 0) atomicMin(((unsigned long long*)p.ptr[10]) + slot, a_);
    }
    {
      unsigned long long a_ = 0ULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = lcell[(size_t)w_ * lstride + 256 + lane]; a_ = t_ > a_ ? t_ : a_; }
      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ > a_ ? t_ : a_; }
      if (lane == 0) atomicMax(((unsigned long long*)p.ptr[11]) + slot, a_);
    }
    {
      unsigned long long a_ = 0ULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = lcell[(size_t)w_ * lstride + 288 + lane]; a_ = t_ > a_ ? t_ : a_; }
      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ > a_ ? t_ : a_; }
      if (lane == 0) atomicMax(((unsigned long long*)p.ptr[12]) + slot, a_);
    }
    {
      unsigned long long a_ = 0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += lcell[(size_t)w_ * lstride + 320 + lane];
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0) atomicAdd(((unsigned long long*)p.ptr[13]) + slot, a_);
    }
    {
      double a_ = 0.0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += __longlong_as_double((long long)lcell[(size_t)w_ * lstride + 352 + lane]);
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0) atomicAdd((double*)(((unsigned long long*)p.ptr[14]) + slot), a_);
    }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 0] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[15]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 1] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[16]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 2] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[17]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 3] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[18]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 4] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[19]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 5] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[20]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 6] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[21]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 7] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[22]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 8] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[23]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 9] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[24]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 10] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[25]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 11] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[26]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 12] : 0ULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);
      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[27]) + slot, a_); }
    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + 13] : 0xFFFFFFFFFFFFFFFFULL;
      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ < a_ ? t_ : a_; }
      if (lane == 0) atomicMin(((unsigned long long*)p.ptr[28]) + slot, a_); }
  }
};
extern "C" __global__ void __launch_bounds__(256) b200_k(const Prog::P p) { agg_body<Prog>(p); }

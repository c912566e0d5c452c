// XXH3-64 (seeded, default secret) for sm_100a device code.
//
// Replaces xxhash_rust::xxh3::xxh3_64_with_seed as called by the reference at
// src/daft-core/src/kernels/hashing.rs:164-170.  Bit-exact with the canonical xxHash v0.8
// XXH3_64bits_withSeed for every input length class; the hot classes for this path (1-3 B flag
// strings, 4 B dates, 8 B keys, 9-16 B names) are straight-line register code with the secret
// folded into immediates.  This header is compiled both ahead of time (hash.cu, join.cu, ...)
// and textually prepended to every JIT-generated kernel.
#pragma once

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

inline __host__ __device__ __noinline__ uint64_t hash_long(const uint8_t *in, int64_t len, uint64_t seed) {
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

inline __host__ __device__ __noinline__ uint64_t hash_17plus(const uint8_t *in, int64_t len, uint64_t seed) {
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

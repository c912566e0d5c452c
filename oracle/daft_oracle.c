/*
 * daft_oracle.c -- CPU restatement of the reference's columnar hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under daft_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / CPU baseline.
 *
 * Parity pin: the hashing section is checked bit-for-bit against the Python
 * `xxhash` package (the same check the reference's own tests make,
 * tests/series/test_hash.py:29-124) and against the literal goldens in
 * tests/recordbatch/test_hash.py:15-18,30-33,45 (see tests/test_oracle_hash.py).
 * Group-by / join / sort semantics are pinned against the literal expected
 * outputs of tests/recordbatch/test_{table_aggs,joins,sorting}.py restated in
 * tests/golden/.  The reference engine itself (Rust) cannot be built in this
 * image (no rustc/cargo), so there is no oracle/_ref binary.
 *
 * Third-party algorithm restated here: XXH3-64 from xxhash-rust 0.8.15
 * (Cargo.lock:9290) == canonical xxHash v0.8 XXH3_64bits_withSeed with the
 * default kSecret.  Call sites in the reference:
 *   src/daft-core/src/kernels/hashing.rs:17,164-170.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* dtype codes -- MUST match include/daft_b200.h (b200_dtype)                 */
/* ------------------------------------------------------------------------- */
enum {
    ORC_NULL = 0, ORC_BOOL = 1,
    ORC_I8 = 2, ORC_I16 = 3, ORC_I32 = 4, ORC_I64 = 5,
    ORC_U8 = 6, ORC_U16 = 7, ORC_U32 = 8, ORC_U64 = 9,
    ORC_F32 = 10, ORC_F64 = 11,
    ORC_DATE32 = 12,      /* physical i32 */
    ORC_TIMESTAMP = 13,   /* physical i64, no tz */
    ORC_LARGE_UTF8 = 14, ORC_LARGE_BINARY = 15
};

typedef struct {
    int32_t dtype;
    int32_t _pad;
    int64_t length;
    const uint8_t *validity; /* Arrow LSB bitmap or NULL = all valid */
    const int64_t *offsets;  /* LargeUtf8/LargeBinary: length+1 entries */
    const void *data;        /* values (bool: bitmap; strings: bytes) */
} orc_col;

static inline int bit_get(const uint8_t *bm, int64_t i) {
    return (bm[i >> 3] >> (i & 7)) & 1;
}
static inline int col_valid(const orc_col *c, int64_t i) {
    if (c->dtype == ORC_NULL) return 0;
    return c->validity ? bit_get(c->validity, i) : 1;
}
static int dtype_width(int32_t dt) {
    switch (dt) {
    case ORC_I8: case ORC_U8: return 1;
    case ORC_I16: case ORC_U16: return 2;
    case ORC_I32: case ORC_U32: case ORC_F32: case ORC_DATE32: return 4;
    case ORC_I64: case ORC_U64: case ORC_F64: case ORC_TIMESTAMP: return 8;
    default: return 0;
    }
}

/* ========================================================================= */
/* XXH3-64 (seeded, default secret)                                           */
/* ========================================================================= */
static const uint8_t kSecret[192] = {
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

#define P32_1 0x9E3779B1U
#define P32_2 0x85EBCA77U
#define P32_3 0xC2B2AE3DU
#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_3 0x165667B19E3779F9ULL
#define P64_4 0x85EBCA77C2B2AE63ULL
#define P64_5 0x27D4EB2F165667C5ULL
#define PMX1 0x165667919E3779F9ULL
#define PMX2 0x9FB21C651E98DF25ULL

static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t bswap64(uint64_t x) { return __builtin_bswap64(x); }
static inline uint32_t bswap32(uint32_t x) { return __builtin_bswap32(x); }
static inline uint64_t mul128_fold64(uint64_t a, uint64_t b) {
    __uint128_t p = (__uint128_t)a * b;
    return (uint64_t)p ^ (uint64_t)(p >> 64);
}
static inline uint64_t xxh64_avalanche(uint64_t h) {
    h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32; return h;
}
static inline uint64_t xxh3_avalanche(uint64_t h) {
    h ^= h >> 37; h *= PMX1; h ^= h >> 32; return h;
}
static inline uint64_t rrmxmx(uint64_t h, uint64_t len) {
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= PMX2;
    h ^= (h >> 35) + len;
    h *= PMX2;
    return h ^ (h >> 28);
}
static inline uint64_t mix16(const uint8_t *in, const uint8_t *sec, uint64_t seed) {
    return mul128_fold64(rd64(in) ^ (rd64(sec) + seed), rd64(in + 8) ^ (rd64(sec + 8) - seed));
}
static void accumulate_512(uint64_t *acc, const uint8_t *in, const uint8_t *sec) {
    for (int i = 0; i < 8; i++) {
        uint64_t dv = rd64(in + 8 * i);
        uint64_t dk = dv ^ rd64(sec + 8 * i);
        acc[i ^ 1] += dv;
        acc[i] += (uint64_t)(uint32_t)dk * (dk >> 32);
    }
}
static void scramble_acc(uint64_t *acc, const uint8_t *sec) {
    for (int i = 0; i < 8; i++) {
        uint64_t a = acc[i];
        a ^= a >> 47;
        a ^= rd64(sec + 8 * i);
        a *= P32_1;
        acc[i] = a;
    }
}
static uint64_t hash_long(const uint8_t *in, size_t len, uint64_t seed) {
    uint8_t custom[192];
    const uint8_t *sec = kSecret;
    if (seed != 0) {
        for (int i = 0; i < 12; i++) {
            uint64_t lo = rd64(kSecret + 16 * i) + seed;
            uint64_t hi = rd64(kSecret + 16 * i + 8) - seed;
            memcpy(custom + 16 * i, &lo, 8);
            memcpy(custom + 16 * i + 8, &hi, 8);
        }
        sec = custom;
    }
    uint64_t acc[8] = {P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1};
    const size_t nb_stripes_per_block = (192 - 64) / 8; /* 16 */
    const size_t block_len = 64 * nb_stripes_per_block; /* 1024 */
    const size_t nb_blocks = (len - 1) / block_len;
    for (size_t n = 0; n < nb_blocks; n++) {
        for (size_t s = 0; s < nb_stripes_per_block; s++)
            accumulate_512(acc, in + n * block_len + s * 64, sec + s * 8);
        scramble_acc(acc, sec + 192 - 64);
    }
    const size_t nb_stripes = ((len - 1) - block_len * nb_blocks) / 64;
    for (size_t s = 0; s < nb_stripes; s++)
        accumulate_512(acc, in + nb_blocks * block_len + s * 64, sec + s * 8);
    accumulate_512(acc, in + len - 64, sec + 192 - 64 - 7);
    uint64_t r = (uint64_t)len * P64_1;
    for (int i = 0; i < 4; i++)
        r += mul128_fold64(acc[2 * i] ^ rd64(sec + 11 + 16 * i), acc[2 * i + 1] ^ rd64(sec + 11 + 16 * i + 8));
    return xxh3_avalanche(r);
}

ORC_EXPORT uint64_t orc_xxh3_64(const uint8_t *in, size_t len, uint64_t seed) {
    const uint8_t *s = kSecret;
    if (len == 0) return xxh64_avalanche(seed ^ (rd64(s + 56) ^ rd64(s + 64)));
    if (len <= 3) {
        uint8_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t comb = ((uint32_t)c1 << 16) | ((uint32_t)c2 << 24) | (uint32_t)c3 | ((uint32_t)len << 8);
        uint64_t bitflip = (uint64_t)(rd32(s) ^ rd32(s + 4)) + seed;
        return xxh64_avalanche((uint64_t)comb ^ bitflip);
    }
    if (len <= 8) {
        seed ^= (uint64_t)bswap32((uint32_t)seed) << 32;
        uint32_t i1 = rd32(in), i2 = rd32(in + len - 4);
        uint64_t bitflip = (rd64(s + 8) ^ rd64(s + 16)) - seed;
        uint64_t in64 = (uint64_t)i2 + ((uint64_t)i1 << 32);
        return rrmxmx(in64 ^ bitflip, len);
    }
    if (len <= 16) {
        uint64_t bf1 = (rd64(s + 24) ^ rd64(s + 32)) + seed;
        uint64_t bf2 = (rd64(s + 40) ^ rd64(s + 48)) - seed;
        uint64_t lo = rd64(in) ^ bf1, hi = rd64(in + len - 8) ^ bf2;
        uint64_t acc = (uint64_t)len + bswap64(lo) + hi + mul128_fold64(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (len <= 128) {
        uint64_t acc = (uint64_t)len * P64_1;
        if (len > 32) {
            if (len > 64) {
                if (len > 96) {
                    acc += mix16(in + 48, s + 96, seed);
                    acc += mix16(in + len - 64, s + 112, seed);
                }
                acc += mix16(in + 32, s + 64, seed);
                acc += mix16(in + len - 48, s + 80, seed);
            }
            acc += mix16(in + 16, s + 32, seed);
            acc += mix16(in + len - 32, s + 48, seed);
        }
        acc += mix16(in, s, seed);
        acc += mix16(in + len - 16, s + 16, seed);
        return xxh3_avalanche(acc);
    }
    if (len <= 240) {
        uint64_t acc = (uint64_t)len * P64_1;
        int nb = (int)len / 16;
        for (int i = 0; i < 8; i++) acc += mix16(in + 16 * i, s + 16 * i, seed);
        acc = xxh3_avalanche(acc);
        for (int i = 8; i < nb; i++) acc += mix16(in + 16 * i, s + 16 * (i - 8) + 3, seed);
        acc += mix16(in + len - 16, s + 136 - 17, seed);
        return xxh3_avalanche(acc);
    }
    return hash_long(in, len, seed);
}

/* ========================================================================= */
/* a1: per-column hash   (src/daft-core/src/kernels/hashing.rs:304-445)        */
/* ========================================================================= */
/*
 * Rules restated:
 *  - primitives: hash LE bytes of the native width (hashing.rs:462-501);
 *    null WITHOUT seed -> xxh3(b""), null WITH seed -> ALSO xxh3(b"", 0)
 *    (SEEDED_NULL_USES_DEFAULT, hashing.rs:73-83,231-240,493).
 *  - bool: ASCII "1"/"0" (hashing.rs:512-552), null as primitives.
 *  - LargeUtf8/LargeBinary/Null arrays go through hash_bytes_iter_inner
 *    (hashing.rs:243-265,554-612): a null value is the EMPTY SLICE hashed WITH
 *    the row seed (tests/series/test_hash.py:61-73 pins this).
 */
ORC_EXPORT int orc_hash_column(const orc_col *c, const uint64_t *seed, uint64_t *out) {
    const int64_t n = c->length;
    const uint64_t null_default = orc_xxh3_64((const uint8_t *)"", 0, 0);
    int w = dtype_width(c->dtype);
    if (w) {
        const uint8_t *v = (const uint8_t *)c->data;
        for (int64_t i = 0; i < n; i++) {
            if (!col_valid(c, i)) out[i] = null_default;
            else out[i] = orc_xxh3_64(v + (size_t)i * w, w, seed ? seed[i] : 0);
        }
        return 0;
    }
    switch (c->dtype) {
    case ORC_BOOL: {
        const uint8_t *v = (const uint8_t *)c->data;
        for (int64_t i = 0; i < n; i++) {
            if (!col_valid(c, i)) out[i] = null_default;
            else {
                uint8_t ch = bit_get(v, i) ? '1' : '0';
                out[i] = orc_xxh3_64(&ch, 1, seed ? seed[i] : 0);
            }
        }
        return 0;
    }
    case ORC_NULL:
        for (int64_t i = 0; i < n; i++) out[i] = orc_xxh3_64((const uint8_t *)"", 0, seed ? seed[i] : 0);
        return 0;
    case ORC_LARGE_UTF8:
    case ORC_LARGE_BINARY: {
        const uint8_t *d = (const uint8_t *)c->data;
        for (int64_t i = 0; i < n; i++) {
            uint64_t sd = seed ? seed[i] : 0;
            if (!col_valid(c, i)) out[i] = orc_xxh3_64((const uint8_t *)"", 0, sd);
            else out[i] = orc_xxh3_64(d + c->offsets[i], (size_t)(c->offsets[i + 1] - c->offsets[i]), sd);
        }
        return 0;
    }
    default:
        return -1;
    }
}

/* a2: RecordBatch::hash_rows (src/daft-recordbatch/src/ops/hash.rs:14-26):
 * h = hash(col0, None); for c in cols[1..]: h = hash(c, seed = h). */
ORC_EXPORT int orc_hash_rows(const orc_col *cols, int ncols, uint64_t *out) {
    if (ncols < 1) return -1;
    int rc = orc_hash_column(&cols[0], NULL, out);
    for (int k = 1; k < ncols && rc == 0; k++) {
        int64_t n = cols[k].length;
        uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
        memcpy(tmp, out, sizeof(uint64_t) * (size_t)n);
        rc = orc_hash_column(&cols[k], tmp, out);
        free(tmp);
    }
    return rc;
}

/* ========================================================================= */
/* value equality used by group-by and join comparators                       */
/* (src/daft-core/src/array/ops/arrow/comparison.rs:58-121)                    */
/* ========================================================================= */
static int values_equal(const orc_col *a, int64_t i, const orc_col *b, int64_t j, int nan_equal) {
    switch (a->dtype) {
    case ORC_BOOL:
        return bit_get((const uint8_t *)a->data, i) == bit_get((const uint8_t *)b->data, j);
    case ORC_F32: {
        float x = ((const float *)a->data)[i], y = ((const float *)b->data)[j];
        if (isnan(x) && isnan(y)) return nan_equal;
        return x == y;
    }
    case ORC_F64: {
        double x = ((const double *)a->data)[i], y = ((const double *)b->data)[j];
        if (isnan(x) && isnan(y)) return nan_equal;
        return x == y;
    }
    case ORC_LARGE_UTF8:
    case ORC_LARGE_BINARY: {
        int64_t la = a->offsets[i + 1] - a->offsets[i], lb = b->offsets[j + 1] - b->offsets[j];
        if (la != lb) return 0;
        return memcmp((const uint8_t *)a->data + a->offsets[i], (const uint8_t *)b->data + b->offsets[j], (size_t)la) == 0;
    }
    case ORC_NULL:
        return 1;
    default: {
        int w = dtype_width(a->dtype);
        return memcmp((const uint8_t *)a->data + (size_t)i * w, (const uint8_t *)b->data + (size_t)j * w, (size_t)w) == 0;
    }
    }
}
/* build_is_equal: (valid,valid)->value eq; (null,null)->nulls_equal; else false */
static int rows_equal(const orc_col *a, int64_t i, const orc_col *b, int64_t j, int ncols,
                      const uint8_t *nulls_equal, int nan_equal) {
    for (int k = 0; k < ncols; k++) {
        int va = col_valid(&a[k], i), vb = col_valid(&b[k], j);
        if (va && vb) { if (!values_equal(&a[k], i, &b[k], j, nan_equal)) return 0; }
        else if (!va && !vb) { if (!nulls_equal[k]) return 0; }
        else return 0;
    }
    return 1;
}

/* tiny open-addressing table keyed by the precomputed row hash (stand-in for
 * hashbrown with IdentityBuildHasher; iteration order is never used). */
typedef struct { uint64_t *hash; int64_t *val; uint64_t mask; } otable;
static int otable_init(otable *t, int64_t n) {
    uint64_t cap = 16; while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    t->hash = (uint64_t *)malloc(cap * 8); t->val = (int64_t *)malloc(cap * 8); t->mask = cap - 1;
    if (!t->hash || !t->val) return -1;
    for (uint64_t i = 0; i < cap; i++) t->val[i] = -1;
    return 0;
}
static void otable_free(otable *t) { free(t->hash); free(t->val); }

/* ========================================================================= */
/* a4: phase 1 of agg_generic_hash_path                                       */
/* (src/daft-recordbatch/src/ops/inline_agg.rs:748-824): dense group ids in    */
/* first-occurrence order, group sizes, first-row index per group.             */
/* Equality: row hashes equal AND comparator(nulls_equal=true,nans_equal=true) */
/* ========================================================================= */
ORC_EXPORT int64_t orc_group_ids(const orc_col *keys, int nkeys, uint32_t *group_ids,
                                 uint64_t *first_rows, uint64_t *group_sizes) {
    const int64_t n = keys[0].length;
    uint64_t *h = (uint64_t *)malloc(8 * (size_t)(n ? n : 1));
    if (orc_hash_rows(keys, nkeys, h)) { free(h); return -1; }
    uint8_t ne[64]; memset(ne, 1, sizeof ne);
    otable t; if (otable_init(&t, n)) { free(h); return -1; }
    int64_t ngroups = 0;
    for (int64_t i = 0; i < n; i++) {
        uint64_t s = h[i] & t.mask;
        for (;;) {
            int64_t g = t.val[s];
            if (g < 0) {
                t.val[s] = ngroups; t.hash[s] = h[i];
                first_rows[ngroups] = (uint64_t)i; group_sizes[ngroups] = 1;
                group_ids[i] = (uint32_t)ngroups; ngroups++;
                break;
            }
            if (t.hash[s] == h[i] && rows_equal(keys, i, keys, (int64_t)first_rows[g], nkeys, ne, 1)) {
                group_ids[i] = (uint32_t)g; group_sizes[g]++;
                break;
            }
            s = (s + 1) & t.mask;
        }
    }
    otable_free(&t); free(h);
    return ngroups;
}

/* a4 phase 2: accumulators (inline_agg.rs:36-298).  Sequential row order;
 * Option<T> accumulator -> out_valid[g]=0 when no valid row was seen. */
#define DEF_SUM(NAME, T)                                                                       \
    ORC_EXPORT void NAME(const T *v, const uint8_t *validity, const uint32_t *gid, int64_t n,   \
                         int64_t ng, T *out, uint8_t *out_valid) {                              \
        for (int64_t g = 0; g < ng; g++) { out[g] = 0; out_valid[g] = 0; }                      \
        for (int64_t i = 0; i < n; i++) {                                                       \
            if (validity && !bit_get(validity, i)) continue;                                    \
            uint32_t g = gid[i];                                                                \
            if (out_valid[g]) out[g] = (T)(out[g] + v[i]); else { out[g] = v[i]; out_valid[g] = 1; } \
        }                                                                                       \
    }
DEF_SUM(orc_sum_i64, int64_t)
DEF_SUM(orc_sum_u64, uint64_t)
DEF_SUM(orc_sum_f32, float)
DEF_SUM(orc_sum_f64, double)

/* min/max: ints std::cmp::min/max; floats left fold `if a<b {a} else {b}` /
 * `if a>b {a} else {b}` (inline_agg.rs:271-298) -- NaN-order dependent. */
#define DEF_MINMAX(NAME, T, CMP)                                                                \
    ORC_EXPORT void NAME(const T *v, const uint8_t *validity, const uint32_t *gid, int64_t n,   \
                         int64_t ng, T *out, uint8_t *out_valid) {                              \
        for (int64_t g = 0; g < ng; g++) { out[g] = 0; out_valid[g] = 0; }                      \
        for (int64_t i = 0; i < n; i++) {                                                       \
            if (validity && !bit_get(validity, i)) continue;                                    \
            uint32_t g = gid[i];                                                                \
            T a = out[g], b = v[i];                                                             \
            if (out_valid[g]) out[g] = (a CMP b) ? a : b; else { out[g] = b; out_valid[g] = 1; } \
        }                                                                                       \
    }
DEF_MINMAX(orc_min_i8, int8_t, <)   DEF_MINMAX(orc_max_i8, int8_t, >)
DEF_MINMAX(orc_min_i16, int16_t, <) DEF_MINMAX(orc_max_i16, int16_t, >)
DEF_MINMAX(orc_min_i32, int32_t, <) DEF_MINMAX(orc_max_i32, int32_t, >)
DEF_MINMAX(orc_min_i64, int64_t, <) DEF_MINMAX(orc_max_i64, int64_t, >)
DEF_MINMAX(orc_min_u8, uint8_t, <)   DEF_MINMAX(orc_max_u8, uint8_t, >)
DEF_MINMAX(orc_min_u16, uint16_t, <) DEF_MINMAX(orc_max_u16, uint16_t, >)
DEF_MINMAX(orc_min_u32, uint32_t, <) DEF_MINMAX(orc_max_u32, uint32_t, >)
DEF_MINMAX(orc_min_u64, uint64_t, <) DEF_MINMAX(orc_max_u64, uint64_t, >)
DEF_MINMAX(orc_min_f32, float, <)   DEF_MINMAX(orc_max_f32, float, >)
DEF_MINMAX(orc_min_f64, double, <)  DEF_MINMAX(orc_max_f64, double, >)

/* CountAccum (inline_agg.rs:36-124): mode 0=All 1=Valid 2=Null */
ORC_EXPORT void orc_count(const uint8_t *validity, int is_null_type, int mode, const uint32_t *gid,
                          int64_t n, int64_t ng, uint64_t *out) {
    for (int64_t g = 0; g < ng; g++) out[g] = 0;
    for (int64_t i = 0; i < n; i++) {
        int valid = is_null_type ? 0 : (validity ? bit_get(validity, i) : 1);
        uint64_t inc = mode == 0 ? 1 : (mode == 1 ? (uint64_t)valid : (uint64_t)!valid);
        out[gid[i]] += inc;
    }
}

/* ========================================================================= */
/* a7: ProbeTable build + probe                                               */
/* (src/daft-recordbatch/src/probeable/probe_table.rs:56-185)                  */
/* One build table (table_idx 0).  For each distinct key a list of build row   */
/* ids in insertion order; probe returns, per probe row, that list (CSR).      */
/* Equality: hashes equal AND comparator(nulls_equal[k], nans_equal=false).    */
/* ========================================================================= */
ORC_EXPORT int64_t orc_join_probe(const orc_col *build, const orc_col *probe, int nkeys,
                                  const uint8_t *nulls_equal,
                                  int64_t *match_offsets /* n_probe+1 */,
                                  int64_t *match_rows /* capacity cap */, int64_t cap) {
    const int64_t nb = build[0].length, np = probe[0].length;
    uint64_t *hb = (uint64_t *)malloc(8 * (size_t)(nb ? nb : 1));
    uint64_t *hp = (uint64_t *)malloc(8 * (size_t)(np ? np : 1));
    if (orc_hash_rows(build, nkeys, hb) || orc_hash_rows(probe, nkeys, hp)) { free(hb); free(hp); return -1; }
    otable t; otable_init(&t, nb);
    /* chain: head row per slot (= first inserted row), next[] links in insertion order, tail[] */
    int64_t *next = (int64_t *)malloc(8 * (size_t)(nb ? nb : 1));
    int64_t *tail = (int64_t *)malloc(8 * (size_t)(nb ? nb : 1));
    for (int64_t i = 0; i < nb; i++) {
        next[i] = -1;
        uint64_t s = hb[i] & t.mask;
        for (;;) {
            int64_t head = t.val[s];
            if (head < 0) { t.val[s] = i; t.hash[s] = hb[i]; tail[i] = i; break; }
            if (t.hash[s] == hb[i] && rows_equal(build, i, build, head, nkeys, nulls_equal, 0)) {
                next[tail[head]] = i; tail[head] = i; break;
            }
            s = (s + 1) & t.mask;
        }
    }
    int64_t total = 0;
    for (int64_t i = 0; i < np; i++) {
        match_offsets[i] = total;
        uint64_t s = hp[i] & t.mask;
        for (;;) {
            int64_t head = t.val[s];
            if (head < 0) break;
            if (t.hash[s] == hp[i] && rows_equal(build, head, probe, i, nkeys, nulls_equal, 0)) {
                for (int64_t r = head; r >= 0; r = next[r]) {
                    if (total < cap) match_rows[total] = r;
                    total++;
                }
                break;
            }
            s = (s + 1) & t.mask;
        }
    }
    match_offsets[np] = total;
    otable_free(&t); free(next); free(tail); free(hb); free(hp);
    return total;
}

/* ========================================================================= */
/* a9: argsort  (src/daft-core/src/array/ops/arrow/sort/primitive/common.rs    */
/* :24-141, src/daft-core/src/array/ops/sort.rs:69-104,171-215,                */
/* src/daft-core/src/kernels/search_sorted.rs:238-245)                         */
/* Multi-key comparator sort; per key descending + nulls_first; floats: all    */
/* NaN equal and greater than everything.  The reference sort is UNSTABLE; the */
/* oracle's canonical form breaks ties by ascending row index.                 */
/* ========================================================================= */
typedef struct { const orc_col *keys; int nkeys; const uint8_t *desc; const uint8_t *nulls_first; } sort_ctx;

static int cmp_values(const orc_col *c, int64_t i, int64_t j) {
    switch (c->dtype) {
#define CMP_T(T) { T x = ((const T *)c->data)[i], y = ((const T *)c->data)[j]; return x < y ? -1 : (x > y ? 1 : 0); }
    case ORC_I8: CMP_T(int8_t)
    case ORC_I16: CMP_T(int16_t)
    case ORC_I32: case ORC_DATE32: CMP_T(int32_t)
    case ORC_I64: case ORC_TIMESTAMP: CMP_T(int64_t)
    case ORC_U8: CMP_T(uint8_t)
    case ORC_U16: CMP_T(uint16_t)
    case ORC_U32: CMP_T(uint32_t)
    case ORC_U64: CMP_T(uint64_t)
#undef CMP_T
    case ORC_F32: {
        float x = ((const float *)c->data)[i], y = ((const float *)c->data)[j];
        int nx = isnan(x), ny = isnan(y);
        if (nx || ny) return nx - ny;
        return x < y ? -1 : (x > y ? 1 : 0);
    }
    case ORC_F64: {
        double x = ((const double *)c->data)[i], y = ((const double *)c->data)[j];
        int nx = isnan(x), ny = isnan(y);
        if (nx || ny) return nx - ny;
        return x < y ? -1 : (x > y ? 1 : 0);
    }
    case ORC_BOOL: {
        int x = bit_get((const uint8_t *)c->data, i), y = bit_get((const uint8_t *)c->data, j);
        return x - y;
    }
    case ORC_LARGE_UTF8:
    case ORC_LARGE_BINARY: {
        int64_t la = c->offsets[i + 1] - c->offsets[i], lb = c->offsets[j + 1] - c->offsets[j];
        int r = memcmp((const uint8_t *)c->data + c->offsets[i], (const uint8_t *)c->data + c->offsets[j],
                       (size_t)(la < lb ? la : lb));
        if (r) return r < 0 ? -1 : 1;
        return la < lb ? -1 : (la > lb ? 1 : 0);
    }
    default:
        return 0;
    }
}
static int sort_cmp(const void *pa, const void *pb, void *vctx) {
    const sort_ctx *c = (const sort_ctx *)vctx;
    int64_t i = *(const int64_t *)pa, j = *(const int64_t *)pb;
    for (int k = 0; k < c->nkeys; k++) {
        const orc_col *col = &c->keys[k];
        int vi = col_valid(col, i), vj = col_valid(col, j);
        int r;
        if (vi && vj) { r = cmp_values(col, i, j); if (c->desc[k]) r = -r; }
        else if (!vi && !vj) r = 0;
        else if (!vi) r = c->nulls_first[k] ? -1 : 1;
        else r = c->nulls_first[k] ? 1 : -1;
        if (r) return r;
    }
    return i < j ? -1 : (i > j ? 1 : 0);
}
ORC_EXPORT void orc_argsort(const orc_col *keys, int nkeys, const uint8_t *desc, const uint8_t *nulls_first,
                            int64_t *out_idx) {
    int64_t n = keys[0].length;
    for (int64_t i = 0; i < n; i++) out_idx[i] = i;
    sort_ctx c = {keys, nkeys, desc, nulls_first};
    qsort_r(out_idx, (size_t)n, sizeof(int64_t), sort_cmp, &c);
}

ORC_EXPORT int orc_abi_version(void) { return 1; }

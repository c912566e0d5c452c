// a7/a8: hash join build + probe.
//
// Replaces ProbeTable / ProbeTableBuilder (src/daft-recordbatch/src/probeable/probe_table.rs:14-213:
// a sequential hashbrown insert of (table<<36|row) ids into per-key Vec<u64>) and the row-at-a-time
// GrowableRecordBatch materialisation of src/daft-local-execution/src/join/inner_join.rs:11-85.
//
// Build (all on the device, no sequential step):
//   1. row hashes of the build keys: XXH3 hash_rows (the chain the reference uses) in general; a single integer-like
//      key column without nulls takes a two-multiply mixer instead (the table's hash is internal: the reference's
//      ProbeTable never exposes it, only the match sets are observable);
//   2. stable radix sort of the build row ids by the TOP B bits of the hash (B = ceil(log2 rows): 2-4 byte passes
//      instead of 8) -> the rows of a bucket are contiguous and in INSERTION ORDER (the reference's match-list order,
//      probe_table.rs:178-181);
//   3. a direct-address table of 2^B 32-byte buckets {start, end, row0, row1, tag0, tag1}: the first TWO rows of a bucket
//      live in the bucket itself (row id + tag = the key value for inline keys, the row hash otherwise), so a probe reads
//      exactly one 32-byte sector unless its bucket holds three or more rows (8% of the occupied buckets at one row per
//      bucket).  No probing, no CAS;
//   4. a blocked Bloom filter (one 64-bit word per key, 2 bits, ~8 bits / build row: L1 / L2 resident), so probe rows
//      without a partner never touch the table (DRAM sectors for large tables, L2 bandwidth for small ones).
// Probe (units of 128 probe rows, one warp per unit; output rows in probe-row order, matches in build insertion order):
//   pass 0 counts the matches of every unit and leaves ONE BIT per probe row ("produces output"); the unit counts are
//   scanned; pass 1 revisits only the rows whose bit is set, recomputes their matches and writes the (probe row, build
//   row) pairs at their final positions (warp scan inside the unit).  Nothing else per probe row is materialised between
//   the passes: no hash column, no per-row counts / offsets / first-match arrays.  Warps never synchronise with each
//   other and fetch the keys of their next unit while they work on the current one.
//   * inline keys: the key is hashed in registers and compared against the tags in the bucket; the probe key may be
//     read THROUGH a selection vector (a view column), so a probe chain never gathers its keys;
//   * anything else: hash_rows over the key columns, then every candidate row is verified by value: nulls match only
//     where null_equals_null, floats compare with IEEE == (NaN never matches, nans_equal=false, probe_table.rs:48-51);
//   * every lane keeps four probe rows in flight (keys, then Bloom words, then bucket sectors: independent loads);
//   * buckets of 32+ rows (heavy duplicate keys) are walked by the whole warp: 32 build rows per step, matches ranked
//     with a ballot so insertion order is kept.
// Join outputs are VIEWS (common.h): both sides' columns become selection vectors over the inputs, so a chain of joins
// moves row ids only and a column is gathered when (and if) a later operator reads it.
// Left / right / outer joins mark matched build rows in a bitmap (1 = matched; the reference keeps the
// complement, src/daft-local-execution/src/join/index_bitmap.rs:5-60) that finalize turns into the
// unmatched-build-rows output.
#include <algorithm>

#include "common.h"
#include "kernels.cuh"
#include "ops.h"

namespace b200 {

constexpr int MAX_JOIN_KEYS = 8;

struct KeyDesc {
    int32_t dtype;
    int32_t nulls_equal;
    const void *data;
    const uint8_t *validity;
    const int64_t *offsets;
};
struct KeySet {
    int32_t nkeys;
    KeyDesc k[MAX_JOIN_KEYS];
};

// One 32-byte sector per bucket: its span in `sorted_rows` plus its first TWO rows inline -- row id and a tag that is
// either the key value itself (inline keys: the match is decided without touching the build side) or the row's full
// hash (a cheap reject before the key compare).  Row ids are 32-bit: a probe table holds fewer than 2^32 rows.
struct __align__(32) JoinBucket {
    unsigned int start;  // first position in `sorted_rows`
    unsigned int end;    // one past the last (start == end: empty bucket)
    unsigned int row0, row1;
    unsigned long long tag0, tag1;
};

}  // namespace b200

struct b200_probe_table {
    int32_t nkeys = 0;
    std::vector<uint8_t> nulls_equal;
    bool track_indices = true;
    bool built = false;
    std::vector<std::unique_ptr<b200::Batch>> pending;  // added tables (shallow copies share device buffers)
    std::vector<std::vector<int32_t>> pending_keys;
    std::unique_ptr<b200::Batch> build;                 // concatenated build side (all columns; key columns plain)
    std::vector<int32_t> key_cols;
    b200::Buf sorted_rows;                              // build row ids sorted by hash (stable)
    b200::Buf slots;                                    // JoinBucket[1 << bits]
    int bits = 0;                                       // bucket = hash >> (64 - bits)
    int inline_width = 0;                               // > 0: tags are key values of that byte width (see JoinBucket)
    b200::Buf bloom;                                    // 64-bit words, null for L2-resident tables
    uint64_t bloom_mask = 0;
    // direct-address mode (unique, dense integer keys): row id by key - min, plus one presence bit per key value
    b200::Buf direct_rows, direct_bits;
    uint64_t direct_min = 0, direct_range = 0;
    int direct_extra = 0;  // direct table on the FIRST key of a multi-column key: the other nkeys - 1 are compared at the hit
    int64_t num_rows = 0;
};
struct b200_join_state {
    b200::Buf matched;  // bit per build row
    int64_t n = 0;
};

namespace b200 {

static int jgrid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8)); }

static KeySet make_keyset(const Batch &b, const std::vector<int32_t> &cols, const std::vector<uint8_t> &ne) {
    KeySet ks;
    ks.nkeys = (int32_t)cols.size();
    for (size_t i = 0; i < cols.size(); i++) {
        if (cols[i] < 0 || cols[i] >= (int32_t)b.cols.size()) fail("join: key column index out of range");
        const Column &c = b.cols[cols[i]];
        require_plain(c);
        ks.k[i].dtype = c.dtype;
        ks.k[i].nulls_equal = ne[i];
        ks.k[i].data = c.data ? c.data->ptr : nullptr;
        ks.k[i].validity = c.null_count > 0 ? as<uint8_t>(c.validity) : nullptr;
        ks.k[i].offsets = as<int64_t>(c.offsets);
    }
    return ks;
}

__device__ __forceinline__ bool key_valid(const KeyDesc &k, int64_t r) {
    if (k.dtype == B200_NULL) return false;
    return k.validity == nullptr || bit_at(k.validity, r);
}
// build_is_equal (src/daft-core/src/array/ops/arrow/comparison.rs:58-83)
__device__ __forceinline__ bool keys_match(const KeySet &a, int64_t i, const KeySet &b, int64_t j) {
    for (int k = 0; k < a.nkeys; k++) {
        const KeyDesc &x = a.k[k], &y = b.k[k];
        const bool vx = key_valid(x, i), vy = key_valid(y, j);
        if (vx != vy) return false;
        if (!vx) {
            if (!x.nulls_equal) return false;
            continue;
        }
        bool eq;
        switch (x.dtype) {
            case B200_BOOL: eq = bit_at((const uint8_t *)x.data, i) == bit_at((const uint8_t *)y.data, j); break;
            case B200_I8: case B200_U8: eq = ((const uint8_t *)x.data)[i] == ((const uint8_t *)y.data)[j]; break;
            case B200_I16: case B200_U16: eq = ((const uint16_t *)x.data)[i] == ((const uint16_t *)y.data)[j]; break;
            case B200_I32: case B200_U32: case B200_DATE32: eq = ((const uint32_t *)x.data)[i] == ((const uint32_t *)y.data)[j]; break;
            case B200_I64: case B200_U64: case B200_TIMESTAMP: eq = ((const uint64_t *)x.data)[i] == ((const uint64_t *)y.data)[j]; break;
            case B200_F32: eq = ((const float *)x.data)[i] == ((const float *)y.data)[j]; break;
            case B200_F64: eq = ((const double *)x.data)[i] == ((const double *)y.data)[j]; break;
            case B200_LARGE_UTF8: case B200_LARGE_BINARY: {
                const int64_t oa = x.offsets[i], la = x.offsets[i + 1] - oa, ob = y.offsets[j], lb = y.offsets[j + 1] - ob;
                eq = la == lb;
                const uint8_t *pa = (const uint8_t *)x.data + oa, *pb = (const uint8_t *)y.data + ob;
                for (int64_t t = 0; eq && t < la; t++) eq = pa[t] == pb[t];
                break;
            }
            default: eq = false;
        }
        if (!eq) return false;
    }
    return true;
}

// ---- build ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long key_bits(const void *data, int width, int64_t r) {
    switch (width) {
        case 1: return ((const uint8_t *)data)[r];
        case 2: return ((const uint16_t *)data)[r];
        case 4: return ((const uint32_t *)data)[r];
        default: return ((const uint64_t *)data)[r];
    }
}
// hash of an inline key (zero-extended to 64 bits): two multiplies, every output bit depends on every input bit
__device__ __forceinline__ uint64_t ihash(unsigned long long v) {
    uint64_t h = v * 0x9E3779B97F4A7C15ULL;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ULL;
    return h ^ (h >> 29);
}
__global__ void hash_inline_keys(const void *__restrict__ data, int width, int64_t n, uint64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = ihash(key_bits(data, width, i));
}
// The head row of a bucket (first position whose top bits differ from its predecessor) walks forward to the bucket's
// end -- buckets hold about one row -- and writes the whole 32-byte entry with two 16-byte stores.  Runs longer than
// BB_WALK are closed by their tail row instead (it sees no head within BB_WALK positions behind it).
constexpr int BB_WALK = 32;
__global__ void bucket_bounds(const uint64_t *__restrict__ sorted_hash, const int64_t *__restrict__ sorted_rows, int64_t n, int shift,
                              const void *__restrict__ key_data, int inline_width, JoinBucket *__restrict__ table) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t h = sorted_hash[i], b = h >> shift;
        const bool head = i == 0 || (sorted_hash[i - 1] >> shift) != b;
        if (head) {
            int64_t j = i + 1;
            while (j < n && j - i <= BB_WALK && (sorted_hash[j] >> shift) == b) j++;
            const bool closed = j == n || (sorted_hash[j] >> shift) != b;
            const int64_t r0 = sorted_rows[i];
            const bool two = j - i >= 2;
            const int64_t r1 = two ? sorted_rows[i + 1] : 0;
            uint4 lo;
            lo.x = (unsigned)i;
            lo.y = closed ? (unsigned)j : (unsigned)i;  // a long run: the tail row stores `end`
            lo.z = (unsigned)r0;
            lo.w = (unsigned)r1;
            ulonglong2 hi;
            hi.x = inline_width ? key_bits(key_data, inline_width, r0) : h;
            hi.y = two ? (inline_width ? key_bits(key_data, inline_width, r1) : sorted_hash[i + 1]) : 0ULL;
            if (closed) {
                *reinterpret_cast<uint4 *>(&table[b]) = lo;
            } else {  // (the tail row owns `end`: no store of ours may overwrite it)
                table[b].start = lo.x;
                table[b].row0 = lo.z;
                table[b].row1 = lo.w;
            }
            *reinterpret_cast<ulonglong2 *>(&table[b].tag0) = hi;
        } else {
            const bool tail = i == n - 1 || (sorted_hash[i + 1] >> shift) != b;
            if (tail) {
                // distance to the head: only runs longer than BB_WALK + 1 rows are closed here
                int64_t back = 1;
                while (back <= BB_WALK + 1 && i - back >= 0 && (sorted_hash[i - back] >> shift) == b) back++;
                if (back > BB_WALK + 1) table[b].end = (unsigned)(i + 1);
            }
        }
    }
}

// Blocked Bloom filter: word = bits 18.. of the hash, two bit positions from its low 10 bits.
// one bit in each 32-bit half of the word (see bloom_pass)
__device__ __forceinline__ unsigned long long bloom_need(uint64_t h) {
    return (1ULL << (h & 31)) | (1ULL << (32 + ((h >> 5) & 31)));
}
__global__ void bloom_build(const uint64_t *__restrict__ hashes, int64_t n, unsigned long long *__restrict__ words, uint64_t mask) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t h = hashes[i];
        atomicOr(&words[(h >> 18) & mask], bloom_need(h));
    }
}

// ---- probe -----------------------------------------------------------------------------------------
constexpr int PT = 256;          // threads per CTA
constexpr int UR = 4;            // probe rows per lane per step of the count pass: independent loads in flight
constexpr int CHUNK = 1024;      // probe rows per warp work item = 32 bitmap words = one count
constexpr int WR = 2;            // hit rows per lane per step of the write pass
constexpr int LONG_BUCKET = 32;  // buckets at least this long are walked by the whole warp

struct ProbeArgs {
    const JoinBucket *table;
    const int64_t *sorted_rows;
    const unsigned long long *bloom;
    unsigned long long bloom_mask;
    int shift;
    int _pad;
    // single integer-like key, hashed in registers; the probe key may sit behind a selection vector
    const void *bkey;
    const void *pkey;
    const uint8_t *pvalid;
    const int64_t *psel;
    // generic keys: hash_rows column of the probe side + value comparison
    const uint64_t *phash;
    // direct-address table (unique dense integer keys): presence bitmap + build row by (key - dmin)
    const unsigned int *dbits, *drows;
    unsigned long long dmin, drange;
    // generic mode: all keys.  Direct table on a key prefix: the keys AFTER the first (the probe side's gathered plain)
    KeySet bk, pk;
};

template <int W>
__device__ __forceinline__ unsigned long long key_w(const void *data, int64_t r) {
    if (W == 1) return ((const uint8_t *)data)[r];
    if (W == 2) return ((const uint16_t *)data)[r];
    if (W == 4) return ((const uint32_t *)data)[r];
    return ((const uint64_t *)data)[r];
}
// the probe-side key stream is read once: evict-first, so it does not push the table / Bloom words out of L2
template <int W>
__device__ __forceinline__ unsigned long long key_stream(const void *data, int64_t r) {
    if (W == 1) return __ldcs((const uint8_t *)data + r);
    if (W == 2) return __ldcs((const uint16_t *)data + r);
    if (W == 4) return __ldcs((const uint32_t *)data + r);
    return __ldcs((const unsigned long long *)data + r);
}
__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    const int lo = __shfl_sync(0xffffffffu, (int)(v & 0xffffffffLL), src), hi = __shfl_sync(0xffffffffu, (int)(v >> 32), src);
    return ((long long)hi << 32) | (unsigned int)lo;
}
// Bloom test on a 64-bit word: one bit in each 32-bit half (two 32-bit shifts instead of three 64-bit ones)
__device__ __forceinline__ bool bloom_pass(unsigned long long w, uint64_t h) {
    const unsigned lo = (unsigned)w, hi = (unsigned)(w >> 32);
    return ((lo >> (h & 31)) & (hi >> ((h >> 5) & 31)) & 1u) != 0;
}

// does build row `br` match probe row `i` (key value `v`)?
template <int W>
__device__ __forceinline__ bool row_matches(const ProbeArgs &a, int64_t br, unsigned long long v, int64_t i) {
    return W ? key_w<W>(a.bkey, br) == v : keys_match(a.bk, br, a.pk, i);
}
// matches of the two inline entries: bit 0 = row0, bit 1 = row1
template <int W>
__device__ __forceinline__ unsigned inline_matches(const ProbeArgs &a, const JoinBucket &e, unsigned len, unsigned long long v, uint64_t h, int64_t i) {
    unsigned m = 0;
    if (W) {
        m = (len >= 1 && e.tag0 == v ? 1u : 0u) | (len >= 2 && e.tag1 == v ? 2u : 0u);
    } else {
        if (len >= 1 && e.tag0 == h && keys_match(a.bk, e.row0, a.pk, i)) m |= 1u;
        if (len >= 2 && e.tag1 == h && keys_match(a.bk, e.row1, a.pk, i)) m |= 2u;
    }
    return m;
}
__device__ __forceinline__ JoinBucket load_bucket(const ProbeArgs &a, uint64_t h) {
    const JoinBucket *b = &a.table[h >> a.shift];
    const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(b));
    const ulonglong2 hi = __ldg(reinterpret_cast<const ulonglong2 *>(&b->tag0));
    JoinBucket e;
    e.start = lo.x;
    e.end = lo.y;
    e.row0 = lo.z;
    e.row1 = lo.w;
    e.tag0 = hi.x;
    e.tag1 = hi.y;
    return e;
}
// matches of probe row i beyond the two inline entries of a short bucket (thread-serial)
template <int W>
__device__ __forceinline__ unsigned tail_matches(const ProbeArgs &a, const JoinBucket &e, unsigned long long v, int64_t i) {
    unsigned c = 0;
    for (unsigned t = e.start + 2; t < e.end; t++) c += row_matches<W>(a, a.sorted_rows[t], v, i) ? 1u : 0u;
    return c;
}
// matches of ONE probe row (owner lane `src`) in a long bucket, counted by the whole warp
template <int W>
__device__ __forceinline__ unsigned warp_count_long(const ProbeArgs &a, unsigned start, unsigned end, unsigned long long v, int64_t i, int lane) {
    unsigned c = 0;
    for (unsigned t0 = start; t0 < end; t0 += 32) {
        const unsigned t = t0 + lane;
        const bool mt = t < end && row_matches<W>(a, a.sorted_rows[t], v, i);
        c += __popc(__ballot_sync(0xffffffffu, mt));
    }
    return c;
}

// key (inline: value, generic: row hash) of probe row i; false = the row cannot match (null / negative selection)
template <int W, bool SEL, bool NULLS>
__device__ __forceinline__ bool load_key(const ProbeArgs &a, int64_t i, unsigned long long &v) {
    if (W) {
        int64_t row = i;
        if (SEL) {
            row = __ldcs(a.psel + i);
            if (row < 0) return false;
        }
        if (NULLS && !bit_at(a.pvalid, row)) return false;
        v = key_stream<W>(a.pkey, row);
        return true;
    }
    v = __ldcs(a.phash + i);
    return true;
}

// matches of one probe row per lane, any bucket length (kept out of line: the registers of the rare path must not
// burden the straight-line step); row0 = probe row of lane 0
template <int W>
__device__ __noinline__ unsigned count_slow_row(const ProbeArgs &a, JoinBucket e, unsigned long long v, unsigned im, int64_t i, int64_t row0, int lane) {
    const unsigned len = e.end - e.start;
    unsigned c = __popc(im);
    if (len > 2 && len < LONG_BUCKET) c += tail_matches<W>(a, e, v, i);
    unsigned pending = __ballot_sync(0xffffffffu, len >= LONG_BUCKET);
    while (pending) {  // long buckets: the whole warp walks them, one owner lane at a time
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const unsigned s = __shfl_sync(0xffffffffu, e.start, src) + 2, en = __shfl_sync(0xffffffffu, e.end, src);
        const unsigned long long vv = (unsigned long long)shfl_ll((long long)v, src);
        const unsigned cc = warp_count_long<W>(a, s, en, vv, row0 + src, lane);
        if (lane == src) c += cc;
    }
    return c;
}

// One step of the count pass: UR rows per lane.  The common case -- every probed bucket holds at most two rows, which
// live inline -- is straight-line code: hash, Bloom word, (predicated) bucket sector, two tag compares, three ballots.
// Anything else (a bucket of three or more rows) takes the slow tail for the whole step.
template <int W, bool FULL, bool KEEP>
__device__ __forceinline__ long long count_step(const ProbeArgs &a, int64_t base, int64_t n, int lane, const unsigned long long (&v)[UR], const bool (&hit0)[UR],
                                                uint32_t *__restrict__ bits) {
    uint64_t h[UR];
    unsigned long long bw[UR];
    bool hit[UR];
#pragma unroll
    for (int r = 0; r < UR; r++) {
        h[r] = W ? ihash(v[r]) : v[r];
        bw[r] = __ldg(a.bloom + ((h[r] >> 18) & a.bloom_mask));  // (a row without a key reads the word of hash(0): harmless)
    }
#pragma unroll
    for (int r = 0; r < UR; r++) hit[r] = hit0[r] && bloom_pass(bw[r], h[r]);
    JoinBucket e[UR];
#pragma unroll
    for (int r = 0; r < UR; r++) {
        e[r].start = e[r].end = 0;
        e[r].tag0 = e[r].tag1 = 0;
        e[r].row0 = e[r].row1 = 0;
        if (hit[r]) e[r] = load_bucket(a, h[r]);
    }
    unsigned im[UR];
    bool rare = false;
#pragma unroll
    for (int r = 0; r < UR; r++) {
        const unsigned len = e[r].end - e[r].start;
        im[r] = inline_matches<W>(a, e[r], len, v[r], h[r], base + r * 32 + lane);
        rare = rare || len > 2;
    }
    long long total = 0;
    if (!__any_sync(0xffffffffu, rare)) {
#pragma unroll
        for (int r = 0; r < UR; r++) {
            const unsigned m0 = __ballot_sync(0xffffffffu, (im[r] & 1u) != 0), m1 = __ballot_sync(0xffffffffu, (im[r] & 2u) != 0);
            unsigned out = m0 | m1;
            unsigned extra = 0;
            if (KEEP) {  // unmatched probe rows produce a null-extended row
                const unsigned live = FULL ? 0xffffffffu : __ballot_sync(0xffffffffu, base + r * 32 + lane < n);
                extra = __popc(live & ~out);
                out = live;
            }
            if (lane == 0 && (FULL || base + r * 32 < n)) bits[(base >> 5) + r] = out;
            total += __popc(m0) + __popc(m1) + extra;
        }
        return total;
    }
    // ---- slow tail: some bucket of this step holds three or more rows
#pragma unroll
    for (int r = 0; r < UR; r++) {
        const int64_t i = base + r * 32 + lane;
        unsigned c = count_slow_row<W>(a, e[r], v[r], im[r], i, base + r * 32, lane);
        if (KEEP && c == 0 && (FULL || i < n)) c = 1;
        const unsigned m = __ballot_sync(0xffffffffu, c > 0);
        if (lane == 0 && (FULL || base + r * 32 < n)) bits[(base >> 5) + r] = m;
        unsigned x = c;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        total += x;
    }
    return total;
}

// ---- pass 0: one warp per chunk of 1024 probe rows: 32 "produces output" bitmap words and one output-row count.
// Eight steps of 4 rows per lane; the keys of the next step are fetched before the current one is processed.
template <int W, bool SEL, bool NULLS, bool KEEP, bool FULL>
__device__ __forceinline__ long long count_chunk(const ProbeArgs &a, int64_t cbase, int64_t n, int lane, uint32_t *__restrict__ bits) {
    constexpr int STEPS = CHUNK / (32 * UR);
    long long total = 0;
    unsigned long long nv[UR];
    bool nh[UR];
#pragma unroll
    for (int r = 0; r < UR; r++) {
        const int64_t i = cbase + r * 32 + lane;
        nv[r] = 0;
        nh[r] = (FULL || i < n) && load_key<W, SEL, NULLS>(a, i, nv[r]);
    }
#pragma unroll 1
    for (int step = 0; step < STEPS; step++) {
        const int64_t base = cbase + (int64_t)step * (32 * UR);
        unsigned long long v[UR];
        bool hit[UR];
#pragma unroll
        for (int r = 0; r < UR; r++) {
            v[r] = nv[r];
            hit[r] = nh[r];
        }
        if (step + 1 < STEPS) {
#pragma unroll
            for (int r = 0; r < UR; r++) {
                const int64_t i = base + (32 * UR) + r * 32 + lane;
                nv[r] = 0;
                nh[r] = (FULL || i < n) && load_key<W, SEL, NULLS>(a, i, nv[r]);
            }
        }
        total += count_step<W, FULL, KEEP>(a, base, n, lane, v, hit, bits);
    }
    return total;
}
template <int W, bool SEL, bool NULLS, bool KEEP>
__global__ void __launch_bounds__(PT, 3) probe_count_kernel(const __grid_constant__ ProbeArgs a, int64_t n, int64_t nchunks, int64_t *__restrict__ chunk_cnt,
                                                            uint32_t *__restrict__ bits) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * (PT / 32) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (PT / 32);
    for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
        const int64_t cbase = chunk * CHUNK;
        const long long total = cbase + CHUNK <= n ? count_chunk<W, SEL, NULLS, KEEP, true>(a, cbase, n, lane, bits)
                                                   : count_chunk<W, SEL, NULLS, KEEP, false>(a, cbase, n, lane, bits);
        if (lane == 0) chunk_cnt[chunk] = total;
    }
}

// ---- pass 1: one warp per chunk.  The set bits of the chunk's 32 bitmap words are expanded into a dense list of row
// offsets (shared memory), so every lane works on a row that produces output: WR rows per lane per step.
template <int W, bool SEL, bool NULLS>
__global__ void __launch_bounds__(PT, 3) probe_write_kernel(const __grid_constant__ ProbeArgs a, int64_t n, int64_t nchunks,
                                                            const int64_t *__restrict__ chunk_off, const uint32_t *__restrict__ bits,
                                                            int64_t *__restrict__ out_probe, int64_t *__restrict__ out_build,
                                                            unsigned int *__restrict__ matched_bits) {
    __shared__ unsigned short s_list[PT / 32][CHUNK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * (PT / 32) + warp, nw = (int64_t)gridDim.x * (PT / 32);
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned short *list = s_list[warp];
    for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
        long long o = chunk_off[chunk];
        if (chunk_off[chunk + 1] == o) continue;  // nothing to write (uniform for the warp)
        const int64_t cbase = chunk * CHUNK;
        // ---- dense list of the rows whose bit is set
        const int64_t word = (cbase >> 5) + lane;
        unsigned w = (cbase + (int64_t)lane * 32 < n) ? bits[word] : 0u;
        unsigned pre = __popc(w);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += y;
        }
        const unsigned hits = __shfl_sync(0xffffffffu, pre, 31);
        pre -= __popc(w);
        __syncwarp();
        while (w) {
            const int b = __ffs(w) - 1;
            w &= w - 1;
            list[pre++] = (unsigned short)(lane * 32 + b);
        }
        __syncwarp();
        // ---- WR hit rows per lane per step, in row order
        for (unsigned j0 = 0; j0 < hits; j0 += 32 * WR) {
            unsigned long long v[WR];
            int64_t row[WR];
            bool act[WR], hit[WR];
#pragma unroll
            for (int r = 0; r < WR; r++) {
                const unsigned j = j0 + r * 32 + lane;
                act[r] = j < hits;
                row[r] = cbase + (act[r] ? list[j] : 0);
                v[r] = 0;
                hit[r] = act[r] && load_key<W, SEL, NULLS>(a, row[r], v[r]);
            }
            uint64_t h[WR];
            JoinBucket e[WR];
#pragma unroll
            for (int r = 0; r < WR; r++) {
                h[r] = W ? ihash(v[r]) : v[r];
                e[r].start = e[r].end = 0;
                if (hit[r]) e[r] = load_bucket(a, h[r]);
            }
#pragma unroll
            for (int r = 0; r < WR; r++) {
                const int64_t i = row[r];
                const unsigned len = e[r].end - e[r].start;
                const unsigned im = inline_matches<W>(a, e[r], len, v[r], h[r], i);
                unsigned c = __popc(im);
                const bool is_long = len >= LONG_BUCKET;
                if (len > 2 && !is_long) c += tail_matches<W>(a, e[r], v[r], i);
                unsigned pending = __ballot_sync(0xffffffffu, is_long);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    const unsigned s = __shfl_sync(0xffffffffu, e[r].start, src) + 2, en = __shfl_sync(0xffffffffu, e[r].end, src);
                    const unsigned long long vv = (unsigned long long)shfl_ll((long long)v[r], src);
                    const unsigned cc = warp_count_long<W>(a, s, en, vv, shfl_ll(i, src), lane);
                    if (lane == src) c += cc;
                }
                const bool unmatched = act[r] && c == 0;  // its bit was set all the same: a kept probe row
                if (unmatched) c = 1;
                // exclusive scan of c over the lanes: ballots when no row has two matches
                long long off;
                if (__all_sync(0xffffffffu, c <= 1)) {
                    const unsigned m = __ballot_sync(0xffffffffu, c == 1);
                    off = o + __popc(m & lt_mask);
                    o += __popc(m);
                } else {
                    long long x = c;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const long long y = shfl_ll(x, max(lane - d, 0));
                        if (lane >= d) x += y;
                    }
                    off = o + x - c;
                    o += shfl_ll(x, 31);
                }
                if (unmatched) {
                    out_probe[off] = i;
                    out_build[off] = -1;
                } else if (c > 0 && !is_long) {
                    long long p = off;
                    if (im & 1u) {
                        out_probe[p] = i;
                        out_build[p] = e[r].row0;
                        if (matched_bits) atomicOr(&matched_bits[e[r].row0 >> 5], 1u << (e[r].row0 & 31));
                        p++;
                    }
                    if (im & 2u) {
                        out_probe[p] = i;
                        out_build[p] = e[r].row1;
                        if (matched_bits) atomicOr(&matched_bits[e[r].row1 >> 5], 1u << (e[r].row1 & 31));
                        p++;
                    }
                    if (len > 2) {
                        for (unsigned t = e[r].start + 2; t < e[r].end; t++) {
                            const int64_t br = a.sorted_rows[t];
                            if (row_matches<W>(a, br, v[r], i)) {
                                out_probe[p] = i;
                                out_build[p] = br;
                                if (matched_bits) atomicOr(&matched_bits[br >> 5], 1u << (br & 31));
                                p++;
                            }
                        }
                    }
                }
                // long buckets: the whole warp writes them, one owner lane at a time, insertion order kept by ballots
                pending = __ballot_sync(0xffffffffu, is_long && c > 0 && !unmatched);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    const unsigned s = __shfl_sync(0xffffffffu, e[r].start, src), en = __shfl_sync(0xffffffffu, e[r].end, src);
                    const unsigned long long vv = (unsigned long long)shfl_ll((long long)v[r], src);
                    const int64_t ii = shfl_ll(i, src);
                    long long oo = shfl_ll(off, src);
                    for (unsigned t0 = s; t0 < en; t0 += 32) {
                        const unsigned t = t0 + lane;
                        int64_t br = -1;
                        bool mt = false;
                        if (t < en) {
                            br = a.sorted_rows[t];
                            mt = row_matches<W>(a, br, vv, ii);
                        }
                        const unsigned mm = __ballot_sync(0xffffffffu, mt);
                        if (mt) {
                            const long long p = oo + __popc(mm & lt_mask);
                            out_probe[p] = ii;
                            out_build[p] = br;
                            if (matched_bits) atomicOr(&matched_bits[br >> 5], 1u << (br & 31));
                        }
                        oo += __popc(mm);
                    }
                }
            }
        }
    }
}

// ---- direct-address mode -----------------------------------------------------------------------------------------
// Join keys are very often dense integers with unique build values (every TPC-H foreign key -> primary key join).  When
// the build keys are unique and their range is at most 32x their count, the table is an array indexed by key - min:
// one presence BIT per key value (a 1 M-row table: 125 KB, L1 / L2 resident -- the exact filter a Bloom filter
// approximates) and one 32-bit build row per key value.  No hashing: a probe is a subtract, a range check, a bit test
// and, for the rows that pass, one 4-byte load.  Same two passes, same outputs as the bucket table.
__global__ void key_minmax(const void *__restrict__ data, int width, int64_t n, unsigned long long *__restrict__ mm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long lo = ~0ULL, hi = 0ULL;
    for (; i < n; i += stride) {
        const unsigned long long v = key_bits(data, width, i);
        lo = min(lo, v);
        hi = max(hi, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
        lo = min(lo, (unsigned long long)shfl_ll((long long)lo, (threadIdx.x & 31) ^ d));
        hi = max(hi, (unsigned long long)shfl_ll((long long)hi, (threadIdx.x & 31) ^ d));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&mm[0], lo);
        atomicMax(&mm[1], hi);
    }
}
__global__ void direct_fill(const void *__restrict__ data, int width, int64_t n, unsigned long long dmin, unsigned int *__restrict__ rows,
                            unsigned int *__restrict__ bits, int *__restrict__ dup) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const unsigned long long idx = key_bits(data, width, i) - dmin;
        const unsigned int old = atomicMin(&rows[idx], (unsigned int)i);  // the first row of a key wins (insertion order)
        if (old != 0xFFFFFFFFu) *dup = 1;                                // a second row with this key: not unique
        atomicOr(&bits[idx >> 5], 1u << (idx & 31));
    }
}

// pass 0 also STASHES the build row of every output row, densely per chunk (stash[chunk * CHUNK + j] = build row of the
// chunk's j-th output row, 0xFFFFFFFF = none): pass 1 then needs neither the keys nor the table again.
// EXTRA: the table addresses the first column of a multi-column key; a hit still has to agree on the other columns
// (a.bk / a.pk hold those, the probe side's indexed by the probe row itself).
template <int W, bool SEL, bool NULLS, bool KEEP, bool EXTRA>
__global__ void __launch_bounds__(PT, 4) direct_count_kernel(const __grid_constant__ ProbeArgs a, int64_t n, int64_t nchunks, int64_t *__restrict__ chunk_cnt,
                                                             uint32_t *__restrict__ bits, uint32_t *__restrict__ stash) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * (PT / 32) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (PT / 32);
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr int DR = 4;  // rows per lane per step; the keys of the next step are in flight while this one is processed
    constexpr int STEPS = CHUNK / (32 * DR);
    for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
        const int64_t cbase = chunk * CHUNK;
        const bool full = cbase + CHUNK <= n;
        unsigned total = 0;
        unsigned long long nv[DR];
        bool nh[DR];
#pragma unroll
        for (int r = 0; r < DR; r++) {
            const int64_t i = cbase + r * 32 + lane;
            nv[r] = 0;
            nh[r] = (full || i < n) && load_key<W, SEL, NULLS>(a, i, nv[r]);
        }
#pragma unroll 1
        for (int step = 0; step < STEPS; step++) {
            const int64_t base = cbase + (int64_t)step * (32 * DR);
            unsigned long long idx[DR];
            bool hit[DR];
#pragma unroll
            for (int r = 0; r < DR; r++) {
                idx[r] = nv[r] - a.dmin;
                hit[r] = nh[r] && idx[r] < a.drange;
            }
            if (step + 1 < STEPS) {
#pragma unroll
                for (int r = 0; r < DR; r++) {
                    const int64_t i = base + (32 * DR) + r * 32 + lane;
                    nv[r] = 0;
                    nh[r] = (full || i < n) && load_key<W, SEL, NULLS>(a, i, nv[r]);
                }
            }
            unsigned word[DR];
#pragma unroll
            for (int r = 0; r < DR; r++) word[r] = hit[r] ? __ldg(a.dbits + (idx[r] >> 5)) : 0u;
            unsigned row[DR];
#pragma unroll
            for (int r = 0; r < DR; r++) {
                hit[r] = hit[r] && ((word[r] >> (idx[r] & 31)) & 1u);
                row[r] = hit[r] && (stash || EXTRA) ? __ldg(a.drows + idx[r]) : 0xFFFFFFFFu;
            }
            if (EXTRA) {
#pragma unroll
                for (int r = 0; r < DR; r++)
                    if (hit[r] && !keys_match(a.bk, (int64_t)row[r], a.pk, base + r * 32 + lane)) {
                        hit[r] = false;
                        row[r] = 0xFFFFFFFFu;
                    }
            }
#pragma unroll
            for (int r = 0; r < DR; r++) {
                unsigned m = __ballot_sync(0xffffffffu, hit[r]);
                if (KEEP) m = full ? 0xffffffffu : __ballot_sync(0xffffffffu, base + r * 32 + lane < n);
                if (lane == 0 && (full || base + r * 32 < n)) bits[(base >> 5) + r] = m;
                if (stash && ((m >> lane) & 1u)) stash[cbase + total + __popc(m & lt_mask)] = row[r];
                total += __popc(m);
            }
        }
        if (lane == 0) chunk_cnt[chunk] = total;
    }
}

// pass 1: probe rows from the bitmap, build rows from the stash -- a pure streaming pass
__global__ void __launch_bounds__(PT, 4) direct_write_kernel(int64_t n, int64_t nchunks, const int64_t *__restrict__ chunk_off,
                                                             const uint32_t *__restrict__ bits, const uint32_t *__restrict__ stash,
                                                             int64_t *__restrict__ out_probe, int64_t *__restrict__ out_build,
                                                             unsigned int *__restrict__ matched_bits) {
    __shared__ unsigned short s_list[PT / 32][CHUNK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * (PT / 32) + warp, nw = (int64_t)gridDim.x * (PT / 32);
    unsigned short *list = s_list[warp];
    for (int64_t chunk = gw; chunk < nchunks; chunk += nw) {
        const long long o = chunk_off[chunk];
        if (chunk_off[chunk + 1] == o) continue;
        const int64_t cbase = chunk * CHUNK;
        unsigned w = (cbase + (int64_t)lane * 32 < n) ? bits[(cbase >> 5) + lane] : 0u;
        unsigned pre = __popc(w);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, pre, d);
            if (lane >= d) pre += y;
        }
        const unsigned hits = __shfl_sync(0xffffffffu, pre, 31);
        pre -= __popc(w);
        __syncwarp();
        while (w) {
            const int b = __ffs(w) - 1;
            w &= w - 1;
            list[pre++] = (unsigned short)(lane * 32 + b);
        }
        __syncwarp();
        for (unsigned j = lane; j < hits; j += 32) {
            const unsigned r = stash[cbase + j];
            const long long br = r == 0xFFFFFFFFu ? -1LL : (long long)r;
            out_probe[o + j] = cbase + list[j];
            out_build[o + j] = br;
            if (matched_bits && br >= 0) atomicOr(&matched_bits[br >> 5], 1u << (br & 31));
        }
    }
}

// ---- host --------------------------------------------------------------------------------------------
b200_probe_table *probe_table_create(int32_t nkeys, const uint8_t *null_equals_null, bool track_indices) {
    if (nkeys < 1 || nkeys > MAX_JOIN_KEYS) fail("join: between 1 and 8 key columns are supported");
    auto *pt = new b200_probe_table;
    pt->nkeys = nkeys;
    pt->nulls_equal.assign((size_t)nkeys, 0);
    if (null_equals_null)
        for (int i = 0; i < nkeys; i++) pt->nulls_equal[i] = null_equals_null[i] ? 1 : 0;
    pt->track_indices = track_indices;
    return pt;
}

void probe_table_add(b200_probe_table *pt, const Batch &build, const int32_t *key_cols) {
    if (!pt) fail("null probe table");
    if (pt->built) fail("probe table already built");
    pt->pending.push_back(std::make_unique<Batch>(build));
    pt->pending_keys.emplace_back(key_cols, key_cols + pt->nkeys);
    if (pt->pending_keys.back() != pt->pending_keys.front()) fail("join: all build tables must use the same key columns");
}

void probe_table_build(b200_probe_table *pt) {
    require_device();
    if (!pt) fail("null probe table");
    if (pt->built) return;
    if (pt->pending.empty()) fail("join: probe table has no build tables");
    cudaStream_t st = ctx().stream;
    pt->key_cols = pt->pending_keys[0];
    if (pt->pending.size() == 1) {
        // payload columns of a large build side may stay views (a join output as build side): only the keys are read
        // here.  A small one is gathered now: every output column then shares the one build-row selection vector.
        const bool small = pt->pending[0]->num_rows <= ((int64_t)1 << 22);
        pt->build = std::make_unique<Batch>(materialize_batch(*pt->pending[0], small ? nullptr : &pt->key_cols));
    } else {
        auto b = std::make_unique<Batch>();
        const size_t nc = pt->pending[0]->cols.size();
        std::vector<Batch> plain;
        for (auto &p : pt->pending) {
            if (p->cols.size() != nc) fail("join: build tables have different schemas");
            b->num_rows += p->num_rows;
            plain.push_back(materialize_batch(*p));
        }
        for (size_t c = 0; c < nc; c++) {
            std::vector<const Column *> cs;
            for (auto &p : plain) cs.push_back(&p.cols[c]);
            b->cols.push_back(concat_columns(cs));
        }
        pt->build = std::move(b);
    }
    pt->pending.clear();
    const Batch &b = *pt->build;
    const int64_t n = b.num_rows;
    pt->num_rows = n;
    // reference limits (probe_table.rs:22-25,128-129) are 2^36 rows per table; ours is the same order
    KernelTimer kt_build("join_build");
    if (n >= ((int64_t)1 << 32) - 1) fail("join: a probe table holds fewer than 2^32 rows");
    // inline key tags: one integer-like key column without nulls (floats keep IEEE ==, strings need their bytes)
    const Column &k0 = b.cols[pt->key_cols[0]];
    const bool int_like = k0.dtype == B200_I8 || k0.dtype == B200_I16 || k0.dtype == B200_I32 || k0.dtype == B200_I64 || k0.dtype == B200_U8 ||
                          k0.dtype == B200_U16 || k0.dtype == B200_U32 || k0.dtype == B200_U64 || k0.dtype == B200_DATE32 || k0.dtype == B200_TIMESTAMP;
    pt->inline_width = (pt->nkeys == 1 && int_like && k0.null_count == 0) ? dtype_width(k0.dtype) : 0;
    // a multi-column key whose first column alone is unique and dense (TPC-H Q5: (C_CUSTKEY, C_NATIONKEY)) gets the
    // direct-address table on that column; the probe compares the remaining columns at the one candidate row
    const int direct_width = (int_like && k0.null_count == 0) ? dtype_width(k0.dtype) : 0;
    if (direct_width && n > 0) {  // unique dense integer keys: the direct-address table needs neither hashes nor a sort
        const char *env = getenv("B200_JOIN_DIRECT");
        if (!(env && *env == '0')) {
            Buf mm = dev_alloc(16);
            const unsigned long long init[2] = {~0ULL, 0ULL};
            B200_CUDA(cudaMemcpyAsync(mm->ptr, init, 16, cudaMemcpyHostToDevice, st));
            key_minmax<<<jgrid(n), 256, 0, st>>>(k0.data->ptr, direct_width, n, as<unsigned long long>(mm));
            count_launch();
            unsigned long long h[2];
            B200_CUDA(cudaMemcpyAsync(h, mm->ptr, 16, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            const unsigned long long range = h[1] - h[0] + 1;
            if (h[1] >= h[0] && range != 0 && range <= (unsigned long long)std::max<int64_t>(n * 32, 1 << 16) && range <= (1ULL << 31)) {
                Buf rows = dev_alloc((size_t)range * 4), dbits = dev_alloc_zero((size_t)(range / 32 + 2) * 4), dup = dev_alloc_zero(8);
                fill_u8(rows->ptr, 0xFF, (size_t)range * 4);
                direct_fill<<<jgrid(n), 256, 0, st>>>(k0.data->ptr, direct_width, n, h[0], as<unsigned int>(rows), as<unsigned int>(dbits), as<int>(dup));
                count_launch();
                if (read_i64(as<int64_t>(dup)) == 0) {
                    pt->direct_rows = rows;
                    pt->direct_bits = dbits;
                    pt->direct_min = h[0];
                    pt->direct_range = range;
                    pt->direct_extra = pt->nkeys - 1;
                    pt->inline_width = direct_width;
                    op_done();
                    pt->built = true;
                    return;
                }
            }
        }
    }
    Buf hashes;
    if (pt->inline_width) {
        hashes = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
        if (n > 0) {
            hash_inline_keys<<<jgrid(n), 256, 0, st>>>(k0.data->ptr, pt->inline_width, n, as<uint64_t>(hashes));
            count_launch();
        }
    } else {
        hashes = hash_rows(b, pt->key_cols.data(), pt->nkeys, nullptr);
    }
    Buf rows = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    int bits = 4;
    while (bits < 40 && ((int64_t)1 << bits) < n) bits++;  // at most one row per bucket on average
    pt->bits = bits;
    pt->slots = dev_alloc_zero(((size_t)1 << bits) * sizeof(JoinBucket));
    if (n > 0) {
        fill_iota(as<int64_t>(rows), n);
        const int sort_bits = (bits + 7) / 8 * 8;  // whole byte passes over the top of the hash
        radix_sort_pairs(hashes, rows, n, 64 - sort_bits, 64, false);
        bucket_bounds<<<jgrid(n), 256, 0, st>>>(as<uint64_t>(hashes), as<int64_t>(rows), n, 64 - bits, k0.data ? k0.data->ptr : nullptr,
                                                pt->inline_width, as<JoinBucket>(pt->slots));
        count_launch();
        // Bloom filter, ~8 bits per build row: probe rows without a partner stop at one 8-byte word (L1 / L2 resident)
        // instead of a 32-byte table sector -- in DRAM for large tables, L2 bandwidth for small ones
        const char *env = getenv("B200_JOIN_BLOOM");
        const bool want = env ? *env != '0' : true;
        if (want) {
            uint64_t words = 1024;
            while ((int64_t)words * 8 < n) words <<= 1;
            pt->bloom = dev_alloc_zero(words * 8);
            pt->bloom_mask = words - 1;
            bloom_build<<<jgrid(n), 256, 0, st>>>(as<uint64_t>(hashes), n, as<unsigned long long>(pt->bloom), pt->bloom_mask);
            count_launch();
        }
    }
    if (!pt->bloom) {  // the probe kernels always consult the filter: a one-word filter that passes everything
        pt->bloom = dev_alloc(8);
        fill_u8(pt->bloom->ptr, 0xFF, 8);
        pt->bloom_mask = 0;
    }
    pt->sorted_rows = rows;
    op_done();
    pt->built = true;
}

int64_t probe_table_num_rows(const b200_probe_table *pt) { return pt ? pt->num_rows : -1; }
void probe_table_free(b200_probe_table *pt) { delete pt; }

static void check_probe(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols) {
    if (!pt || !pt->built) fail("join: probe table is not built");
    for (int k = 0; k < pt->nkeys; k++) {
        if (key_cols[k] < 0 || key_cols[k] >= (int32_t)probe.cols.size()) fail("join: probe key column index out of range");
        const int32_t a = pt->build->cols[pt->key_cols[k]].dtype, b = probe.cols[key_cols[k]].dtype;
        // join keys must have identical dtypes: the same number hashes differently as Int32 vs Int64
        // (src/daft-local-execution/src/pipeline.rs:1215-1221)
        if (a != b) fail(std::string("join: key dtype mismatch ") + dtype_name(a) + " vs " + dtype_name(b));
    }
}

// Everything a probe launch needs; keeps the buffers it points into alive.
struct ProbePlan {
    ProbeArgs args;
    Batch keys;  // generic mode: the probe batch with plain key columns
    Buf hash;
    Column key0;  // inline mode: the (possibly view) key column
};
static void make_probe_plan(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols, ProbePlan &pp) {
    ProbeArgs &a = pp.args;
    memset(&a, 0, sizeof(a));
    a.table = as<JoinBucket>(pt->slots);
    a.sorted_rows = as<int64_t>(pt->sorted_rows);
    a.bloom = as<unsigned long long>(pt->bloom);
    a.bloom_mask = pt->bloom_mask;
    a.shift = 64 - pt->bits;
    a.dbits = as<unsigned int>(pt->direct_bits);
    a.drows = as<unsigned int>(pt->direct_rows);
    a.dmin = pt->direct_min;
    a.drange = pt->direct_range;
    if (pt->inline_width) {
        const Column &bk = pt->build->cols[pt->key_cols[0]];
        a.bkey = bk.data ? bk.data->ptr : nullptr;
        pp.key0 = probe.cols[key_cols[0]];
        const Column &pk = pp.key0;
        const Column &src = pk.is_view() ? *pk.base : pk;
        require_plain(src);
        a.pkey = src.data ? src.data->ptr : nullptr;
        a.pvalid = src.null_count > 0 ? as<uint8_t>(src.validity) : nullptr;
        a.psel = pk.is_view() ? as<int64_t>(pk.sel) : nullptr;
        if (pt->direct_extra > 0) {  // the other key columns, gathered plain: compared at the hits only
            std::vector<int32_t> rest(key_cols + 1, key_cols + pt->nkeys), brest(pt->key_cols.begin() + 1, pt->key_cols.end());
            std::vector<uint8_t> ne(pt->nulls_equal.begin() + 1, pt->nulls_equal.end());
            pp.keys = materialize_batch(probe, &rest);
            a.bk = make_keyset(*pt->build, brest, ne);
            a.pk = make_keyset(pp.keys, rest, ne);
        }
    } else {
        std::vector<int32_t> pk(key_cols, key_cols + pt->nkeys);
        pp.keys = materialize_batch(probe, &pk);
        pp.hash = hash_rows(pp.keys, key_cols, pt->nkeys, nullptr);
        a.phash = as<uint64_t>(pp.hash);
        a.bk = make_keyset(*pt->build, pt->key_cols, pt->nulls_equal);
        a.pk = make_keyset(pp.keys, pk, pt->nulls_equal);
    }
}

// kernel variants: key width (0 = generic keys) x selection vector x nullable probe key x kept unmatched probe rows
template <int W>
static void launch_count_w(bool sel, bool nulls, bool keep, int grid, cudaStream_t st, const ProbeArgs &a, int64_t n, int64_t nchunks, int64_t *chunk_cnt,
                           uint32_t *bits, uint32_t *stash) {
#define B200_COUNT(S, N, K)                                                                                   \
    do {                                                                                                      \
        if (W != 0 && a.drows && a.bk.nkeys > 0) direct_count_kernel<(W ? W : 8), S, N, K, true><<<grid, PT, 0, st>>>(a, n, nchunks, chunk_cnt, bits, stash); \
        else if (W != 0 && a.drows) direct_count_kernel<(W ? W : 8), S, N, K, false><<<grid, PT, 0, st>>>(a, n, nchunks, chunk_cnt, bits, stash); \
        else probe_count_kernel<W, S, N, K><<<grid, PT, 0, st>>>(a, n, nchunks, chunk_cnt, bits);             \
    } while (0)
    if (sel) {
        if (nulls) { if (keep) { B200_COUNT(true, true, true); } else { B200_COUNT(true, true, false); } }
        else { if (keep) { B200_COUNT(true, false, true); } else { B200_COUNT(true, false, false); } }
    } else {
        if (nulls) { if (keep) { B200_COUNT(false, true, true); } else { B200_COUNT(false, true, false); } }
        else { if (keep) { B200_COUNT(false, false, true); } else { B200_COUNT(false, false, false); } }
    }
#undef B200_COUNT
}
// stash (direct-address tables only): room for one 32-bit build row per probe row, filled densely per chunk
static void launch_count(int width, const ProbeArgs &a, int64_t n, int64_t nchunks, int64_t *chunk_cnt, uint32_t *bits, int keep, uint32_t *stash = nullptr) {
    cudaStream_t st = ctx().stream;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nchunks + PT / 32 - 1) / (PT / 32), (int64_t)ctx().sm_count * (a.drows ? 4 : 3)));
    const bool sel = a.psel != nullptr, nulls = a.pvalid != nullptr;
    KernelTimer kt(a.drows ? "join_probe_count_direct" : "join_probe_count");
    switch (width) {
        case 1: launch_count_w<1>(sel, nulls, keep != 0, grid, st, a, n, nchunks, chunk_cnt, bits, stash); break;
        case 2: launch_count_w<2>(sel, nulls, keep != 0, grid, st, a, n, nchunks, chunk_cnt, bits, stash); break;
        case 4: launch_count_w<4>(sel, nulls, keep != 0, grid, st, a, n, nchunks, chunk_cnt, bits, stash); break;
        case 8: launch_count_w<8>(sel, nulls, keep != 0, grid, st, a, n, nchunks, chunk_cnt, bits, stash); break;
        default: launch_count_w<0>(false, false, keep != 0, grid, st, a, n, nchunks, chunk_cnt, bits, stash); break;
    }
    count_launch();
    B200_CUDA(cudaGetLastError());
}
template <int W>
static void launch_write_w(bool sel, bool nulls, int grid, cudaStream_t st, const ProbeArgs &a, int64_t n, int64_t nchunks, const int64_t *chunk_off,
                           const uint32_t *bits, int64_t *out_probe, int64_t *out_build, unsigned int *matched_bits) {
#define B200_WRITE(S, N) probe_write_kernel<W, S, N><<<grid, PT, 0, st>>>(a, n, nchunks, chunk_off, bits, out_probe, out_build, matched_bits)
    if (sel) { if (nulls) { B200_WRITE(true, true); } else { B200_WRITE(true, false); } }
    else { if (nulls) { B200_WRITE(false, true); } else { B200_WRITE(false, false); } }
#undef B200_WRITE
}
static void launch_write(int width, const ProbeArgs &a, int64_t n, int64_t nchunks, const int64_t *chunk_off, const uint32_t *bits, int64_t *out_probe,
                         int64_t *out_build, unsigned int *matched_bits, const uint32_t *stash = nullptr) {
    cudaStream_t st = ctx().stream;
    if (stash) {
        const int g = (int)std::max<int64_t>(1, std::min<int64_t>((nchunks + PT / 32 - 1) / (PT / 32), (int64_t)ctx().sm_count * 4));
        KernelTimer kt("join_probe_write_direct");
        direct_write_kernel<<<g, PT, 0, st>>>(n, nchunks, chunk_off, bits, stash, out_probe, out_build, matched_bits);
        count_launch();
        B200_CUDA(cudaGetLastError());
        return;
    }
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nchunks + PT / 32 - 1) / (PT / 32), (int64_t)ctx().sm_count * (a.drows ? 4 : 3)));
    const bool sel = a.psel != nullptr, nulls = a.pvalid != nullptr;
    KernelTimer kt(a.drows ? "join_probe_write_direct" : "join_probe_write");
    switch (width) {
        case 1: launch_write_w<1>(sel, nulls, grid, st, a, n, nchunks, chunk_off, bits, out_probe, out_build, matched_bits); break;
        case 2: launch_write_w<2>(sel, nulls, grid, st, a, n, nchunks, chunk_off, bits, out_probe, out_build, matched_bits); break;
        case 4: launch_write_w<4>(sel, nulls, grid, st, a, n, nchunks, chunk_off, bits, out_probe, out_build, matched_bits); break;
        case 8: launch_write_w<8>(sel, nulls, grid, st, a, n, nchunks, chunk_off, bits, out_probe, out_build, matched_bits); break;
        default: launch_write_w<0>(false, false, grid, st, a, n, nchunks, chunk_off, bits, out_probe, out_build, matched_bits); break;
    }
    count_launch();
    B200_CUDA(cudaGetLastError());
}

Batch *probe_exists(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols) {
    require_device();
    check_probe(pt, probe, key_cols);
    const int64_t n = probe.num_rows;
    Column c;
    c.name = "exists";
    c.dtype = B200_BOOL;
    c.length = n;
    c.data = dev_alloc_zero((size_t)((n + 31) / 32 + 1) * 4);
    c.data_bytes = (n + 7) / 8;
    if (n > 0) {
        ProbePlan pp;
        make_probe_plan(pt, probe, key_cols, pp);
        const int64_t nchunks = (n + CHUNK - 1) / CHUNK;
        Buf counts = dev_alloc((size_t)(nchunks + 1) * 8);
        launch_count(pt->inline_width, pp.args, n, nchunks, as<int64_t>(counts), as<uint32_t>(c.data), 0);
    }
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    out->cols.push_back(std::move(c));
    op_done();
    return out.release();
}

// (probe_idx, build_idx) pairs; build_idx = -1 for a kept unmatched probe row
static void probe_pairs(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols, bool keep_unmatched_probe,
                        unsigned int *matched_bits, Buf &out_probe, Buf &out_build, int64_t &total) {
    check_probe(pt, probe, key_cols);
    const int64_t n = probe.num_rows;
    total = 0;
    if (n == 0) {
        out_probe = dev_alloc(8);
        out_build = dev_alloc(8);
        return;
    }
    ProbePlan pp;
    make_probe_plan(pt, probe, key_cols, pp);
    const int64_t nchunks = (n + CHUNK - 1) / CHUNK;
    Buf counts = dev_alloc((size_t)(nchunks + 1) * 8), offs = dev_alloc((size_t)(nchunks + 1) * 8);
    Buf bits = dev_alloc((size_t)((n + 31) / 32 + 1) * 4);  // one "produces output" bit per probe row, pass 0 -> pass 1
    Buf stash = pp.args.drows ? dev_alloc((size_t)nchunks * CHUNK * 4) : Buf();
    launch_count(pt->inline_width, pp.args, n, nchunks, as<int64_t>(counts), as<uint32_t>(bits), keep_unmatched_probe ? 1 : 0, as<uint32_t>(stash));
    exclusive_scan_i64(as<int64_t>(counts), as<int64_t>(offs), nchunks, as<int64_t>(offs) + nchunks);
    total = read_i64(as<int64_t>(offs) + nchunks);
    out_probe = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    out_build = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    if (total > 0)
        launch_write(pt->inline_width, pp.args, n, nchunks, as<int64_t>(offs), as<uint32_t>(bits), as<int64_t>(out_probe), as<int64_t>(out_build),
                     matched_bits, as<uint32_t>(stash));
}

Batch *probe_indices(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols) {
    require_device();
    Buf pi, bi;
    int64_t total = 0;
    probe_pairs(pt, probe, key_cols, false, nullptr, pi, bi, total);
    auto out = std::make_unique<Batch>();
    out->num_rows = total;
    out->cols.push_back(make_u64_column("probe_idx", pi, total));
    out->cols.push_back(make_u64_column("build_idx", bi, total));
    op_done();
    return out.release();
}

b200_join_state *join_state_create(const b200_probe_table *pt) {
    require_device();
    if (!pt || !pt->built) fail("join: probe table is not built");
    auto *s = new b200_join_state;
    s->n = pt->num_rows;
    s->matched = dev_alloc_zero((size_t)((s->n + 31) / 32 + 1) * 4);
    return s;
}
void join_state_free(b200_join_state *st) { delete st; }

static void append_cols(Batch &dst, std::vector<Column> src) {
    for (auto &c : src) dst.cols.push_back(std::move(c));
}

Batch *hash_join_probe(const b200_probe_table *pt, b200_join_state *st, const Batch &probe, const int32_t *key_cols,
                       int32_t join_type, bool build_on_left) {
    require_device();
    if (!pt || !pt->built) fail("join: probe table is not built");
    const bool probe_is_left = !build_on_left;
    if (join_type == B200_JOIN_SEMI || join_type == B200_JOIN_ANTI) {
        // probe side = left side (anti_semi_join.rs); output = probe columns of rows that (do not) find a match
        if (!probe_is_left) fail("semi/anti join with the build side on the left must use probe_exists + finalize on the host");
        std::unique_ptr<Batch> ex(probe_exists(pt, probe, key_cols));
        const int64_t n = probe.num_rows;
        // reuse the filter machinery: bitmap -> indices
        int64_t total = 0;
        Buf idx = bitmap_select(as<uint32_t>(ex->cols[0].data), n, join_type == B200_JOIN_ANTI, &total);
        auto out = std::make_unique<Batch>();
        out->num_rows = total;
        append_cols(*out, make_view_columns(probe, idx, total, false));
        op_done();
        return out.release();
    }
    const bool keep_probe = join_type == B200_JOIN_OUTER || (join_type == B200_JOIN_LEFT && probe_is_left) ||
                            (join_type == B200_JOIN_RIGHT && !probe_is_left);
    const bool keep_build = join_type == B200_JOIN_OUTER || (join_type == B200_JOIN_LEFT && !probe_is_left) ||
                            (join_type == B200_JOIN_RIGHT && probe_is_left);
    if (keep_build && !st) fail("join: this join type needs a b200_join_state to track matched build rows");
    Buf pi, bi;
    int64_t total = 0;
    probe_pairs(pt, probe, key_cols, keep_probe, keep_build ? as<unsigned int>(st->matched) : nullptr, pi, bi, total);
    // late materialisation: both sides become selection vectors over the inputs
    std::vector<Column> pcols = make_view_columns(probe, pi, total, false);
    std::vector<Column> bcols = make_view_columns(*pt->build, bi, total, keep_probe);
    auto out = std::make_unique<Batch>();
    out->num_rows = total;
    if (build_on_left) {
        append_cols(*out, std::move(bcols));
        append_cols(*out, std::move(pcols));
    } else {
        append_cols(*out, std::move(pcols));
        append_cols(*out, std::move(bcols));
    }
    op_done();
    return out.release();
}

Batch *hash_join_finalize(const b200_probe_table *pt, b200_join_state *st, const Batch &probe_like, bool build_on_left) {
    require_device();
    if (!pt || !pt->built || !st) fail("join finalize: need a built probe table and a join state");
    int64_t total = 0;
    Buf idx = bitmap_select(as<uint32_t>(st->matched), st->n, true, &total);
    const Batch build_plain = materialize_batch(*pt->build);
    std::unique_ptr<Batch> bcols(gather_batch(build_plain, as<int64_t>(idx), total, false));
    // null-extended probe side
    Buf neg = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    fill_u8(neg->ptr, 0xFF, (size_t)total * 8);
    std::unique_ptr<Batch> pcols(gather_batch(probe_like, as<int64_t>(neg), total, true));
    auto out = std::make_unique<Batch>();
    out->num_rows = total;
    if (build_on_left) {
        append_cols(*out, std::move(bcols->cols));
        append_cols(*out, std::move(pcols->cols));
    } else {
        append_cols(*out, std::move(pcols->cols));
        append_cols(*out, std::move(bcols->cols));
    }
    op_done();
    return out.release();
}

}  // namespace b200

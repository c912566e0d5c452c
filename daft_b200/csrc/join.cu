// a7/a8: hash join build + probe.
//
// Replaces ProbeTable / ProbeTableBuilder (src/daft-recordbatch/src/probeable/probe_table.rs:14-213:
// a sequential hashbrown insert of (table<<36|row) ids into per-key Vec<u64>) and the row-at-a-time
// GrowableRecordBatch materialisation of src/daft-local-execution/src/join/inner_join.rs:11-85.
//
// Build (all on the device, no sequential step):
//   1. hash_rows over the build keys (the same XXH3 chain the reference uses);
//   2. stable radix sort of the build row ids by the TOP B bits of the hash (B = ceil(log2 rows): 2-4 byte passes
//      instead of 8) -> the rows of a bucket are contiguous and in INSERTION ORDER (the reference's match-list order,
//      probe_table.rs:178-181);
//   3. a direct-address table of 2^B 32-byte entries {start, end, first row, tag}: bucket heads / tails write their own
//      position.  No probing, no CAS: a probe reads exactly one 32-byte sector;
//   4. tables that do not fit L2 also get a blocked Bloom filter (one 64-bit word per key, 3 bits, ~8 bits / build row:
//      it stays L2-resident), so probe rows without a partner never touch the table in DRAM.
// Probe (two passes over tiles of 1024 probe rows, output rows in probe-row order, matches in build insertion order):
//   pass 0 counts the matches of every tile, the tile counts are scanned, pass 1 recomputes the matches and writes the
//   (probe row, build row) pairs at their final positions (block-wide scan inside the tile).  Nothing per probe row is
//   materialised between the passes: no hash column, no per-row counts / offsets / first-match arrays.
//   * single integer-like key without build-side nulls: the key is hashed in registers (XXH3 of its 1/2/4/8 bytes, the
//     value hash_rows would produce) and compared against the tag the bucket carries inline -- one sector per probe;
//     the probe key may be read THROUGH a selection vector (a view column), so a probe chain never gathers its keys;
//   * anything else: hash_rows over the key columns, then every bucket row is verified by value: nulls match only where
//     null_equals_null, floats compare with IEEE == (NaN never matches, nans_equal=false, probe_table.rs:48-51);
//   * every thread keeps four probe rows in flight (keys, then Bloom words, then bucket sectors: independent loads);
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
#include "xxh3.cuh"

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

// One 32-byte sector per bucket: besides its span in `sorted_rows` the entry carries the bucket's FIRST row inline --
// its row id and a tag that is either the key value itself (single integer-like key column without nulls: the
// match is decided without touching the build side) or the row's full hash (a cheap reject before the key compare).
// A probe into a one-row bucket (the common case at <= 1 row per bucket) therefore costs one random sector.
struct __align__(32) JoinBucket {
    long long start;      // first position in `sorted_rows`
    long long end;        // one past the last (start == end == 0: empty bucket)
    long long first_row;  // sorted_rows[start]
    unsigned long long tag;
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
            const int64_t r = sorted_rows[i];
            longlong2 lo, hi;
            lo.x = i;
            lo.y = closed ? j : 0;  // a long run: the tail row stores `end`
            hi.x = r;
            hi.y = (long long)(inline_width ? key_bits(key_data, inline_width, r) : h);
            longlong2 *e = reinterpret_cast<longlong2 *>(&table[b]);
            if (closed) e[0] = lo;
            else table[b].start = i;  // (the tail row owns `end`: no store of ours may overwrite it)
            e[1] = hi;
        } else {
            const bool tail = i == n - 1 || (sorted_hash[i + 1] >> shift) != b;
            if (tail) {
                // distance to the head: only runs longer than BB_WALK + 1 rows are closed here
                int64_t back = 1;
                while (back <= BB_WALK + 1 && i - back >= 0 && (sorted_hash[i - back] >> shift) == b) back++;
                if (back > BB_WALK + 1) table[b].end = i + 1;
            }
        }
    }
}

// Blocked Bloom filter: word = bits 18.. of the hash, three bit positions from its low 18 bits.
__device__ __forceinline__ unsigned long long bloom_need(uint64_t h) {
    return (1ULL << (h & 63)) | (1ULL << ((h >> 6) & 63)) | (1ULL << ((h >> 12) & 63));
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
constexpr int PR = 4;            // probe rows per thread per tile: independent loads in flight
constexpr int PTILE = PT * PR;   // rows per tile
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
    KeySet bk, pk;
};

template <int W>
__device__ __forceinline__ unsigned long long key_w(const void *data, int64_t r) {
    if (W == 1) return ((const uint8_t *)data)[r];
    if (W == 2) return ((const uint16_t *)data)[r];
    if (W == 4) return ((const uint32_t *)data)[r];
    return ((const uint64_t *)data)[r];
}
template <int W>
__device__ __forceinline__ uint64_t hash_w(unsigned long long v) {  // == hash_rows of the key column (hash.cu hash_value<T>, seed 0)
    if (W == 1) return xxh::hash_u8((uint8_t)v, 0);
    if (W == 2) return xxh::hash_u16((uint16_t)v, 0);
    if (W == 4) return xxh::hash_u32((uint32_t)v, 0);
    return xxh::hash_u64(v, 0);
}
__device__ __forceinline__ long long shfl_ll(long long v, int src) {
    const int lo = __shfl_sync(0xffffffffu, (int)(v & 0xffffffffLL), src), hi = __shfl_sync(0xffffffffu, (int)(v >> 32), src);
    return ((long long)hi << 32) | (unsigned int)lo;
}

// thread-serial walk of a short bucket: f(build_row) for every match, in insertion order
template <int W, class F>
__device__ __forceinline__ void walk_bucket(const ProbeArgs &a, const JoinBucket &e, unsigned long long v, uint64_t h, int64_t i, F &&f) {
    if (W) {
        if (e.tag == v) f(e.first_row);
        for (long long t = e.start + 1; t < e.end; t++) {
            const int64_t br = a.sorted_rows[t];
            if (key_w<W>(a.bkey, br) == v) f(br);
        }
    } else {
        if (e.tag == h && keys_match(a.bk, e.first_row, a.pk, i)) f(e.first_row);
        for (long long t = e.start + 1; t < e.end; t++) {
            const int64_t br = a.sorted_rows[t];
            if (keys_match(a.bk, br, a.pk, i)) f(br);
        }
    }
}

// MODE 0: tile_cnt[tile] = matches of the tile (kept unmatched probe rows count 1)
// MODE 1: write (probe row, build row) pairs from tile_off[tile] on
// MODE 2: out_bits = "probe row has a match" bitmap
template <int W, int MODE>
__global__ void __launch_bounds__(PT, 2) probe_kernel(const __grid_constant__ ProbeArgs a, int64_t n, int64_t ntiles, const int64_t *__restrict__ tile_off,
                                                      int64_t *__restrict__ tile_cnt, int64_t *__restrict__ out_probe, int64_t *__restrict__ out_build,
                                                      unsigned int *__restrict__ matched_bits, uint32_t *__restrict__ out_bits, int keep_unmatched) {
    __shared__ long long s_tot[PR][PT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t base = tile * PTILE;
        long long o_tile = 0;
        if (MODE == 1) {
            o_tile = tile_off[tile];
            if (tile_off[tile + 1] == o_tile) continue;  // nothing to write for this tile (uniform for the CTA)
        }
        JoinBucket e[PR];
        unsigned long long v[PR];
        uint64_t h[PR];
        bool hit[PR];
        // ---- keys of all rows, then hashes (+ Bloom words), then bucket sectors
#pragma unroll
        for (int r = 0; r < PR; r++) {
            const int64_t i = base + (int64_t)r * PT + tid;
            hit[r] = i < n;
            v[r] = 0;
            h[r] = 0;
            if (W) {
                if (hit[r]) {
                    const int64_t row = a.psel ? a.psel[i] : i;
                    hit[r] = row >= 0 && (a.pvalid == nullptr || bit_at(a.pvalid, row));
                    if (hit[r]) v[r] = key_w<W>(a.pkey, row);
                }
            } else if (hit[r]) {
                h[r] = a.phash[i];
            }
        }
        if (W) {
#pragma unroll
            for (int r = 0; r < PR; r++) h[r] = hash_w<W>(v[r]);
        }
        if (a.bloom) {
            unsigned long long bw[PR];
#pragma unroll
            for (int r = 0; r < PR; r++) bw[r] = hit[r] ? a.bloom[(h[r] >> 18) & a.bloom_mask] : 0ULL;
#pragma unroll
            for (int r = 0; r < PR; r++) {
                const unsigned long long need = bloom_need(h[r]);
                hit[r] = hit[r] && (bw[r] & need) == need;
            }
        }
#pragma unroll
        for (int r = 0; r < PR; r++) {
            e[r].start = e[r].end = 0;
            if (hit[r]) {
                const longlong2 *q = reinterpret_cast<const longlong2 *>(&a.table[h[r] >> a.shift]);
                const longlong2 lo = q[0], hi = q[1];
                e[r].start = lo.x;
                e[r].end = lo.y;
                e[r].first_row = hi.x;
                e[r].tag = (unsigned long long)hi.y;
            }
        }
        // ---- match counts (and the first match, which settles the common one-match case without a second walk)
        long long cnt[PR], first[PR];
#pragma unroll
        for (int r = 0; r < PR; r++) {
            const int64_t i = base + (int64_t)r * PT + tid;
            cnt[r] = 0;
            first[r] = -1;
            const long long len = e[r].end - e[r].start;
            if (len > 0 && len < LONG_BUCKET) {
                long long c = 0, f = -1;
                walk_bucket<W>(a, e[r], v[r], h[r], i, [&](int64_t br) {
                    if (c == 0) f = br;
                    c++;
                });
                cnt[r] = c;
                first[r] = f;
            }
            // long buckets: the whole warp walks them, one owner lane at a time
            unsigned pending = __ballot_sync(0xffffffffu, len >= LONG_BUCKET);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const long long s = shfl_ll(e[r].start, src), en = shfl_ll(e[r].end, src);
                const unsigned long long vv = (unsigned long long)shfl_ll((long long)v[r], src);
                const int64_t ii = base + (int64_t)r * PT + (warp << 5) + src;
                long long c = 0, f = -1;
                for (long long t0 = s; t0 < en; t0 += 32) {
                    const long long t = t0 + lane;
                    int64_t br = -1;
                    bool mt = false;
                    if (t < en) {
                        br = a.sorted_rows[t];
                        mt = W ? key_w<W>(a.bkey, br) == vv : keys_match(a.bk, br, a.pk, ii);
                    }
                    const unsigned mm = __ballot_sync(0xffffffffu, mt);
                    if (mm) {
                        const long long fb = shfl_ll(br, __ffs(mm) - 1);
                        if (c == 0) f = fb;
                        c += __popc(mm);
                    }
                }
                if (lane == src) {
                    cnt[r] = c;
                    first[r] = f;
                }
            }
            if (keep_unmatched && i < n && cnt[r] == 0) cnt[r] = 1;  // first stays -1: a null-extended row
        }
        if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < PR; r++) {
                const unsigned m = __ballot_sync(0xffffffffu, cnt[r] > 0);
                const int64_t w0 = base + (int64_t)r * PT + (warp << 5);
                if (lane == 0 && w0 < n) out_bits[w0 >> 5] = m;
            }
            continue;
        }
        // ---- block-wide exclusive scan in row order (r major, thread minor)
        long long inc[PR];
#pragma unroll
        for (int r = 0; r < PR; r++) {
            long long x = cnt[r];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long y = shfl_ll(x, max(lane - d, 0));
                if (lane >= d) x += y;
            }
            inc[r] = x;
            if (lane == 31) s_tot[r][warp] = x;
        }
        __syncthreads();
        long long run = 0, off[PR];
#pragma unroll
        for (int r = 0; r < PR; r++) {
            long long before = 0, all = 0;
#pragma unroll
            for (int w = 0; w < PT / 32; w++) {
                const long long t = s_tot[r][w];
                all += t;
                if (w < warp) before += t;
            }
            off[r] = run + before + inc[r] - cnt[r];
            run += all;
        }
        __syncthreads();  // s_tot is reused by the next tile
        if (MODE == 0) {
            if (tid == 0) tile_cnt[tile] = run;
            continue;
        }
        // ---- MODE 1: write the pairs
#pragma unroll
        for (int r = 0; r < PR; r++) {
            const int64_t i = base + (int64_t)r * PT + tid;
            const long long o = o_tile + off[r];
            const long long len = e[r].end - e[r].start;
            if (cnt[r] == 1) {  // one match, or a kept unmatched probe row (first == -1)
                out_probe[o] = i;
                out_build[o] = first[r];
                if (matched_bits && first[r] >= 0) atomicOr(&matched_bits[first[r] >> 5], 1u << (first[r] & 31));
            } else if (cnt[r] > 1 && len < LONG_BUCKET) {
                long long w = 0;
                walk_bucket<W>(a, e[r], v[r], h[r], i, [&](int64_t br) {
                    out_probe[o + w] = i;
                    out_build[o + w] = br;
                    if (matched_bits) atomicOr(&matched_bits[br >> 5], 1u << (br & 31));
                    w++;
                });
            }
            unsigned pending = __ballot_sync(0xffffffffu, cnt[r] > 1 && len >= LONG_BUCKET);
            while (pending) {
                const int src = __ffs(pending) - 1;
                pending &= pending - 1;
                const long long s = shfl_ll(e[r].start, src), en = shfl_ll(e[r].end, src);
                const unsigned long long vv = (unsigned long long)shfl_ll((long long)v[r], src);
                const int64_t ii = base + (int64_t)r * PT + (warp << 5) + src;
                long long oo = shfl_ll(o, src);
                for (long long t0 = s; t0 < en; t0 += 32) {
                    const long long t = t0 + lane;
                    int64_t br = -1;
                    bool mt = false;
                    if (t < en) {
                        br = a.sorted_rows[t];
                        mt = W ? key_w<W>(a.bkey, br) == vv : keys_match(a.bk, br, a.pk, ii);
                    }
                    const unsigned mm = __ballot_sync(0xffffffffu, mt);
                    if (mt) {
                        const long long p = oo + __popc(mm & ((1u << lane) - 1u));
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
        // payload columns may stay views (a join output as build side): only the keys are read here
        pt->build = std::make_unique<Batch>(materialize_batch(*pt->pending[0], &pt->key_cols));
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
    Buf hashes = hash_rows(b, pt->key_cols.data(), pt->nkeys, nullptr);
    Buf rows = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    int bits = 4;
    while (bits < 40 && ((int64_t)1 << bits) < n) bits++;  // at most one row per bucket on average
    pt->bits = bits;
    pt->slots = dev_alloc_zero(((size_t)1 << bits) * sizeof(JoinBucket));
    if (n > 0) {
        fill_iota(as<int64_t>(rows), n);
        const int sort_bits = (bits + 7) / 8 * 8;  // whole byte passes over the top of the hash
        radix_sort_pairs(hashes, rows, n, 64 - sort_bits, 64, false);
        // inline key tags: one integer-like key column without nulls (floats keep IEEE ==, strings need their bytes)
        const Column &k0 = b.cols[pt->key_cols[0]];
        const bool int_like = k0.dtype == B200_I8 || k0.dtype == B200_I16 || k0.dtype == B200_I32 || k0.dtype == B200_I64 || k0.dtype == B200_U8 ||
                              k0.dtype == B200_U16 || k0.dtype == B200_U32 || k0.dtype == B200_U64 || k0.dtype == B200_DATE32 || k0.dtype == B200_TIMESTAMP;
        pt->inline_width = (pt->nkeys == 1 && int_like && k0.null_count == 0) ? dtype_width(k0.dtype) : 0;
        bucket_bounds<<<jgrid(n), 256, 0, st>>>(as<uint64_t>(hashes), as<int64_t>(rows), n, 64 - bits, k0.data ? k0.data->ptr : nullptr,
                                                pt->inline_width, as<JoinBucket>(pt->slots));
        count_launch();
        // a table that cannot stay in L2 (126 MB) gets the Bloom filter: ~8 bits per build row
        const char *env = getenv("B200_JOIN_BLOOM");
        const bool want = env ? *env == '1' : (((size_t)1 << bits) * sizeof(JoinBucket) > ((size_t)48 << 20));
        if (want) {
            uint64_t words = 1024;
            while ((int64_t)words * 8 < n) words <<= 1;
            pt->bloom = dev_alloc_zero(words * 8);
            pt->bloom_mask = words - 1;
            bloom_build<<<jgrid(n), 256, 0, st>>>(as<uint64_t>(hashes), n, as<unsigned long long>(pt->bloom), pt->bloom_mask);
            count_launch();
        }
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
    } else {
        std::vector<int32_t> pk(key_cols, key_cols + pt->nkeys);
        pp.keys = materialize_batch(probe, &pk);
        pp.hash = hash_rows(pp.keys, key_cols, pt->nkeys, nullptr);
        a.phash = as<uint64_t>(pp.hash);
        a.bk = make_keyset(*pt->build, pt->key_cols, pt->nulls_equal);
        a.pk = make_keyset(pp.keys, pk, pt->nulls_equal);
    }
}

template <int MODE>
static void launch_probe(int width, const ProbeArgs &a, int64_t n, int64_t ntiles, const int64_t *tile_off, int64_t *tile_cnt, int64_t *out_probe,
                         int64_t *out_build, unsigned int *matched_bits, uint32_t *out_bits, int keep) {
    cudaStream_t st = ctx().stream;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx().sm_count * 2));
    switch (width) {
        case 1: probe_kernel<1, MODE><<<grid, PT, 0, st>>>(a, n, ntiles, tile_off, tile_cnt, out_probe, out_build, matched_bits, out_bits, keep); break;
        case 2: probe_kernel<2, MODE><<<grid, PT, 0, st>>>(a, n, ntiles, tile_off, tile_cnt, out_probe, out_build, matched_bits, out_bits, keep); break;
        case 4: probe_kernel<4, MODE><<<grid, PT, 0, st>>>(a, n, ntiles, tile_off, tile_cnt, out_probe, out_build, matched_bits, out_bits, keep); break;
        case 8: probe_kernel<8, MODE><<<grid, PT, 0, st>>>(a, n, ntiles, tile_off, tile_cnt, out_probe, out_build, matched_bits, out_bits, keep); break;
        default: probe_kernel<0, MODE><<<grid, PT, 0, st>>>(a, n, ntiles, tile_off, tile_cnt, out_probe, out_build, matched_bits, out_bits, keep); break;
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
    c.data = dev_alloc((size_t)((n + 31) / 32 + 1) * 4);
    c.data_bytes = (n + 7) / 8;
    if (n > 0) {
        ProbePlan pp;
        make_probe_plan(pt, probe, key_cols, pp);
        launch_probe<2>(pt->inline_width, pp.args, n, (n + PTILE - 1) / PTILE, nullptr, nullptr, nullptr, nullptr, nullptr, as<uint32_t>(c.data), 0);
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
    const int64_t ntiles = (n + PTILE - 1) / PTILE;
    Buf counts = dev_alloc((size_t)(ntiles + 1) * 8), offs = dev_alloc((size_t)(ntiles + 1) * 8);
    const int keep = keep_unmatched_probe ? 1 : 0;
    launch_probe<0>(pt->inline_width, pp.args, n, ntiles, nullptr, as<int64_t>(counts), nullptr, nullptr, nullptr, nullptr, keep);
    exclusive_scan_i64(as<int64_t>(counts), as<int64_t>(offs), ntiles, as<int64_t>(offs) + ntiles);
    total = read_i64(as<int64_t>(offs) + ntiles);
    out_probe = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    out_build = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    if (total > 0)
        launch_probe<1>(pt->inline_width, pp.args, n, ntiles, as<int64_t>(offs), nullptr, as<int64_t>(out_probe), as<int64_t>(out_build), matched_bits,
                        nullptr, keep);
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

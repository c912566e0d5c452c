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
//   3. a direct-address table of 2^B {start, end} entries: bucket heads / tails write their own position.  No
//      probing, no CAS, no stored hashes: a probe reads exactly one 16-byte entry.
// Probe (count -> scan -> write, output rows in probe-row order): bucket = top B bits of the probe row's hash; every
//   row of the bucket is verified against the probe row's key values: nulls match only where null_equals_null, floats
//   compare with IEEE == (NaN never matches, nans_equal=false, probe_table.rs:48-51).  The count pass remembers each
//   probe row's first match, so the write pass repeats the bucket walk only for rows with several matches.
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
    std::unique_ptr<b200::Batch> build;                 // concatenated build side (all columns)
    std::vector<int32_t> key_cols;
    b200::Buf sorted_rows;                              // build row ids sorted by hash (stable)
    b200::Buf slots;                                    // JoinBucket[1 << bits]
    int bits = 0;                                       // bucket = hash >> (64 - bits)
    int inline_width = 0;                               // > 0: tags are key values of that byte width (see JoinBucket)
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

// Walk one bucket for probe row i: calls f(build_row) for every match, in insertion order.
template <class F>
__device__ __forceinline__ void bucket_matches(const JoinBucket &e, uint64_t h, const KeySet &bk, const KeySet &pk, int64_t i,
                                               const int64_t *__restrict__ sorted_rows, int inline_width, F &&f) {
    if (e.end == e.start) return;
    if (inline_width) {
        if (pk.k[0].validity != nullptr && !bit_at(pk.k[0].validity, i)) return;  // build side has no nulls in this mode
        const unsigned long long v = key_bits(pk.k[0].data, inline_width, i);
        if (e.tag == v) f(e.first_row);
        for (long long t = e.start + 1; t < e.end; t++) {
            const int64_t br = sorted_rows[t];
            if (key_bits(bk.k[0].data, inline_width, br) == v) f(br);
        }
    } else {
        if (e.tag == h && keys_match(bk, e.first_row, pk, i)) f(e.first_row);
        for (long long t = e.start + 1; t < e.end; t++) {
            const int64_t br = sorted_rows[t];
            if (keys_match(bk, br, pk, i)) f(br);
        }
    }
}

// ---- probe -----------------------------------------------------------------------------------------
// counts[i] = matches of probe row i (1 for a kept unmatched row), first[i] = its first matching build row or -1
__global__ void __launch_bounds__(256) probe_count_kernel(const uint64_t *__restrict__ probe_hash, int64_t n_probe, KeySet bk, KeySet pk,
                                                           const JoinBucket *__restrict__ table, int shift,
                                                           const int64_t *__restrict__ sorted_rows, int inline_width, int keep_unmatched_probe,
                                                           int64_t *__restrict__ counts, int64_t *__restrict__ first) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_probe; i += stride) {
        const uint64_t h = probe_hash[i];
        const JoinBucket e = table[h >> shift];
        int64_t nmatch = 0, f = -1;
        bucket_matches(e, h, bk, pk, i, sorted_rows, inline_width, [&](int64_t br) {
            if (nmatch == 0) f = br;
            nmatch++;
        });
        counts[i] = (nmatch == 0 && keep_unmatched_probe) ? 1 : nmatch;
        first[i] = f;
    }
}
__global__ void __launch_bounds__(256) probe_write_kernel(const uint64_t *__restrict__ probe_hash, int64_t n_probe, KeySet bk, KeySet pk,
                                                           const JoinBucket *__restrict__ table, int shift,
                                                           const int64_t *__restrict__ sorted_rows, int inline_width,
                                                           const int64_t *__restrict__ out_off,
                                                           const int64_t *__restrict__ first, int64_t *__restrict__ out_probe,
                                                           int64_t *__restrict__ out_build, unsigned int *__restrict__ matched_bits) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_probe; i += stride) {
        const int64_t o = out_off[i], cnt = out_off[i + 1] - o;
        if (cnt == 0) continue;
        const int64_t f = first[i];
        if (cnt == 1) {  // one match, or a kept unmatched probe row (f == -1)
            out_probe[o] = i;
            out_build[o] = f;
            if (matched_bits && f >= 0) atomicOr(&matched_bits[f >> 5], 1u << (f & 31));
            continue;
        }
        const uint64_t h = probe_hash[i];
        const JoinBucket e = table[h >> shift];
        int64_t w = 0;
        bucket_matches(e, h, bk, pk, i, sorted_rows, inline_width, [&](int64_t br) {
            out_probe[o + w] = i;
            out_build[o + w] = br;
            if (matched_bits) atomicOr(&matched_bits[br >> 5], 1u << (br & 31));
            w++;
        });
    }
}

__global__ void exists_kernel(const uint64_t *__restrict__ probe_hash, int64_t n_probe, KeySet bk, KeySet pk,
                              const JoinBucket *__restrict__ table, int shift, const int64_t *__restrict__ sorted_rows, int inline_width,
                              uint32_t *__restrict__ out_bits) {
    const int64_t n_round = (n_probe + 31) & ~(int64_t)31;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_round; i += stride) {
        bool found = false;
        if (i < n_probe) {
            const uint64_t h = probe_hash[i];
            const JoinBucket e = table[h >> shift];
            bucket_matches(e, h, bk, pk, i, sorted_rows, inline_width, [&](int64_t) { found = true; });
        }
        const unsigned m = __ballot_sync(0xffffffffu, found);
        if ((threadIdx.x & 31) == 0) out_bits[i >> 5] = m;
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
    if (pt->pending.size() == 1) {
        pt->build = std::move(pt->pending[0]);
    } else {
        auto b = std::make_unique<Batch>();
        const size_t nc = pt->pending[0]->cols.size();
        for (auto &p : pt->pending) {
            if (p->cols.size() != nc) fail("join: build tables have different schemas");
            b->num_rows += p->num_rows;
        }
        for (size_t c = 0; c < nc; c++) {
            std::vector<const Column *> cs;
            for (auto &p : pt->pending) cs.push_back(&p->cols[c]);
            b->cols.push_back(concat_columns(cs));
        }
        pt->build = std::move(b);
    }
    pt->key_cols = pt->pending_keys[0];
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

Batch *probe_exists(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols) {
    require_device();
    check_probe(pt, probe, key_cols);
    const int64_t n = probe.num_rows;
    std::vector<int32_t> pk(key_cols, key_cols + pt->nkeys);
    Buf h = hash_rows(probe, key_cols, pt->nkeys, nullptr);
    Column c;
    c.name = "exists";
    c.dtype = B200_BOOL;
    c.length = n;
    c.data = dev_alloc((size_t)((n + 31) / 32 + 1) * 4);
    c.data_bytes = (n + 7) / 8;
    if (n > 0) {
        exists_kernel<<<jgrid(n), 256, 0, ctx().stream>>>(as<uint64_t>(h), n, make_keyset(*pt->build, pt->key_cols, pt->nulls_equal),
                                                           make_keyset(probe, pk, pt->nulls_equal), as<JoinBucket>(pt->slots),
                                                           64 - pt->bits, as<int64_t>(pt->sorted_rows), pt->inline_width, as<uint32_t>(c.data));
        count_launch();
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
    cudaStream_t st = ctx().stream;
    const int64_t n = probe.num_rows;
    std::vector<int32_t> pk(key_cols, key_cols + pt->nkeys);
    total = 0;
    if (n == 0) {
        out_probe = dev_alloc(8);
        out_build = dev_alloc(8);
        return;
    }
    Buf h = hash_rows(probe, key_cols, pt->nkeys, nullptr);
    const KeySet bk = make_keyset(*pt->build, pt->key_cols, pt->nulls_equal), pks = make_keyset(probe, pk, pt->nulls_equal);
    Buf counts = dev_alloc((size_t)(n + 1) * 8), offs = dev_alloc((size_t)(n + 1) * 8), first = dev_alloc((size_t)n * 8);
    probe_count_kernel<<<jgrid(n), 256, 0, st>>>(as<uint64_t>(h), n, bk, pks, as<JoinBucket>(pt->slots), 64 - pt->bits,
                                                 as<int64_t>(pt->sorted_rows), pt->inline_width, keep_unmatched_probe ? 1 : 0, as<int64_t>(counts),
                                                 as<int64_t>(first));
    count_launch();
    exclusive_scan_i64(as<int64_t>(counts), as<int64_t>(offs), n, as<int64_t>(offs) + n);
    total = read_i64(as<int64_t>(offs) + n);
    out_probe = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    out_build = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    if (total > 0) {
        probe_write_kernel<<<jgrid(n), 256, 0, st>>>(as<uint64_t>(h), n, bk, pks, as<JoinBucket>(pt->slots), 64 - pt->bits,
                                                     as<int64_t>(pt->sorted_rows), pt->inline_width, as<int64_t>(offs), as<int64_t>(first),
                                                     as<int64_t>(out_probe),
                                                     as<int64_t>(out_build), matched_bits);
        count_launch();
    }
    B200_CUDA(cudaGetLastError());
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

static void append_cols(Batch &dst, std::unique_ptr<Batch> src) {
    for (auto &c : src->cols) dst.cols.push_back(std::move(c));
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
        Batch *out = gather_batch(probe, as<int64_t>(idx), total, false);
        op_done();
        return out;
    }
    const bool keep_probe = join_type == B200_JOIN_OUTER || (join_type == B200_JOIN_LEFT && probe_is_left) ||
                            (join_type == B200_JOIN_RIGHT && !probe_is_left);
    const bool keep_build = join_type == B200_JOIN_OUTER || (join_type == B200_JOIN_LEFT && !probe_is_left) ||
                            (join_type == B200_JOIN_RIGHT && probe_is_left);
    if (keep_build && !st) fail("join: this join type needs a b200_join_state to track matched build rows");
    Buf pi, bi;
    int64_t total = 0;
    probe_pairs(pt, probe, key_cols, keep_probe, keep_build ? as<unsigned int>(st->matched) : nullptr, pi, bi, total);
    std::unique_ptr<Batch> pcols(gather_batch(probe, as<int64_t>(pi), total, false));
    std::unique_ptr<Batch> bcols(gather_batch(*pt->build, as<int64_t>(bi), total, keep_probe));
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
    std::unique_ptr<Batch> bcols(gather_batch(*pt->build, as<int64_t>(idx), total, false));
    // null-extended probe side
    Buf neg = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    fill_u8(neg->ptr, 0xFF, (size_t)total * 8);
    std::unique_ptr<Batch> pcols(gather_batch(probe_like, as<int64_t>(neg), total, true));
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

}  // namespace b200

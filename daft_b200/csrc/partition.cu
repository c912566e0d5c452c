// a3: hash partition (RecordBatch::partition_by_hash + partition_by_index,
// src/daft-recordbatch/src/ops/partition.rs:12-70) and a13: dedup (ops/agg.rs:342-353).
//
// target = hash_rows(keys) % P with u64 arithmetic; rows keep their input order inside a
// partition.  Implemented as a stable radix partition of the row ids on the partition id (one
// 8-bit pass for P <= 256, two for P <= 65536) followed by one gather per output column.
#include <algorithm>

#include "common.h"
#include "jit.h"
#include "kernels.cuh"
#include "ops.h"
#include "xxh3.cuh"

namespace b200 {

__global__ void partition_ids(const uint64_t *__restrict__ hashes, int64_t n, uint64_t P, uint64_t *__restrict__ pid,
                              int64_t *__restrict__ rowid, unsigned long long *__restrict__ counts) {
    extern __shared__ unsigned int s_cnt[];  // P entries when P <= 4096
    const bool use_smem = P <= 4096;
    if (use_smem) {
        for (int i = threadIdx.x; i < (int)P; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t p = hashes[i] % P;
        pid[i] = p;
        rowid[i] = i;
        if (use_smem) atomicAdd(&s_cnt[p], 1u);
        else atomicAdd(&counts[p], 1ULL);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < (int)P; i += blockDim.x)
            if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
    }
}

Buf partition_permutation(const uint64_t *hashes, int64_t n, int32_t P, std::vector<int64_t> &offsets) {
    cudaStream_t st = ctx().stream;
    offsets.assign((size_t)P + 1, 0);
    Buf rowid = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    if (n == 0) return rowid;
    Buf pid = dev_alloc((size_t)n * 8), counts = dev_alloc_zero((size_t)P * 8);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8));
    const size_t smem = P <= 4096 ? (size_t)P * 4 : 0;
    partition_ids<<<grid, 256, smem, st>>>(hashes, n, (uint64_t)P, as<uint64_t>(pid), as<int64_t>(rowid), as<unsigned long long>(counts));
    count_launch();
    std::vector<unsigned long long> h((size_t)P);
    B200_CUDA(cudaMemcpyAsync(h.data(), counts->ptr, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    int bits = 8;
    while ((1LL << bits) < P) bits += 8;
    radix_sort_pairs(pid, rowid, n, 0, bits, false);
    B200_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < P; p++) offsets[p + 1] = offsets[p] + (int64_t)h[p];
    return rowid;
}

// ---- direct scatter for P <= 256 ----------------------------------------------------------------------------------
// Two passes over the hashes instead of (partition id, row id) pairs + a radix pass + one gather per column:
// pass 1 counts rows per (partition, CTA chunk); after one exclusive scan, pass 2 recomputes hash % P, ranks the row
// inside its 256-row step with warp match_any (stable: earlier rows first) and copies every fixed-width never-null
// column straight to its final position in a buffer that holds all partitions back to back (the per-partition
// columns are views into it).  Other columns (nullable, Boolean, strings) are gathered through the row-id list the
// same pass emits.  Traffic per row: 8 B hash twice + payload once in, once out.
constexpr int PS_THREADS = 256, PS_WARPS = PS_THREADS / 32, PS_MAXP = 256, PS_MAXCOLS = 24;
struct PartCol {
    const void *src;
    void *out;
    int width;
};
struct PartCols {
    PartCol c[PS_MAXCOLS];
};

// hash % P without the 64-bit division (about a hundred instructions on the SM): Lemire's fastmod with M = ceil(2^128 / P),
// a % P = ((M * a mod 2^128) * P) >> 128 -- exact for every 64-bit a and P.
struct PartKey {
    const uint64_t *hashes;
    const void *key;
    unsigned long long m_lo, m_hi, P;
    unsigned long long pow2_mask;  // P - 1 when P is a power of two (then hash % P is a mask), else ~0
};
__device__ __forceinline__ uint32_t fastmod_u64(uint64_t a, const PartKey &k) {
    if (k.pow2_mask != ~0ULL) return (uint32_t)(a & k.pow2_mask);
    const uint64_t l_lo = k.m_lo * a;
    const uint64_t l_hi = __umul64hi(k.m_lo, a) + k.m_hi * a;
    const uint64_t t = __umul64hi(l_lo, k.P);
    const uint64_t u_lo = l_hi * k.P, u_hi = __umul64hi(l_hi, k.P);
    return (uint32_t)(u_hi + ((u_lo + t < u_lo) ? 1ULL : 0ULL));
}
// W = 0: the u64 per row is given; W = 4 / 8: XXH3-64 of the key value, computed here (xxh3.cuh hash_u32 / hash_u64 are what
// hash_rows runs for these dtypes, seed 0)
template <int W>
__device__ __forceinline__ uint32_t part_of(const PartKey &k, int64_t i) {
    uint64_t h;
    if (W == 0) h = __ldg(k.hashes + i);
    else if (W == 4) h = xxh::hash_u32(__ldg((const uint32_t *)k.key + i), 0);
    else h = xxh::hash_u64(__ldg((const unsigned long long *)k.key + i), 0);
    return fastmod_u64(h, k);
}
static PartKey make_part_key(const PartSrc &src, int32_t P) {
    PartKey k;
    k.hashes = src.hashes;
    k.key = src.key;
    k.P = (unsigned long long)P;
    const unsigned __int128 M = (~(unsigned __int128)0) / (unsigned __int128)P + 1;  // P = 1: wraps to 0, and a % 1 = 0
    k.m_lo = (unsigned long long)M;
    k.m_hi = (unsigned long long)(M >> 64);
    k.pow2_mask = (P & (P - 1)) == 0 ? (unsigned long long)(P - 1) : ~0ULL;
    return k;
}

constexpr int PH_ROWS = 4;  // rows per thread and step of the histogram pass: four independent loads in flight
// The pass also leaves every row's partition as ONE byte (P <= 256): the scatter pass reads that instead of hashing again.
template <int W>
__global__ void __launch_bounds__(PS_THREADS) partition_hist(const __grid_constant__ PartKey key, int64_t n, int64_t chunk,
                                                             int64_t *__restrict__ hist /* [partition][block] */, uint8_t *__restrict__ pid) {
    __shared__ unsigned int s_cnt[PS_MAXP];
    for (int i = threadIdx.x; i < PS_MAXP; i += PS_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const int64_t begin = (int64_t)blockIdx.x * chunk, end = min(begin + chunk, n);
    // warp-aggregated: with a handful of partitions every lane would hit the same few shared-memory words
    const int lane = threadIdx.x & 31;
    for (int64_t i0 = begin; i0 < end; i0 += PS_THREADS * PH_ROWS) {
        uint32_t d[PH_ROWS];
#pragma unroll
        for (int r = 0; r < PH_ROWS; r++) {
            const int64_t i = i0 + r * PS_THREADS + threadIdx.x;
            d[r] = i < end ? part_of<W>(key, i) : 0xFFFFFFFFu;
            if (i < end) pid[i] = (uint8_t)d[r];
        }
#pragma unroll
        for (int r = 0; r < PH_ROWS; r++) {
            const unsigned peers = __match_any_sync(0xffffffffu, d[r]);
            if (d[r] != 0xFFFFFFFFu && (peers & ((1u << lane) - 1u)) == 0) atomicAdd(&s_cnt[d[r]], (unsigned)__popc(peers));
        }
    }
    __syncthreads();
    const int Pn = (int)key.P;
    for (int i = threadIdx.x; i < Pn; i += PS_THREADS) hist[(int64_t)i * gridDim.x + blockIdx.x] = s_cnt[i];
}

// PEER: partition d is not written behind partition d - 1 in a local buffer but straight into rank d's receive buffer
// (peer memory over NVLink): column c of the row at position `pos` of partition d goes to dest.out[c][d] + (pos -
// dest.start[d]) * width, where out[c][d] already points at this sender's slot in rank d's buffer.
constexpr int PS_MAXPEERS = 8, PS_PEERCOLS = 16;
struct PeerDest {
    uint8_t *out[PS_PEERCOLS][PS_MAXPEERS];
    long long start[PS_MAXPEERS];
};
template <bool PEER>
__global__ void __launch_bounds__(PS_THREADS) partition_scatter(const uint8_t *__restrict__ pid, int Pn, int64_t n, int64_t chunk,
                                                                const int64_t *__restrict__ base /* scanned hist */,
                                                                const __grid_constant__ PartCols cols, int ncols, int64_t *__restrict__ rowid_out,
                                                                const PeerDest *__restrict__ dest) {
    __shared__ int64_t s_base[PS_MAXP];                 // next output position per partition for this CTA
    __shared__ unsigned int s_wcnt[PS_WARPS][PS_MAXP];  // per-warp counts -> exclusive prefix over warps
    __shared__ unsigned int s_tile[PS_MAXP];
    for (int i = threadIdx.x; i < Pn; i += PS_THREADS) s_base[i] = base[(int64_t)i * gridDim.x + blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t begin = (int64_t)blockIdx.x * chunk, end = min(begin + chunk, n);
    for (int64_t t0 = begin; t0 < end; t0 += PS_THREADS) {
        for (int i = threadIdx.x; i < PS_WARPS * Pn; i += PS_THREADS) s_wcnt[i / Pn][i % Pn] = 0;
        __syncthreads();
        const int64_t i = t0 + threadIdx.x;
        const bool valid = i < end;
        uint32_t d = 0xFFFFFFFFu;  // out-of-range lanes only match each other
        if (valid) d = pid[i];
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));  // stable: lower lanes (earlier rows) first
        if (valid && rank == 0) s_wcnt[warp][d] = __popc(peers);
        __syncthreads();
        for (int dd = threadIdx.x; dd < Pn; dd += PS_THREADS) {
            unsigned int run = 0;
#pragma unroll
            for (int w = 0; w < PS_WARPS; w++) {
                const unsigned int c = s_wcnt[w][dd];
                s_wcnt[w][dd] = run;
                run += c;
            }
            s_tile[dd] = run;
        }
        __syncthreads();
        if (valid) {
            const int64_t pos = s_base[d] + (int64_t)s_wcnt[warp][d] + rank;
            if (rowid_out) rowid_out[pos] = i;
#pragma unroll 4
            for (int c = 0; c < ncols; c++) {
                const PartCol &pc = cols.c[c];
                void *o = pc.out;
                int64_t at = pos;
                if (PEER) {
                    o = dest->out[c][d];
                    at = pos - dest->start[d];
                }
                switch (pc.width) {
                    case 1: ((uint8_t *)o)[at] = ((const uint8_t *)pc.src)[i]; break;
                    case 2: ((uint16_t *)o)[at] = ((const uint16_t *)pc.src)[i]; break;
                    case 4: ((uint32_t *)o)[at] = ((const uint32_t *)pc.src)[i]; break;
                    default: ((uint64_t *)o)[at] = ((const uint64_t *)pc.src)[i]; break;
                }
            }
        }
        __syncthreads();
        for (int dd = threadIdx.x; dd < Pn; dd += PS_THREADS) s_base[dd] += s_tile[dd];
    }
    if (PEER) __threadfence_system();  // the rows must be visible to the peer GPUs before this rank signals completion
}


// ---- scatter staged through shared memory (local partitions and peer stores) -------------------------------------------------------------------------
// The plain scatter above stores each row where it belongs: with W destinations a warp's store touches W short runs.  In
// local HBM the L2 merges them; over NVLink every run is its own small write packet.  Here a CTA orders a tile of 1024 rows
// by destination in shared memory first (stable: tile order is row order), so each destination receives ONE contiguous run
// per column and tile -- 128 rows at W = 8, a 1 KB burst of 8-byte values -- written by consecutive lanes.
constexpr int PP_ITEMS = 4, PP_TILE = PS_THREADS * PP_ITEMS;
template <bool PEER>
__global__ void __launch_bounds__(PS_THREADS) partition_scatter_staged(const uint8_t *__restrict__ pid, int Pn, int64_t n, int64_t chunk,
                                                                      const int64_t *__restrict__ base, const __grid_constant__ PartCols cols,
                                                                      int ncols, int64_t *__restrict__ rowid_out, const PeerDest *__restrict__ dest) {
    static_assert(PS_MAXP == PS_THREADS, "one thread per partition in the prefix step");
    __shared__ int64_t s_base[PS_MAXP];
    __shared__ unsigned int s_wcnt[PS_WARPS][PS_MAXP];
    __shared__ unsigned int s_tile[PS_MAXP], s_toff[PS_MAXP], s_wtot[PS_WARPS];
    __shared__ unsigned long long s_val[PP_TILE];
    __shared__ uint8_t s_dof[PP_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_base[threadIdx.x] = threadIdx.x < Pn ? base[(int64_t)threadIdx.x * gridDim.x + blockIdx.x] : 0;
    const int64_t begin = (int64_t)blockIdx.x * chunk, end = min(begin + chunk, n);
    for (int64_t t0 = begin; t0 < end; t0 += PP_TILE) {
        for (int i = lane; i < Pn; i += 32) s_wcnt[warp][i] = 0;
        __syncwarp();
        const int64_t wbase = t0 + (int64_t)warp * 32 * PP_ITEMS + lane;  // (warp, item, lane) order is row order
        uint32_t d[PP_ITEMS];
        unsigned short rank[PP_ITEMS];
#pragma unroll
        for (int j = 0; j < PP_ITEMS; j++) {
            const int64_t i = wbase + j * 32;
            const bool valid = i < end;
            d[j] = valid ? pid[i] : 0xFFFFFFFFu;
            const unsigned peers = __match_any_sync(0xffffffffu, d[j]);
            const int lt = __popc(peers & ((1u << lane) - 1u));
            const unsigned int before = valid ? s_wcnt[warp][d[j]] : 0u;
            __syncwarp();
            if (valid && lt == 0) s_wcnt[warp][d[j]] = before + __popc(peers);
            __syncwarp();
            rank[j] = (unsigned short)(before + lt);
        }
        __syncthreads();
        {
            const int dd = threadIdx.x;
            unsigned int run = 0;
            if (dd < Pn) {
#pragma unroll
                for (int w = 0; w < PS_WARPS; w++) {
                    const unsigned int c = s_wcnt[w][dd];
                    s_wcnt[w][dd] = run;
                    run += c;
                }
            }
            s_tile[dd] = run;
            unsigned int x = run;  // exclusive scan over the partitions: where a destination's run starts inside the tile
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_wtot[warp] = x;
            __syncthreads();
            unsigned int wpre = 0;
#pragma unroll
            for (int w = 0; w < PS_WARPS; w++) wpre += w < warp ? s_wtot[w] : 0u;
            s_toff[dd] = wpre + x - run;
        }
        __syncthreads();
        unsigned short lp[PP_ITEMS];
#pragma unroll
        for (int j = 0; j < PP_ITEMS; j++) {
            lp[j] = 0;
            if (wbase + j * 32 < end) {
                lp[j] = (unsigned short)(s_toff[d[j]] + s_wcnt[warp][d[j]] + rank[j]);
                s_dof[lp[j]] = (uint8_t)d[j];
                if (rowid_out) rowid_out[s_base[d[j]] + (int64_t)(lp[j] - s_toff[d[j]])] = wbase + j * 32;
            }
        }
        const int cnt = (int)min((int64_t)PP_TILE, end - t0);
        for (int c = 0; c < ncols; c++) {
            const PartCol &pc = cols.c[c];
            __syncthreads();  // s_dof complete / the previous column's s_val consumed
#pragma unroll
            for (int j = 0; j < PP_ITEMS; j++) {
                const int64_t i = wbase + j * 32;
                if (i < end) {
                    unsigned long long v;
                    switch (pc.width) {
                        case 1: v = ((const uint8_t *)pc.src)[i]; break;
                        case 2: v = ((const uint16_t *)pc.src)[i]; break;
                        case 4: v = ((const uint32_t *)pc.src)[i]; break;
                        default: v = ((const unsigned long long *)pc.src)[i]; break;
                    }
                    s_val[lp[j]] = v;
                }
            }
            __syncthreads();
            for (int e = threadIdx.x; e < cnt; e += PS_THREADS) {
                const int dd = s_dof[e];
                int64_t at = s_base[dd] + (int64_t)(e - (int)s_toff[dd]);
                uint8_t *o = (uint8_t *)pc.out;
                if (PEER) {
                    at -= dest->start[dd];
                    o = dest->out[c][dd];
                }
                const unsigned long long v = s_val[e];
                switch (pc.width) {
                    case 1: ((uint8_t *)o)[at] = (uint8_t)v; break;
                    case 2: ((uint16_t *)o)[at] = (uint16_t)v; break;
                    case 4: ((uint32_t *)o)[at] = (uint32_t)v; break;
                    default: ((unsigned long long *)o)[at] = v; break;
                }
            }
        }
        __syncthreads();
        s_base[threadIdx.x] += s_tile[threadIdx.x];
    }
    if (PEER) __threadfence_system();  // the rows must be visible to the peer GPUs before this rank signals completion
}

static void launch_hist(const PartSrc &src, int32_t P, int64_t n, int64_t chunk, int64_t nblocks, int64_t *hist, uint8_t *pid) {
    const PartKey k = make_part_key(src, P);
    cudaStream_t st = ctx().stream;
    if (src.key && src.width == 8) partition_hist<8><<<(unsigned)nblocks, PS_THREADS, 0, st>>>(k, n, chunk, hist, pid);
    else if (src.key && src.width == 4) partition_hist<4><<<(unsigned)nblocks, PS_THREADS, 0, st>>>(k, n, chunk, hist, pid);
    else partition_hist<0><<<(unsigned)nblocks, PS_THREADS, 0, st>>>(k, n, chunk, hist, pid);
    count_launch();
}
template <bool PEER>
static void launch_scatter(const uint8_t *pid, int32_t P, int64_t n, int64_t chunk, int64_t nblocks, const int64_t *base, const PartCols &pcs, int npc,
                           int64_t *rowid, const PeerDest *dest) {
    partition_scatter<PEER><<<(unsigned)nblocks, PS_THREADS, 0, ctx().stream>>>(pid, (int)P, n, chunk, base, pcs, npc, rowid, dest);
    count_launch();
}
// One never-null integer key of 4 or 8 bytes: its XXH3 is a few multiplies, cheaper than writing and re-reading a hash column
PartSrc part_src_for_keys(const Batch &b, const int32_t *key_cols, int32_t nkeys, bool allow_fused) {
    PartSrc src;
    const char *env = getenv("B200_PARTITION_FUSED");
    if (allow_fused && nkeys == 1 && !(env && *env == '0')) {
        if (key_cols[0] < 0 || key_cols[0] >= (int32_t)b.cols.size()) fail("partition: key column index out of range");
        const Column &k = b.cols[key_cols[0]];
        const bool int_like = is_integer(k.dtype) || k.dtype == B200_DATE32 || k.dtype == B200_TIMESTAMP;
        const int w = dtype_width(k.dtype);
        if (int_like && (w == 4 || w == 8) && k.null_count == 0 && !k.is_view() && k.data) {
            src.key = k.data->ptr;
            src.width = w;
            return src;
        }
    }
    src.keep = hash_rows(b, key_cols, nkeys, nullptr);
    src.hashes = as<uint64_t>(src.keep);
    return src;
}

__global__ void pick_strided(const int64_t *__restrict__ src, int64_t stride, int n, int64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[(int64_t)i * stride];
}

// all partitions back to back in ONE batch (rows of partition p at [offsets[p], offsets[p + 1])): what the exchange sends
Batch *partition_staged(const Batch &b, const PartSrc &src, int32_t P, std::vector<int64_t> &offsets, Buf *rowid_out) {
    cudaStream_t st = ctx().stream;
    const int64_t n = b.num_rows;
    KernelTimer kt("partition_direct");
    int64_t nblocks = std::max<int64_t>(1, std::min<int64_t>((n + 2047) / 2048, (int64_t)ctx().sm_count * 16));
    int64_t chunk = (n + nblocks - 1) / nblocks;
    chunk = (chunk + PS_THREADS - 1) / PS_THREADS * PS_THREADS;
    nblocks = (n + chunk - 1) / chunk;
    Buf hist = dev_alloc((size_t)P * nblocks * 8), base = dev_alloc((size_t)(P * nblocks + 1) * 8), starts = dev_alloc((size_t)(P + 1) * 8);
    Buf pid = dev_alloc((size_t)std::max<int64_t>(n, 1));
    launch_hist(src, P, n, chunk, nblocks, as<int64_t>(hist), as<uint8_t>(pid));
    exclusive_scan_i64(as<int64_t>(hist), as<int64_t>(base), (int64_t)P * nblocks, nullptr);
    pick_strided<<<(P + 255) / 256, 256, 0, st>>>(as<int64_t>(base), nblocks, P, as<int64_t>(starts));
    count_launch();
    offsets.assign((size_t)P + 1, 0);
    B200_CUDA(cudaMemcpyAsync(offsets.data(), starts->ptr, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    // columns the kernel places itself vs. columns gathered through the row-id list
    PartCols pcs;
    int npc = 0;
    std::vector<Buf> whole(b.cols.size());
    bool need_rowid = false;
    for (size_t c = 0; c < b.cols.size(); c++) {
        const Column &sc = b.cols[c];
        const int w = dtype_width(sc.dtype);
        if (w && sc.null_count == 0 && npc < PS_MAXCOLS) {
            whole[c] = dev_alloc((size_t)n * w);
            pcs.c[npc++] = PartCol{sc.data->ptr, whole[c]->ptr, w};
        } else if (sc.dtype != B200_NULL) {
            need_rowid = true;
        }
    }
    Buf rowid = need_rowid ? dev_alloc((size_t)n * 8) : Buf();
    {
        // local HBM: the plain scatter wins (60 M rows x 3 columns, P = 8: 1.09 ms against 1.29 ms staged -- the L2 merges the short
        // runs and the staging only adds barriers); over NVLink the staged one does (exchange: 3.27 -> 2.25 ms at 8 GPUs)
        const char *env = getenv("B200_SCATTER_STAGED");
        if (!(env && *env == '1')) {
            launch_scatter<false>(as<uint8_t>(pid), P, n, chunk, nblocks, as<int64_t>(base), pcs, npc, as<int64_t>(rowid), nullptr);
        } else {
            partition_scatter_staged<false><<<(unsigned)nblocks, PS_THREADS, 0, st>>>(as<uint8_t>(pid), (int)P, n, chunk, as<int64_t>(base), pcs, npc,
                                                                                     as<int64_t>(rowid), nullptr);
            count_launch();
        }
    }
    B200_CUDA(cudaStreamSynchronize(st));
    offsets[(size_t)P] = n;
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    for (size_t c = 0; c < b.cols.size(); c++) {
        const Column &sc = b.cols[c];
        if (whole[c]) {
            Column oc;
            oc.name = sc.name;
            oc.dtype = sc.dtype;
            oc.length = n;
            oc.data = whole[c];
            oc.data_bytes = n * dtype_width(sc.dtype);
            out->cols.push_back(std::move(oc));
        } else if (rowid_out) {
            out->cols.push_back(Column());  // the caller gathers these itself, partition by partition (rowid_out)
        } else {
            out->cols.push_back(gather_column(sc, as<int64_t>(rowid), n, false));
        }
    }
    if (rowid_out) *rowid_out = rowid;
    return out.release();
}

// ---- the same scatter in two steps for the peer-store exchange (exchange.cu): counts first (the ranks agree on the
// receive layout), then the scatter straight into the peers' buffers
PartitionPlan partition_count(const PartSrc &src, int64_t n, int32_t P) {
    cudaStream_t st = ctx().stream;
    PartitionPlan pl;
    pl.n = n;
    pl.P = P;
    int64_t nblocks = std::max<int64_t>(1, std::min<int64_t>((n + 2047) / 2048, (int64_t)ctx().sm_count * 16));
    int64_t chunk = (n + nblocks - 1) / nblocks;
    chunk = std::max<int64_t>(PS_THREADS, (chunk + PS_THREADS - 1) / PS_THREADS * PS_THREADS);
    nblocks = std::max<int64_t>(1, (n + chunk - 1) / chunk);
    pl.nblocks = nblocks;
    pl.chunk = chunk;
    Buf hist = dev_alloc((size_t)P * nblocks * 8), starts = dev_alloc((size_t)(P + 1) * 8);
    pl.base = dev_alloc((size_t)(P * nblocks + 1) * 8);
    pl.pid = dev_alloc((size_t)std::max<int64_t>(n, 1));
    launch_hist(src, P, n, chunk, nblocks, as<int64_t>(hist), as<uint8_t>(pl.pid));
    exclusive_scan_i64(as<int64_t>(hist), as<int64_t>(pl.base), (int64_t)P * nblocks, nullptr);
    pick_strided<<<(P + 255) / 256, 256, 0, st>>>(as<int64_t>(pl.base), nblocks, P, as<int64_t>(starts));
    count_launch();
    pl.offsets.assign((size_t)P + 1, 0);
    B200_CUDA(cudaMemcpyAsync(pl.offsets.data(), starts->ptr, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    pl.offsets[(size_t)P] = n;
    return pl;
}
// cols[c] of `b` (fixed width, no nulls) -> dest_ptrs[c][d] (already offset to this sender's slot in rank d's buffer)
void partition_scatter_to_peers(const PartitionPlan &pl, const Batch &b, const std::vector<int> &cols,
                                const std::vector<std::vector<void *>> &dest_ptrs, Buf *rowid_out) {
    cudaStream_t st = ctx().stream;
    if ((int)cols.size() > PS_PEERCOLS || pl.P > PS_MAXPEERS) fail("peer scatter: too many columns / peers");
    PartCols pcs;
    PeerDest pd;
    memset(&pd, 0, sizeof(pd));
    for (size_t k = 0; k < cols.size(); k++) {
        const Column &sc = b.cols[cols[k]];
        pcs.c[k] = PartCol{sc.data->ptr, nullptr, dtype_width(sc.dtype)};
        for (int d = 0; d < pl.P; d++) pd.out[k][d] = (uint8_t *)dest_ptrs[k][d];
    }
    for (int d = 0; d < pl.P; d++) pd.start[d] = pl.offsets[d];
    Buf dpd = dev_alloc(sizeof(PeerDest));
    B200_CUDA(cudaMemcpyAsync(dpd->ptr, &pd, sizeof(PeerDest), cudaMemcpyHostToDevice, st));
    Buf rowid = rowid_out ? dev_alloc((size_t)std::max<int64_t>(pl.n, 1) * 8) : Buf();
    KernelTimer kt("partition_scatter_peer");
    const char *env = getenv("B200_PEER_SCATTER_STAGED");
    if (env && *env == '0') {
        launch_scatter<true>(as<uint8_t>(pl.pid), pl.P, pl.n, pl.chunk, pl.nblocks, as<int64_t>(pl.base), pcs, (int)cols.size(), as<int64_t>(rowid), as<PeerDest>(dpd));
    } else {
        partition_scatter_staged<true><<<(unsigned)pl.nblocks, PS_THREADS, 0, st>>>(as<uint8_t>(pl.pid), (int)pl.P, pl.n, pl.chunk, as<int64_t>(pl.base), pcs,
                                                                                   (int)cols.size(), as<int64_t>(rowid), as<PeerDest>(dpd));
        count_launch();
    }
    B200_CUDA(cudaStreamSynchronize(st));  // `pd` is a stack object; the caller's barrier follows anyway
    if (rowid_out) *rowid_out = rowid;
}

static std::vector<Batch *> partition_direct(const Batch &b, const PartSrc &src, int32_t P) {
    std::vector<int64_t> offsets;
    Buf rowid;
    std::unique_ptr<Batch> staged(partition_staged(b, src, P, offsets, &rowid));
    std::vector<Batch *> out;
    try {
        for (int p = 0; p < P; p++) {
            auto pb = std::make_unique<Batch>();
            pb->num_rows = offsets[p + 1] - offsets[p];
            for (size_t c = 0; c < b.cols.size(); c++) {
                const Column &sc = staged->cols[c];
                if (sc.data || b.cols[c].dtype == B200_NULL) {  // placed by the scatter kernel: a zero-copy slice
                    Column oc = b.cols[c].dtype == B200_NULL ? b.cols[c] : slice_column(sc, offsets[p], offsets[p + 1]);
                    oc.length = pb->num_rows;
                    if (oc.dtype == B200_NULL) oc.null_count = pb->num_rows;
                    pb->cols.push_back(std::move(oc));
                } else {
                    pb->cols.push_back(gather_column(b.cols[c], as<int64_t>(rowid) + offsets[p], pb->num_rows, false));
                }
            }
            out.push_back(pb.release());
        }
        op_done();
    } catch (...) {
        for (auto *x : out) delete x;
        throw;
    }
    return out;
}

// RepartitionSpec::Random (RecordBatch::partition_by_random, src/daft-recordbatch/src/ops/partition.rs:72-86): a target
// per row drawn from a seeded generator.  The reference draws from rand::StdRng (a third-party ChaCha12 stream, absent
// from the reference tree); what its tests pin is conservation and determinism per (seed, input)
// (tests/recordbatch/test_partitioning.py:124-147,199-214), which a counter-based generator keeps: target(i) =
// mix64(mix64(seed) ^ (salt + i)) % P -- independent of how the rows are cut into launches or ranks (salt = the row's
// global position).
__device__ __forceinline__ uint64_t rnd_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
__global__ void random_targets_kernel(uint64_t seed, uint64_t salt, int64_t n, uint64_t P, uint64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const uint64_t s = rnd_mix64(seed);
    for (; i < n; i += stride) out[i] = rnd_mix64(s ^ (salt + (uint64_t)i)) % P;
}
Buf random_targets(int64_t n, int32_t P, uint64_t seed, uint64_t salt) {
    Buf t = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    if (n > 0) {
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8));
        random_targets_kernel<<<grid, 256, 0, ctx().stream>>>(seed, salt, n, (uint64_t)P, as<uint64_t>(t));
        count_launch();
    }
    return t;
}
static std::vector<Batch *> partition_by_targets(const Batch &b, const Buf &targets, int32_t P) {
    if (P <= PS_MAXP && b.num_rows >= 4096) {
        PartSrc src;
        src.hashes = as<uint64_t>(targets);
        return partition_direct(b, src, P);
    }
    std::vector<int64_t> offsets;
    Buf perm = partition_permutation(as<uint64_t>(targets), b.num_rows, P, offsets);
    std::vector<Batch *> out;
    try {
        for (int p = 0; p < P; p++) out.push_back(gather_batch(b, as<int64_t>(perm) + offsets[p], offsets[p + 1] - offsets[p], false));
        op_done();
    } catch (...) {
        for (auto *x : out) delete x;
        throw;
    }
    return out;
}
std::vector<Batch *> op_partition_by_random(const Batch &b, int32_t P, uint64_t seed) {
    require_device();
    if (P <= 0) fail("Can not partition a Table by 0 partitions");
    return partition_by_targets(b, random_targets(b.num_rows, P, seed, 0), P);
}

std::vector<Batch *> op_partition_by_hash(const Batch &b, const int32_t *key_cols, int32_t nkeys, int32_t P) {
    require_device();
    if (P <= 0) fail("Can not partition a Table by 0 partitions");
    if (P <= PS_MAXP && b.num_rows >= 4096) return partition_direct(b, part_src_for_keys(b, key_cols, nkeys, true), P);
    Buf h = hash_rows(b, key_cols, nkeys, nullptr);
    std::vector<int64_t> offsets;
    Buf perm = partition_permutation(as<uint64_t>(h), b.num_rows, P, offsets);
    std::vector<Batch *> out;
    try {
        for (int p = 0; p < P; p++)
            out.push_back(gather_batch(b, as<int64_t>(perm) + offsets[p], offsets[p + 1] - offsets[p], false));
        op_done();
    } catch (...) {
        for (auto *x : out) delete x;
        throw;
    }
    return out;
}

// RecordBatch::partition_by_range (src/daft-recordbatch/src/ops/partition.rs:88-101): target = search_sorted(boundaries, keys)
// -> boundaries.len() + 1 partitions, input order preserved inside a partition
std::vector<Batch *> op_partition_by_range(const Batch &b, const int32_t *key_cols, int32_t nkeys, const Batch &boundaries,
                                           const uint8_t *descending) {
    require_device();
    const int32_t P = (int32_t)boundaries.num_rows + 1;
    if (boundaries.num_rows == 0) return {new Batch(b)};
    Buf target = op_search_sorted(boundaries, b, key_cols, nkeys, descending);  // in [0, P): usable as "hash % P"
    if (P <= PS_MAXP && b.num_rows >= 4096) {
        PartSrc src;
        src.hashes = as<uint64_t>(target);
        return partition_direct(b, src, P);
    }
    std::vector<int64_t> offsets;
    Buf perm = partition_permutation(as<uint64_t>(target), b.num_rows, P, offsets);
    std::vector<Batch *> out;
    try {
        for (int p = 0; p < P; p++) out.push_back(gather_batch(b, as<int64_t>(perm) + offsets[p], offsets[p + 1] - offsets[p], false));
        op_done();
    } catch (...) {
        for (auto *x : out) delete x;
        throw;
    }
    return out;
}

// distinct rows over the key columns, first occurrence kept, first-occurrence order

Batch *op_dedup(const Batch &b, const int32_t *key_cols, int32_t nkeys) {
    require_device();
    if (nkeys < 1) fail("Attempting to dedup RecordBatch on no columns");
    // AGG plan: no filter, keys = the columns, no aggregations
    std::vector<uint8_t> plan;
    auto put32 = [&](uint32_t v) { plan.insert(plan.end(), (uint8_t *)&v, (uint8_t *)&v + 4); };
    plan.push_back(0);
    put32((uint32_t)nkeys);
    for (int k = 0; k < nkeys; k++) {
        const std::string nm = "k" + std::to_string(k);
        put32((uint32_t)nm.size());
        plan.insert(plan.end(), nm.begin(), nm.end());
        plan.push_back(0);  // COL
        put32((uint32_t)key_cols[k]);
    }
    put32(0);
    Buf first;
    int64_t ng = 0;
    std::unique_ptr<Batch> keys(op_agg_first_rows(b, plan.data(), plan.size(), &first, &ng));
    Batch *out = gather_batch(b, as<int64_t>(first), ng, false);
    op_done();
    return out;
}

}  // namespace b200

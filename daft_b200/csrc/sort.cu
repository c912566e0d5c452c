// a9: argsort / sort / top-N as a stable LSD radix sort (replaces the single-threaded comparator
// sort of src/daft-core/src/array/ops/arrow/sort/primitive/common.rs:24-141).
//
// Every sort key column is first mapped to an order-preserving u64 ("ordered key"): integers are
// biased, floats are canonicalised (-0.0 -> +0.0, every NaN -> one value above +inf, matching
// cmp_float at src/daft-core/src/kernels/search_sorted.rs:238-245), strings are sorted by 7-byte
// chunks of 9-bit digits (0 = past the end), all rows on the first chunk and only the still-tied runs on the later ones.  Multi-key sorts run the keys from last to
// first; because each pass is stable the result is exactly the stable sort under the reference's
// comparator (the reference sort is unstable, so any tie order is admissible).
//
// Per pass: shared-memory histogram -> scan over (digit, CTA) -> stable scatter using
// warp match_any ranks.  HBM traffic per pass = 16 B read (hist pass reads keys only: 8 B) + 16 B
// written per row; passes whose digit is constant over the whole input are skipped.
#include <algorithm>

#include "common.h"
#include "kernels.cuh"
#include "ops.h"

namespace b200 {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;

template <int BITS>
__global__ void __launch_bounds__(RS_THREADS) radix_hist(const uint64_t *__restrict__ keys, int64_t n, int64_t chunk, int shift,
                                                          uint32_t xor_mask, int64_t *__restrict__ hist /* [digit][block] */,
                                                          unsigned long long *__restrict__ totals) {
    constexpr int R = 1 << BITS;
    __shared__ unsigned int s_cnt[R];
    for (int i = threadIdx.x; i < R; i += RS_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const int64_t begin = (int64_t)blockIdx.x * chunk, end = min(begin + chunk, n);
    for (int64_t i = begin + threadIdx.x; i < end; i += RS_THREADS) {
        uint32_t d = ((uint32_t)(keys[i] >> shift) & (R - 1)) ^ xor_mask;
        atomicAdd(&s_cnt[d], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < R; i += RS_THREADS) {
        hist[(int64_t)i * gridDim.x + blockIdx.x] = s_cnt[i];
        if (s_cnt[i]) atomicAdd(&totals[i], (unsigned long long)s_cnt[i]);
    }
}

// Stable scatter of one 8-bit digit.  A CTA walks its chunk in tiles of RS_THREADS * RS_ITEMS rows; inside a tile
// warp w owns rows [w * 32 * RS_ITEMS, (w + 1) * 32 * RS_ITEMS) and item j of lane l is row j * 32 + l of that
// span, so (warp, item, lane) order IS row order.  All keys / values of a thread are loaded up front (memory-level
// parallelism); ranks inside the warp come from match_any + a warp-private digit counter (no block barrier per
// item).  The tile is then ordered by digit in shared memory and written out by consecutive threads, so every
// digit's rows leave as one contiguous, coalesced run instead of 32 scattered 8-byte stores per instruction.
constexpr int RS_ITEMS = 8;
#ifndef RS_MINB
#define RS_MINB 4
#endif
template <int BITS>
__global__ void __launch_bounds__(RS_THREADS, RS_MINB) radix_scatter(const uint64_t *__restrict__ keys, const int64_t *__restrict__ vals,
                                                             int64_t n, int64_t chunk, int shift, uint32_t xor_mask,
                                                             const int64_t *__restrict__ base /* scanned hist */,
                                                             uint64_t *__restrict__ keys_out, int64_t *__restrict__ vals_out) {
    constexpr int R = 1 << BITS;
    constexpr int TILE = RS_THREADS * RS_ITEMS;
    static_assert(R == RS_THREADS, "one thread per digit in the prefix step");
    __shared__ int64_t s_base[R];                 // next output position per digit for this CTA
    __shared__ unsigned int s_wcnt[RS_WARPS][R];  // per-warp digit counts -> exclusive prefix over warps
    __shared__ unsigned int s_tile[R];            // digit totals of the current tile
    __shared__ unsigned int s_toff[R];            // exclusive prefix of s_tile: where a digit's run starts inside the tile
    __shared__ unsigned int s_wtot[RS_WARPS];
    __shared__ uint64_t s_k[TILE];
    __shared__ int64_t s_v[TILE];
    s_base[threadIdx.x] = base[(int64_t)threadIdx.x * gridDim.x + blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t begin = (int64_t)blockIdx.x * chunk, end = min(begin + chunk, n);
    for (int64_t t0 = begin; t0 < end; t0 += TILE) {
#pragma unroll
        for (int i = 0; i < R / 32; i++) s_wcnt[warp][i * 32 + lane] = 0;
        __syncwarp();
        const int64_t wbase = t0 + (int64_t)warp * 32 * RS_ITEMS + lane;
        uint64_t k[RS_ITEMS];
        uint32_t d[RS_ITEMS];
        unsigned short rank[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const int64_t i = wbase + j * 32;
            k[j] = i < end ? keys[i] : 0;
        }
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const bool valid = wbase + j * 32 < end;
            d[j] = valid ? (((uint32_t)(k[j] >> shift) & (R - 1)) ^ xor_mask) : 0xFFFFFFFFu;  // invalid lanes only match each other
            const unsigned peers = __match_any_sync(0xffffffffu, d[j]);
            const int lt = __popc(peers & ((1u << lane) - 1u));  // stable: lower lanes (earlier rows) first
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
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                const unsigned int c = s_wcnt[w][dd];
                s_wcnt[w][dd] = run;
                run += c;
            }
            s_tile[dd] = run;
            // exclusive scan of the digit totals over the 256 digits
            unsigned int x = run;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_wtot[warp] = x;
            __syncthreads();
            unsigned int wpre = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) wpre += w < warp ? s_wtot[w] : 0u;
            s_toff[dd] = wpre + x - run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            if (wbase + j * 32 < end) {
                // the row ids are loaded here, after the ranking: they are only passed through, and holding them in
                // registers across the ranking costs a resident CTA per SM
                const unsigned int lp = s_toff[d[j]] + s_wcnt[warp][d[j]] + rank[j];
                s_k[lp] = k[j];
                s_v[lp] = vals[wbase + j * 32];
            }
        }
        __syncthreads();
        const int cnt = (int)min((int64_t)TILE, end - t0);
        for (int e = threadIdx.x; e < cnt; e += RS_THREADS) {
            const uint64_t kk = s_k[e];
            const uint32_t dd = ((uint32_t)(kk >> shift) & (R - 1)) ^ xor_mask;
            const int64_t pos = s_base[dd] + (int64_t)(e - (int)s_toff[dd]);
            keys_out[pos] = kk;
            vals_out[pos] = s_v[e];
        }
        __syncthreads();
        s_base[threadIdx.x] += s_tile[threadIdx.x];
        // the next tile reads s_base only after its own block barriers
    }
}

template <int BITS>
static void radix_pass(const uint64_t *kin, const int64_t *vin, uint64_t *kout, int64_t *vout, int64_t n, int shift,
                       bool invert, bool *skipped) {
    constexpr int R = 1 << BITS;
    cudaStream_t st = ctx().stream;
    int64_t nblocks = std::min<int64_t>((n + 2047) / 2048, (int64_t)ctx().sm_count * 8);
    nblocks = std::max<int64_t>(nblocks, 1);
    int64_t chunk = (n + nblocks - 1) / nblocks;
    chunk = (chunk + RS_THREADS * RS_ITEMS - 1) / (RS_THREADS * RS_ITEMS) * (RS_THREADS * RS_ITEMS);
    nblocks = (n + chunk - 1) / chunk;
    Buf hist = dev_alloc((size_t)R * nblocks * 8), base = dev_alloc((size_t)R * nblocks * 8), totals = dev_alloc_zero((size_t)R * 8);
    const uint32_t xm = invert ? (uint32_t)(R - 1) : 0u;
    radix_hist<BITS><<<(unsigned)nblocks, RS_THREADS, 0, st>>>(kin, n, chunk, shift, xm, as<int64_t>(hist), as<unsigned long long>(totals));
    count_launch();
    std::vector<unsigned long long> tot(R);
    B200_CUDA(cudaMemcpyAsync(tot.data(), totals->ptr, (size_t)R * 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    for (int i = 0; i < R; i++)
        if (tot[i] == (unsigned long long)n) {
            *skipped = true;  // constant digit: the pass would be the identity permutation
            return;
        }
    *skipped = false;
    exclusive_scan_i64(as<int64_t>(hist), as<int64_t>(base), (int64_t)R * nblocks, nullptr);
    radix_scatter<BITS><<<(unsigned)nblocks, RS_THREADS, 0, st>>>(kin, vin, n, chunk, shift, xm, as<int64_t>(base), kout, vout);
    count_launch();
    B200_CUDA(cudaGetLastError());
}

// ---- onesweep: one histogram pass for ALL digits, then one pass per digit that reads the pairs once and writes them once.
// The per-(digit, CTA) histogram + scan of the two-kernel scheme above is replaced by a decoupled look-back inside the
// scatter: tiles take their index from an atomic counter (so a tile only ever waits for tiles that already run), publish
// their digit counts, and sum their predecessors' counts until they meet a tile that already knows its inclusive prefix.
// Traffic per pass: 16 B read + 16 B written per row (+ 1 KB of status words per 2048-row tile) instead of 24 + 16.
constexpr int OS_MAXPASS = 8;
__global__ void __launch_bounds__(RS_THREADS) radix_hist_all(const uint64_t *__restrict__ keys, int64_t n, int begin_bit, int npass,
                                                              unsigned long long *__restrict__ totals /* [npass][256] */) {
    __shared__ unsigned int s_cnt[OS_MAXPASS][256];
    for (int i = threadIdx.x; i < OS_MAXPASS * 256; i += RS_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * RS_THREADS;
    // whole warps stay in the loop together (the votes below need every lane)
    for (int64_t i0 = (int64_t)blockIdx.x * RS_THREADS + (threadIdx.x & ~31); i0 < n; i0 += stride) {
        const int64_t i = i0 + lane;
        const bool valid = i < n;
        const uint64_t k = valid ? keys[i] : 0;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        // bits in which the warp's keys differ: a digit with none of them is the same for all 32 rows (the high bytes of most
        // integer columns) and costs one add per warp instead of 32 colliding ones
        const uint64_t k0 = ((uint64_t)(uint32_t)__shfl_sync(0xffffffffu, (int)(k >> 32), 0) << 32) | (uint32_t)__shfl_sync(0xffffffffu, (int)k, 0);
        const uint64_t x = valid ? (k ^ k0) : 0ULL;
        const uint64_t diff = ((uint64_t)__reduce_or_sync(0xffffffffu, (unsigned)(x >> 32)) << 32) | __reduce_or_sync(0xffffffffu, (unsigned)x);
        for (int p = 0; p < npass; p++) {
            const int sh = begin_bit + 8 * p;
            const uint32_t d = (uint32_t)(k >> sh) & 255u;
            if (((diff >> sh) & 255ULL) == 0) {
                if (lane == 0) atomicAdd(&s_cnt[p][d], (unsigned)__popc(vmask));
            } else if (valid) {
                atomicAdd(&s_cnt[p][d], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += RS_THREADS) {
        const unsigned int c = (&s_cnt[0][0])[i];
        if (c) atomicAdd(&totals[i], (unsigned long long)c);
    }
}

constexpr uint32_t OS_AGG = 1u << 30, OS_INC = 2u << 30, OS_VAL = (1u << 30) - 1;
__global__ void __launch_bounds__(RS_THREADS) radix_onesweep(const uint64_t *__restrict__ keys, const int64_t *__restrict__ vals, int64_t n, int shift,
                                                              uint32_t xor_mask, const int64_t *__restrict__ digit_base /* [256] */,
                                                              unsigned int *__restrict__ tile_counter, volatile uint32_t *__restrict__ status /* [tiles][256] */,
                                                              uint64_t *__restrict__ keys_out, int64_t *__restrict__ vals_out) {
    constexpr int R = 256;
    constexpr int TILE = RS_THREADS * RS_ITEMS;
    __shared__ int64_t s_base[R];
    __shared__ unsigned int s_wcnt[RS_WARPS][R];
    __shared__ unsigned int s_toff[R];
    __shared__ unsigned int s_wtot[RS_WARPS];
    __shared__ unsigned int s_tile_id;
    __shared__ uint64_t s_k[TILE];
    __shared__ int64_t s_v[TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ntiles = (n + TILE - 1) / TILE;
    for (;;) {
        __syncthreads();  // the previous tile's s_k / s_v / s_tile_id are no longer read
        if (threadIdx.x == 0) s_tile_id = atomicAdd(tile_counter, 1u);
        __syncthreads();
        const int64_t tile = s_tile_id;
        if (tile >= ntiles) break;
        const int64_t t0 = tile * TILE, end = min(t0 + (int64_t)TILE, n);
#pragma unroll
        for (int i = 0; i < R / 32; i++) s_wcnt[warp][i * 32 + lane] = 0;
        __syncwarp();
        const int64_t wbase = t0 + (int64_t)warp * 32 * RS_ITEMS + lane;
        uint64_t k[RS_ITEMS];
        int64_t v[RS_ITEMS];
        uint32_t d[RS_ITEMS];
        unsigned short rank[RS_ITEMS];
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const int64_t i = wbase + j * 32;
            k[j] = 0;
            v[j] = 0;
            if (i < end) {
                k[j] = keys[i];
                v[j] = vals[i];
            }
        }
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            const bool valid = wbase + j * 32 < end;
            d[j] = valid ? (((uint32_t)(k[j] >> shift) & (R - 1)) ^ xor_mask) : 0xFFFFFFFFu;
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
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) {
                const unsigned int c = s_wcnt[w][dd];
                s_wcnt[w][dd] = run;
                run += c;
            }
            // publish this tile's count of digit dd, then sum the predecessors' (decoupled look-back)
            volatile uint32_t *mine = status + tile * R + dd;
            if (tile == 0) {
                *mine = OS_INC | run;
                s_base[dd] = digit_base[dd];
            } else {
                *mine = OS_AGG | run;
                uint32_t excl = 0;
                for (int64_t t = tile - 1;; t--) {
                    uint32_t sv;
                    do { sv = status[t * R + dd]; } while ((sv >> 30) == 0);
                    excl += sv & OS_VAL;
                    if ((sv >> 30) == 2) break;
                }
                *mine = OS_INC | (excl + run);
                s_base[dd] = digit_base[dd] + (int64_t)excl;
            }
            // exclusive scan of the digit totals over the 256 digits: where a digit's run starts inside the tile
            unsigned int x = run;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            if (lane == 31) s_wtot[warp] = x;
            __syncthreads();
            unsigned int wpre = 0;
#pragma unroll
            for (int w = 0; w < RS_WARPS; w++) wpre += w < warp ? s_wtot[w] : 0u;
            s_toff[dd] = wpre + x - run;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < RS_ITEMS; j++) {
            if (wbase + j * 32 < end) {
                const unsigned int lp = s_toff[d[j]] + s_wcnt[warp][d[j]] + rank[j];
                s_k[lp] = k[j];
                s_v[lp] = v[j];
            }
        }
        __syncthreads();
        const int cnt = (int)(end - t0);
        for (int e = threadIdx.x; e < cnt; e += RS_THREADS) {
            const uint64_t kk = s_k[e];
            const uint32_t dd = ((uint32_t)(kk >> shift) & (R - 1)) ^ xor_mask;
            const int64_t pos = s_base[dd] + (int64_t)(e - (int)s_toff[dd]);
            keys_out[pos] = kk;
            vals_out[pos] = s_v[e];
        }
    }
}

static bool radix_sort_onesweep(Buf &keys, Buf &vals, int64_t n, int begin_bit, int end_bit, bool descending) {
    const int npass = (end_bit - begin_bit + 7) / 8;
    // Opt-in (B200_SORT_ONESWEEP=1).  Measured on B200, 60 M (key, row id) pairs: a onesweep pass takes 0.76-0.88 ms against
    // 0.07 + 0.55-0.9 ms for histogram + scatter of the two-kernel scheme, and the 8-digit histogram 0.39-0.52 ms -- 4.2 ms
    // either way for an Int64 argsort, slower for Float64 (8 live digits).  Both scatter kernels run at 2.2-2.5 TB/s: the
    // tile work (ranking, shared-memory reorder, three block barriers at 24 warps / SM) bounds them, not the histogram read.
    const char *env = getenv("B200_SORT_ONESWEEP");
    if (!(env && *env == '1') || npass < 1 || npass > OS_MAXPASS || n >= ((int64_t)1 << 30) || n < (1 << 16)) return false;
    cudaStream_t st = ctx().stream;
    constexpr int TILE = RS_THREADS * RS_ITEMS;
    const int64_t ntiles = (n + TILE - 1) / TILE;
    Buf totals = dev_alloc_zero((size_t)npass * 256 * 8);
    const int hgrid = (int)std::min<int64_t>((n + RS_THREADS - 1) / RS_THREADS, (int64_t)ctx().sm_count * 8);
    radix_hist_all<<<hgrid, RS_THREADS, 0, st>>>(as<uint64_t>(keys), n, begin_bit, npass, as<unsigned long long>(totals));
    count_launch();
    std::vector<unsigned long long> tot((size_t)npass * 256);
    B200_CUDA(cudaMemcpyAsync(tot.data(), totals->ptr, tot.size() * 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));  // the only host round trip of the sort
    // digit bases of every pass that permutes anything (a pass whose digit is constant is the identity)
    std::vector<int> live;
    std::vector<int64_t> bases;
    const uint32_t xm = descending ? 255u : 0u;
    for (int p = 0; p < npass; p++) {
        bool constant = false;
        for (int d = 0; d < 256; d++) constant = constant || tot[(size_t)p * 256 + d] == (unsigned long long)n;
        if (constant) continue;
        live.push_back(p);
        int64_t run = 0;
        for (int dd = 0; dd < 256; dd++) {  // dd = digit as the kernel sees it (xor-ed for a descending sort)
            bases.push_back(run);
            run += (int64_t)tot[(size_t)p * 256 + (dd ^ (int)xm)];
        }
    }
    if (live.empty()) return true;
    Buf dbase = dev_alloc(bases.size() * 8);
    B200_CUDA(cudaMemcpyAsync(dbase->ptr, bases.data(), bases.size() * 8, cudaMemcpyHostToDevice, st));
    Buf k2 = dev_alloc((size_t)n * 8), v2 = dev_alloc((size_t)n * 8);
    const size_t status_bytes = (size_t)ntiles * 256 * 4 + 256;
    Buf status = dev_alloc(status_bytes);
    int occ = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, radix_onesweep, RS_THREADS, 0));
    const int grid = (int)std::min<int64_t>(ntiles, (int64_t)ctx().sm_count * std::max(1, occ));
    for (size_t li = 0; li < live.size(); li++) {
        fill_u8(status->ptr, 0, status_bytes);  // tile statuses, and the tile counter behind them
        unsigned int *counter = reinterpret_cast<unsigned int *>(as<uint8_t>(status) + (size_t)ntiles * 256 * 4);
        radix_onesweep<<<grid, RS_THREADS, 0, st>>>(as<uint64_t>(keys), as<int64_t>(vals), n, begin_bit + 8 * live[li], xm, as<int64_t>(dbase) + li * 256,
                                                    counter, as<uint32_t>(status), as<uint64_t>(k2), as<int64_t>(v2));
        count_launch();
        std::swap(keys, k2);
        std::swap(vals, v2);
    }
    // (`bases` is pageable host memory: cudaMemcpyAsync has staged it before returning, so it may go out of scope here)
    B200_CUDA(cudaGetLastError());
    return true;
}

void radix_sort_pairs(Buf &keys, Buf &vals, int64_t n, int begin_bit, int end_bit, bool descending) {
    if (n <= 1) return;
    if (radix_sort_onesweep(keys, vals, n, begin_bit, end_bit, descending)) return;
    Buf k2 = dev_alloc((size_t)n * 8), v2 = dev_alloc((size_t)n * 8);
    for (int shift = begin_bit; shift < end_bit; shift += 8) {
        bool skipped = false;
        radix_pass<8>(as<uint64_t>(keys), as<int64_t>(vals), as<uint64_t>(k2), as<int64_t>(v2), n, shift, descending, &skipped);
        if (!skipped) {
            std::swap(keys, k2);
            std::swap(vals, v2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ordered keys
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t ordered_f64(double v) {
    if (v != v) return 0xFFFFFFFFFFFFFFFFULL;   // all NaN equal, greatest
    v = v + 0.0;                                 // -0.0 -> +0.0 (equal under the reference comparator)
    uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

// out[i] = ordered key of src[idx[i]] (0 when null); nullflag[i] = 1 when null
__global__ void make_ordered_keys(int32_t dtype, const void *__restrict__ data, const uint8_t *__restrict__ validity,
                                  const int64_t *__restrict__ idx, int64_t n, uint64_t *__restrict__ out,
                                  uint64_t *__restrict__ nullflag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t r = idx[i];
        const bool ok = validity == nullptr || bit_at(validity, r);
        uint64_t k = 0;
        if (ok) {
            switch (dtype) {
                case B200_BOOL: k = bit_at((const uint8_t *)data, r); break;
                case B200_I8: k = (uint64_t)((int64_t)((const int8_t *)data)[r]) ^ 0x8000000000000000ULL; break;
                case B200_I16: k = (uint64_t)((int64_t)((const int16_t *)data)[r]) ^ 0x8000000000000000ULL; break;
                case B200_I32: case B200_DATE32: k = (uint64_t)((int64_t)((const int32_t *)data)[r]) ^ 0x8000000000000000ULL; break;
                case B200_I64: case B200_TIMESTAMP: k = (uint64_t)(((const int64_t *)data)[r]) ^ 0x8000000000000000ULL; break;
                case B200_U8: k = ((const uint8_t *)data)[r]; break;
                case B200_U16: k = ((const uint16_t *)data)[r]; break;
                case B200_U32: k = ((const uint32_t *)data)[r]; break;
                case B200_U64: k = ((const uint64_t *)data)[r]; break;
                case B200_F32: k = ordered_f64((double)((const float *)data)[r]); break;
                case B200_F64: k = ordered_f64(((const double *)data)[r]); break;
                default: break;
            }
        }
        out[i] = k;
        if (nullflag) nullflag[i] = ok ? 0 : 1;
    }
}

// ---- strings: sort by a 7-byte chunk, then refine only the rows that still tie ------------------------------------------
// A byte-position LSD sort costs one pass over ALL rows per byte of the LONGEST value: one comment of a few hundred bytes and
// the sort runs hundreds of passes.  Instead: chunk r of a value = its bytes [7r, 7r + 7) as seven 9-bit digits (0 = past the
// end, byte + 1 otherwise) in one u64, most significant first -- comparing chunks in order is comparing the strings.  Round 0
// sorts all rows by chunk 0.  A run of equal chunks whose LAST digit is 0 holds equal strings (they all ended inside the
// chunk); every other run of two or more rows is unresolved.  Round r sorts only the unresolved rows, by (run, chunk r), and
// writes them back to the positions they occupied -- the runs are contiguous, so the order of everything else stands.
// Every sort is a stable LSD radix sort, so rows with equal strings keep the order earlier sort keys gave them.
__global__ void make_string_chunk(const int64_t *__restrict__ offsets, const uint8_t *__restrict__ bytes, const uint8_t *__restrict__ validity,
                                  const int64_t *__restrict__ rows, int64_t m, int64_t chunk, uint64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        const int64_t r = rows[i];
        uint64_t k = 0;
        if (validity == nullptr || bit_at(validity, r)) {
            const int64_t o = offsets[r], len = offsets[r + 1] - o, at = chunk * 7;
#pragma unroll
            for (int j = 0; j < 7; j++) k = (k << 9) | (at + j < len ? (uint64_t)bytes[o + at + j] + 1 : 0ULL);
        }
        out[i] = k;
    }
}
// after a round: head[i] = 1 where a new run starts (run id changes or chunk changes); unresolved[i] = 1 when the row's run
// has at least two rows and its chunk does not end the string
__global__ void string_runs(const uint64_t *__restrict__ chunk, const uint64_t *__restrict__ run, int64_t m, int64_t *__restrict__ head,
                            int64_t *__restrict__ unresolved) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) {
        const uint64_t c = chunk[i];
        const bool same_prev = i > 0 && chunk[i - 1] == c && (run == nullptr || run[i - 1] == run[i]);
        const bool same_next = i + 1 < m && chunk[i + 1] == c && (run == nullptr || run[i + 1] == run[i]);
        head[i] = same_prev ? 0 : 1;
        unresolved[i] = ((same_prev || same_next) && (c & 511ULL) != 0) ? 1 : 0;
    }
}
// compaction of the unresolved rows: their position in the full order, their row id and their new run id
__global__ void string_compact(const int64_t *__restrict__ unresolved, const int64_t *__restrict__ dst, const int64_t *__restrict__ run_scan,
                               const int64_t *__restrict__ pos_in, const int64_t *__restrict__ rows, int64_t m, int64_t *__restrict__ pos_out,
                               int64_t *__restrict__ rows_out, uint64_t *__restrict__ run_out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride)
        if (unresolved[i]) {
            const int64_t d = dst[i];
            pos_out[d] = pos_in ? pos_in[i] : i;
            rows_out[d] = rows[i];
            run_out[d] = (uint64_t)run_scan[i];
        }
}
__global__ void scatter_rows(const int64_t *__restrict__ pos, const int64_t *__restrict__ rows, int64_t m, int64_t *__restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) idx[pos[i]] = rows[i];
}
__global__ void gather_u64_by(const uint64_t *__restrict__ src, const int64_t *__restrict__ at, int64_t m, uint64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) out[i] = src[at[i]];
}

__global__ void add_i64(const int64_t *__restrict__ a, const int64_t *__restrict__ b, int64_t m, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) out[i] = a[i] + b[i];
}
__global__ void gather_i64_by(const int64_t *__restrict__ src, const int64_t *__restrict__ at, int64_t m, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < m; i += stride) out[i] = src[at[i]];
}
__global__ void make_null_flag(const uint8_t *__restrict__ validity, const int64_t *__restrict__ idx, int64_t n, uint64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = bit_at(validity, idx[i]) ? 0 : 1;
}
__global__ void max_str_len(const int64_t *__restrict__ offsets, int64_t n, unsigned long long *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long m = 0;
    for (; i < n; i += stride) m = max(m, (unsigned long long)(offsets[i + 1] - offsets[i]));
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}
__global__ void iota_kernel(int64_t *out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = i;
}

void fill_iota(int64_t *out, int64_t n) {
    if (n <= 0) return;
    const int g = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8));
    iota_kernel<<<g, 256, 0, ctx().stream>>>(out, n);
    count_launch();
}

// ------------------------------------------------------------------------------------------------
// tiny inputs (post-aggregation results: TPC-H sorts 4..1000 rows): one CTA, rank by comparison, one launch
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_SORT_MAX = 2048;
constexpr int SMALL_SORT_KEYS = 8;
struct SortKeyDesc {
    int32_t dtype, desc, nulls_first, _pad;
    const void *data;
    const uint8_t *validity;
    const int64_t *offsets;
};
struct SortKeySet {
    int32_t nkeys, _pad;
    SortKeyDesc k[SMALL_SORT_KEYS];
};

// three-way compare of row i of table A with row j of table B under the reference comparator (NaN greatest, nulls by
// nulls_first); flags and dtypes come from A, both tables have the same key dtypes
__device__ __forceinline__ int small_cmp_n(const SortKeySet &ka, int64_t i, const SortKeySet &kb, int64_t j, int nk) {
    for (int q = 0; q < nk; q++) {
        const SortKeyDesc &k = ka.k[q];
        const SortKeyDesc &kj = kb.k[q];
        if (k.dtype == B200_NULL) continue;
        const bool vi = k.validity == nullptr || bit_at(k.validity, i), vj = kj.validity == nullptr || bit_at(kj.validity, j);
        int r = 0;
        if (vi && vj) {
            switch (k.dtype) {
#define CMP_T(T) { const T a = ((const T *)k.data)[i], b = ((const T *)kj.data)[j]; r = a < b ? -1 : (a > b ? 1 : 0); } break;
                case B200_I8: CMP_T(int8_t)
                case B200_I16: CMP_T(int16_t)
                case B200_I32: case B200_DATE32: CMP_T(int32_t)
                case B200_I64: case B200_TIMESTAMP: CMP_T(int64_t)
                case B200_U8: CMP_T(uint8_t)
                case B200_U16: CMP_T(uint16_t)
                case B200_U32: CMP_T(uint32_t)
                case B200_U64: CMP_T(uint64_t)
#undef CMP_T
                case B200_BOOL: r = (int)bit_at((const uint8_t *)k.data, i) - (int)bit_at((const uint8_t *)kj.data, j); break;
                case B200_F32: case B200_F64: {
                    const double a = k.dtype == B200_F32 ? (double)((const float *)k.data)[i] : ((const double *)k.data)[i];
                    const double b = k.dtype == B200_F32 ? (double)((const float *)kj.data)[j] : ((const double *)kj.data)[j];
                    const bool na = a != a, nb = b != b;
                    if (na || nb) r = (int)na - (int)nb;
                    else r = a < b ? -1 : (a > b ? 1 : 0);
                    break;
                }
                case B200_LARGE_UTF8: case B200_LARGE_BINARY: {
                    const int64_t oa = k.offsets[i], la = k.offsets[i + 1] - oa, ob = kj.offsets[j], lb = kj.offsets[j + 1] - ob;
                    const uint8_t *pa = (const uint8_t *)k.data + oa, *pb = (const uint8_t *)kj.data + ob;
                    const int64_t m = la < lb ? la : lb;
                    for (int64_t t = 0; t < m && r == 0; t++) r = (int)pa[t] - (int)pb[t];
                    if (r == 0) r = la < lb ? -1 : (la > lb ? 1 : 0);
                    break;
                }
                default: break;
            }
            if (k.desc) r = -r;
        } else if (vi != vj) {
            r = (!vi) ? (k.nulls_first ? -1 : 1) : (k.nulls_first ? 1 : -1);
        }
        if (r) return r;
    }
    return 0;
}

__device__ __forceinline__ int small_cmp2(const SortKeySet &ka, int64_t i, const SortKeySet &kb, int64_t j) { return small_cmp_n(ka, i, kb, j, ka.nkeys); }
__device__ __forceinline__ int small_cmp(const SortKeySet &ks, int64_t i, int64_t j) { return small_cmp2(ks, i, ks, j); }

// RecordBatch::search_sorted (src/daft-recordbatch/src/ops/search_sorted.rs:13-36, kernels/search_sorted.rs:420-433): for every key
// row the number of rows of the sorted table that compare <= it (upper bound; nulls_first = descending per column)
__global__ void search_sorted_kernel(SortKeySet sorted, SortKeySet keys, int64_t nsorted, int64_t nkeys_rows, uint64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nkeys_rows; i += stride) {
        int64_t left = 0, right = nsorted;
        while (left < right) {
            const int64_t mid = left + ((right - left) >> 1);
            if (small_cmp2(sorted, mid, keys, i) <= 0) left = mid + 1;
            else right = mid;
        }
        out[i] = (uint64_t)left;
    }
}

__global__ void __launch_bounds__(1024) small_argsort_kernel(SortKeySet ks, int64_t n, int64_t *__restrict__ out_idx) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        int64_t rank = 0;
        for (int64_t j = 0; j < n; j++) {
            const int c = small_cmp(ks, j, i);
            rank += (c < 0) || (c == 0 && j < i);  // stable: ties keep row order
        }
        out_idx[rank] = i;
    }
}

static int sgrid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8)); }

static int key_bits(int32_t dt) {
    switch (dt) {
        case B200_BOOL: return 8;
        case B200_U8: return 8;
        case B200_U16: return 16;
        case B200_U32: return 32;
        default: return 64;  // signed ints carry the bias in bit 63; floats use all 64
    }
}

static SortKeySet make_sort_keyset(const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending, const uint8_t *nulls_first) {
    if (nkeys < 1 || nkeys > SMALL_SORT_KEYS) fail("between 1 and 8 key columns are supported");
    SortKeySet ks;
    ks.nkeys = nkeys;
    for (int k = 0; k < nkeys; k++) {
        const int32_t ci = key_cols[k];
        if (ci < 0 || ci >= (int32_t)b.cols.size()) fail("key column index out of range");
        const Column &c = b.cols[ci];
        const bool desc = descending && descending[k];
        ks.k[k] = {c.dtype, desc ? 1 : 0, (nulls_first ? nulls_first[k] != 0 : desc) ? 1 : 0, 0, c.data ? c.data->ptr : nullptr,
                   c.null_count > 0 ? as<uint8_t>(c.validity) : nullptr, as<int64_t>(c.offsets)};
    }
    return ks;
}

// positions[i] = number of rows of `sorted` (over its first nkeys columns) that compare <= key row i of `b`
Buf op_search_sorted(const Batch &sorted, const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending) {
    require_device();
    if ((int32_t)sorted.cols.size() != nkeys) fail("search_sorted: the sorted table must have exactly the key columns");
    std::vector<int32_t> bcols(nkeys);
    for (int k = 0; k < nkeys; k++) {
        bcols[k] = k;
        const int32_t ci = key_cols[k];
        if (ci < 0 || ci >= (int32_t)b.cols.size()) fail("search_sorted: key column index out of range");
        if (sorted.cols[k].dtype != b.cols[ci].dtype)
            fail(std::string("Schema Mismatch in search_sorted: data: ") + dtype_name(sorted.cols[k].dtype) + " vs keys: " + dtype_name(b.cols[ci].dtype));
    }
    // SortOptions::new(reversed, reversed): nulls first exactly when descending (kernels/search_sorted.rs:401-406)
    const SortKeySet ks_sorted = make_sort_keyset(sorted, bcols.data(), nkeys, descending, descending);
    const SortKeySet ks_keys = make_sort_keyset(b, key_cols, nkeys, descending, descending);
    const int64_t n = b.num_rows;
    Buf out = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    if (n > 0) {
        search_sorted_kernel<<<sgrid(n), 256, 0, ctx().stream>>>(ks_sorted, ks_keys, sorted.num_rows, n, as<uint64_t>(out));
        count_launch();
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// Sort-merge, asof and cross joins: the joins the reference runs on SORTED inputs, re-composed from the sort,
// the binary search above and the view gather.  A sequential two-pointer merge (src/daft-recordbatch/src/ops/joins/
// merge_join.rs:45-330, asof_join.rs:12-160) has no parallel form; on the device every left row binary-searches the sorted
// right keys on its own (the searched levels of a sorted key column stay in L2), which yields the same matches:
//   merge join : the run [lo, hi) of right rows whose keys equal the left row's (a null in any key matches nothing),
//   asof join  : the last right row with the same by-keys and on-key <= the left row's (null on-keys never match).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool any_key_null(const SortKeySet &ks, int64_t i, int nkeys) {
    for (int q = 0; q < nkeys; q++)
        if (ks.k[q].dtype == B200_NULL || (ks.k[q].validity != nullptr && !bit_at(ks.k[q].validity, i))) return true;
    return false;
}
// number of rows of `sorted` that compare < (STRICT) or <= row i of `keys`
template <bool STRICT>
__device__ __forceinline__ int64_t bound_of(const SortKeySet &sorted, int64_t nsorted, const SortKeySet &keys, int64_t i) {
    int64_t left = 0, right = nsorted;
    while (left < right) {
        const int64_t mid = left + ((right - left) >> 1);
        const int c = small_cmp2(sorted, mid, keys, i);
        if (STRICT ? c < 0 : c <= 0) left = mid + 1;
        else right = mid;
    }
    return left;
}
__global__ void merge_ranges_kernel(SortKeySet right, SortKeySet left, int64_t nright, int64_t nleft, int64_t *__restrict__ lo,
                                    int64_t *__restrict__ cnt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nleft; i += stride) {
        int64_t b = 0, c = 0;
        if (!any_key_null(left, i, left.nkeys)) {
            b = bound_of<true>(right, nright, left, i);
            c = bound_of<false>(right, nright, left, i) - b;  // right rows with a null key sort last and compare greater
        }
        lo[i] = b;
        cnt[i] = c;
    }
}
// one thread per output pair: the left row owning output position p is found by binary search in the scanned counts
__global__ void merge_expand_kernel(const int64_t *__restrict__ offs, const int64_t *__restrict__ lo, int64_t nleft, int64_t total,
                                    int64_t *__restrict__ lidx, int64_t *__restrict__ ridx) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; p < total; p += stride) {
        int64_t a = 0, b = nleft;  // last i with offs[i] <= p
        while (b - a > 1) {
            const int64_t mid = a + ((b - a) >> 1);
            if (offs[mid] <= p) a = mid;
            else b = mid;
        }
        lidx[p] = a;
        ridx[p] = lo[a] + (p - offs[a]);
    }
}
__global__ void asof_backward_kernel(SortKeySet right, SortKeySet left, int64_t nright, int64_t nleft, int32_t nby, int64_t *__restrict__ ridx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nleft; i += stride) {
        int64_t r = -1;
        if (!any_key_null(left, i, left.nkeys)) {
            const int64_t cand = bound_of<false>(right, nright, left, i) - 1;  // the last right row <= (by..., on) of the left row
            if (cand >= 0 && !any_key_null(right, cand, right.nkeys)) {
                if (nby == 0 || small_cmp_n(right, cand, left, i, nby) == 0) r = cand;   // same by-group?
            }
        }
        ridx[i] = r;
    }
}
__global__ void cross_indices_kernel(int64_t nl, int64_t nr, int outer_left, int64_t *__restrict__ lidx, int64_t *__restrict__ ridx) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, total = nl * nr;
    for (; p < total; p += stride) {
        // the outer-loop side repeats each of its rows |inner| times in a row, the inner side repeats as a whole
        lidx[p] = outer_left ? p / nr : p % nl;
        ridx[p] = outer_left ? p % nr : p / nl;
    }
}

static void check_join_keys(const Batch &l, const int32_t *lk, const Batch &r, const int32_t *rk, int32_t nkeys, const char *what) {
    if (nkeys < 1) fail(std::string(what) + ": No columns were passed in to join on");
    for (int k = 0; k < nkeys; k++) {
        if (lk[k] < 0 || lk[k] >= (int32_t)l.cols.size() || rk[k] < 0 || rk[k] >= (int32_t)r.cols.size()) fail(std::string(what) + ": key column index out of range");
        if (l.cols[lk[k]].dtype != r.cols[rk[k]].dtype)
            fail(std::string(what) + ": Types between left and right keys do not match: " + dtype_name(l.cols[lk[k]].dtype) + " vs " + dtype_name(r.cols[rk[k]].dtype));
    }
}
static Batch *joined_views(const Batch &left, const Buf &lidx, const Batch &right, const Buf &ridx, int64_t n, bool right_may_be_null) {
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    for (Column &c : make_view_columns(left, lidx, n, false)) out->cols.push_back(std::move(c));
    for (Column &c : make_view_columns(right, ridx, n, right_may_be_null)) out->cols.push_back(std::move(c));
    return out.release();
}

// inner join of two batches SORTED ascending (nulls last) on their key columns; output = [left columns..., right columns...] as
// views, left-major: pairs of one left row are contiguous, in right order
Batch *op_merge_join_sorted(const Batch &left, const int32_t *left_keys, const Batch &right, const int32_t *right_keys, int32_t nkeys) {
    require_device();
    check_join_keys(left, left_keys, right, right_keys, nkeys, "sort_merge_join");
    std::vector<uint8_t> zeros((size_t)nkeys, 0);
    const SortKeySet kl = make_sort_keyset(left, left_keys, nkeys, zeros.data(), zeros.data());
    const SortKeySet kr = make_sort_keyset(right, right_keys, nkeys, zeros.data(), zeros.data());
    const int64_t nl = left.num_rows, nr = right.num_rows;
    cudaStream_t st = ctx().stream;
    int64_t total = 0;
    Buf lo, offs;
    if (nl > 0 && nr > 0) {
        lo = dev_alloc((size_t)nl * 8);
        Buf cnt = dev_alloc((size_t)nl * 8);
        offs = dev_alloc((size_t)(nl + 1) * 8);
        merge_ranges_kernel<<<sgrid(nl), 256, 0, st>>>(kr, kl, nr, nl, as<int64_t>(lo), as<int64_t>(cnt));
        count_launch();
        exclusive_scan_i64(as<int64_t>(cnt), as<int64_t>(offs), nl, as<int64_t>(offs) + nl);
        total = read_i64(as<int64_t>(offs) + nl);
    }
    Buf lidx = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8), ridx = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    if (total > 0) {
        merge_expand_kernel<<<sgrid(total), 256, 0, st>>>(as<int64_t>(offs), as<int64_t>(lo), nl, total, as<int64_t>(lidx), as<int64_t>(ridx));
        count_launch();
    }
    Batch *out = joined_views(left, lidx, right, ridx, total, false);
    op_done();
    return out;
}

// backward asof join: keys = by columns..., then the on column (nby = nkeys - 1); `right` SORTED ascending (nulls last) on its keys,
// `left` in any order.  Output = every left row in its own order with the matched right row (a null row where none matches).
Batch *op_asof_join_sorted(const Batch &left, const int32_t *left_keys, const Batch &right, const int32_t *right_keys, int32_t nkeys) {
    require_device();
    check_join_keys(left, left_keys, right, right_keys, nkeys, "asof_join");
    std::vector<uint8_t> zeros((size_t)nkeys, 0);
    const SortKeySet kl = make_sort_keyset(left, left_keys, nkeys, zeros.data(), zeros.data());
    const SortKeySet kr = make_sort_keyset(right, right_keys, nkeys, zeros.data(), zeros.data());
    const int64_t nl = left.num_rows;
    Buf lidx = dev_alloc((size_t)std::max<int64_t>(nl, 1) * 8), ridx = dev_alloc((size_t)std::max<int64_t>(nl, 1) * 8);
    if (nl > 0) {
        fill_iota(as<int64_t>(lidx), nl);
        asof_backward_kernel<<<sgrid(nl), 256, 0, ctx().stream>>>(kr, kl, right.num_rows, nl, nkeys - 1, as<int64_t>(ridx));
        count_launch();
    }
    Batch *out = joined_views(left, lidx, right, ridx, nl, true);
    op_done();
    return out;
}

// RecordBatch::cross_join (src/daft-recordbatch/src/ops/joins/mod.rs:231-284): every pair, the outer-loop side's rows in order
Batch *op_cross_join(const Batch &left, const Batch &right, bool outer_loop_left) {
    require_device();
    const int64_t nl = left.num_rows, nr = right.num_rows;
    if (nl > 0 && nr > ((int64_t)1 << 40) / nl) fail("cross_join: the result would exceed 2^40 rows");
    const int64_t total = nl * nr;
    Buf lidx = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8), ridx = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    if (total > 0) {
        cross_indices_kernel<<<sgrid(total), 256, 0, ctx().stream>>>(nl, nr, outer_loop_left ? 1 : 0, as<int64_t>(lidx), as<int64_t>(ridx));
        count_launch();
    }
    Batch *out = joined_views(left, lidx, right, ridx, total, false);
    op_done();
    return out;
}

// ------------------------------------------------------------------------------------------------
// Window functions over a batch SORTED by (partition keys, order keys) (RecordBatch::window_row_number / window_rank /
// window_offset, src/daft-recordbatch/src/ops/window.rs:636-790, which the reference runs per partition batch inside
// WindowPartitionAndOrderBySink, src/daft-local-execution/src/sinks/window_partition_and_order_by.rs).  Here the whole sorted
// table goes through segmented passes: boundary flags by comparing neighbours -> inclusive scans give every row its partition
// id and peer-group id -> the first row of every partition / peer group is scattered by id -> ranks are differences.
//   row_number = i - first(partition) + 1        rank = first(peer group) - first(partition) + 1
//   dense_rank = peer group id - peer group id of the partition's first row + 1
// Keys compare as in the reference's build_multi_array_is_equal(nulls_equal = true, nans_equal = true): null == null.
// ------------------------------------------------------------------------------------------------
__global__ void window_flags_kernel(SortKeySet part, SortKeySet order, int64_t n, int64_t *__restrict__ new_part, int64_t *__restrict__ new_peer) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        // small_cmp2 orders null == null as equal and NaN == NaN as equal: the window's notion of "same partition / same peers"
        const bool np_ = i == 0 || (part.nkeys > 0 && small_cmp2(part, i - 1, part, i) != 0);
        const bool nq = np_ || (order.nkeys > 0 && small_cmp2(order, i - 1, order, i) != 0);
        new_part[i] = np_ ? 1 : 0;
        new_peer[i] = nq ? 1 : 0;
    }
}
// ids are the inclusive scans of the flags minus one; first[id] = the row where the flag is set
__global__ void window_firsts_kernel(const int64_t *__restrict__ flag_part, const int64_t *__restrict__ excl_part, const int64_t *__restrict__ flag_peer,
                                     const int64_t *__restrict__ excl_peer, int64_t n, int64_t *__restrict__ first_of_part,
                                     int64_t *__restrict__ first_of_peer) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        if (flag_part[i]) first_of_part[excl_part[i]] = i;   // exclusive scan at a flagged row = its id
        if (flag_peer[i]) first_of_peer[excl_peer[i]] = i;
    }
}
__global__ void window_ranks_kernel(const int64_t *__restrict__ flag_part, const int64_t *__restrict__ excl_part, const int64_t *__restrict__ flag_peer,
                                    const int64_t *__restrict__ excl_peer, const int64_t *__restrict__ first_of_part,
                                    const int64_t *__restrict__ first_of_peer, int64_t n, uint64_t *__restrict__ row_number,
                                    uint64_t *__restrict__ rank, uint64_t *__restrict__ dense_rank, int64_t *__restrict__ part_begin,
                                    int64_t *__restrict__ part_id) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t pid = excl_part[i] + flag_part[i] - 1, gid = excl_peer[i] + flag_peer[i] - 1;
        const int64_t p0 = first_of_part[pid];
        const int64_t g0 = excl_peer[p0];   // peer-group id of the partition's first row (that row is flagged: exclusive scan = id)
        row_number[i] = (uint64_t)(i - p0 + 1);
        rank[i] = (uint64_t)(first_of_peer[gid] - p0 + 1);
        dense_rank[i] = (uint64_t)(gid - g0 + 1);
        part_begin[i] = p0;
        part_id[i] = pid;
    }
}
// lead / lag: row i reads row i + offset when that row is in the same partition, else -1 (a null row / the default)
__global__ void window_offset_kernel(const int64_t *__restrict__ part_id, int64_t n, int64_t offset, int64_t *__restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t j = i + offset;
        idx[i] = (j >= 0 && j < n && part_id[j] == part_id[i]) ? j : -1;
    }
}

static Column make_i64_column(const std::string &name, Buf data, int64_t n) {
    Column c = make_u64_column(name, data, n);
    c.dtype = B200_I64;
    return c;
}
// `sorted`: ordered by its partition key columns, then its order key columns.  Returns five columns of sorted.num_rows rows:
// row_number, rank, dense_rank (UInt64), partition_begin (Int64: row index of the partition's first row), partition_id (Int64).
Batch *op_window_ranks(const Batch &sorted, const int32_t *part_cols, int32_t npart, const int32_t *order_cols, int32_t norder) {
    require_device();
    if (npart < 0 || norder < 0 || npart > SMALL_SORT_KEYS || norder > SMALL_SORT_KEYS) fail("window: at most 8 partition and 8 order keys");
    SortKeySet kp, ko;
    kp.nkeys = 0;
    ko.nkeys = 0;
    if (npart > 0) kp = make_sort_keyset(sorted, part_cols, npart, nullptr, nullptr);
    if (norder > 0) ko = make_sort_keyset(sorted, order_cols, norder, nullptr, nullptr);
    const int64_t n = sorted.num_rows;
    const size_t bytes = (size_t)std::max<int64_t>(n, 1) * 8;
    Buf rn = dev_alloc(bytes), rk = dev_alloc(bytes), dr = dev_alloc(bytes), pb = dev_alloc(bytes), pid = dev_alloc(bytes);
    if (n > 0) {
        cudaStream_t st = ctx().stream;
        Buf fp = dev_alloc(bytes), fq = dev_alloc(bytes), ep = dev_alloc(bytes + 8), eq = dev_alloc(bytes + 8), first_p = dev_alloc(bytes),
            first_q = dev_alloc(bytes);
        window_flags_kernel<<<sgrid(n), 256, 0, st>>>(kp, ko, n, as<int64_t>(fp), as<int64_t>(fq));
        count_launch();
        exclusive_scan_i64(as<int64_t>(fp), as<int64_t>(ep), n, nullptr);
        exclusive_scan_i64(as<int64_t>(fq), as<int64_t>(eq), n, nullptr);
        window_firsts_kernel<<<sgrid(n), 256, 0, st>>>(as<int64_t>(fp), as<int64_t>(ep), as<int64_t>(fq), as<int64_t>(eq), n, as<int64_t>(first_p),
                                                       as<int64_t>(first_q));
        count_launch();
        window_ranks_kernel<<<sgrid(n), 256, 0, st>>>(as<int64_t>(fp), as<int64_t>(ep), as<int64_t>(fq), as<int64_t>(eq), as<int64_t>(first_p),
                                                      as<int64_t>(first_q), n, as<uint64_t>(rn), as<uint64_t>(rk), as<uint64_t>(dr), as<int64_t>(pb),
                                                      as<int64_t>(pid));
        count_launch();
    }
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    out->cols.push_back(make_u64_column("row_number", rn, n));
    out->cols.push_back(make_u64_column("rank", rk, n));
    out->cols.push_back(make_u64_column("dense_rank", dr, n));
    out->cols.push_back(make_i64_column("partition_begin", pb, n));
    out->cols.push_back(make_i64_column("partition_id", pid, n));
    op_done();
    return out.release();
}
// rows of `values` (sorted like the batch op_window_ranks saw) shifted by `offset` inside their partitions: views with null rows
// where the shifted position leaves the partition, plus an Int64 column "__src" holding the source row (-1 = outside)
Batch *op_window_offset(const Batch &values, const Column &partition_id, int64_t offset) {
    require_device();
    require_plain(partition_id);
    const int64_t n = values.num_rows;
    if (partition_id.length != n || partition_id.dtype != B200_I64) fail("window_offset: partition ids must be Int64, one per row");
    Buf idx = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    if (n > 0) {
        window_offset_kernel<<<sgrid(n), 256, 0, ctx().stream>>>(as<int64_t>(partition_id.data), n, offset, as<int64_t>(idx));
        count_launch();
    }
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    for (Column &c : make_view_columns(values, idx, n, true)) out->cols.push_back(std::move(c));
    out->cols.push_back(make_i64_column("__src", idx, n));
    op_done();
    return out.release();
}

Buf op_argsort(const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending, const uint8_t *nulls_first) {
    require_device();
    if (nkeys < 1) fail("argsort: need at least one sort key");
    const int64_t n = b.num_rows;
    cudaStream_t st = ctx().stream;
    Buf idx = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    if (n == 0) return idx;
    iota_kernel<<<sgrid(n), 256, 0, st>>>(as<int64_t>(idx), n);
    count_launch();
    if (n == 1) return idx;
    if (n <= SMALL_SORT_MAX && nkeys <= SMALL_SORT_KEYS) {
        SortKeySet ks;
        ks.nkeys = nkeys;
        for (int k = 0; k < nkeys; k++) {
            const int32_t ci = key_cols[k];
            if (ci < 0 || ci >= (int32_t)b.cols.size()) fail("argsort: key column index out of range");
            const Column &c = b.cols[ci];
            const bool desc = descending && descending[k];
            ks.k[k] = {c.dtype, desc ? 1 : 0, (nulls_first ? nulls_first[k] != 0 : desc) ? 1 : 0, 0, c.data ? c.data->ptr : nullptr,
                       c.null_count > 0 ? as<uint8_t>(c.validity) : nullptr, as<int64_t>(c.offsets)};
        }
        small_argsort_kernel<<<1, 1024, 0, st>>>(ks, n, as<int64_t>(idx));
        count_launch();
        op_done();
        return idx;
    }
    Buf keys = dev_alloc((size_t)n * 8);
    for (int k = nkeys - 1; k >= 0; k--) {
        const int32_t ci = key_cols[k];
        if (ci < 0 || ci >= (int32_t)b.cols.size()) fail("argsort: key column index out of range");
        const Column &c = b.cols[ci];
        const bool desc = descending && descending[k];
        const bool nf = nulls_first ? nulls_first[k] != 0 : desc;
        if (c.dtype == B200_NULL) continue;  // every row ties
        const uint8_t *validity = c.null_count > 0 ? as<uint8_t>(c.validity) : nullptr;
        if (is_string(c.dtype)) {
            const int64_t *off = as<int64_t>(c.offsets);
            const uint8_t *dat = as<uint8_t>(c.data);
            // values shorter than a chunk leave its low digits zero for every row: the radix passes start above them
            Buf ml = dev_alloc_zero(8);
            max_str_len<<<sgrid(n), 256, 0, st>>>(off, c.length, as<unsigned long long>(ml));
            count_launch();
            const int64_t maxlen = read_i64(as<int64_t>(ml));
            const int low_bit = (int)(9 * std::max<int64_t>(0, 7 - maxlen)) / 8 * 8;
            // round 0: all rows by their first chunk
            make_string_chunk<<<sgrid(n), 256, 0, st>>>(off, dat, validity, as<int64_t>(idx), n, 0, as<uint64_t>(keys));
            count_launch();
            radix_sort_pairs(keys, idx, n, low_bit, 63, desc);
            // refine: the rows still tied after chunk r - 1 are sorted by (run, chunk r) and put back where they were
            Buf pos, rows, run, chunks = keys;  // of the unresolved subset: position in idx, row id, run id, last sorted chunk
            int64_t m = n;
            for (int64_t round = 1; m > 1; round++) {
                Buf head = dev_alloc((size_t)m * 8), unres = dev_alloc((size_t)m * 8), hscan = dev_alloc((size_t)(m + 1) * 8),
                    uscan = dev_alloc((size_t)(m + 1) * 8);
                string_runs<<<sgrid(m), 256, 0, st>>>(as<uint64_t>(chunks), as<uint64_t>(run), m, as<int64_t>(head), as<int64_t>(unres));
                count_launch();
                exclusive_scan_i64(as<int64_t>(head), as<int64_t>(hscan), m, nullptr);
                exclusive_scan_i64(as<int64_t>(unres), as<int64_t>(uscan), m, as<int64_t>(uscan) + m);
                const int64_t m2 = read_i64(as<int64_t>(uscan) + m);
                if (m2 == 0) break;
                Buf pos2 = dev_alloc((size_t)m2 * 8), rows2 = dev_alloc((size_t)m2 * 8), run2 = dev_alloc((size_t)m2 * 8);
                // new run id = number of heads up to and including the row's own run head: constant over a run, ascending
                Buf runid = dev_alloc((size_t)m * 8);
                add_i64<<<sgrid(m), 256, 0, st>>>(as<int64_t>(hscan), as<int64_t>(head), m, as<int64_t>(runid));
                count_launch();
                string_compact<<<sgrid(m), 256, 0, st>>>(as<int64_t>(unres), as<int64_t>(uscan), as<int64_t>(runid), as<int64_t>(pos),
                                                         round == 1 ? as<int64_t>(idx) : as<int64_t>(rows), m, as<int64_t>(pos2), as<int64_t>(rows2),
                                                         as<uint64_t>(run2));
                count_launch();
                Buf ch2 = dev_alloc((size_t)m2 * 8);
                make_string_chunk<<<sgrid(m2), 256, 0, st>>>(off, dat, validity, as<int64_t>(rows2), m2, round, as<uint64_t>(ch2));
                count_launch();
                // stable sort by (run, chunk): chunk first, then run; the run ids travel as a second payload through an index
                Buf ord = dev_alloc((size_t)m2 * 8);
                fill_iota(as<int64_t>(ord), m2);
                radix_sort_pairs(ch2, ord, m2, 0, 63, desc);
                Buf run_by = dev_alloc((size_t)m2 * 8);
                gather_u64_by<<<sgrid(m2), 256, 0, st>>>(as<uint64_t>(run2), as<int64_t>(ord), m2, as<uint64_t>(run_by));
                count_launch();
                Buf ord2 = dev_alloc((size_t)m2 * 8);
                fill_iota(as<int64_t>(ord2), m2);
                int rbits = 1;
                while (((int64_t)1 << rbits) <= m && rbits < 63) rbits++;
                radix_sort_pairs(run_by, ord2, m2, 0, rbits, false);   // ord2[j] = position in the chunk-sorted order
                // final order of the subset: element j = rows2[ord[ord2[j]]], chunk = ch2[ord2[j]], run = run_by[j]
                Buf rows_s = dev_alloc((size_t)m2 * 8), ch_s = dev_alloc((size_t)m2 * 8), comp = dev_alloc((size_t)m2 * 8);
                gather_i64_by<<<sgrid(m2), 256, 0, st>>>(as<int64_t>(ord), as<int64_t>(ord2), m2, as<int64_t>(comp));
                count_launch();
                gather_i64_by<<<sgrid(m2), 256, 0, st>>>(as<int64_t>(rows2), as<int64_t>(comp), m2, as<int64_t>(rows_s));
                count_launch();
                gather_u64_by<<<sgrid(m2), 256, 0, st>>>(as<uint64_t>(ch2), as<int64_t>(ord2), m2, as<uint64_t>(ch_s));
                count_launch();
                scatter_rows<<<sgrid(m2), 256, 0, st>>>(as<int64_t>(pos2), as<int64_t>(rows_s), m2, as<int64_t>(idx));
                count_launch();
                pos = pos2;
                rows = rows_s;
                run = run_by;
                chunks = ch_s;
                m = m2;
            }
        } else {
            make_ordered_keys<<<sgrid(n), 256, 0, st>>>(c.dtype, c.data ? c.data->ptr : nullptr, validity, as<int64_t>(idx), n,
                                                         as<uint64_t>(keys), nullptr);
            count_launch();
            radix_sort_pairs(keys, idx, n, 0, key_bits(c.dtype), desc);
        }
        if (validity) {
            make_null_flag<<<sgrid(n), 256, 0, st>>>(validity, as<int64_t>(idx), n, as<uint64_t>(keys));
            count_launch();
            // nulls_first: nulls (flag 1) must sort first -> descending on the flag
            radix_sort_pairs(keys, idx, n, 0, 8, nf);
        }
    }
    op_done();
    return idx;
}

// ------------------------------------------------------------------------------------------------
// top-N by selection (RecordBatch::top_n = sort + slice in the reference, src/daft-recordbatch/src/ops/sort.rs:47-61):
// instead of sorting all rows, find the rank-K threshold of the FIRST sort key with an MSB-first radix select
// (11-bit digits, one 8 B/row read per pass), compact the rows at or below the threshold (ties included), and run
// the full multi-key comparator sort on those few candidates only.
// ------------------------------------------------------------------------------------------------
constexpr int SEL_BITS = 11, SEL_BINS = 1 << SEL_BITS;

// 64-bit selection key of row r: ascending order of the key == output order of the first sort key; nulls are
// folded onto the smallest / greatest value (a monotone coarsening: rows it cannot separate stay candidates)
__device__ __forceinline__ uint64_t select_key(int32_t dtype, const void *data, const uint8_t *validity, int64_t r, bool desc, bool nulls_first) {
    if (validity != nullptr && !bit_at(validity, r)) return nulls_first ? 0ULL : 0xFFFFFFFFFFFFFFFFULL;
    uint64_t k = 0;
    switch (dtype) {
        case B200_BOOL: k = bit_at((const uint8_t *)data, r); break;
        case B200_I8: k = (uint64_t)((int64_t)((const int8_t *)data)[r]) ^ 0x8000000000000000ULL; break;
        case B200_I16: k = (uint64_t)((int64_t)((const int16_t *)data)[r]) ^ 0x8000000000000000ULL; break;
        case B200_I32: case B200_DATE32: k = (uint64_t)((int64_t)((const int32_t *)data)[r]) ^ 0x8000000000000000ULL; break;
        case B200_I64: case B200_TIMESTAMP: k = (uint64_t)(((const int64_t *)data)[r]) ^ 0x8000000000000000ULL; break;
        case B200_U8: k = ((const uint8_t *)data)[r]; break;
        case B200_U16: k = ((const uint16_t *)data)[r]; break;
        case B200_U32: k = ((const uint32_t *)data)[r]; break;
        case B200_U64: k = ((const uint64_t *)data)[r]; break;
        case B200_F32: k = ordered_f64((double)((const float *)data)[r]); break;
        case B200_F64: k = ordered_f64(((const double *)data)[r]); break;
        default: break;
    }
    return desc ? ~k : k;
}

// histogram of the SEL_BITS-bit digit at `shift` over the rows whose higher bits equal `prefix`
__global__ void __launch_bounds__(256) select_hist(int32_t dtype, const void *__restrict__ data, const uint8_t *__restrict__ validity, int64_t n,
                                                   int desc, int nulls_first, uint64_t prefix, int shift, int bits, int first,
                                                   unsigned long long *__restrict__ hist) {
    __shared__ unsigned int h[SEL_BINS];
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t k = select_key(dtype, data, validity, i, desc != 0, nulls_first != 0);
        const bool in = first || (k >> (shift + bits)) == prefix;
        if (in) atomicAdd(&h[(k >> shift) & ((1u << bits) - 1u)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SEL_BINS; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], (unsigned long long)h[i]);
}

// row ids (ascending) of the rows with (key >> shift) <= bound
__global__ void __launch_bounds__(256) select_compact(int32_t dtype, const void *__restrict__ data, const uint8_t *__restrict__ validity, int64_t n,
                                                      int desc, int nulls_first, uint64_t bound, int shift, unsigned long long *__restrict__ counter,
                                                      int64_t *__restrict__ out, int64_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    const int lane = threadIdx.x & 31;
    for (; i < n_round; i += stride) {
        bool take = false;
        if (i < n) take = (select_key(dtype, data, validity, i, desc != 0, nulls_first != 0) >> shift) <= bound;
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (m) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
            if (take && pos < cap) out[pos] = i;
        }
    }
}

__global__ void map_indices(const int64_t *__restrict__ cand, const int64_t *__restrict__ local, int64_t n, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = cand[local[i]];
}

// first `k` entries of the argsort (k clamped to the row count)
Buf op_top_indices(const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending, const uint8_t *nulls_first, int64_t k) {
    require_device();
    if (nkeys < 1) fail("top_n: need at least one sort key");
    const int64_t n = b.num_rows;
    k = std::max<int64_t>(0, std::min(k, n));
    const int32_t c0 = key_cols[0];
    if (c0 < 0 || c0 >= (int32_t)b.cols.size()) fail("top_n: key column index out of range");
    const Column &kc = b.cols[c0];
    const bool selectable = n > (1 << 16) && k > 0 && k <= n / 16 && !is_string(kc.dtype) && kc.dtype != B200_NULL;
    if (!selectable) return op_argsort(b, key_cols, nkeys, descending, nulls_first);
    cudaStream_t st = ctx().stream;
    const bool desc = descending && descending[0];
    const bool nf = nulls_first ? nulls_first[0] != 0 : desc;
    const uint8_t *validity = kc.null_count > 0 ? as<uint8_t>(kc.validity) : nullptr;
    const void *data = kc.data ? kc.data->ptr : nullptr;
    const int total_bits = 64;
    Buf hist = dev_alloc((size_t)SEL_BINS * 8);
    std::vector<unsigned long long> h(SEL_BINS);
    uint64_t prefix = 0;
    int64_t below = 0;          // rows strictly before the selected bucket
    int64_t bucket = n;         // rows inside it
    int shift = total_bits;     // bits of the key not yet fixed
    int64_t rank = k;           // rank of the threshold inside the current bucket (1-based)
    const int64_t slack = 4096;
    while (shift > 0 && bucket > slack) {
        const int bits = std::min(SEL_BITS, shift);
        const int nshift = shift - bits;
        fill_u8(hist->ptr, 0, (size_t)SEL_BINS * 8);
        select_hist<<<sgrid(n), 256, 0, st>>>(kc.dtype, data, validity, n, desc ? 1 : 0, nf ? 1 : 0, prefix, nshift, bits, shift == total_bits ? 1 : 0,
                                              as<unsigned long long>(hist));
        count_launch();
        B200_CUDA(cudaMemcpyAsync(h.data(), hist->ptr, (size_t)SEL_BINS * 8, cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        int64_t acc = 0;
        int d = 0;
        for (; d < SEL_BINS; d++) {
            if (acc + (int64_t)h[d] >= rank) break;
            acc += (int64_t)h[d];
        }
        if (d == SEL_BINS) fail("top_n: selection histogram is inconsistent");
        below += acc;
        rank -= acc;
        bucket = (int64_t)h[d];
        prefix = (prefix << bits) | (uint64_t)d;
        shift = nshift;
    }
    const int64_t ncand = below + bucket;
    if (ncand > n / 2) return op_argsort(b, key_cols, nkeys, descending, nulls_first);  // massive ties: nothing was saved
    Buf cand = dev_alloc((size_t)ncand * 8), counter = dev_alloc_zero(8);
    select_compact<<<sgrid(n), 256, 0, st>>>(kc.dtype, data, validity, n, desc ? 1 : 0, nf ? 1 : 0, prefix, shift, as<unsigned long long>(counter),
                                             as<int64_t>(cand), ncand);
    count_launch();
    if (read_i64(as<int64_t>(counter)) != ncand) fail("top_n: candidate count mismatch");
    // candidates arrive in arbitrary warp order: ties are unordered in the reference too (unstable sort)
    Batch keys_only;
    keys_only.num_rows = n;
    std::vector<int32_t> kidx(nkeys);
    for (int i = 0; i < nkeys; i++) {
        const int32_t ci = key_cols[i];
        if (ci < 0 || ci >= (int32_t)b.cols.size()) fail("top_n: key column index out of range");
        keys_only.cols.push_back(b.cols[ci]);
        kidx[i] = i;
    }
    std::unique_ptr<Batch> cb(gather_batch(keys_only, as<int64_t>(cand), ncand, false));
    Buf local = op_argsort(*cb, kidx.data(), nkeys, descending, nulls_first);
    Buf out = dev_alloc((size_t)std::max<int64_t>(k, 1) * 8);
    map_indices<<<(unsigned)((k + 255) / 256), 256, 0, st>>>(as<int64_t>(cand), as<int64_t>(local), k, as<int64_t>(out));
    count_launch();
    op_done();
    return out;
}

// group slots ordered by their first row (first-occurrence group order, inline_agg.rs:804-807)
__global__ void __launch_bounds__(1024) small_sort_pairs_kernel(const unsigned long long *__restrict__ keys,
                                                                const int64_t *__restrict__ vals, int64_t n,
                                                                unsigned long long *__restrict__ keys_out, int64_t *__restrict__ vals_out) {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long k = keys[i];
        int64_t rank = 0;
        for (int64_t j = 0; j < n; j++) rank += (keys[j] < k) || (keys[j] == k && j < i);
        keys_out[rank] = k;
        vals_out[rank] = vals[i];
    }
}

Buf sort_slots_by_minrow(Buf slots, Buf minrow, int64_t ng) {
    if (ng <= 1) return slots;
    if (ng <= SMALL_SORT_MAX) {
        Buf k2 = dev_alloc((size_t)ng * 8), v2 = dev_alloc((size_t)ng * 8);
        small_sort_pairs_kernel<<<1, 1024, 0, ctx().stream>>>(as<unsigned long long>(minrow), as<int64_t>(slots), ng,
                                                             as<unsigned long long>(k2), as<int64_t>(v2));
        count_launch();
        B200_CUDA(cudaMemcpyAsync(minrow->ptr, k2->ptr, (size_t)ng * 8, cudaMemcpyDeviceToDevice, ctx().stream));
        return v2;
    }
    // minrow < 2^40 in practice; skipped passes make the high bytes free
    Buf keys = minrow;
    radix_sort_pairs(keys, slots, ng, 0, 64, false);
    if (keys.get() != minrow.get())
        B200_CUDA(cudaMemcpyAsync(minrow->ptr, keys->ptr, (size_t)ng * 8, cudaMemcpyDeviceToDevice, ctx().stream));
    return slots;
}

}  // namespace b200

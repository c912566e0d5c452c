// Library context, stream-ordered device memory, and the small fixed-function kernels every
// operator leans on: exclusive scan, bit counting, row gather (take), slice, concat.
#include <chrono>
#include <map>
#include <mutex>

#include "common.h"
#include "kernels.cuh"

namespace b200 {

std::atomic<int64_t> g_launches{0};

Context &ctx() {
    static Context c;
    return c;
}

void require_device() {
    if (!ctx().ready) fail("daft_b200: no CUDA device initialised (call b200_init); there is no CPU fallback");
}

// ---- per-kernel timing ---------------------------------------------------------------------------
namespace {
struct TimedSample {
    std::string name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};
std::mutex g_timer_mu;
bool g_timer_on = false;
std::vector<TimedSample> g_samples;
std::vector<cudaEvent_t> g_event_pool;
cudaEvent_t take_event() {
    if (!g_event_pool.empty()) {
        cudaEvent_t e = g_event_pool.back();
        g_event_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    B200_CUDA(cudaEventCreate(&e));
    return e;
}
}  // namespace
void kernel_timing_enable(bool on) {
    std::lock_guard<std::mutex> g(g_timer_mu);
    g_timer_on = on;
}
bool kernel_timing_enabled() { return g_timer_on; }
KernelTimer::KernelTimer(const char *name) {
    if (!g_timer_on) return;
    std::lock_guard<std::mutex> g(g_timer_mu);
    if (g_samples.size() >= 65536) return;  // bounded: a forgotten enable must not grow without limit
    TimedSample s;
    s.name = name;
    s.e0 = take_event();
    s.e1 = take_event();
    B200_CUDA(cudaEventRecord(s.e0, ctx().stream));
    slot = (int)g_samples.size();
    g_samples.push_back(std::move(s));
}
KernelTimer::~KernelTimer() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> g(g_timer_mu);
    if (slot < (int)g_samples.size()) cudaEventRecord(g_samples[slot].e1, ctx().stream);
}
std::string kernel_timing_report() {
    std::lock_guard<std::mutex> g(g_timer_mu);
    if (ctx().ready) B200_CUDA(cudaStreamSynchronize(ctx().stream));
    std::vector<std::string> names;
    std::vector<double> total;
    std::vector<int64_t> count;
    for (auto &s : g_samples) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.e0, s.e1) != cudaSuccess) ms = 0.f;
        size_t k = 0;
        while (k < names.size() && names[k] != s.name) k++;
        if (k == names.size()) {
            names.push_back(s.name);
            total.push_back(0.0);
            count.push_back(0);
        }
        total[k] += ms;
        count[k] += 1;
        g_event_pool.push_back(s.e0);
        g_event_pool.push_back(s.e1);
    }
    g_samples.clear();
    std::string out = "{";
    for (size_t k = 0; k < names.size(); k++) {
        char buf[96];
        snprintf(buf, sizeof buf, "\"launches\": %lld, \"ms\": %.6f", (long long)count[k], total[k]);
        out += (k ? ", \"" : "\"") + names[k] + "\": {" + buf + "}";
    }
    return out + "}";
}

const char *dtype_name(int32_t dt) {
    static const char *names[] = {"null", "bool", "int8", "int16", "int32", "int64", "uint8", "uint16", "uint32",
                                  "uint64", "float32", "float64", "date32", "timestamp", "large_utf8", "large_binary", "list"};
    return (dt >= 0 && dt < 17) ? names[dt] : "?";
}

// ---- device memory ---------------------------------------------------------------------------------------------
// Small blocks come straight from the stream-ordered pool (cudaMallocAsync).  Blocks of 1 MB and more go through a
// size-class cache in front of it: a freed block is parked in the bin of its (rounded) size and handed to the next
// request of that class.  Everything this library does is ordered on ONE stream, so a parked block may be reused at
// once -- the next user's kernels queue behind the previous user's.  Without the cache a repeated multi-GB pattern
// (join tables, selection vectors) makes the CUDA pool re-map physical memory every few iterations: 40 ms stalls on a
// 10 ms query.  Cached bytes are capped; a failed allocation empties the cache and retries.
namespace {
constexpr size_t kCacheMin = (size_t)1 << 20;
std::mutex g_cache_mu;
std::map<size_t, std::vector<void *>> g_cache;  // rounded size -> parked blocks
size_t g_cached_bytes = 0;
size_t cache_limit() {
    static const size_t lim = [] {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) return (size_t)16 << 30;
        return total_b / 2;
    }();
    return lim;
}
// size classes: 8 steps per power of two (<= 12.5 % slack)
size_t size_class(size_t bytes) {
    size_t p = kCacheMin;
    while (p < bytes) p <<= 1;
    const size_t step = p >> 4;  // p / 16: classes at p/2 + k * p/16
    const size_t lo = p >> 1;
    if (bytes <= lo) return lo;
    return lo + (bytes - lo + step - 1) / step * step;
}
void cache_trim_locked(size_t keep) {
    while (g_cached_bytes > keep && !g_cache.empty()) {
        auto it = std::prev(g_cache.end());  // largest class first
        void *p = it->second.back();
        it->second.pop_back();
        g_cached_bytes -= it->first;
        if (it->second.empty()) g_cache.erase(it);
        cudaFreeAsync(p, ctx().stream);
    }
}
}  // namespace

DevAlloc::~DevAlloc() {
    if (!owned || !ptr) return;
    if (cached_bytes) {
        std::lock_guard<std::mutex> g(g_cache_mu);
        g_cache[cached_bytes].push_back(ptr);
        g_cached_bytes += cached_bytes;
        if (g_cached_bytes > cache_limit()) cache_trim_locked(cache_limit() / 2);
        return;
    }
    cudaFreeAsync(ptr, ctx().stream);
}

Buf dev_alloc(size_t bytes) {
    require_device();
    auto a = std::make_shared<DevAlloc>();
    a->bytes = bytes;
    // +16 B slack: word-granular readers (unaligned string loads, bitmap words) may touch the tail
    const size_t want = bytes + 16;
    static const bool trace = [] { const char *e = getenv("B200_TRACE_ALLOC"); return e && *e == '1'; }();
    static const bool use_cache = [] { const char *e = getenv("B200_ALLOC_CACHE"); return !(e && *e == '0'); }();
    size_t cls = 0;
    if (use_cache && want >= kCacheMin) {
        cls = size_class(want);
        std::lock_guard<std::mutex> g(g_cache_mu);
        auto it = g_cache.find(cls);
        if (it != g_cache.end() && !it->second.empty()) {
            a->ptr = it->second.back();
            it->second.pop_back();
            if (it->second.empty()) g_cache.erase(it);
            g_cached_bytes -= cls;
            a->cached_bytes = cls;
            return a;
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMallocAsync(&a->ptr, cls ? cls : want, ctx().stream);
    if (e != cudaSuccess) {  // give the parked blocks back and try once more
        cudaGetLastError();
        {
            std::lock_guard<std::mutex> g(g_cache_mu);
            cache_trim_locked(0);
        }
        B200_CUDA(cudaStreamSynchronize(ctx().stream));
        B200_CUDA(cudaMallocAsync(&a->ptr, cls ? cls : want, ctx().stream));
    }
    a->cached_bytes = cls;
    if (trace) {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms > 0.5) fprintf(stderr, "[b200 alloc] %.1f MB took %.2f ms\n", bytes / 1e6, ms);
    }
    return a;
}
// Bytes an allocation could still obtain: what the driver reports free, plus the blocks parked in the size-class cache (a failed
// allocation releases them, see dev_alloc) and the part of the stream-ordered pool that is reserved but not in use (the pool hands
// it out again without asking the driver) -- cudaMemGetInfo alone counts both as used.
size_t dev_available_bytes() {
    size_t free_b = 0, total_b = 0;
    B200_CUDA(cudaMemGetInfo(&free_b, &total_b));
    size_t avail = free_b;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        avail += g_cached_bytes;
    }
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx().device) == cudaSuccess && pool) {
        uint64_t reserved = 0, used = 0;
        if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
            avail += (size_t)(reserved - used);
    }
    cudaGetLastError();
    return std::min(avail, total_b);
}
Buf dev_alloc_zero(size_t bytes) {
    Buf b = dev_alloc(bytes);
    B200_CUDA(cudaMemsetAsync(b->ptr, 0, bytes + 16, ctx().stream));
    return b;
}
Buf dev_borrow(const void *ptr, size_t bytes) {
    auto a = std::make_shared<DevAlloc>();
    a->ptr = const_cast<void *>(ptr);
    a->bytes = bytes;
    a->owned = false;
    return a;
}
Buf dev_sub(const Buf &parent, size_t offset, size_t bytes) {
    auto *v = new DevAlloc;
    v->ptr = static_cast<char *>(parent->ptr) + offset;
    v->bytes = bytes;
    v->owned = false;
    return Buf(v, [parent](DevAlloc *p) { delete p; });  // the view keeps its parent allocation alive
}
void op_done() {
    static const bool blocking = [] { const char *e = getenv("B200_SYNC_OPS"); return e && *e == '1'; }();
    if (blocking) B200_CUDA(cudaStreamSynchronize(ctx().stream));
    B200_CUDA(cudaPeekAtLastError());
}
void fill_u8(void *ptr, int value, size_t bytes) {
    if (bytes) B200_CUDA(cudaMemsetAsync(ptr, value, bytes, ctx().stream));
}

int64_t read_i64(const int64_t *dev_ptr) {
    int64_t v = 0;
    B200_CUDA(cudaMemcpyAsync(&v, dev_ptr, 8, cudaMemcpyDeviceToHost, ctx().stream));
    B200_CUDA(cudaStreamSynchronize(ctx().stream));
    return v;
}

// ------------------------------------------------------------------------------------------------
// exclusive scan (int64): reduce-then-scan, 2048 items per CTA, recursive on the CTA sums
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const int64_t *__restrict__ in, int64_t n,
                                                                int64_t *__restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    s = block_reduce_sum_i64<SCAN_THREADS>(s);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = s;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const int64_t *__restrict__ in, int64_t n,
                                                            const int64_t *__restrict__ tile_offsets,
                                                            int64_t *__restrict__ out, int64_t *total) {
    __shared__ int64_t warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        v[k] = i < n ? in[i] : 0;
        s += v[k];
    }
    // exclusive scan of per-thread sums across the CTA
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    int64_t woff = 0;
    for (int w = 0; w < warp; w++) woff += warp_tot[w];
    int64_t run = (tile_offsets ? tile_offsets[blockIdx.x] : 0) + woff + inc - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        int64_t i = base + k;
        if (i < n) out[i] = run;
        run += v[k];
    }
    if (total && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total = run;
}

void exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, int64_t *total_dev) {
    cudaStream_t st = ctx().stream;
    if (n <= 0) {
        if (total_dev) B200_CUDA(cudaMemsetAsync(total_dev, 0, 8, st));
        return;
    }
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles == 1) {
        scan_tiles<<<1, SCAN_THREADS, 0, st>>>(in, n, nullptr, out, total_dev);
        count_launch();
        return;
    }
    Buf sums = dev_alloc(tiles * 8), offs = dev_alloc(tiles * 8);
    scan_tile_sums<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, as<int64_t>(sums));
    count_launch();
    exclusive_scan_i64(as<int64_t>(sums), as<int64_t>(offs), tiles, nullptr);
    scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, as<int64_t>(offs), out, total_dev);
    count_launch();
}

// ------------------------------------------------------------------------------------------------
__global__ void popcount_kernel(const uint32_t *__restrict__ words, int64_t nwords, int tail_bits,
                                unsigned long long *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t s = 0;
    for (; i < nwords; i += stride) {
        uint32_t w = words[i];
        if (i == nwords - 1 && tail_bits) w &= (1u << tail_bits) - 1u;
        s += __popc(w);
    }
    s = block_reduce_sum_i64<256>(s);
    if (threadIdx.x == 0 && s) atomicAdd(out, (unsigned long long)s);
}

int64_t count_set_bits(const uint8_t *bitmap, int64_t nbits) {
    if (nbits <= 0) return 0;
    Buf out = dev_alloc_zero(8);
    int64_t nwords = (nbits + 31) / 32;
    int grid = (int)std::min<int64_t>((nwords + 255) / 256, ctx().sm_count * 8);
    popcount_kernel<<<grid, 256, 0, ctx().stream>>>(reinterpret_cast<const uint32_t *>(bitmap), nwords, (int)(nbits & 31),
                                                    as<unsigned long long>(out));
    count_launch();
    return read_i64(as<int64_t>(out));
}

// ------------------------------------------------------------------------------------------------
// gather (RecordBatch::take; src/daft-recordbatch/src/lib.rs take + growable/ in the reference)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void gather_fixed(const T *__restrict__ src, const int64_t *__restrict__ idx, int64_t n, T *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int64_t j = idx[i];
        out[i] = j >= 0 ? src[j] : T(0);
    }
}

// out bit i = (idx[i] >= 0) && bit(src, idx[i]);  src == nullptr means "all ones"
__global__ void gather_bits(const uint8_t *__restrict__ src, const int64_t *__restrict__ idx, int64_t n,
                            uint32_t *__restrict__ out_words) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n + 31) & ~int64_t(31);
    for (; i < n_round; i += stride) {
        bool bit = false;
        if (i < n) {
            int64_t j = idx[i];
            bit = j >= 0 && (src == nullptr || ((src[j >> 3] >> (j & 7)) & 1));
        }
        uint32_t w = __ballot_sync(0xffffffffu, bit);
        if ((threadIdx.x & 31) == 0) out_words[i >> 5] = w;
    }
}

__global__ void gather_str_len(const int64_t *__restrict__ offsets, const int64_t *__restrict__ idx, int64_t n,
                               int64_t *__restrict__ len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int64_t j = idx[i];
        len[i] = j >= 0 ? offsets[j + 1] - offsets[j] : 0;
    }
}
// list gather: positions of the gathered rows' elements in the flat child (8 lanes per row, like the string copy)
__global__ void gather_list_child_idx(const int64_t *__restrict__ src_off, const int64_t *__restrict__ idx, int64_t n,
                                      const int64_t *__restrict__ out_off, int64_t *__restrict__ child_idx) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
    const int sub = (int)(t & 7);
    for (int64_t i = t >> 3; i < n; i += stride) {
        const int64_t j = idx[i];
        if (j < 0) continue;
        const int64_t b = src_off[j], len = src_off[j + 1] - b, o = out_off[i];
        for (int64_t e = sub; e < len; e += 8) child_idx[o + e] = b + e;
    }
}
// 8 lanes cooperate on one row so that long strings still copy with some parallelism
__global__ void gather_str_copy(const int64_t *__restrict__ src_off, const uint8_t *__restrict__ src,
                                const int64_t *__restrict__ idx, int64_t n, const int64_t *__restrict__ out_off,
                                uint8_t *__restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 3;
    const int sub = (int)(t & 7);
    for (int64_t i = t >> 3; i < n; i += stride) {
        int64_t j = idx[i];
        if (j < 0) continue;
        int64_t s = src_off[j], len = src_off[j + 1] - s, d = out_off[i];
        for (int64_t k = sub; k < len; k += 8) out[d + k] = src[s + k];
    }
}

static int grid_for(int64_t n, int threads = 256) {
    int64_t g = (n + threads - 1) / threads;
    int64_t cap = (int64_t)ctx().sm_count * 16;
    return (int)std::max<int64_t>(1, std::min(g, cap));
}

Column gather_column(const Column &src, const int64_t *idx, int64_t n, bool idx_may_be_negative) {
    require_plain(src);
    cudaStream_t st = ctx().stream;
    Column out;
    out.name = src.name;
    out.dtype = src.dtype;
    out.length = n;
    if (src.dtype == B200_NULL) {
        out.null_count = n;
        return out;
    }
    const bool need_validity = src.null_count > 0 || idx_may_be_negative;
    if (need_validity && n > 0) {
        out.validity = dev_alloc(((n + 31) / 32) * 4);
        gather_bits<<<grid_for(n), 256, 0, st>>>(as<uint8_t>(src.validity), idx, n, as<uint32_t>(out.validity));
        count_launch();
        out.null_count = n - count_set_bits(as<uint8_t>(out.validity), n);
        if (out.null_count == 0) out.validity.reset();
    }
    int w = dtype_width(src.dtype);
    if (w) {
        out.data = dev_alloc((size_t)n * w);
        out.data_bytes = n * w;
        if (n) {
            int g = grid_for(n);
            switch (w) {
                case 1: gather_fixed<uint8_t><<<g, 256, 0, st>>>(as<uint8_t>(src.data), idx, n, as<uint8_t>(out.data)); break;
                case 2: gather_fixed<uint16_t><<<g, 256, 0, st>>>(as<uint16_t>(src.data), idx, n, as<uint16_t>(out.data)); break;
                case 4: gather_fixed<uint32_t><<<g, 256, 0, st>>>(as<uint32_t>(src.data), idx, n, as<uint32_t>(out.data)); break;
                default: gather_fixed<uint64_t><<<g, 256, 0, st>>>(as<uint64_t>(src.data), idx, n, as<uint64_t>(out.data)); break;
            }
            count_launch();
        }
    } else if (src.dtype == B200_BOOL) {
        out.data = dev_alloc(((n + 31) / 32) * 4);
        out.data_bytes = (n + 7) / 8;
        if (n) {
            gather_bits<<<grid_for(n), 256, 0, st>>>(as<uint8_t>(src.data), idx, n, as<uint32_t>(out.data));
            count_launch();
        }
    } else if (is_string(src.dtype)) {
        Buf len = dev_alloc((size_t)(n + 1) * 8);
        out.offsets = dev_alloc((size_t)(n + 1) * 8);
        if (n) {
            gather_str_len<<<grid_for(n), 256, 0, st>>>(as<int64_t>(src.offsets), idx, n, as<int64_t>(len));
            count_launch();
        }
        exclusive_scan_i64(as<int64_t>(len), as<int64_t>(out.offsets), n, as<int64_t>(out.offsets) + n);
        int64_t total = read_i64(as<int64_t>(out.offsets) + n);
        out.data = dev_alloc((size_t)total);
        out.data_bytes = total;
        if (n && total) {
            gather_str_copy<<<grid_for(n * 8), 256, 0, st>>>(as<int64_t>(src.offsets), as<uint8_t>(src.data), idx, n,
                                                              as<int64_t>(out.offsets), as<uint8_t>(out.data));
            count_launch();
        }
    } else if (src.dtype == B200_LIST) {
        if (!src.child) fail("gather: list column without a child");
        Buf len = dev_alloc((size_t)(n + 1) * 8);
        out.offsets = dev_alloc((size_t)(n + 1) * 8);
        if (n) {
            gather_str_len<<<grid_for(n), 256, 0, st>>>(as<int64_t>(src.offsets), idx, n, as<int64_t>(len));
            count_launch();
        }
        exclusive_scan_i64(as<int64_t>(len), as<int64_t>(out.offsets), n, as<int64_t>(out.offsets) + n);
        const int64_t total = read_i64(as<int64_t>(out.offsets) + n);
        Buf cidx = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
        if (n && total) {
            gather_list_child_idx<<<grid_for(n * 8), 256, 0, st>>>(as<int64_t>(src.offsets), idx, n, as<int64_t>(out.offsets), as<int64_t>(cidx));
            count_launch();
        }
        out.child = std::make_shared<const Column>(gather_column(*src.child, as<int64_t>(cidx), total, false));
    } else {
        fail(std::string("gather: unsupported dtype ") + dtype_name(src.dtype));
    }
    return out;
}

// Every fixed-width, never-null column of a batch gathered by ONE launch: the index is read once per row and the
// per-column copies are independent loads (memory-level parallelism across columns).
struct GatherDesc {
    const void *src;
    void *out;
    int width;
};
constexpr int kGatherMaxCols = 24;
struct GatherSet {
    GatherDesc d[kGatherMaxCols];
};
__global__ void gather_fixed_multi(const __grid_constant__ GatherSet gs, int ncols, const int64_t *__restrict__ idx, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t j = idx[i];
#pragma unroll 4
        for (int c = 0; c < ncols; c++) {
            const GatherDesc &d = gs.d[c];
            switch (d.width) {
                case 1: ((uint8_t *)d.out)[i] = j >= 0 ? ((const uint8_t *)d.src)[j] : (uint8_t)0; break;
                case 2: ((uint16_t *)d.out)[i] = j >= 0 ? ((const uint16_t *)d.src)[j] : (uint16_t)0; break;
                case 4: ((uint32_t *)d.out)[i] = j >= 0 ? ((const uint32_t *)d.src)[j] : 0u; break;
                default: ((uint64_t *)d.out)[i] = j >= 0 ? ((const uint64_t *)d.src)[j] : 0ULL; break;
            }
        }
    }
}

// String columns of a SMALL permutation gather (n <= 2048, e.g. the sort of an aggregation result): lengths, the
// offsets scan and the byte copy in one single-block launch per batch (blockIdx.x = string column); the output
// byte count of a permutation equals the input's, so no host round trip is needed to size the buffer.
struct GatherStrDesc {
    const int64_t *src_off;
    const uint8_t *src;
    int64_t *out_off;
    uint8_t *out;
};
struct GatherStrSet {
    GatherStrDesc d[kGatherMaxCols];
};
constexpr int kSmallGather = 2048;
__global__ void __launch_bounds__(1024) gather_str_small(const __grid_constant__ GatherStrSet gs, const int64_t *__restrict__ idx, int n) {
    __shared__ int64_t s_warp[32];
    const GatherStrDesc &d = gs.d[blockIdx.x];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // two consecutive rows per thread
    int64_t len0 = 0, len1 = 0, j0 = -1, j1 = -1;
    if (2 * t < n) { j0 = idx[2 * t]; len0 = d.src_off[j0 + 1] - d.src_off[j0]; }
    if (2 * t + 1 < n) { j1 = idx[2 * t + 1]; len1 = d.src_off[j1 + 1] - d.src_off[j1]; }
    int64_t x = len0 + len1;
    for (int k = 1; k < 32; k <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, k);
        if (lane >= k) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int64_t w = s_warp[lane];
        for (int k = 1; k < 32; k <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, w, k);
            if (lane >= k) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const int64_t excl = x - (len0 + len1) + (warp ? s_warp[warp - 1] : 0);
    if (2 * t < n) d.out_off[2 * t] = excl;
    if (2 * t + 1 < n) d.out_off[2 * t + 1] = excl + len0;
    if (2 * t == n - 1 || 2 * t + 1 == n - 1) d.out_off[n] = excl + len0 + len1;
    if (j0 >= 0) {
        const int64_t s = d.src_off[j0];
        for (int64_t k = 0; k < len0; k++) d.out[excl + k] = d.src[s + k];
    }
    if (j1 >= 0) {
        const int64_t s = d.src_off[j1];
        for (int64_t k = 0; k < len1; k++) d.out[excl + len0 + k] = d.src[s + k];
    }
}

Batch *gather_batch(const Batch &src, const int64_t *idx, int64_t n, bool idx_may_be_negative, bool permutation) {
    for (const Column &c : src.cols) require_plain(c);
    KernelTimer kt("gather_batch");
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    cudaStream_t st = ctx().stream;
    // fused paths first; whatever they do not cover goes column by column
    std::vector<char> done(src.cols.size(), 0);
    out->cols.resize(src.cols.size());
    if (n > 0) {
        GatherSet gs;
        int nf = 0;
        GatherStrSet ss;
        int ns = 0;
        const bool small_perm = permutation && !idx_may_be_negative && n <= kSmallGather && n == src.num_rows;
        auto flush_fixed = [&] {
            if (!nf) return;
            gather_fixed_multi<<<grid_for(n), 256, 0, st>>>(gs, nf, idx, n);
            count_launch();
            nf = 0;
        };
        for (size_t c = 0; c < src.cols.size(); c++) {
            const Column &sc = src.cols[c];
            const int w = dtype_width(sc.dtype);
            const bool no_validity = sc.null_count == 0 && !idx_may_be_negative;
            if (w && no_validity) {
                Column oc;
                oc.name = sc.name;
                oc.dtype = sc.dtype;
                oc.length = n;
                oc.data = dev_alloc((size_t)n * w);
                oc.data_bytes = n * w;
                gs.d[nf++] = GatherDesc{as<uint8_t>(sc.data), as<uint8_t>(oc.data), w};
                out->cols[c] = std::move(oc);
                done[c] = 1;
                if (nf == kGatherMaxCols) flush_fixed();
            } else if (is_string(sc.dtype) && no_validity && small_perm && ns < kGatherMaxCols) {
                Column oc;
                oc.name = sc.name;
                oc.dtype = sc.dtype;
                oc.length = n;
                oc.offsets = dev_alloc((size_t)(n + 1) * 8);
                oc.data = dev_alloc((size_t)std::max<int64_t>(sc.data_bytes, 1));
                oc.data_bytes = sc.data_bytes;
                ss.d[ns++] = GatherStrDesc{as<int64_t>(sc.offsets), as<uint8_t>(sc.data), as<int64_t>(oc.offsets), as<uint8_t>(oc.data)};
                out->cols[c] = std::move(oc);
                done[c] = 1;
            }
        }
        flush_fixed();
        if (ns) {
            gather_str_small<<<ns, 1024, 0, st>>>(ss, idx, (int)n);
            count_launch();
        }
    }
    for (size_t c = 0; c < src.cols.size(); c++)
        if (!done[c]) out->cols[c] = gather_column(src.cols[c], idx, n, idx_may_be_negative);
    return out.release();
}

// ------------------------------------------------------------------------------------------------
// views: selection vectors over plain columns (late materialisation of join outputs)
// ------------------------------------------------------------------------------------------------
struct ViewState {
    std::mutex mu;
    bool done = false;
    Column plain;
};

// out[i] = idx[i] >= 0 ? inner[idx[i]] : -1   (selection vector of a view of a view)
__global__ void compose_sel(const int64_t *__restrict__ inner, const int64_t *__restrict__ idx, int64_t n, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int64_t j = idx[i];
        out[i] = j >= 0 ? inner[j] : -1;
    }
}

static bool view_done(const Column &c) {
    if (!c.is_view()) return true;
    std::lock_guard<std::mutex> g(c.state->mu);
    return c.state->done;
}

Column materialize(const Column &c) {
    if (!c.is_view()) return c;
    std::lock_guard<std::mutex> g(c.state->mu);
    if (!c.state->done) {
        c.state->plain = gather_column(*c.base, as<int64_t>(c.sel), c.length, c.sel_may_be_negative);
        c.state->done = true;
    }
    Column out = c.state->plain;
    out.name = c.name;
    return out;
}

Batch materialize_batch(const Batch &b, const std::vector<int32_t> *cols) {
    Batch out = b;
    std::vector<int32_t> all;
    if (!cols) {
        for (size_t i = 0; i < b.cols.size(); i++) all.push_back((int32_t)i);
        cols = &all;
    }
    // pending views grouped by selection vector: one fused gather per group
    std::vector<int32_t> pending;
    for (int32_t ci : *cols) {
        if (ci < 0 || ci >= (int32_t)b.cols.size()) fail("column index out of range");
        if (b.cols[ci].is_view() && !view_done(b.cols[ci])) pending.push_back(ci);
    }
    std::vector<char> taken(pending.size(), 0);
    for (size_t a = 0; a < pending.size(); a++) {
        if (taken[a]) continue;
        const Column &ca = b.cols[pending[a]];
        Batch group;
        std::vector<size_t> members;
        for (size_t k = a; k < pending.size(); k++) {
            const Column &ck = b.cols[pending[k]];
            if (taken[k] || ck.sel.get() != ca.sel.get() || ck.length != ca.length) continue;
            bool dup = false;  // the same view listed twice (select of one column under two names)
            for (size_t m : members) dup |= b.cols[pending[m]].state.get() == ck.state.get();
            if (dup) continue;
            taken[k] = 1;
            members.push_back(k);
            group.cols.push_back(*ck.base);
        }
        bool neg = false;
        for (size_t m : members) neg |= b.cols[pending[m]].sel_may_be_negative;
        group.num_rows = group.cols.empty() ? 0 : group.cols[0].length;
        std::unique_ptr<Batch> g(gather_batch(group, as<int64_t>(ca.sel), ca.length, neg));
        for (size_t m = 0; m < members.size(); m++) {
            const Column &cm = b.cols[pending[members[m]]];
            std::lock_guard<std::mutex> lk(cm.state->mu);
            if (!cm.state->done) {
                cm.state->plain = std::move(g->cols[m]);
                cm.state->done = true;
            }
        }
    }
    for (int32_t ci : *cols)
        if (out.cols[ci].is_view()) out.cols[ci] = materialize(b.cols[ci]);
    return out;
}

std::vector<Column> make_view_columns(const Batch &src, const Buf &idx, int64_t n, bool idx_may_be_negative) {
    std::vector<Column> out;
    std::vector<std::pair<const DevAlloc *, Buf>> composed;  // inner selection vector -> composed vector
    for (const Column &sc : src.cols) {
        Column v;
        v.name = sc.name;
        v.dtype = sc.dtype;
        v.length = n;
        v.state = std::make_shared<ViewState>();
        if (sc.is_view() && !view_done(sc)) {
            Buf comp;
            for (auto &pr : composed)
                if (pr.first == sc.sel.get()) comp = pr.second;
            if (!comp) {
                comp = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
                KernelTimer kt("compose_sel");
                if (n > 0) {
                    compose_sel<<<grid_for(n), 256, 0, ctx().stream>>>(as<int64_t>(sc.sel), as<int64_t>(idx), n, as<int64_t>(comp));
                    count_launch();
                }
                composed.emplace_back(sc.sel.get(), comp);
            }
            v.base = sc.base;
            v.sel = comp;
            v.sel_may_be_negative = idx_may_be_negative || sc.sel_may_be_negative;
        } else {
            v.base = std::make_shared<const Column>(materialize(sc));
            v.sel = idx;
            v.sel_may_be_negative = idx_may_be_negative;
        }
        out.push_back(std::move(v));
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// index column (any integer dtype, nullable) -> int64 with null -> -1
template <typename T>
__global__ void widen_indices(const T *__restrict__ src, const uint8_t *__restrict__ validity, int64_t n,
                              int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        bool ok = validity == nullptr || ((validity[i >> 3] >> (i & 7)) & 1);
        out[i] = ok ? (int64_t)src[i] : -1;
    }
}

Buf indices_to_i64(const Column &c) {
    const int64_t n = c.length;
    cudaStream_t st = ctx().stream;
    if ((c.dtype == B200_I64 || c.dtype == B200_U64) && c.null_count == 0) return c.data;
    Buf out = dev_alloc((size_t)n * 8);
    if (n == 0) return out;
    const uint8_t *v = c.null_count ? as<uint8_t>(c.validity) : nullptr;
    int g = grid_for(n);
    switch (c.dtype) {
        case B200_I8: widen_indices<int8_t><<<g, 256, 0, st>>>(as<int8_t>(c.data), v, n, as<int64_t>(out)); break;
        case B200_I16: widen_indices<int16_t><<<g, 256, 0, st>>>(as<int16_t>(c.data), v, n, as<int64_t>(out)); break;
        case B200_I32: widen_indices<int32_t><<<g, 256, 0, st>>>(as<int32_t>(c.data), v, n, as<int64_t>(out)); break;
        case B200_I64: widen_indices<int64_t><<<g, 256, 0, st>>>(as<int64_t>(c.data), v, n, as<int64_t>(out)); break;
        case B200_U8: widen_indices<uint8_t><<<g, 256, 0, st>>>(as<uint8_t>(c.data), v, n, as<int64_t>(out)); break;
        case B200_U16: widen_indices<uint16_t><<<g, 256, 0, st>>>(as<uint16_t>(c.data), v, n, as<int64_t>(out)); break;
        case B200_U32: widen_indices<uint32_t><<<g, 256, 0, st>>>(as<uint32_t>(c.data), v, n, as<int64_t>(out)); break;
        case B200_U64: widen_indices<uint64_t><<<g, 256, 0, st>>>(as<uint64_t>(c.data), v, n, as<int64_t>(out)); break;
        default: fail(std::string("take: indices must be an integer column, got ") + dtype_name(c.dtype));
    }
    count_launch();
    return out;
}

// Caller-supplied take indices are validated before the gather dereferences them: the reference raises "index out of bounds"
// (tests/recordbatch/test_take.py:31-34); indices produced inside the library (sort, join, partition) skip this.
__global__ void check_indices_kernel(const int64_t *__restrict__ idx, int64_t n, int64_t limit, bool allow_negative, int *flag) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (; i < n; i += stride) {
        const int64_t j = idx[i];
        bad |= j >= limit || (j < 0 && !allow_negative);
    }
    if (bad) *flag = 1;
}

void check_take_indices(const int64_t *idx, int64_t n, int64_t limit, bool null_marked) {
    if (n == 0) return;
    cudaStream_t st = ctx().stream;
    Buf flag = dev_alloc_zero(sizeof(int));
    check_indices_kernel<<<grid_for(n), 256, 0, st>>>(idx, n, limit, null_marked, as<int>(flag));
    count_launch();
    B200_CUDA(cudaGetLastError());
    int h = 0;
    B200_CUDA(cudaMemcpyAsync(&h, as<int>(flag), sizeof(int), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (h) fail("take: index out of bounds (the batch has " + std::to_string(limit) + " rows)");
}

Column make_u64_column(const std::string &name, Buf data, int64_t n) {
    Column c;
    c.name = name;
    c.dtype = B200_U64;
    c.length = n;
    c.data = std::move(data);
    c.data_bytes = n * 8;
    return c;
}

// ------------------------------------------------------------------------------------------------
// slice / concat
// ------------------------------------------------------------------------------------------------
__global__ void iota_i64(int64_t *out, int64_t n, int64_t start) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = start + i;
}

Column slice_column(const Column &src, int64_t start, int64_t end) {
    // byte-aligned fixed-width slices without nulls are views; everything else goes through gather
    const int64_t n = end - start;
    if (src.is_view() && !view_done(src)) {  // a slice of a selection vector
        Column out = src;
        out.length = n;
        out.sel = dev_sub(src.sel, (size_t)start * 8, (size_t)n * 8);
        out.state = std::make_shared<ViewState>();
        return out;
    }
    if (src.is_view()) return slice_column(materialize(src), start, end);
    int w = dtype_width(src.dtype);
    if (w && src.null_count == 0) {
        Column out = src;
        out.length = n;
        out.data = dev_borrow(as<uint8_t>(src.data) + start * w, (size_t)n * w);
        // keep the parent alive
        auto parent = src.data;
        out.data = Buf(out.data.get(), [parent, keep = out.data](DevAlloc *) {});
        out.data_bytes = n * w;
        return out;
    }
    Buf idx = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    if (n) {
        iota_i64<<<grid_for(n), 256, 0, ctx().stream>>>(as<int64_t>(idx), n, start);
        count_launch();
    }
    return gather_column(src, as<int64_t>(idx), n, false);
}

__global__ void shift_offsets(const int64_t *__restrict__ src, int64_t n, int64_t delta, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = src[i] - src[0] + delta;
}
// copy `nbits` bits from src (bit 0 aligned) into dst starting at bit `dst_bit`; src==nullptr -> ones.
// One thread per destination byte: each output byte is owned by exactly one (part, thread) pair only
// when parts are byte aligned, so the host runs parts one after another on the same stream and the
// kernel ORs into bytes that straddle a boundary.
__global__ void copy_bits(const uint8_t *__restrict__ src, int64_t nbits, uint8_t *__restrict__ dst, int64_t dst_bit) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // bit index within part, step 1 (simple, rarely hot)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nbits; i += stride) {
        bool bit = src == nullptr || ((src[i >> 3] >> (i & 7)) & 1);
        if (bit) {
            int64_t d = dst_bit + i;
            atomicOr(reinterpret_cast<unsigned int *>(dst) + (d >> 5), 1u << (d & 31));
        }
    }
}

Column concat_columns(const std::vector<const Column *> &cols) {
    for (auto *c : cols) require_plain(*c);
    cudaStream_t st = ctx().stream;
    Column out;
    out.name = cols[0]->name;
    out.dtype = cols[0]->dtype;
    int64_t n = 0, nulls = 0, bytes = 0;
    for (auto *c : cols) {
        if (c->dtype != out.dtype) fail("concat: dtype mismatch in column " + out.name);
        n += c->length;
        nulls += c->null_count;
        bytes += c->data_bytes;
    }
    out.length = n;
    out.null_count = nulls;
    if (out.dtype == B200_NULL) return out;
    if (nulls > 0) {
        out.validity = dev_alloc_zero(((n + 31) / 32) * 4);
        int64_t pos = 0;
        for (auto *c : cols) {
            if (c->length) {
                copy_bits<<<grid_for(c->length), 256, 0, st>>>(c->null_count ? as<uint8_t>(c->validity) : nullptr, c->length,
                                                                as<uint8_t>(out.validity), pos);
                count_launch();
            }
            pos += c->length;
        }
    }
    int w = dtype_width(out.dtype);
    if (w) {
        out.data = dev_alloc((size_t)n * w);
        out.data_bytes = n * w;
        int64_t pos = 0;
        for (auto *c : cols) {
            if (c->length)
                B200_CUDA(cudaMemcpyAsync(as<uint8_t>(out.data) + pos * w, as<uint8_t>(c->data), (size_t)c->length * w,
                                          cudaMemcpyDeviceToDevice, st));
            pos += c->length;
        }
    } else if (out.dtype == B200_BOOL) {
        out.data = dev_alloc_zero(((n + 31) / 32) * 4);
        out.data_bytes = (n + 7) / 8;
        int64_t pos = 0;
        for (auto *c : cols) {
            if (c->length) {
                copy_bits<<<grid_for(c->length), 256, 0, st>>>(as<uint8_t>(c->data), c->length, as<uint8_t>(out.data), pos);
                count_launch();
            }
            pos += c->length;
        }
    } else if (is_string(out.dtype)) {
        out.offsets = dev_alloc((size_t)(n + 1) * 8);
        out.data = dev_alloc((size_t)bytes);
        out.data_bytes = bytes;
        int64_t pos = 0, bpos = 0;
        for (auto *c : cols) {
            if (c->length) {
                shift_offsets<<<grid_for(c->length), 256, 0, st>>>(as<int64_t>(c->offsets), c->length, bpos,
                                                                    as<int64_t>(out.offsets) + pos);
                count_launch();
                // data of a (possibly sliced) column starts at offsets[0]; data_bytes counts from there
                if (c->data_bytes)
                    B200_CUDA(cudaMemcpyAsync(as<uint8_t>(out.data) + bpos, as<uint8_t>(c->data), (size_t)c->data_bytes,
                                              cudaMemcpyDeviceToDevice, st));
            }
            pos += c->length;
            bpos += c->data_bytes;
        }
        B200_CUDA(cudaMemcpyAsync(as<int64_t>(out.offsets) + n, &bytes, 8, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaStreamSynchronize(st));  // `bytes` is a stack variable
    } else if (out.dtype == B200_LIST) {
        // lists built by this library start at offset 0 and own exactly their child (aggregates, gathers): the parts' offsets
        // are shifted by the elements before them and the children concatenated
        out.offsets = dev_alloc((size_t)(n + 1) * 8);
        int64_t pos = 0, epos = 0;
        std::vector<const Column *> kids;
        for (auto *c : cols) {
            if (!c->child) fail("concat: list column without a child");
            if (c->length) {
                shift_offsets<<<grid_for(c->length), 256, 0, st>>>(as<int64_t>(c->offsets), c->length, epos, as<int64_t>(out.offsets) + pos);
                count_launch();
            }
            pos += c->length;
            epos += c->child->length;
            kids.push_back(c->child.get());
        }
        B200_CUDA(cudaMemcpyAsync(as<int64_t>(out.offsets) + n, &epos, 8, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaStreamSynchronize(st));  // `epos` is a stack variable
        out.child = std::make_shared<const Column>(concat_columns(kids));
    } else {
        fail(std::string("concat: unsupported dtype ") + dtype_name(out.dtype));
    }
    return out;
}

}  // namespace b200

// a1/a2: per-column XXH3-64 row hashing with Daft's null and seed-chaining rules.
//
// Replaces kernels::hashing::hash (src/daft-core/src/kernels/hashing.rs:304-445) and
// RecordBatch::hash_rows (src/daft-recordbatch/src/ops/hash.rs:14-26).  Rules reproduced:
//   - primitives hash the native-width little-endian bytes (hashing.rs:462-501); a null hashes to
//     xxh3_64(b"") and IGNORES the seed (SEEDED_NULL_USES_DEFAULT, hashing.rs:73-83,231-240);
//   - bool hashes ASCII "1"/"0" (hashing.rs:512-552), null as primitives;
//   - LargeUtf8/LargeBinary/Null arrays: a null is the empty slice hashed WITH the row seed
//     (hash_bytes_iter_inner, hashing.rs:243-265; pinned by tests/series/test_hash.py:61-73);
//   - hash_rows chains columns: the previous column's hash is the next column's per-row seed.
//
// HBM-bound: 8 B (or 4/2/1 B) read + 8 B seed read + 8 B written per row per key column; each
// warp reads 32 consecutive values (coalesced 256 B) per instruction.
#include <cstdlib>
#include "common.h"
#include "kernels.cuh"
#include "xxh3.cuh"

namespace b200 {

constexpr uint64_t kNullHash = 0x2D06800538D394C2ULL;  // xxh3_64(b"", 0) = 3244421341483603138
static_assert(kNullHash == 3244421341483603138ULL, "null hash golden (tests/recordbatch/test_hash.py:17)");

template <typename T>
__device__ __forceinline__ uint64_t hash_value(T v, uint64_t seed);
template <> __device__ __forceinline__ uint64_t hash_value<uint8_t>(uint8_t v, uint64_t s) { return xxh::hash_u8(v, s); }
template <> __device__ __forceinline__ uint64_t hash_value<uint16_t>(uint16_t v, uint64_t s) { return xxh::hash_u16(v, s); }
template <> __device__ __forceinline__ uint64_t hash_value<uint32_t>(uint32_t v, uint64_t s) { return xxh::hash_u32(v, s); }
template <> __device__ __forceinline__ uint64_t hash_value<uint64_t>(uint64_t v, uint64_t s) { return xxh::hash_u64(v, s); }

template <typename T>
__global__ void __launch_bounds__(256) hash_fixed_kernel(const T *__restrict__ data, const uint8_t *__restrict__ validity,
                                                          const uint64_t *seed, uint64_t *out,
                                                          int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t s = seed ? seed[i] : 0;
        uint64_t h;
        if (validity && !bit_at(validity, i)) h = kNullHash;
        else h = hash_value<T>(data[i], s);
        out[i] = h;
    }
}

__global__ void __launch_bounds__(256) hash_bool_kernel(const uint8_t *__restrict__ data, const uint8_t *__restrict__ validity,
                                                         const uint64_t *seed, uint64_t *out,
                                                         int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t s = seed ? seed[i] : 0;
        uint64_t h;
        if (validity && !bit_at(validity, i)) h = kNullHash;
        else h = xxh::hash_u8(bit_at(data, i) ? '1' : '0', s);
        out[i] = h;
    }
}

__global__ void __launch_bounds__(256) hash_null_kernel(const uint64_t *seed, uint64_t *out, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = xxh::hash_empty(seed ? seed[i] : 0);
}

// 8 bytes at any alignment from three aligned words (reads at most 3 bytes before p and 11 past p + 8: inside the
// allocation's 16-byte tail slack, see common.h)
__device__ __forceinline__ uint64_t ld_le8(const uint8_t *p) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const unsigned sh = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3) * 8u;
    const uint32_t a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);
    return (uint64_t)__funnelshift_r(a, b, sh) | ((uint64_t)__funnelshift_r(b, c, sh) << 32);
}

// Strings / binary: HB_U rows per thread, phase 1 = offsets (and seeds) of all rows, phase 2 = the first 8 bytes of
// all rows (branch-free aligned words), phase 3 = the hashes.  Values up to 8 bytes (flags, codes, short names) are
// hashed from registers; 9..16 bytes need one more load; longer values take the generic byte reader.
template <int HB_U, int MINB>
__global__ void __launch_bounds__(256, MINB) hash_bytes_kernel(const int64_t *__restrict__ offsets, const uint8_t *__restrict__ data,
                                                          const uint8_t *__restrict__ validity, const uint64_t *seed,
                                                          uint64_t *out, int64_t n) {
    const int64_t tile = (int64_t)blockDim.x * HB_U;
    for (int64_t base = (int64_t)blockIdx.x * tile; base < n; base += (int64_t)gridDim.x * tile) {
        int64_t o[HB_U], len[HB_U];
        uint64_t s[HB_U], first8[HB_U];
        bool ok[HB_U];
#pragma unroll
        for (int u = 0; u < HB_U; u++) {
            const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
            o[u] = 0;
            len[u] = 0;
            s[u] = 0;
            ok[u] = false;
            if (i < n) {
                o[u] = offsets[i];
                len[u] = offsets[i + 1] - o[u];
                s[u] = seed ? seed[i] : 0;
                ok[u] = !(validity && !bit_at(validity, i));
            }
        }
#pragma unroll
        for (int u = 0; u < HB_U; u++) first8[u] = (ok[u] && len[u] > 0 && len[u] <= 16) ? ld_le8(data + o[u]) : 0ULL;
#pragma unroll
        for (int u = 0; u < HB_U; u++) {
            const int64_t i = base + (int64_t)u * blockDim.x + threadIdx.x;
            if (i >= n) continue;
            const int64_t L = len[u];
            uint64_t h;
            if (!ok[u] || L == 0) {
                h = xxh::hash_empty(s[u]);  // a null string hashes the empty slice WITH the seed (hashing.rs:243-265)
            } else if (L <= 3) {
                const uint32_t w = (uint32_t)first8[u];
                h = xxh::hash_1to3(w & 0xff, (w >> (8 * (unsigned)(L >> 1))) & 0xff, (w >> (8 * (unsigned)(L - 1))) & 0xff, (uint32_t)L, s[u]);
            } else if (L <= 8) {
                h = xxh::hash_4to8((uint32_t)first8[u], (uint32_t)(first8[u] >> (8 * (unsigned)(L - 4))), (uint32_t)L, s[u]);
            } else if (L <= 16) {
                h = xxh::hash_9to16(first8[u], ld_le8(data + o[u] + L - 8), (uint64_t)L, s[u]);
            } else {
                h = xxh::hash_bytes(data + o[u], L, s[u]);
            }
            out[i] = h;
        }
    }
}


void hash_column(const Column &c, const uint64_t *seed, uint64_t *out) {
    const int64_t n = c.length;
    if (n == 0) return;
    cudaStream_t st = ctx().stream;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 16);
    const uint8_t *v = c.null_count > 0 ? as<uint8_t>(c.validity) : nullptr;
    switch (dtype_width(c.dtype)) {
        case 1: hash_fixed_kernel<uint8_t><<<grid, 256, 0, st>>>(as<uint8_t>(c.data), v, seed, out, n); break;
        case 2: hash_fixed_kernel<uint16_t><<<grid, 256, 0, st>>>(as<uint16_t>(c.data), v, seed, out, n); break;
        case 4: hash_fixed_kernel<uint32_t><<<grid, 256, 0, st>>>(as<uint32_t>(c.data), v, seed, out, n); break;
        case 8: hash_fixed_kernel<uint64_t><<<grid, 256, 0, st>>>(as<uint64_t>(c.data), v, seed, out, n); break;
        default:
            if (c.dtype == B200_BOOL) hash_bool_kernel<<<grid, 256, 0, st>>>(as<uint8_t>(c.data), v, seed, out, n);
            else if (c.dtype == B200_NULL) hash_null_kernel<<<grid, 256, 0, st>>>(seed, out, n);
            else if (is_string(c.dtype))
            {
                // 4 rows per thread at <= 64 registers (4 CTAs / SM) measured best on B200: 0.76 ms for two 60 M-row flag columns
                // vs 1.03 ms for one row per thread.  Tried and slower: staging a warp's 128 values through shared memory with
                // coalesced loads (0.89 ms), a speculative fixed-length fetch (spills).
                // The spill bytes ptxas reports for this kernel sit around the out-of-line call for values of more than 16 bytes
                // (caller-saved registers): every HB_U x MINB shape has them, none executes them on short values.
                hash_bytes_kernel<4, 4><<<grid, 256, 0, st>>>(as<int64_t>(c.offsets), as<uint8_t>(c.data), v, seed, out, n);
            }
            else fail(std::string("Hash not implemented for type ") + dtype_name(c.dtype));
    }
    count_launch();
    B200_CUDA(cudaGetLastError());
}

Buf hash_rows(const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint64_t *seed) {
    if (nkeys < 1) fail("hash_rows: need at least one key column");
    const int64_t n = b.num_rows;
    Buf out = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    for (int k = 0; k < nkeys; k++) {
        int32_t ci = key_cols[k];
        if (ci < 0 || ci >= (int32_t)b.cols.size()) fail("hash_rows: key column index out of range");
        // in-place chaining is safe: thread i reads seed[i] before writing out[i]
        hash_column(b.cols[ci], k == 0 ? seed : as<uint64_t>(out), as<uint64_t>(out));
    }
    return out;
}

}  // namespace b200

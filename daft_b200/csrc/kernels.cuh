// Device helpers shared by the ahead-of-time kernels.
#pragma once
#include <cstdint>

namespace b200 {

__device__ __forceinline__ int64_t warp_reduce_sum_i64(int64_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

template <int THREADS>
__device__ __forceinline__ int64_t block_reduce_sum_i64(int64_t v) {
    __shared__ int64_t s_part[THREADS / 32];
    v = warp_reduce_sum_i64(v);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    int64_t r = 0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < THREADS / 32 ? s_part[threadIdx.x] : 0;
        r = warp_reduce_sum_i64(r);
    }
    __syncthreads();
    return r;  // valid in warp 0
}

__device__ __forceinline__ bool bit_at(const uint8_t *bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }


// streaming loads/stores that do not pollute L1 (data touched once)
__device__ __forceinline__ uint64_t ldg_stream_u64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

}  // namespace b200

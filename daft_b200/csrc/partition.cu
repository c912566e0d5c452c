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

std::vector<Batch *> op_partition_by_hash(const Batch &b, const int32_t *key_cols, int32_t nkeys, int32_t P) {
    require_device();
    if (P <= 0) fail("Can not partition a Table by 0 partitions");
    Buf h = hash_rows(b, key_cols, nkeys, nullptr);
    std::vector<int64_t> offsets;
    Buf perm = partition_permutation(as<uint64_t>(h), b.num_rows, P, offsets);
    std::vector<Batch *> out;
    try {
        for (int p = 0; p < P; p++)
            out.push_back(gather_batch(b, as<int64_t>(perm) + offsets[p], offsets[p + 1] - offsets[p], false));
        B200_CUDA(cudaStreamSynchronize(ctx().stream));
    } catch (...) {
        for (auto *x : out) delete x;
        throw;
    }
    return out;
}

// distinct rows over the key columns, first occurrence kept, first-occurrence order
Batch *op_agg_first_rows(const Batch &in, const uint8_t *plan, size_t len, Buf *first_rows, int64_t *ngroups);  // ops_jit.cu

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
    B200_CUDA(cudaStreamSynchronize(ctx().stream));
    return out;
}

}  // namespace b200

// Host orchestration of the JIT-specialised operators: Project, Filter, fused HashAggregate.
#include <algorithm>
#include <mutex>

#include "common.h"
#include "jit.h"
#include "kernels.cuh"
#include "ops.h"

namespace b200 {

using jit::Compiled;
using jit::LaunchArgs;

static int stream_grid(int64_t n, int threads, int per_sm) {
    int64_t g = (n + threads - 1) / threads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)ctx().sm_count * per_sm));
}

// ------------------------------------------------------------------------------------------------
// Project (RecordBatch::eval_expression_list, src/daft-recordbatch/src/lib.rs:1603-1631)
// optional take_idx: evaluate the expressions at rows take_idx[i] (negative -> null row)
// ------------------------------------------------------------------------------------------------
static Batch *op_project_flat(const Batch &in, const uint8_t *plan, size_t len, const int64_t *take_idx, int64_t take_n, bool take_nonneg);

// List columns are not readable by expressions, but a projection may carry them along as bare column references (the final
// projection of a two-stage list aggregate does): those are passed through -- shared, or gathered with `take` -- and only the
// other expressions are compiled.
Batch *op_project(const Batch &in, const uint8_t *plan, size_t len, const int64_t *take_idx, int64_t take_n, bool take_nonneg) {
    bool any_list = false;
    for (const Column &c : in.cols) any_list = any_list || c.dtype == B200_LIST;
    if (!any_list) return op_project_flat(in, plan, len, take_idx, take_n, take_nonneg);
    const std::vector<jit::ProjectEntry> ents = jit::project_entries(plan, len);
    std::vector<uint8_t> rest(4);
    uint32_t nrest = 0;
    std::vector<int> where(ents.size(), -1);
    for (size_t i = 0; i < ents.size(); i++) {
        const int32_t bc = ents[i].bare_col;
        if (bc >= 0 && bc < (int32_t)in.cols.size() && in.cols[bc].dtype == B200_LIST) continue;
        where[i] = (int)nrest++;
        rest.insert(rest.end(), plan + ents[i].begin, plan + ents[i].end);
    }
    memcpy(rest.data(), &nrest, 4);
    std::unique_ptr<Batch> flat;
    if (nrest) flat.reset(op_project_flat(in, rest.data(), rest.size(), take_idx, take_n, take_nonneg));
    auto out = std::make_unique<Batch>();
    out->num_rows = take_idx ? take_n : in.num_rows;
    for (size_t i = 0; i < ents.size(); i++) {
        if (where[i] >= 0) {
            out->cols.push_back(std::move(flat->cols[(size_t)where[i]]));
        } else {
            Column c = take_idx ? gather_column(materialize(in.cols[ents[i].bare_col]), take_idx, take_n, !take_nonneg) : in.cols[ents[i].bare_col];
            c.name = ents[i].name;
            out->cols.push_back(std::move(c));
        }
    }
    return out.release();
}

static Batch *op_project_flat(const Batch &in, const uint8_t *plan, size_t len, const int64_t *take_idx, int64_t take_n, bool take_nonneg) {
    require_device();
    const bool with_take = take_idx != nullptr;
    const Compiled &k = jit::get(B200_OP_PROJECT, plan, len, in, with_take);
    const int64_t n = with_take ? take_n : in.num_rows;
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    const size_t nout = k.outs.size();
    std::vector<Buf> lens(nout);
    for (size_t i = 0; i < nout; i++) {
        Column c;
        c.name = k.outs[i].name;
        c.dtype = k.outs[i].dtype;
        c.length = n;
        if (c.dtype == B200_NULL) {
            c.null_count = n;
        } else {
            if (k.outs[i].nullable) c.validity = dev_alloc((size_t)((n + 31) / 32) * 4);
            int w = dtype_width(c.dtype);
            if (w) {
                c.data = dev_alloc((size_t)n * w);
                c.data_bytes = n * w;
            } else if (c.dtype == B200_BOOL) {
                c.data = dev_alloc((size_t)((n + 31) / 32) * 4);
                c.data_bytes = (n + 7) / 8;
            } else {
                lens[i] = dev_alloc((size_t)(n + 1) * 8);
                c.offsets = dev_alloc((size_t)(n + 1) * 8);
            }
        }
        out->cols.push_back(std::move(c));
    }
    auto fill = [&](LaunchArgs &a) {
        a.ptr.assign(k.ptrs.size(), nullptr);
        a.ival.assign(k.ints.size(), 0);
        for (size_t s = 0; s < k.ptrs.size(); s++) {
            const auto &ps = k.ptrs[s];
            switch (ps.kind) {
                case jit::P_COL_DATA: a.ptr[s] = in.cols[ps.index].data ? in.cols[ps.index].data->ptr : nullptr; break;
                case jit::P_COL_VALID: a.ptr[s] = in.cols[ps.index].validity ? in.cols[ps.index].validity->ptr : nullptr; break;
                case jit::P_COL_OFFSETS: a.ptr[s] = in.cols[ps.index].offsets ? in.cols[ps.index].offsets->ptr : nullptr; break;
                case jit::P_OUT_DATA: a.ptr[s] = out->cols[ps.index].data ? out->cols[ps.index].data->ptr : nullptr; break;
                case jit::P_OUT_VALID: a.ptr[s] = out->cols[ps.index].validity ? out->cols[ps.index].validity->ptr : nullptr; break;
                case jit::P_OUT_LEN: a.ptr[s] = lens[ps.index] ? lens[ps.index]->ptr : nullptr; break;
                case jit::P_OUT_OFFSETS: a.ptr[s] = out->cols[ps.index].offsets ? out->cols[ps.index].offsets->ptr : nullptr; break;
                case jit::P_TAKE: a.ptr[s] = const_cast<int64_t *>(take_idx); break;
                default: fail("project: unexpected parameter kind");
            }
        }
        for (size_t s = 0; s < k.ints.size(); s++)
            if (k.ints[s] == jit::I_N) a.ival[s] = n;
    };
    if (n > 0) {
        LaunchArgs a;
        fill(a);
        const int grid = stream_grid(n, 256, 8);
        jit::launch(k, a, grid, 256, 0, 0, false);
        if (k.has_string_out) {
            for (size_t i = 0; i < nout; i++) {
                Column &c = out->cols[i];
                if (!is_string(c.dtype)) continue;
                exclusive_scan_i64(as<int64_t>(lens[i]), as<int64_t>(c.offsets), n, as<int64_t>(c.offsets) + n);
                int64_t total = read_i64(as<int64_t>(c.offsets) + n);
                c.data = dev_alloc((size_t)total);
                c.data_bytes = total;
            }
            LaunchArgs a2;
            fill(a2);
            jit::launch(k, a2, grid, 256, 0, 1, false);
        }
    } else {
        for (auto &c : out->cols)
            if (is_string(c.dtype)) {
                fill_u8(c.offsets->ptr, 0, 8);
                c.data = dev_alloc(0);
            }
    }
    for (size_t i = 0; i < nout; i++) {
        Column &c = out->cols[i];
        if (c.dtype != B200_NULL && c.validity) {
            if (with_take && take_nonneg && !k.outs[i].expr_nullable) {
                c.validity.reset();  // every bit is set by construction: skip the count (and its sync)
                continue;
            }
            c.null_count = n - count_set_bits(as<uint8_t>(c.validity), n);
            if (c.null_count == 0) c.validity.reset();
        }
    }
    return out.release();
}

// ------------------------------------------------------------------------------------------------
// Filter (RecordBatch::filter, lib.rs:520-570): predicate -> bitmap -> row indices -> gather
// ------------------------------------------------------------------------------------------------
__global__ void word_popc(const uint32_t *__restrict__ words, int64_t nwords, int64_t nbits, uint32_t flip, int64_t *__restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nwords; i += stride) {
        uint32_t w = words[i] ^ flip;
        if (i == nwords - 1 && (nbits & 31)) w &= (1u << (nbits & 31)) - 1u;
        counts[i] = __popc(w);
    }
}
__global__ void expand_bits(const uint32_t *__restrict__ words, int64_t nwords, int64_t nbits, uint32_t flip,
                            const int64_t *__restrict__ offsets, int64_t *__restrict__ out_idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nwords; i += stride) {
        uint32_t w = words[i] ^ flip;
        if (i == nwords - 1 && (nbits & 31)) w &= (1u << (nbits & 31)) - 1u;
        int64_t o = offsets[i];
        while (w) {
            int b = __ffs(w) - 1;
            out_idx[o++] = i * 32 + b;
            w &= w - 1;
        }
    }
}

Buf bitmap_select(const uint32_t *words, int64_t n, bool invert, int64_t *count_out) {
    const int64_t nwords = (n + 31) / 32;
    const uint32_t flip = invert ? 0xFFFFFFFFu : 0u;
    Buf counts = dev_alloc((size_t)(nwords + 1) * 8), offs = dev_alloc((size_t)(nwords + 1) * 8);
    cudaStream_t st = ctx().stream;
    const int grid = stream_grid(nwords, 256, 8);
    word_popc<<<grid, 256, 0, st>>>(words, nwords, n, flip, as<int64_t>(counts));
    count_launch();
    exclusive_scan_i64(as<int64_t>(counts), as<int64_t>(offs), nwords, as<int64_t>(offs) + nwords);
    const int64_t total = read_i64(as<int64_t>(offs) + nwords);
    Buf idx = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    if (total) {
        expand_bits<<<grid, 256, 0, st>>>(words, nwords, n, flip, as<int64_t>(offs), as<int64_t>(idx));
        count_launch();
    }
    *count_out = total;
    return idx;
}

Batch *op_filter(const Batch &in, const uint8_t *plan, size_t len) {
    require_device();
    cudaStream_t st = ctx().stream;
    const int64_t n = in.num_rows;
    if (n == 0) return gather_batch(in, nullptr, 0, false);
    const Compiled &k = jit::get(B200_OP_FILTER, plan, len, in, false);
    const int64_t ntiles = (n + 2047) / 2048;
    Buf counts = dev_alloc((size_t)(ntiles + 1) * 8), offs = dev_alloc((size_t)(ntiles + 1) * 8);
    auto out = std::make_unique<Batch>();
    Buf idx;
    auto fill = [&](LaunchArgs &a, Buf &cnt) {
        a.ptr.assign(k.ptrs.size(), nullptr);
        a.ival.assign(k.ints.size(), 0);
        for (size_t s = 0; s < k.ptrs.size(); s++) {
            const auto &ps = k.ptrs[s];
            switch (ps.kind) {
                case jit::P_COL_DATA: a.ptr[s] = in.cols[ps.index].data ? in.cols[ps.index].data->ptr : nullptr; break;
                case jit::P_COL_VALID: a.ptr[s] = in.cols[ps.index].validity ? in.cols[ps.index].validity->ptr : nullptr; break;
                case jit::P_COL_OFFSETS: a.ptr[s] = in.cols[ps.index].offsets ? in.cols[ps.index].offsets->ptr : nullptr; break;
                case jit::P_OUT_LEN: a.ptr[s] = cnt->ptr; break;
                case jit::P_OUT_DATA: a.ptr[s] = out->cols.size() > (size_t)ps.index && out->cols[ps.index].data ? out->cols[ps.index].data->ptr : nullptr; break;
                case jit::P_TAKE: a.ptr[s] = idx ? idx->ptr : nullptr; break;
                default: fail("filter: unexpected parameter kind");
            }
        }
        for (size_t s = 0; s < k.ints.size(); s++)
            if (k.ints[s] == jit::I_N) a.ival[s] = n;
    };
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)ctx().sm_count * 8));
    {
        LaunchArgs a;
        fill(a, counts);
        jit::launch(k, a, grid, 256, 0, 0, false);
    }
    exclusive_scan_i64(as<int64_t>(counts), as<int64_t>(offs), ntiles, as<int64_t>(offs) + ntiles);
    const int64_t total = read_i64(as<int64_t>(offs) + ntiles);
    out->num_rows = total;
    for (size_t c = 0; c < in.cols.size(); c++) {
        Column oc;
        oc.name = in.cols[c].name;
        oc.dtype = in.cols[c].dtype;
        oc.length = total;
        if (!k.outs[c].nullable) {  // compacted by the kernel
            const int w = dtype_width(oc.dtype);
            oc.data = dev_alloc((size_t)std::max<int64_t>(total, 1) * w);
            oc.data_bytes = total * w;
        }
        out->cols.push_back(std::move(oc));
    }
    if (k.with_take) idx = dev_alloc((size_t)std::max<int64_t>(total, 1) * 8);
    if (total > 0) {
        LaunchArgs a;
        fill(a, offs);
        jit::launch(k, a, grid, 256, 0, 1, false);
    }
    for (size_t c = 0; c < in.cols.size(); c++)
        if (k.outs[c].nullable) out->cols[c] = gather_column(in.cols[c], as<int64_t>(idx), total, false);
    op_done();
    return out.release();
}

// ------------------------------------------------------------------------------------------------
// fused HashAggregate
// ------------------------------------------------------------------------------------------------
// records are `astride` words apart; emitted slot ids are ELEMENT indices (record * astride), see agg_find_slot
__global__ void compact_slots(const unsigned long long *__restrict__ nrows, int64_t cap, int astride, int ungrouped, unsigned long long *counter,
                              int64_t *__restrict__ out_slot) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t cap_round = (cap + 31) & ~(int64_t)31;
    const int lane = threadIdx.x & 31;
    for (; i < cap_round; i += stride) {
        // a record is a group iff some row was accumulated into it; one counter update per warp, not per group
        const bool used = i < cap && (ungrouped ? (i == 0) : (nrows[i * astride] != 0));
        const unsigned m = __ballot_sync(0xffffffffu, used);
        if (m) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (used) out_slot[base + __popc(m & ((1u << lane) - 1u))] = i * astride;
        }
    }
}
// every record starts as [0, identity of cell 0, identity of cell 1, ...] (0 for sums / counts, all ones for minima)
struct TableInit {
    unsigned long long word[40];
};
__global__ void init_table(unsigned long long *__restrict__ table, int64_t nrecords, int astride, const __grid_constant__ TableInit init) {
    // one record per thread, 16-byte stores (astride is a multiple of 4 words, records are 32-byte aligned)
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; r < nrecords; r += stride) {
        ulonglong2 *rec = reinterpret_cast<ulonglong2 *>(table + r * astride);
        for (int w = 0; w < astride; w += 2) rec[w >> 1] = make_ulonglong2(init.word[w], init.word[w + 1]);
    }
}

// Tier-1 tables (65536 slots) almost always hold a handful of groups: ONE single-CTA launch finds the occupied slots,
// orders them by first row (first-occurrence group order, inline_agg.rs:804-807) and emits slot ids + first rows.
// counter[0] = number of groups, or -1 when there are more than SMALL_GROUPS (the host then takes the general path).
constexpr int SMALL_GROUPS = 1024;
// Collective mode (tags != nullptr: the table was all-reduced across ranks, exchange.cu allreduce_agg_table): a group
// whose tag differs from the all-reduced non-zero MIN (two keys met in one slot), whose code is inexact (bit 63), or
// that has no local first row (minrow still empty) cannot be finished from the merged table: status 3.
__global__ void __launch_bounds__(1024) small_groups_epilogue(const unsigned long long *__restrict__ nrows, const unsigned long long *__restrict__ minrow_cell,
                                                              int64_t cap, int astride, int ungrouped, long long *__restrict__ counter,
                                                              int64_t *__restrict__ out_slot, unsigned long long *__restrict__ out_minrow,
                                                              const unsigned long long *__restrict__ tags = nullptr,
                                                              const unsigned long long *__restrict__ tagmin = nullptr, int *status = nullptr) {
    __shared__ int64_t s_slot[SMALL_GROUPS];
    __shared__ unsigned long long s_min[SMALL_GROUPS];
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // eight independent loads in flight per thread: one CTA scans the 65 536-slot table in a few microseconds
    for (int64_t i0 = threadIdx.x; i0 < cap; i0 += (int64_t)blockDim.x * 8) {
        unsigned long long nr[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            nr[u] = (i < cap && !ungrouped) ? nrows[i * astride] : 0ULL;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int64_t i = i0 + (int64_t)u * blockDim.x;
            const bool used = i < cap && (ungrouped ? (i == 0) : (nr[u] != 0));
            if (used) {
                const int pos = atomicAdd(&s_n, 1);
                const unsigned long long mr = minrow_cell[i * astride];
                if (pos < SMALL_GROUPS) {
                    s_slot[pos] = i * astride;
                    s_min[pos] = mr;
                }
                if (tags && !ungrouped) {
                    const unsigned long long t = tags[i * astride];
                    if (t != tagmin[i] || (t >> 63) || mr == ~0ULL) atomicMax(status, 3);
                }
            }
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n > SMALL_GROUPS) {
        if (threadIdx.x == 0) counter[0] = -1;
        return;
    }
    // rank by (first row, slot): first rows of distinct groups are distinct, the slot only breaks the ungrouped / empty tie
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const unsigned long long m = s_min[i];
        const int64_t sl = s_slot[i];
        int rank = 0;
        for (int j = 0; j < n; j++) rank += (s_min[j] < m) || (s_min[j] == m && s_slot[j] < sl);
        out_slot[rank] = sl;
        out_minrow[rank] = m;
    }
    if (threadIdx.x == 0) counter[0] = n;
}

__global__ void gather_u64(const unsigned long long *__restrict__ src, const int64_t *__restrict__ idx, int64_t n,
                           unsigned long long *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
// take indices from a "first row" cell: all ones (no such row) -> -1 (a null row)
__global__ void cell_to_take(const unsigned long long *__restrict__ cell, const int64_t *__restrict__ slots, int64_t n, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned long long r = cell[slots[i]];
        out[i] = r == ~0ULL ? -1 : (int64_t)r;
    }
}


// ---- list / set / concat aggregates ----------------------------------------------------------------------------------
// grouped_list / grouped_set / grouped_concat (src/daft-core/src/array/ops/list_agg.rs:30-57, series/ops/agg.rs:531-575,
// array/ops/concat_agg.rs:138-161) walk each group's row indices in order.  Here: every row's group comes from a second pass
// over the finished table (jit_prelude.cuh agg_body_rowslot), a STABLE radix sort by group position puts each group's rows
// together in row order, and the child values are gathered in that order.
__global__ void scatter_rank(const int64_t *__restrict__ slots, int64_t ng, unsigned int *__restrict__ rank) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ng) rank[slots[i]] = (unsigned int)i;
}
// sort key of a row = position of its group in the output, or `ng` (behind every group) for a row that does not take part:
// dropped by the predicate, or a null value where nulls do not count (set, concat)
__global__ void list_sort_keys(const long long *__restrict__ row_slot, const unsigned int *__restrict__ rank, const uint8_t *__restrict__ valid,
                               int64_t n, unsigned long long ng, unsigned long long *__restrict__ keys) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const long long s = row_slot[i];
        keys[i] = (s < 0 || (valid && !bit_at(valid, i))) ? ng : (unsigned long long)rank[s];
    }
}
// bounds[g] = first position of the sorted keys holding a key >= g, for g in [0, ng]: list offsets, bounds[ng] = rows taking part
__global__ void group_bounds(const unsigned long long *__restrict__ keys, int64_t n, long long ng, int64_t *__restrict__ bounds) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i <= n; i += stride) {
        const long long lo = i == 0 ? 0 : (long long)min(keys[i - 1], (unsigned long long)ng) + 1;
        const long long hi = i == n ? ng : (long long)min(keys[i], (unsigned long long)ng);
        for (long long g = lo; g <= hi; g++) bounds[g] = i;
    }
}
__global__ void gather_i64_at(const int64_t *__restrict__ src, const int64_t *__restrict__ idx, int64_t n, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}
__global__ void group_has_rows(const int64_t *__restrict__ bounds, int64_t ng, uint32_t *__restrict__ validity) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w * 32 >= ng) return;
    uint32_t m = 0;
    for (int b = 0; b < 32 && w * 32 + b < ng; b++)
        if (bounds[w * 32 + b + 1] > bounds[w * 32 + b]) m |= 1u << b;
    validity[w] = m;
}

static void put_u32(std::vector<uint8_t> &v, uint32_t x) { v.insert(v.end(), (uint8_t *)&x, (uint8_t *)&x + 4); }
static void put_col_entry(std::vector<uint8_t> &v, const char *name, uint32_t col) {
    put_u32(v, (uint32_t)strlen(name));
    v.insert(v.end(), name, name + strlen(name));
    v.push_back(0);  // T_COL
    put_u32(v, col);
}

struct RowGroups {
    Buf row_slot, rank;  // int64 per input row (-1: dropped), group position per table element
};

// One list-type aggregate over `vals` (the child evaluated at every input row): the rows taking part, ordered by group.
// Returns the rows (int64[total]), their group positions (the sorted keys) and the ng + 1 offsets.
static void rows_by_group(const RowGroups &rg, const uint8_t *valid, int64_t n, int64_t ng, Buf &rows, Buf &keys, Buf &bounds, int64_t &total) {
    cudaStream_t st = ctx().stream;
    bounds = dev_alloc((size_t)(ng + 2) * 8);
    keys = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    rows = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
    if (n > 0) {
        list_sort_keys<<<stream_grid(n, 256, 8), 256, 0, st>>>(as<long long>(rg.row_slot), as<unsigned int>(rg.rank), valid, n, (unsigned long long)ng,
                                                               as<unsigned long long>(keys));
        count_launch();
        fill_iota(as<int64_t>(rows), n);
        int bits = 1;
        while (((int64_t)1 << bits) <= ng) bits++;
        radix_sort_pairs(keys, rows, n, 0, bits, false);  // LSD passes: stable, so a group's rows stay in row order
    }
    group_bounds<<<stream_grid(n + 1, 256, 8), 256, 0, st>>>(as<unsigned long long>(keys), n, (long long)ng, as<int64_t>(bounds));
    count_launch();
    total = read_i64(as<int64_t>(bounds) + ng);
}

struct FinalizeDesc {
    int kind, val_dtype, out_dtype, cnt_is_valid_count, const_zero, _pad;
    const unsigned long long *cell;
    const unsigned long long *cnt;
    void *out;
    uint32_t *out_valid;
    const unsigned long long *nan_cell, *last_cell;  // float min / max: 1 + row of the group's last NaN / last valid value
};
constexpr int MAX_FINALIZE = 24;
struct FinalizeSet {
    int n;
    FinalizeDesc d[MAX_FINALIZE];
};
// cell (8-byte) -> typed output value
__device__ __forceinline__ void finalize_one(int kind, int32_t val_dtype, int32_t out_dtype, const unsigned long long *__restrict__ cell,
                                             const unsigned long long *__restrict__ cnt, int cnt_is_valid_count, int const_zero,
                                             const int64_t *__restrict__ slots, int64_t ng, void *__restrict__ out,
                                             uint32_t *__restrict__ out_valid, int64_t i, const unsigned long long *__restrict__ nan_cell,
                                             const unsigned long long *__restrict__ last_cell) {
    bool valid = false, bool_bit = false;
    if (i < ng) {
        const int64_t s = slots[i];
        const unsigned long long c = cell ? cell[s] : 0ULL;
        const unsigned long long n = cnt ? cnt[s] : 0ULL;
        valid = true;
        if (kind == jit::AGG_COUNT) {
            ((unsigned long long *)out)[i] = const_zero ? 0ULL : n;
        } else {
            if (cnt_is_valid_count) valid = n > 0;
            if (kind == jit::AGG_MEAN) {
                double sum;
                if (val_dtype == B200_F64) sum = __longlong_as_double((long long)c);
                else if (val_dtype == B200_F32) sum = (double)__int_as_float((int)c);
                else if (val_dtype == B200_U64) sum = (double)c;
                else sum = (double)(long long)c;
                ((double *)out)[i] = valid ? sum / (double)n : 0.0;
            } else if (kind == jit::AGG_SUM) {
                if (val_dtype == B200_F32) ((float *)out)[i] = valid ? __int_as_float((int)c) : 0.f;
                else ((unsigned long long *)out)[i] = valid ? c : 0ULL;  // f64 bits / i64 / u64
            } else {  // min / max: decode the order-preserving key
                unsigned long long v = c;
                // the left fold ends on NaN exactly when the group's last valid value is one (inline_agg.rs:271-298)
                const bool ends_nan = nan_cell != nullptr && nan_cell[s] != 0ULL && nan_cell[s] == last_cell[s];
                switch (out_dtype) {
                    case B200_F64: {
                        double d = (ends_nan || c == 0xFFFFFFFFFFFFFFFFULL) ? __longlong_as_double(0x7FF8000000000000LL)
                                                                 : __longlong_as_double((long long)((c >> 63) ? (c & 0x7FFFFFFFFFFFFFFFULL) : ~c));
                        ((double *)out)[i] = valid ? d : 0.0;
                        break;
                    }
                    case B200_F32: {
                        double d = (ends_nan || c == 0xFFFFFFFFFFFFFFFFULL) ? __longlong_as_double(0x7FF8000000000000LL)
                                                                 : __longlong_as_double((long long)((c >> 63) ? (c & 0x7FFFFFFFFFFFFFFFULL) : ~c));
                        ((float *)out)[i] = valid ? (float)d : 0.f;
                        break;
                    }
                    case B200_I8: ((int8_t *)out)[i] = valid ? (int8_t)(long long)(v ^ 0x8000000000000000ULL) : 0; break;
                    case B200_I16: ((int16_t *)out)[i] = valid ? (int16_t)(long long)(v ^ 0x8000000000000000ULL) : 0; break;
                    case B200_I32: case B200_DATE32: ((int32_t *)out)[i] = valid ? (int32_t)(long long)(v ^ 0x8000000000000000ULL) : 0; break;
                    case B200_I64: case B200_TIMESTAMP: ((long long *)out)[i] = valid ? (long long)(v ^ 0x8000000000000000ULL) : 0; break;
                    case B200_U8: ((uint8_t *)out)[i] = valid ? (uint8_t)v : 0; break;
                    case B200_U16: ((uint16_t *)out)[i] = valid ? (uint16_t)v : 0; break;
                    case B200_U32: ((uint32_t *)out)[i] = valid ? (uint32_t)v : 0; break;
                    case B200_BOOL: bool_bit = valid && v != 0; break;  // bool_and / bool_or: a bitmap, written below
                    default: ((unsigned long long *)out)[i] = valid ? v : 0ULL; break;
                }
            }
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (out_valid && (threadIdx.x & 31) == 0) out_valid[i >> 5] = m;
    if (out_dtype == B200_BOOL && kind != jit::AGG_COUNT) {
        const unsigned b = __ballot_sync(0xffffffffu, bool_bit);
        if ((threadIdx.x & 31) == 0) ((uint32_t *)out)[i >> 5] = b;
    }
}

// every aggregate column of the output in ONE launch (grid.y = aggregate)
__global__ void finalize_agg_all(FinalizeSet fs, const int64_t *__restrict__ slots, int64_t ng) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t g_round = (ng + 31) & ~(int64_t)31;
    if (i >= g_round) return;
    const FinalizeDesc &d = fs.d[blockIdx.y];
    finalize_one(d.kind, d.val_dtype, d.out_dtype, d.cell, d.cnt, d.cnt_is_valid_count, d.const_zero, slots, ng, d.out, d.out_valid, i, d.nan_cell,
                 d.last_cell);
}

Buf sort_slots_by_minrow(Buf slots, Buf minrow, int64_t ng);  // sort.cu

// float min / max, NaN-forgetting fold: does some group have a NaN that is not its last valid value?
struct NanPairs {
    int n;
    const unsigned long long *nan_cell[16], *last_cell[16];
};
__global__ void nan_mid_check(const int64_t *__restrict__ slots, int64_t ng, const __grid_constant__ NanPairs np, int *flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ng) return;
    const int64_t s = slots[i];
    for (int k = 0; k < np.n; k++) {
        const unsigned long long a = np.nan_cell[k][s];
        if (a != 0ULL && a != np.last_cell[k][s]) *flag = 1;
    }
}
__global__ void reset_cell(unsigned long long *__restrict__ cell, int64_t nrecords, int astride, unsigned long long value) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < nrecords; i += stride) cell[i * astride] = value;
}


// `coll` != nullptr: the aggregate of a low-cardinality plan over one scan partition per rank -- the group table is
// all-reduced across the ranks before it is read back, so every rank finishes with the global result and no second
// aggregation stage runs.  Returns nullptr (on every rank alike: the decision is taken from all-reduced words) when the
// plan or the data do not allow it; the caller then runs the two-stage plan.
Batch *op_agg_first_rows(const Batch &in, const uint8_t *plan, size_t len, Buf *first_rows, int64_t *ngroups, b200_comm *coll) {
    require_device();
    cudaStream_t st = ctx().stream;
    const int64_t n = in.num_rows;
    // Strategy (the reference samples the first morsel for the same decision, grouped_aggregate.rs:205-245):
    // start with the low-cardinality kernel; it raises flag 2 as soon as most rows miss its shared-memory tier, and
    // the global-table kernel takes over from scratch, growing its table while it reports overflow (flag 1).
    int variant = 0;
    const Compiled *kp = nullptr;
    {
        // the TMA-staged build of the low-cardinality kernel needs 16-byte aligned column buffers and enough rows to
        // fill its pipeline; anything else runs the direct-load build
        const char *e = getenv("B200_AGG_STAGED");
        if (n >= (1 << 16) && !(e && *e == '0')) {
            const Compiled &k2 = jit::get(B200_OP_AGG, plan, len, in, false, 2);
            bool ok = k2.stage_bytes > 0;
            for (const auto &ps : k2.ptrs) {
                const void *q = nullptr;
                if (ps.kind == jit::P_COL_DATA) q = in.cols[ps.index].data ? in.cols[ps.index].data->ptr : nullptr;
                else if (ps.kind == jit::P_COL_OFFSETS) q = in.cols[ps.index].offsets ? in.cols[ps.index].offsets->ptr : nullptr;
                if (((uintptr_t)q & 15) != 0) ok = false;
            }
            if (ok) {
                variant = 2;
                kp = &k2;
            }
        }
        if (!kp) kp = &jit::get(B200_OP_AGG, plan, len, in, false, 0);
    }
    // Table behind the tier-1 kernels: on large inputs they only finish when a CTA meets a handful of groups (anything else
    // is abandoned to the global build), so a small table is enough and keeps its initialisation and the epilogue's scan
    // cheap; small inputs never reach the abandon threshold and may hold tens of thousands of groups.
    int64_t cap = kp->nkeys > 0 ? (n >= (1 << 20) ? (1 << 13) : (1 << 16)) : 1;
    if (kp->nkeys > 0 && n < (1 << 15)) {  // n rows cannot form more than n groups: merges of partials get a tiny table
        cap = 64;
        while (cap < 2 * n) cap <<= 1;
    }
    if (coll) {
        cap = kp->nkeys > 0 ? (1 << 11) : 1;  // the same (small: it is all-reduced) table shape on every rank, whatever its row count
        for (const auto &ag : kp->aggs)
            if (ag.kind == jit::AGG_ANY_VALUE || jit::is_list_agg(ag.kind)) return nullptr;  // first-row values / row lists are per rank: two-stage plan
        for (int op : kp->lane_cell_ops)
            if (op == jit::C_ADD_F32) return nullptr;
        if (kp->has_nan_fold) return nullptr;  // float min / max fold in row order: rows are per rank
    }
    const Compiled *k_first = kp;  // carries the plan's history (the global build is only compiled when needed)
    if (!coll && kp->nkeys > 0 && k_first->last_groups > 0 && n >= (1 << 16)) {  // sticky choice: skip the tier-1 attempt it would abandon again
        variant = 1;
        kp = &jit::get(B200_OP_AGG, plan, len, in, false, 1);
        cap = 1 << 16;
        while (2 * cap < 3 * k_first->last_groups) cap <<= 1;  // load factor <= 2/3: a larger table only costs DRAM locality
    }
    Buf slot_tag, acc, slots, minrow_sorted, flag = dev_alloc_zero(16);  // flag: {int status, pad, u64 group counter}
    int64_t ng = 0;
    bool small_path = false;  // slots / first rows already ordered by the single-CTA epilogue
    size_t cell_step_final = 1;  // layout of the table the successful attempt used
    int astride_final = 1;
    int ncells = 0, cell_nrows = 0, cell_minrow = 0;
    for (int attempt = 0;; attempt++) {
        const Compiled &k = *kp;
        ncells = k.NLANE + k.NWARP;  // value cells + row count + first row, then ballot count cells
        cell_nrows = k.NLANE - 2;
        cell_minrow = k.NLANE - 1;
        // layout: records when the table outgrows L2 (DRAM locality), one array per cell otherwise (parallel L2 atomics)
        if (k.astride > 40) fail("agg: too many accumulator cells per group");
        const bool records = (size_t)cap * (size_t)k.astride * 8 > ((size_t)48 << 20);
        const int astride = records ? k.astride : 1;
        const size_t cell_step = records ? 1 : (size_t)cap;  // words between the bases of consecutive cell arrays
        acc = dev_alloc((size_t)cap * k.astride * 8);
        if (records) {
            TableInit ti;
            for (int w = 0; w < 40; w++) ti.word[w] = 0ULL;
            for (int c = 0; c < k.NLANE; c++)
                if (k.lane_cell_ops[c] == jit::C_MIN) ti.word[1 + c] = ~0ULL;
            init_table<<<stream_grid(cap, 256, 8), 256, 0, st>>>(as<unsigned long long>(acc), cap, k.astride, ti);
            count_launch();
        } else {
            fill_u8(acc->ptr, 0, (size_t)cap * (1 + ncells) * 8);
            for (int c = 0; c < k.NLANE; c++)
                if (k.lane_cell_ops[c] == jit::C_MIN) fill_u8(as<uint8_t>(acc) + (size_t)(1 + c) * cap * 8, 0xFF, (size_t)cap * 8);
        }
        auto cell_base = [&](int c) { return as<unsigned long long>(acc) + (size_t)(1 + c) * cell_step; };
        fill_u8(flag->ptr, 0, 16);
        int status = 0;
        if (n > 0) {
            LaunchArgs a;
            a.ptr.assign(k.ptrs.size(), nullptr);
            a.ival.assign(k.ints.size(), 0);
            for (size_t s = 0; s < k.ptrs.size(); s++) {
                const auto &ps = k.ptrs[s];
                switch (ps.kind) {
                    case jit::P_COL_DATA: a.ptr[s] = in.cols[ps.index].data ? in.cols[ps.index].data->ptr : nullptr; break;
                    case jit::P_COL_VALID: a.ptr[s] = in.cols[ps.index].validity ? in.cols[ps.index].validity->ptr : nullptr; break;
                    case jit::P_COL_OFFSETS: a.ptr[s] = in.cols[ps.index].offsets ? in.cols[ps.index].offsets->ptr : nullptr; break;
                    case jit::P_SLOT_HASH: a.ptr[s] = acc->ptr; break;
                    case jit::P_OVERFLOW: a.ptr[s] = flag->ptr; break;
                    case jit::P_ACC: a.ptr[s] = cell_base(ps.index); break;
                    default: fail("agg: unexpected parameter kind");
                }
            }
            for (size_t s = 0; s < k.ints.size(); s++) {
                const int kind = k.ints[s] % 1000;
                if (kind == jit::I_N) a.ival[s] = n;
                else if (kind == jit::I_MASK) a.ival[s] = cap - 1;
                else if (kind == jit::I_RSTRIDE) a.ival[s] = astride;
            }
            int occ = 0;
            B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void *)k.fn, k.block_threads, k.smem_bytes));
            occ = std::max(1, occ);
            // work units: a 64-row chunk per warp (direct-load builds) or a whole stage per CTA (staged build)
            const int64_t ctas = variant == 2 ? (n + k.stage_rows - 1) / k.stage_rows : ((n + 63) / 64 + 7) / 8;
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)ctx().sm_count * occ));
            jit::launch(k, a, grid, k.block_threads, k.smem_bytes, 0, true);
        }
        Buf tagmin;
        if (coll) {  // merge the tables of all ranks (also the status word: every rank takes the same decision below)
            if (records) fail("agg all-reduce: unexpected record layout");
            tagmin = dev_alloc((size_t)cap * 8);
            std::vector<int> ops(k.lane_cell_ops.begin(), k.lane_cell_ops.end());
            for (int w = 0; w < k.NWARP; w++) ops.push_back(jit::C_ADD_I64);
            std::vector<char> keep(ops.size(), 0);
            keep[(size_t)k.NLANE - 1] = 1;  // first rows index THIS rank's input: they stay local
            allreduce_agg_table(coll, as<unsigned long long>(acc), (size_t)cap, ops, keep, as<unsigned long long>(tagmin), as<int>(flag));
        }
        // compact the occupied slots right behind the kernel: status flag and group count come back in ONE read
        // (an ungrouped aggregate has exactly one group and no failure mode: no read at all)
        slots = dev_alloc((size_t)std::max<int64_t>(cap, 1) * 8);
        const unsigned long long *nrows_c = cell_base(cell_nrows);
        const unsigned long long *minrow_c = cell_base(cell_minrow);
        cell_step_final = cell_step;
        astride_final = astride;
        small_path = cap <= (1 << 16);
        if (small_path) {
            minrow_sorted = dev_alloc((size_t)SMALL_GROUPS * 8);
            small_groups_epilogue<<<1, 1024, 0, st>>>(nrows_c, minrow_c, cap, astride, k.nkeys == 0, as<long long>(flag) + 1, as<int64_t>(slots),
                                                      as<unsigned long long>(minrow_sorted), coll ? as<unsigned long long>(acc) : nullptr,
                                                      as<unsigned long long>(tagmin), as<int>(flag));
        } else {
            compact_slots<<<stream_grid(cap, 256, 8), 256, 0, st>>>(nrows_c, cap, astride, k.nkeys == 0, as<unsigned long long>(flag) + 1, as<int64_t>(slots));
        }
        count_launch();
        if (k.nkeys > 0) {
            long long host[2] = {0, 0};
            B200_CUDA(cudaMemcpyAsync(host, flag->ptr, 16, cudaMemcpyDeviceToHost, st));
            B200_CUDA(cudaStreamSynchronize(st));
            status = (int)(host[0] & 0xffffffffLL);
            ng = host[1];
            if (coll && (status || ng < 0)) return nullptr;  // abandoned / overflow / not mergeable: identical on every rank
            if (!status && ng < 0) {  // more groups than the single-CTA epilogue handles: general path on the same table
                small_path = false;
                fill_u8(as<uint8_t>(flag) + 8, 0, 8);
                compact_slots<<<stream_grid(cap, 256, 8), 256, 0, st>>>(nrows_c, cap, astride, 0, as<unsigned long long>(flag) + 1, as<int64_t>(slots));
                count_launch();
                ng = read_i64(as<int64_t>(flag) + 1);
            }
        } else {
            ng = 1;
        }
        if (!status) break;
        if (attempt > 10) fail("agg: hash table overflow after repeated growth");
        if (variant != 1) {
            variant = 1;
            kp = &jit::get(B200_OP_AGG, plan, len, in, false, 1);
            cap = 1 << 20;
            while (cap < 2 * std::min<int64_t>(n, (int64_t)1 << 23)) cap <<= 1;
        } else {
            cap <<= 3;
        }
        const size_t free_b = dev_available_bytes();
        if ((size_t)cap * 8 * (size_t)(kp->astride + 1) > free_b + (size_t)(kp->astride + 1) * (size_t)(cap >> 3) * 8)
            fail("agg: the group table would not fit in device memory");
    }
    const Compiled &k = *kp;
    k_first->last_groups = (variant == 1) ? ng : 0;

    // ---- compact occupied slots, order groups by first occurrence
    auto cell_ptr = [&](int c) { return as<unsigned long long>(acc) + (size_t)(1 + c) * cell_step_final; };
    const unsigned long long *nrows_cell = cell_ptr(cell_nrows);
    Buf minrow = small_path ? minrow_sorted : dev_alloc((size_t)std::max<int64_t>(ng, 1) * 8);
    if (ng > 0 && !small_path) {
        gather_u64<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(cell_ptr(cell_minrow),
                                                                  as<int64_t>(slots), ng, as<unsigned long long>(minrow));
        count_launch();
        if (k.nkeys > 0 && ng > 1) {
            slots = sort_slots_by_minrow(slots, minrow, ng);  // also permutes `minrow` in place
        }
    }

    // ---- float min / max: the reference's fold forgets what came before a NaN (jit.h AggOut::nan_cell).  The kernel
    // folded the non-NaN candidates of all rows; a group whose last NaN is followed by more values is re-folded over
    // the rows after that NaN by a second pass over the input (rare: only when such a group exists).
    if (k.has_nan_fold && ng > 0 && n > 0) {
        NanPairs np;
        np.n = 0;
        for (const auto &ag : k.aggs)
            if (ag.nan_cell >= 0 && np.n < 16) {
                np.nan_cell[np.n] = cell_ptr(ag.nan_cell);
                np.last_cell[np.n] = cell_ptr(ag.last_cell);
                np.n++;
            }
        Buf nflag = dev_alloc_zero(8);
        nan_mid_check<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(as<int64_t>(slots), ng, np, as<int>(nflag));
        count_launch();
        if (read_i64(as<int64_t>(nflag)) != 0) {
            for (const auto &ag : k.aggs)
                if (ag.nan_cell >= 0) {
                    reset_cell<<<stream_grid(cap, 256, 8), 256, 0, st>>>(cell_ptr(ag.lane_cell), cap, astride_final,
                                                                         k.lane_cell_ops[ag.lane_cell] == jit::C_MIN ? ~0ULL : 0ULL);
                    count_launch();
                }
            const Compiled &k3 = jit::get(B200_OP_AGG, plan, len, in, false, 3);
            LaunchArgs a;
            a.ptr.assign(k3.ptrs.size(), nullptr);
            a.ival.assign(k3.ints.size(), 0);
            for (size_t s2 = 0; s2 < k3.ptrs.size(); s2++) {
                const auto &ps = k3.ptrs[s2];
                switch (ps.kind) {
                    case jit::P_COL_DATA: a.ptr[s2] = in.cols[ps.index].data ? in.cols[ps.index].data->ptr : nullptr; break;
                    case jit::P_COL_VALID: a.ptr[s2] = in.cols[ps.index].validity ? in.cols[ps.index].validity->ptr : nullptr; break;
                    case jit::P_COL_OFFSETS: a.ptr[s2] = in.cols[ps.index].offsets ? in.cols[ps.index].offsets->ptr : nullptr; break;
                    case jit::P_SLOT_HASH: a.ptr[s2] = acc->ptr; break;
                    case jit::P_OVERFLOW: a.ptr[s2] = flag->ptr; break;
                    case jit::P_ACC: a.ptr[s2] = cell_ptr(ps.index); break;
                    default: fail("agg: unexpected parameter kind");
                }
            }
            for (size_t s2 = 0; s2 < k3.ints.size(); s2++) {
                const int kind = k3.ints[s2] % 1000;
                if (kind == jit::I_N) a.ival[s2] = n;
                else if (kind == jit::I_MASK) a.ival[s2] = cap - 1;
                else if (kind == jit::I_RSTRIDE) a.ival[s2] = astride_final;
            }
            jit::launch(k3, a, stream_grid(n, 256, 8), 256, 0, 0, false);
        }
    }

    // ---- output: key columns = key expressions evaluated at each group's first row, then agg columns
    std::unique_ptr<Batch> out;
    if (k.nkeys > 0) {
        out.reset(op_project(in, k.key_project_plan.data(), k.key_project_plan.size(), as<int64_t>(minrow), ng, true));
        // first rows are real rows: drop the "take may produce nulls" pessimism
    } else {
        out = std::make_unique<Batch>();
        out->num_rows = ng;
    }
    // any_value columns: the child expressions evaluated at the groups' first rows (one projection for all of them)
    std::unique_ptr<Batch> anyval;
    if (!k.anyval_project_plan.empty() && ng > 0 && n > 0)
        anyval.reset(op_project(in, k.anyval_project_plan.data(), k.anyval_project_plan.size(), as<int64_t>(minrow), ng, true));
    // ---- list / set / concat: rows ordered by group, child gathered in that order (see rows_by_group above)
    std::vector<Column> list_cols;
    bool any_list_agg = false;
    for (const auto &ag : k.aggs) any_list_agg = any_list_agg || jit::is_list_agg(ag.kind);
    if (any_list_agg) {
        std::unique_ptr<Batch> vals;  // children that are expressions; bare columns are gathered from the input itself
        if (!k.listagg_project_plan.empty()) vals.reset(op_project(in, k.listagg_project_plan.data(), k.listagg_project_plan.size(), nullptr, 0));
        RowGroups rg;
        rg.row_slot = dev_alloc((size_t)std::max<int64_t>(n, 1) * 8);
        rg.rank = dev_alloc((size_t)std::max<int64_t>(cap, 1) * (size_t)astride_final * 4);
        if (ng > 0) {
            scatter_rank<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(as<int64_t>(slots), ng, as<unsigned int>(rg.rank));
            count_launch();
        }
        if (n > 0) {
            const Compiled &k4 = jit::get(B200_OP_AGG, plan, len, in, false, 4);
            LaunchArgs a;
            a.ptr.assign(k4.ptrs.size(), nullptr);
            a.ival.assign(k4.ints.size(), 0);
            for (size_t s2 = 0; s2 < k4.ptrs.size(); s2++) {
                const auto &ps = k4.ptrs[s2];
                switch (ps.kind) {
                    case jit::P_COL_DATA: a.ptr[s2] = in.cols[ps.index].data ? in.cols[ps.index].data->ptr : nullptr; break;
                    case jit::P_COL_VALID: a.ptr[s2] = in.cols[ps.index].validity ? in.cols[ps.index].validity->ptr : nullptr; break;
                    case jit::P_COL_OFFSETS: a.ptr[s2] = in.cols[ps.index].offsets ? in.cols[ps.index].offsets->ptr : nullptr; break;
                    case jit::P_SLOT_HASH: a.ptr[s2] = acc->ptr; break;
                    case jit::P_OVERFLOW: a.ptr[s2] = flag->ptr; break;
                    case jit::P_ACC: a.ptr[s2] = cell_ptr(ps.index); break;
                    case jit::P_SLOT_ROW: a.ptr[s2] = rg.row_slot->ptr; break;
                    default: fail("agg: unexpected parameter kind");
                }
            }
            for (size_t s2 = 0; s2 < k4.ints.size(); s2++) {
                const int kind = k4.ints[s2] % 1000;
                if (kind == jit::I_N) a.ival[s2] = n;
                else if (kind == jit::I_MASK) a.ival[s2] = cap - 1;
                else if (kind == jit::I_RSTRIDE) a.ival[s2] = astride_final;
            }
            jit::launch(k4, a, stream_grid(n, 256, 8), 256, 0, 0, false);
        }
        for (const auto &ag : k.aggs) {
            if (!jit::is_list_agg(ag.kind)) continue;
            const Column &x = ag.child_col >= 0 ? in.cols[ag.child_col] : vals->cols[ag.proj_index];
            const bool skip_nulls = ag.kind != jit::AGG_LIST && x.null_count > 0;
            Buf rows, gkeys, bounds;
            int64_t total = 0;
            if (x.dtype == B200_NULL && ag.kind != jit::AGG_LIST) {  // every value is null: nothing takes part
                bounds = dev_alloc_zero((size_t)(ng + 2) * 8);
                rows = dev_alloc(8);
            } else {
                rows_by_group(rg, skip_nulls ? as<uint8_t>(x.validity) : nullptr, n, ng, rows, gkeys, bounds, total);
            }
            Column ordered = gather_column(x, as<int64_t>(rows), total, false);
            if (ag.kind == jit::AGG_SET && total > 0) {
                // distinct values per group in first-occurrence order (deduplicate_indices, series/ops/agg.rs:23-27): the rows are
                // grouped already, so the first rows of the distinct (group, value) pairs -- ascending positions -- are the answer
                Batch tmp;
                tmp.num_rows = total;
                tmp.cols.push_back(make_u64_column("g", gkeys, total));
                tmp.cols[0].dtype = B200_I64;
                tmp.cols.push_back(ordered);
                std::vector<uint8_t> dp;
                dp.push_back(0);           // no predicate
                put_u32(dp, 2);            // keys: (group, value)
                put_col_entry(dp, "g", 0);
                put_col_entry(dp, "v", 1);
                put_u32(dp, 1);            // one aggregate: count(all)
                dp.push_back((uint8_t)jit::AGG_COUNT);
                dp.push_back(0);
                put_col_entry(dp, "n", 0);
                Buf firsts;
                int64_t nd = 0;
                std::unique_ptr<Batch> dd(op_agg_first_rows(tmp, dp.data(), dp.size(), &firsts, &nd, nullptr));
                Buf g2 = dev_alloc((size_t)std::max<int64_t>(nd, 1) * 8);
                if (nd > 0) {
                    gather_i64_at<<<(unsigned)((nd + 255) / 256), 256, 0, st>>>(as<int64_t>(gkeys), as<int64_t>(firsts), nd, as<int64_t>(g2));
                    count_launch();
                }
                group_bounds<<<stream_grid(nd + 1, 256, 8), 256, 0, st>>>(as<unsigned long long>(g2), nd, (long long)ng, as<int64_t>(bounds));
                count_launch();
                ordered = gather_column(ordered, as<int64_t>(firsts), nd, false);
                total = nd;
            }
            Column c;
            c.name = ag.name;
            c.length = ng;
            if (ag.kind == jit::AGG_CONCAT) {
                // the group's valid strings back to back: offsets of the ordered strings at the group boundaries
                c.dtype = ag.out_dtype;
                c.offsets = dev_alloc((size_t)(ng + 1) * 8);
                if (total > 0) {
                    gather_i64_at<<<(unsigned)((ng + 1 + 255) / 256), 256, 0, st>>>(as<int64_t>(ordered.offsets), as<int64_t>(bounds), ng + 1, as<int64_t>(c.offsets));
                    count_launch();
                    c.data = ordered.data;
                    c.data_bytes = ordered.data_bytes;
                    c.child = ordered.child;  // concat of lists (concat_agg.rs:68-116): the ordered lists' elements, regrouped
                } else {
                    fill_u8(c.offsets->ptr, 0, (size_t)(ng + 1) * 8);
                    if (c.dtype == B200_LIST) c.child = std::make_shared<const Column>(gather_column(*x.child, nullptr, 0, false));
                    else c.data = dev_alloc(8);
                }
                if (x.null_count > 0 && ng > 0) {  // null where no value of the group is valid (concat_agg.rs:140-151)
                    c.validity = dev_alloc_zero((size_t)((ng + 31) / 32 + 1) * 4);
                    group_has_rows<<<(unsigned)(((ng + 31) / 32 + 255) / 256), 256, 0, st>>>(as<int64_t>(bounds), ng, as<uint32_t>(c.validity));
                    count_launch();
                    c.null_count = ng - count_set_bits(as<uint8_t>(c.validity), ng);
                    if (k.nkeys == 0 && total == 0) {
                        // ungrouped: null only when rows survive the predicate and all of them are null; over no rows at all the
                        // result is the empty string (concat_agg.rs:117-131 on the filtered column)
                        Buf r2, k2, b2;
                        int64_t live = 0;
                        rows_by_group(rg, nullptr, n, ng, r2, k2, b2, live);
                        if (live == 0) c.null_count = 0;
                    }
                    if (c.null_count == 0) c.validity.reset();
                }
            } else {
                c.dtype = B200_LIST;
                c.offsets = bounds;
                ordered.name = "item";
                c.child = std::make_shared<const Column>(std::move(ordered));
            }
            list_cols.push_back(std::move(c));
        }
    }
    size_t list_next = 0;
    FinalizeSet fs;
    fs.n = 0;
    std::vector<Column> agg_cols;
    auto flush_finalize = [&]() {
        if (fs.n && ng > 0) {
            dim3 grid((unsigned)((ng + 255) / 256), (unsigned)fs.n);
            finalize_agg_all<<<grid, 256, 0, st>>>(fs, as<int64_t>(slots), ng);
            count_launch();
        }
        fs.n = 0;
    };
    for (const auto &ag : k.aggs) {
        Column c;
        c.name = ag.name;
        c.dtype = ag.out_dtype;
        c.length = ng;
        if (jit::is_list_agg(ag.kind)) {
            agg_cols.push_back(std::move(list_cols[list_next++]));
            continue;
        }
        if (ag.kind == jit::AGG_ANY_VALUE) {
            if (anyval && ag.first_valid_cell >= 0 && ng > 0) {
                // ignore_nulls: the child at the group's first row where it is valid (Series::any_value, series/ops/agg.rs:353-382)
                Buf idx = dev_alloc((size_t)ng * 8);
                cell_to_take<<<(unsigned)((ng + 255) / 256), 256, 0, st>>>(cell_ptr(ag.first_valid_cell), as<int64_t>(slots), ng, as<int64_t>(idx));
                count_launch();
                std::unique_ptr<Batch> av(op_project(in, k.anyval_project_plan.data(), k.anyval_project_plan.size(), as<int64_t>(idx), ng, false));
                c = av->cols[ag.proj_index];
                c.name = ag.name;
            } else if (anyval) {
                c = anyval->cols[ag.proj_index];
                c.name = ag.name;
            } else {  // no rows at all (ungrouped aggregate over an empty input): null
                c.null_count = ng;
                if (c.dtype != B200_NULL) {
                    c.validity = dev_alloc_zero((size_t)((ng + 31) / 32 + 1) * 4);
                    const int w0 = dtype_width(c.dtype);
                    c.data = dev_alloc_zero((size_t)std::max<int64_t>(ng, 1) * std::max(w0, 1) + 8);
                    c.data_bytes = w0 ? ng * w0 : (c.dtype == B200_BOOL ? (ng + 7) / 8 : 0);
                    if (is_string(c.dtype)) c.offsets = dev_alloc_zero((size_t)(ng + 1) * 8);
                }
            }
            agg_cols.push_back(std::move(c));
            continue;
        }
        const int w = dtype_width(c.dtype);
        if (c.dtype == B200_BOOL) {
            c.data = dev_alloc((size_t)((ng + 31) / 32 + 1) * 4);
            c.data_bytes = (ng + 7) / 8;
        } else {
            c.data = dev_alloc((size_t)std::max<int64_t>(ng, 1) * w);
            c.data_bytes = ng * w;
        }
        const bool may_null = ag.kind != jit::AGG_COUNT && (ag.child_nullable || k.nkeys == 0);
        if (may_null) c.validity = dev_alloc((size_t)((ng + 31) / 32 + 1) * 4);
        const unsigned long long *cell = ag.lane_cell >= 0 ? cell_ptr(ag.lane_cell) : nullptr;
        const unsigned long long *cnt = ag.cnt_cell >= 0 ? cell_ptr(k.NLANE + ag.cnt_cell) : nrows_cell;
        // an ungrouped aggregate over zero rows still yields one row: sums null, counts 0
        const int cnt_is_valid = (ag.kind != jit::AGG_COUNT) ? 1 : 0;
        fs.d[fs.n++] = {ag.kind, ag.val_dtype, ag.out_dtype, cnt_is_valid, ag.const_zero ? 1 : 0, 0, cell, cnt, c.data->ptr, as<uint32_t>(c.validity),
                        ag.nan_cell >= 0 ? cell_ptr(ag.nan_cell) : nullptr, ag.last_cell >= 0 ? cell_ptr(ag.last_cell) : nullptr};
        agg_cols.push_back(std::move(c));
        if (fs.n == MAX_FINALIZE) flush_finalize();
    }
    flush_finalize();
    for (auto &c : agg_cols) {
        if (c.validity) {
            c.null_count = ng - count_set_bits(as<uint8_t>(c.validity), ng);
            if (c.null_count == 0) c.validity.reset();
        }
        out->cols.push_back(std::move(c));
    }
    op_done();
    if (first_rows) *first_rows = minrow;
    if (ngroups) *ngroups = ng;
    return out.release();
}

Batch *op_agg(const Batch &in, const uint8_t *plan, size_t len) { return op_agg_first_rows(in, plan, len, nullptr, nullptr, nullptr); }

}  // namespace b200

// ------------------------------------------------------------------------------------------------
// GroupedAggregateSink state (src/daft-local-execution/src/sinks/grouped_aggregate.rs:146-488): the planner hands the
// sink its partial aggregations, final aggregations and final projection; every pushed batch is reduced to a partial
// result that stays in HBM, finalize merges them.
// ------------------------------------------------------------------------------------------------
struct b200_groupby_state {
    std::vector<uint8_t> partial_plan, final_plan, project_plan;
    std::vector<std::unique_ptr<b200::Batch>> partials;
    std::mutex mu;
};

namespace b200 {

b200_groupby_state *groupby_create(const uint8_t *p1, size_t n1, const uint8_t *p2, size_t n2, const uint8_t *p3, size_t n3) {
    if (!p1 || !p2 || !p3) fail("groupby: null plan");
    auto *s = new b200_groupby_state;
    s->partial_plan.assign(p1, p1 + n1);
    s->final_plan.assign(p2, p2 + n2);
    s->project_plan.assign(p3, p3 + n3);
    return s;
}
void groupby_free(b200_groupby_state *s) { delete s; }
void groupby_push(b200_groupby_state *s, const Batch &in) {
    if (!s) fail("null groupby state");
    const std::vector<int32_t> cols = jit::referenced_columns(B200_OP_AGG, s->partial_plan.data(), s->partial_plan.size());
    std::unique_ptr<Batch> part(op_agg(materialize_batch(in, &cols), s->partial_plan.data(), s->partial_plan.size()));
    std::lock_guard<std::mutex> g(s->mu);
    s->partials.push_back(std::move(part));
}
Batch *groupby_finalize(b200_groupby_state *s) {
    if (!s) fail("null groupby state");
    std::lock_guard<std::mutex> g(s->mu);
    if (s->partials.empty()) fail("groupby: finalize before any batch was pushed");
    std::unique_ptr<Batch> all;
    if (s->partials.size() == 1) {
        all = std::make_unique<Batch>(*s->partials[0]);
    } else {
        all = std::make_unique<Batch>();
        const size_t nc = s->partials[0]->cols.size();
        for (auto &p : s->partials) all->num_rows += p->num_rows;
        for (size_t c = 0; c < nc; c++) {
            std::vector<const Column *> cs;
            for (auto &p : s->partials) cs.push_back(&p->cols[c]);
            all->cols.push_back(concat_columns(cs));
        }
    }
    std::unique_ptr<Batch> merged(op_agg(*all, s->final_plan.data(), s->final_plan.size()));
    Batch *out = op_project(*merged, s->project_plan.data(), s->project_plan.size(), nullptr, 0);
    op_done();
    return out;
}

Batch *op_agg_allreduce(b200_comm *comm, const Batch &in, const uint8_t *plan, size_t len) {
    if (!comm) fail("null communicator");
    return op_agg_first_rows(in, plan, len, nullptr, nullptr, comm);
}

}  // namespace b200

// Deterministic, dbgen-shaped synthetic TPC-H columns, generated either straight into HBM or into host
// buffers by the same __host__ __device__ functions (so both sides see identical data).
//
// dbgen itself cannot be obtained here (the reference clones and builds it at test time,
// benchmarking/tpch/data_generation.py:259-265); this generator follows the TPC-H specification's
// value domains and the dtypes the reference infers (data_generation.py:320-331: Int64 keys/quantity,
// Float64 money, Date, Utf8 flags) with a counter-based RNG, so any row range of any table can be
// produced independently on any GPU (scan partitions = contiguous row ranges).
// Differences from dbgen, stated: lineitems per order follow a fixed 28-lines-per-7-orders pattern
// (mean 4, like dbgen's uniform 1..7) instead of a random count; O_ORDERSTATUS follows the lines of that pattern;
// comments draw from a 48-word subset of the text grammar and P_NAME from 32 of the 92 colours; clerk, L_COMMENT,
// PS_COMMENT, N_COMMENT and R_COMMENT are not generated.  L_SUPPKEY is one of the part's four partsupp suppliers, as in dbgen.
#include <thread>

#include "common.h"
#include "ops.h"
#include "tpch_gen_core.h"

namespace b200 {

using namespace tpchgen;

// dictionary-valued string columns: lengths of the drawn words, (scan on the host side of the launch), then the bytes
__global__ void gen_dict_lens(int col, uint64_t seed, int64_t row_begin, int64_t n, const int32_t *__restrict__ dict_off, int64_t *__restrict__ lens) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int k = gen_dict_index(col, seed, row_begin + i);
        lens[i] = dict_off[k + 1] - dict_off[k];
    }
}
__global__ void gen_dict_bytes(int col, uint64_t seed, int64_t row_begin, int64_t n, const int32_t *__restrict__ dict_off,
                               const uint8_t *__restrict__ dict_bytes, const int64_t *__restrict__ offsets, uint8_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const int k = gen_dict_index(col, seed, row_begin + i);
        const int32_t b = dict_off[k], len = dict_off[k + 1] - b;
        const int64_t o = offsets[i];
        for (int32_t t = 0; t < len; t++) out[o + t] = dict_bytes[b + t];
    }
}
// generated strings (names, phones, addresses, comments): the same two passes, the value comes from gen_string
__global__ void gen_string_lens(int col, uint64_t seed, int64_t row_begin, int64_t n, int64_t *__restrict__ lens) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) lens[i] = gen_string(col, seed, row_begin + i, nullptr);
}
__global__ void gen_string_bytes(int col, uint64_t seed, int64_t row_begin, int64_t n, const int64_t *__restrict__ offsets, uint8_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) gen_string(col, seed, row_begin + i, out + offsets[i]);
}
static void dict_tables(int col, std::vector<int32_t> &off, std::string &bytes) {
    off.assign(1, 0);
    bytes.clear();
    for (int k = 0; k < dict_size(col); k++) {
        bytes += dict_word(col, k);
        off.push_back((int32_t)bytes.size());
    }
}

__global__ void gen_kernel(int col, int kind, uint64_t seed, int64_t row_begin, int64_t n, Scale sc, void *data, int64_t *offsets) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) gen_store(col, kind, seed, row_begin + i, i, sc, data, offsets);
    if (kind == 2 && blockIdx.x == 0 && threadIdx.x == 0) offsets[n] = n;
}

int64_t tpch_num_rows(int32_t table, double sf) {
    const Scale s = scale_of(sf);
    switch (table) {
        case B200_TPCH_LINEITEM: {
            // the official cardinalities at the standard scale factors, 6 001 215 x SF otherwise
            if (sf == 1.0) return 6001215;
            if (sf == 10.0) return 59986052;
            if (sf == 100.0) return 600037902;
            return std::max<int64_t>(1, (int64_t)(6001215.0 * sf + 0.5));
        }
        case B200_TPCH_ORDERS: return s.n_orders;
        case B200_TPCH_CUSTOMER: return s.n_customers;
        case B200_TPCH_SUPPLIER: return s.n_suppliers;
        case B200_TPCH_NATION: return 25;
        case B200_TPCH_REGION: return 5;
        case B200_TPCH_PART: return s.n_parts;
        case B200_TPCH_PARTSUPP: return 4 * s.n_parts;
    }
    fail("tpch: unknown table");
}

static std::vector<int> resolve_columns(int32_t table, const char *const *columns, int32_t ncolumns) {
    std::vector<int> out;
    if (!columns || ncolumns <= 0) {
        for (int c = 0; c < NUM_COLS; c++)
            if (kCols[c].table == table) out.push_back(c);
        return out;
    }
    for (int i = 0; i < ncolumns; i++) {
        int found = -1;
        for (int c = 0; c < NUM_COLS; c++)
            if (kCols[c].table == table && strcasecmp(kCols[c].name, columns[i]) == 0) found = c;
        if (found < 0) fail(std::string("tpch: unknown column ") + columns[i]);
        out.push_back(found);
    }
    return out;
}
static int kind_of(int col) {
    if (kCols[col].dtype == B200_DATE32) return 1;
    if (kCols[col].dtype == B200_LARGE_UTF8) return 2;
    return 0;
}

// small dimension-table string columns are built on the host
static void name_column_host(int col, int64_t rb, int64_t re, std::vector<int64_t> &offs, std::string &bytes) {
    offs.assign(1, 0);
    for (int64_t r = rb; r < re; r++) {
        bytes += col == N_NAME ? kNations[r] : kRegions[r];
        offs.push_back((int64_t)bytes.size());
    }
}

Batch *tpch_generate(int32_t table, double sf, int64_t row_begin, int64_t row_end, uint64_t seed, const char *const *columns,
                     int32_t ncolumns) {
    require_device();
    const int64_t total = tpch_num_rows(table, sf);
    if (row_end < 0) row_end = total;
    if (row_begin < 0 || row_end < row_begin || row_end > total) fail("tpch: row range out of bounds");
    const int64_t n = row_end - row_begin;
    const Scale sc = scale_of(sf);
    cudaStream_t st = ctx().stream;
    auto out = std::make_unique<Batch>();
    out->num_rows = n;
    for (int col : resolve_columns(table, columns, ncolumns)) {
        Column c;
        c.name = kCols[col].name;
        c.dtype = kCols[col].dtype;
        c.length = n;
        const int kind = kind_of(col);
        if (is_dict_col(col)) {
            std::vector<int32_t> doff;
            std::string dbytes;
            dict_tables(col, doff, dbytes);
            Buf d_off = dev_alloc(doff.size() * 4), d_bytes = dev_alloc(dbytes.size()), lens = dev_alloc((size_t)(n + 1) * 8);
            B200_CUDA(cudaMemcpyAsync(d_off->ptr, doff.data(), doff.size() * 4, cudaMemcpyHostToDevice, st));
            B200_CUDA(cudaMemcpyAsync(d_bytes->ptr, dbytes.data(), dbytes.size(), cudaMemcpyHostToDevice, st));
            c.offsets = dev_alloc((size_t)(n + 1) * 8);
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8));
            if (n) {
                gen_dict_lens<<<grid, 256, 0, st>>>(col, seed, row_begin, n, as<int32_t>(d_off), as<int64_t>(lens));
                count_launch();
            }
            exclusive_scan_i64(as<int64_t>(lens), as<int64_t>(c.offsets), n, as<int64_t>(c.offsets) + n);
            const int64_t total = read_i64(as<int64_t>(c.offsets) + n);
            c.data = dev_alloc((size_t)std::max<int64_t>(total, 1));
            c.data_bytes = total;
            if (n) {
                gen_dict_bytes<<<grid, 256, 0, st>>>(col, seed, row_begin, n, as<int32_t>(d_off), as<uint8_t>(d_bytes), as<int64_t>(c.offsets),
                                                     as<uint8_t>(c.data));
                count_launch();
            }
            B200_CUDA(cudaStreamSynchronize(st));  // the host dictionary buffers above
        } else if (is_generated_string(col)) {
            Buf lens = dev_alloc((size_t)(n + 1) * 8);
            c.offsets = dev_alloc((size_t)(n + 1) * 8);
            const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8));
            if (n) {
                gen_string_lens<<<grid, 256, 0, st>>>(col, seed, row_begin, n, as<int64_t>(lens));
                count_launch();
            }
            exclusive_scan_i64(as<int64_t>(lens), as<int64_t>(c.offsets), n, as<int64_t>(c.offsets) + n);
            const int64_t total = read_i64(as<int64_t>(c.offsets) + n);
            c.data = dev_alloc((size_t)std::max<int64_t>(total, 1));
            c.data_bytes = total;
            if (n) {
                gen_string_bytes<<<grid, 256, 0, st>>>(col, seed, row_begin, n, as<int64_t>(c.offsets), as<uint8_t>(c.data));
                count_launch();
            }
        } else if (col == N_NAME || col == R_NAME) {
            std::vector<int64_t> offs;
            std::string bytes;
            name_column_host(col, row_begin, row_end, offs, bytes);
            c.offsets = dev_alloc(offs.size() * 8);
            c.data = dev_alloc(bytes.size());
            c.data_bytes = (int64_t)bytes.size();
            B200_CUDA(cudaMemcpyAsync(c.offsets->ptr, offs.data(), offs.size() * 8, cudaMemcpyHostToDevice, st));
            if (!bytes.empty()) B200_CUDA(cudaMemcpyAsync(c.data->ptr, bytes.data(), bytes.size(), cudaMemcpyHostToDevice, st));
            B200_CUDA(cudaStreamSynchronize(st));
        } else {
            if (col == N_REGIONKEY) {
                std::vector<int64_t> v;
                for (int64_t r = row_begin; r < row_end; r++) v.push_back(kNationRegion[r]);
                c.data = dev_alloc((size_t)n * 8);
                c.data_bytes = n * 8;
                B200_CUDA(cudaMemcpyAsync(c.data->ptr, v.data(), (size_t)n * 8, cudaMemcpyHostToDevice, st));
                B200_CUDA(cudaStreamSynchronize(st));
            } else {
                const size_t w = kind == 0 ? 8 : (kind == 1 ? 4 : 1);
                c.data = dev_alloc((size_t)std::max<int64_t>(n, 1) * w);
                c.data_bytes = n * (int64_t)w;
                if (kind == 2) c.offsets = dev_alloc((size_t)(n + 1) * 8);
                const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8));
                gen_kernel<<<grid, 256, 0, st>>>(col, kind, seed, row_begin, n, sc, c.data->ptr, as<int64_t>(c.offsets));
                count_launch();
            }
        }
        out->cols.push_back(std::move(c));
    }
    B200_CUDA(cudaStreamSynchronize(st));
    return out.release();
}

// ---- host generation (same functions, std::thread fan-out) -----------------------------------------
void tpch_generate_host(int32_t table, double sf, int64_t row_begin, int64_t row_end, uint64_t seed, const char *const *columns,
                        int32_t ncolumns, ArrowArray *out_array, ArrowSchema *out_schema) {
    const int64_t total = tpch_num_rows(table, sf);
    if (row_end < 0) row_end = total;
    if (row_begin < 0 || row_end < row_begin || row_end > total) fail("tpch: row range out of bounds");
    const int64_t n = row_end - row_begin;
    const Scale sc = scale_of(sf);
    std::vector<HostCol> cols;
    for (int col : resolve_columns(table, columns, ncolumns)) {
        HostCol h{kCols[col].name, kCols[col].dtype, n, nullptr, nullptr, 0};
        const int kind = kind_of(col);
        if (is_dict_col(col)) {
            std::vector<std::string> words;
            for (int k = 0; k < dict_size(col); k++) words.push_back(dict_word(col, k));
            h.offsets = (int64_t *)malloc((size_t)(n + 1) * 8);
            int64_t total = 0;
            for (int64_t i = 0; i < n; i++) {
                h.offsets[i] = total;
                total += (int64_t)words[(size_t)gen_dict_index(col, seed, row_begin + i)].size();
            }
            h.offsets[n] = total;
            h.data = malloc((size_t)total + 8);
            for (int64_t i = 0; i < n; i++) {
                const std::string &w = words[(size_t)gen_dict_index(col, seed, row_begin + i)];
                memcpy((uint8_t *)h.data + h.offsets[i], w.data(), w.size());
            }
            h.data_bytes = total;
        } else if (is_generated_string(col)) {
            h.offsets = (int64_t *)malloc((size_t)(n + 1) * 8);
            int64_t total = 0;
            for (int64_t i = 0; i < n; i++) {
                h.offsets[i] = total;
                total += gen_string(col, seed, row_begin + i, nullptr);
            }
            h.offsets[n] = total;
            h.data = malloc((size_t)total + 8);
            const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), n / 65536 + 1));
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) {
                const int64_t b = n * t / nt, e = n * (t + 1) / nt;
                th.emplace_back([=, &h] {
                    for (int64_t i = b; i < e; i++) gen_string(col, seed, row_begin + i, (uint8_t *)h.data + h.offsets[i]);
                });
            }
            for (auto &x : th) x.join();
            h.data_bytes = total;
        } else if (col == N_NAME || col == R_NAME) {
            std::vector<int64_t> offs;
            std::string bytes;
            name_column_host(col, row_begin, row_end, offs, bytes);
            h.offsets = (int64_t *)malloc(offs.size() * 8);
            memcpy(h.offsets, offs.data(), offs.size() * 8);
            h.data = malloc(bytes.size() + 8);
            memcpy(h.data, bytes.data(), bytes.size());
            h.data_bytes = (int64_t)bytes.size();
        } else if (col == N_REGIONKEY) {
            h.data = malloc((size_t)n * 8 + 8);
            for (int64_t r = row_begin; r < row_end; r++) ((int64_t *)h.data)[r - row_begin] = kNationRegion[r];
            h.data_bytes = n * 8;
        } else {
            const size_t w = kind == 0 ? 8 : (kind == 1 ? 4 : 1);
            h.data = malloc((size_t)n * w + 8);
            h.data_bytes = n * (int64_t)w;
            if (kind == 2) h.offsets = (int64_t *)malloc((size_t)(n + 1) * 8);
            const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), n / 65536 + 1));
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++) {
                const int64_t b = n * t / nt, e = n * (t + 1) / nt;
                th.emplace_back([=, &h] {
                    for (int64_t i = b; i < e; i++) gen_store(col, kind, seed, row_begin + i, i, sc, h.data, h.offsets);
                });
            }
            for (auto &x : th) x.join();
            if (kind == 2) h.offsets[n] = n;
        }
        cols.push_back(h);
    }
    export_host_columns(cols, n, out_array, out_schema);
}

}  // namespace b200

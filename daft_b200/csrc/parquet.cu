// Parquet -> HBM scan: the reference's stream_parquet (src/daft-parquet/src/read.rs:553) behind its scan-task source
// (src/daft-local-execution/src/sources/scan_task.rs:33), re-designed for the device: the host only reads bytes and page HEADERS --
// footer, column-chunk byte ranges, the page table (parquet_read.h) -- the compressed column chunks go to HBM as they are, and every
// page is decompressed and decoded there (parquet_core.h), straight into the Arrow buffers the operators read.
//
// Parallelism is over PAGES: one CUDA thread runs one page through a stage (Snappy, run-length levels and dictionary indices are
// sequential inside a page; a column chunk of a 1 M-row row group holds tens to hundreds of pages, a scan of many row groups and
// columns thousands).  Bytes moved per value: the compressed page once in, the decompressed page once out and once in, the value
// once out -- all in HBM; PCIe carries the compressed bytes only.
#include <mutex>

#include "common.h"
#include "ops.h"
#include "parquet_read.h"

namespace b200 {

using namespace pq;

// One page per WARP, decoded by its lane 0: the stages are data-dependent sequential walks (Snappy elements, run headers), so 32
// pages in one warp would serialise each other's branches (first version: 149 ms to decompress 58 MB, gpurun_out/r02b_parquet_scan.txt);
// a warp of its own lets every page run at the speed of one thread, and puts 32 x more warps on the SMs for the same page count.
constexpr int PQ_WARPS_PER_BLOCK = 4;
__global__ void __launch_bounds__(32 * PQ_WARPS_PER_BLOCK) pq_stage_kernel(ScanCtx ctx, int stage, int count) {
    const int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (i >= count || (threadIdx.x & 31) != 0) return;
    switch (stage) {
        case 0: stage_decompress(ctx, i); break;
        case 1: stage_dictionary(ctx, i); break;
        case 2: stage_values(ctx, i); break;
        default: stage_string_bytes(ctx, i); break;
    }
}
// one byte per row -> LSB-first bitmap (one 32-bit word per thread), zero bytes counted
__global__ void pq_pack_kernel(const uint8_t *__restrict__ bytes01, int64_t n, uint32_t *__restrict__ words, unsigned long long *__restrict__ zeros) {
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, nwords = (n + 31) / 32;
    unsigned long long z = 0;
    for (; w < nwords; w += stride) {
        uint32_t v = 0;
        const int64_t base = w * 32;
        const int lim = (int)(n - base < 32 ? n - base : 32);
        for (int b = 0; b < lim; b++) {
            const uint32_t bit = bytes01[base + b] ? 1u : 0u;
            v |= bit << b;
            z += bit ^ 1u;
        }
        words[w] = v;
    }
    if (z) atomicAdd(zeros, z);
}

__global__ void pq_convert_kernel(int kind, const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int64_t n, int64_t factor) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) convert_one(kind, src, dst, i, factor);
}

namespace {
struct DeviceExec {
    struct Mem {
        void *ptr = nullptr;
        Buf buf;
    };
    Mem alloc(size_t bytes) {
        Mem m;
        m.buf = dev_alloc(std::max<size_t>(bytes, 8));
        m.ptr = m.buf->ptr;
        return m;
    }
    // the planner reads the column chunks straight into page-locked memory (kept across scans, grown on demand): no second host copy,
    // and the copy to HBM runs at PCIe speed instead of through the driver's bounce buffers
    uint8_t *stage(size_t bytes) {
        static uint8_t *buf = nullptr;
        static size_t cap = 0;
        if (bytes > cap) {
            if (buf) cudaFreeHost(buf);
            buf = nullptr;
            cap = 0;
            const size_t want = std::max<size_t>(bytes + bytes / 4, (size_t)64 << 20);
            B200_CUDA(cudaHostAlloc((void **)&buf, want, cudaHostAllocDefault));
            cap = want;
        }
        return buf;
    }
    void upload(Mem &m, const void *src, size_t bytes) {
        if (bytes) B200_CUDA(cudaMemcpyAsync(m.ptr, src, bytes, cudaMemcpyHostToDevice, ctx().stream));
        B200_CUDA(cudaStreamSynchronize(ctx().stream));   // `src` is pageable host memory owned by the planner
    }
    void download(void *dst, const void *src, size_t bytes) {
        if (bytes) B200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx().stream));
        B200_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    void run(int stage, const ScanCtx &c, int count) {
        if (count <= 0) return;
        KernelTimer t(stage == 0 ? "parquet_decompress" : stage == 1 ? "parquet_dictionary" : stage == 2 ? "parquet_values" : "parquet_string_bytes");
        pq_stage_kernel<<<(count + PQ_WARPS_PER_BLOCK - 1) / PQ_WARPS_PER_BLOCK, 32 * PQ_WARPS_PER_BLOCK, 0, ctx().stream>>>(c, stage, count);
        count_launch();
        B200_CUDA(cudaGetLastError());
    }
    int64_t scan(const int64_t *lengths, int64_t *offsets, int64_t n) {
        exclusive_scan_i64(lengths, offsets, n, offsets + n);
        return read_i64(offsets + n);
    }
    void convert(int kind, const uint8_t *src, uint8_t *dst, int64_t n, int64_t factor) {
        if (n <= 0) return;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8));
        pq_convert_kernel<<<grid, 256, 0, ctx().stream>>>(kind, src, dst, n, factor);
        count_launch();
    }
    int64_t pack(const uint8_t *bytes01, int64_t n, uint8_t *bitmap) {
        if (n <= 0) return 0;
        Buf zeros = dev_alloc_zero(8);
        const int64_t nwords = (n + 31) / 32;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((nwords + 255) / 256, (int64_t)ctx().sm_count * 8));
        pq_pack_kernel<<<grid, 256, 0, ctx().stream>>>(bytes01, n, (uint32_t *)bitmap, as<unsigned long long>(zeros));
        count_launch();
        return read_i64(as<int64_t>(zeros));
    }
};
}  // namespace

Batch *parquet_read(const char *path, const char *const *columns, int32_t ncolumns, int64_t rg_begin, int64_t rg_end) {
    require_device();
    if (!path) fail("parquet: null path");
    // scans share the page-locked staging buffer: one at a time (entry points may be called from several host threads)
    static std::mutex scan_mu;
    std::lock_guard<std::mutex> scan_lock(scan_mu);
    std::vector<std::string> want;
    for (int i = 0; i < ncolumns; i++) want.push_back(columns[i]);
    DeviceExec x;
    int64_t rows = 0;
    std::vector<ColumnResult<DeviceExec>> cols = scan_file(x, path, want, rg_begin, rg_end, &rows);
    auto out = std::make_unique<Batch>();
    out->num_rows = rows;
    for (auto &r : cols) {
        Column c;
        c.name = r.name;
        c.dtype = r.dtype;
        c.length = rows;
        c.null_count = r.null_count;
        c.validity = r.null_count > 0 ? r.validity.buf : nullptr;
        c.offsets = r.offsets.buf;
        c.data = r.data.buf;
        c.data_bytes = r.data_bytes;
        out->cols.push_back(std::move(c));
    }
    op_done();
    return out.release();
}

void parquet_file_info(const char *path, int64_t *row_groups, int64_t *rows) {
    if (!path) fail("parquet: null path");
    FileBytes fb(path);
    const FileMeta fm = read_footer(fb);
    *row_groups = (int64_t)fm.row_groups.size();
    *rows = fm.num_rows;
}

}  // namespace b200

// Host build of the Parquet scan (TEST INFRASTRUCTURE): the planner of daft_b200/csrc/parquet_read.h and the per-page decoders of
// parquet_core.h -- the code the device runs one page per CUDA thread -- driven by a plain loop over malloc'd buffers, so that
// tests/test_parquet_cpu.py can hold every decoder to pyarrow's reader without a GPU.  Nothing under daft_b200/ uses this file.
#include <cstdlib>
#include <memory>

#include "../daft_b200/csrc/parquet_read.h"

using namespace b200::pq;

struct HostExec {
    struct Mem {
        void *ptr = nullptr;
        std::shared_ptr<void> keep;
    };
    Mem alloc(size_t bytes) {
        Mem m;
        m.keep = std::shared_ptr<void>(calloc(bytes + 16, 1), free);
        m.ptr = m.keep.get();
        return m;
    }
    std::vector<uint8_t> staging;
    uint8_t *stage(size_t bytes) {
        staging.resize(bytes);
        return staging.data();
    }
    void upload(Mem &m, const void *src, size_t bytes) {
        if (bytes) memcpy(m.ptr, src, bytes);
    }
    void download(void *dst, const void *src, size_t bytes) {
        if (bytes) memcpy(dst, src, bytes);
    }
    void run(int stage, const ScanCtx &ctx, int count) {
        for (int i = 0; i < count; i++) {
            switch (stage) {
                case 0: stage_decompress(ctx, i); break;
                case 1: stage_dictionary(ctx, i); break;
                case 2: stage_values(ctx, i); break;
                default: stage_string_bytes(ctx, i); break;
            }
        }
    }
    int64_t scan(const int64_t *lengths, int64_t *offsets, int64_t n) {
        int64_t acc = 0;
        for (int64_t i = 0; i < n; i++) {
            offsets[i] = acc;
            acc += lengths[i];
        }
        offsets[n] = acc;
        return acc;
    }
    void convert(int kind, const uint8_t *src, uint8_t *dst, int64_t n, int64_t factor) {
        for (int64_t i = 0; i < n; i++) convert_one(kind, src, dst, i, factor);
    }
    int64_t pack(const uint8_t *bytes01, int64_t n, uint8_t *bitmap) {
        int64_t zeros = 0;
        memset(bitmap, 0, (size_t)((n + 7) / 8));
        for (int64_t i = 0; i < n; i++) {
            if (bytes01[i]) bitmap[i >> 3] |= (uint8_t)(1u << (i & 7));
            else zeros++;
        }
        return zeros;
    }
};

struct HostScan {
    std::vector<ColumnResult<HostExec>> cols;
    int64_t rows = 0;
    std::string error;
};

extern "C" {
__attribute__((visibility("default"))) HostScan *pq_host_read(const char *path, const char *const *columns, int ncolumns, long long rg_begin, long long rg_end) {
    auto *s = new HostScan;
    try {
        HostExec x;
        std::vector<std::string> want;
        for (int i = 0; i < ncolumns; i++) want.push_back(columns[i]);
        s->cols = scan_file(x, path, want, rg_begin, rg_end, &s->rows);
    } catch (const std::exception &e) {
        s->error = e.what();
    }
    return s;
}
__attribute__((visibility("default"))) const char *pq_host_error(HostScan *s) { return s->error.empty() ? nullptr : s->error.c_str(); }
__attribute__((visibility("default"))) long long pq_host_rows(HostScan *s) { return s->rows; }
__attribute__((visibility("default"))) int pq_host_num_columns(HostScan *s) { return (int)s->cols.size(); }
__attribute__((visibility("default"))) const char *pq_host_column(HostScan *s, int i, int *dtype, long long *null_count, long long *data_bytes, void **data, void **validity,
                                                                   void **offsets) {
    const auto &c = s->cols[(size_t)i];
    *dtype = c.dtype;
    *null_count = c.null_count;
    *data_bytes = c.data_bytes;
    *data = c.data.ptr;
    *validity = c.validity.ptr;
    *offsets = c.offsets.ptr;
    return c.name.c_str();
}
__attribute__((visibility("default"))) void pq_host_free(HostScan *s) { delete s; }
// row groups and rows of a file (footer only)
__attribute__((visibility("default"))) long long pq_host_row_groups(const char *path, long long *rows) {
    try {
        FileBytes fb(path);
        const FileMeta fm = read_footer(fb);
        *rows = fm.num_rows;
        return (long long)fm.row_groups.size();
    } catch (const std::exception &) {
        return -1;
    }
}
}

#ifdef PQ_HOST_MAIN
// standalone reader for sanitizer runs: g++ -fsanitize=address,undefined -DPQ_HOST_MAIN tests/pq_host.cpp -o pq_asan; ./pq_asan file.parquet
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    HostScan *s = pq_host_read(argv[1], nullptr, 0, 0, -1);
    if (!s->error.empty()) printf("err %s\n", s->error.c_str());
    else printf("ok %lld rows %zu columns\n", (long long)s->rows, s->cols.size());
    pq_host_free(s);
    return 0;
}
#endif

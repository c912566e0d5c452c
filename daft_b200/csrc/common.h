// Internal host-side types of libdaftb200.so: device buffers, columns, batches, error plumbing.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/daft_b200.h"

namespace b200 {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const std::string &msg) { throw Error(msg); }

#define B200_CUDA(expr)                                                                              \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            ::b200::fail(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                         std::to_string(__LINE__) + " in " #expr);                                   \
    } while (0)

// ---- library context -------------------------------------------------------------------------
struct Context {
    int device = -1;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    bool ready = false;
};
Context &ctx();
void require_device();  // throws unless b200_init succeeded (no CPU fallback)
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- device memory ---------------------------------------------------------------------------
// Stream-ordered allocations from the device's default mempool (release threshold = unlimited so
// freed blocks stay cached in HBM between operators).
struct DevAlloc {
    void *ptr = nullptr;
    size_t bytes = 0;
    bool owned = true;
    size_t cached_bytes = 0;  // > 0: the block belongs to the size-class cache (its rounded size), see runtime.cu
    ~DevAlloc();
};
using Buf = std::shared_ptr<DevAlloc>;
Buf dev_alloc(size_t bytes);                    // uninitialised
size_t dev_available_bytes();  // driver-free + parked cache blocks + unused pool reservation
Buf dev_alloc_zero(size_t bytes);
Buf dev_borrow(const void *ptr, size_t bytes);  // non-owning view
Buf dev_sub(const Buf &parent, size_t offset, size_t bytes);  // view into `parent` that keeps it alive
template <typename T>
inline T *as(const Buf &b) { return b ? reinterpret_cast<T *>(b->ptr) : nullptr; }

// ---- columns / batches -----------------------------------------------------------------------
inline int dtype_width(int32_t dt) {
    switch (dt) {
        case B200_I8: case B200_U8: return 1;
        case B200_I16: case B200_U16: return 2;
        case B200_I32: case B200_U32: case B200_F32: case B200_DATE32: return 4;
        case B200_I64: case B200_U64: case B200_F64: case B200_TIMESTAMP: return 8;
        default: return 0;
    }
}
inline bool is_string(int32_t dt) { return dt == B200_LARGE_UTF8 || dt == B200_LARGE_BINARY; }
inline bool is_float(int32_t dt) { return dt == B200_F32 || dt == B200_F64; }
inline bool is_signed_int(int32_t dt) { return dt >= B200_I8 && dt <= B200_I64; }
inline bool is_unsigned_int(int32_t dt) { return dt >= B200_U8 && dt <= B200_U64; }
inline bool is_integer(int32_t dt) { return is_signed_int(dt) || is_unsigned_int(dt); }
const char *dtype_name(int32_t dt);

struct Column;
// A VIEW column is a selection vector over another column: row i of the view is row sel[i] of `base` (a negative
// index is a null row).  Join outputs are views (late materialisation: a probe chain moves row ids, not payload
// columns; the reference pipelines morsel-sized intermediates through its probe operators for the same reason,
// src/daft-local-execution/src/join/inner_join.rs:11-85).  A view is gathered into a plain column the first time an
// operator needs its values; the result is cached in `state`, which every copy of the view shares.
struct ViewState;
struct Column {
    std::string name;
    int32_t dtype = B200_NULL;
    int64_t length = 0;
    int64_t null_count = 0;  // exact (plain columns)
    Buf validity;            // null when null_count == 0
    Buf offsets;             // strings, lists
    Buf data;
    std::shared_ptr<const Column> child;  // B200_LIST: the flat child column the offsets index
    int64_t data_bytes = 0;
    // view part (all empty for a plain column)
    std::shared_ptr<const Column> base;  // a plain column
    Buf sel;                             // int64[length]
    bool sel_may_be_negative = false;
    std::shared_ptr<ViewState> state;
    bool is_view() const { return (bool)sel; }
    bool nullable() const { return dtype == B200_NULL || null_count > 0; }
};

}  // namespace b200

struct b200_batch {
    std::vector<b200::Column> cols;
    int64_t num_rows = 0;
};

namespace b200 {
using Batch = ::b200_batch;

// ---- small device utilities (util.cu) --------------------------------------------------------
// exclusive scan of int64 `in[n]` -> `out[n]`, returns total via *total_dev (device int64) if non-null
void exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, int64_t *total_dev);
int64_t count_set_bits(const uint8_t *bitmap, int64_t nbits);  // syncs
int64_t read_i64(const int64_t *dev_ptr);                      // syncs
void fill_u8(void *ptr, int value, size_t bytes);
// End of an operator: results are stream-ordered (every buffer of this library lives and dies on ctx().stream), so the
// host does not wait for the device here -- the next operator's launches queue behind this one's.  Launch errors are
// still reported; B200_SYNC_OPS=1 restores a blocking wait (debugging).
void op_done();
// gather rows of `src` at `idx[n]` (int64, -1 => null row) into a new column
Column gather_column(const Column &src, const int64_t *idx, int64_t n, bool idx_may_be_negative);
// `permutation`: idx holds every source row exactly once (sort): string outputs are sized without a host round trip
Batch *gather_batch(const Batch &src, const int64_t *idx, int64_t n, bool idx_may_be_negative, bool permutation = false);
Column slice_column(const Column &src, int64_t start, int64_t end);
Column concat_columns(const std::vector<const Column *> &cols);
// any integer column -> int64 index buffer (null -> -1)
Buf indices_to_i64(const Column &c);
void check_take_indices(const int64_t *idx, int64_t n, int64_t limit, bool null_marked);  // fails on idx >= limit / negative
Column make_u64_column(const std::string &name, Buf data, int64_t n);

// ---- views (runtime.cu) ------------------------------------------------------------------------
// rows idx[0..n) of `src` as view columns (views of views compose their selection vectors: one small gather per
// distinct selection vector of `src`)
std::vector<Column> make_view_columns(const Batch &src, const Buf &idx, int64_t n, bool idx_may_be_negative);
// the plain form of `c` (gathers a view on first use; cached)
Column materialize(const Column &c);
// a batch whose columns `cols` (all when null) are plain; view columns sharing a selection vector are gathered together
Batch materialize_batch(const Batch &b, const std::vector<int32_t> *cols = nullptr);
inline void require_plain(const Column &c) {
    if (c.is_view()) fail("internal error: operator reached an unmaterialised view column '" + c.name + "'");
}

// ---- per-kernel device timing (runtime.cu; bench support) ----------------------------------------
// While enabled (b200_kernel_timing), a KernelTimer brackets the launches in its scope with CUDA events on the library
// stream; kernel_timing_report() waits for the stream and returns {"name": {"launches": n, "ms": total}, ...} as JSON
// text, then forgets the samples.  Disabled = no events, no cost.
void kernel_timing_enable(bool on);
bool kernel_timing_enabled();
std::string kernel_timing_report();
struct KernelTimer {
    explicit KernelTimer(const char *name);
    ~KernelTimer();
    int slot = -1;
};

// ---- hashing (hash.cu) -----------------------------------------------------------------------
// out[i] = xxh3_64(col[i], seed ? seed[i] : 0) with Daft's null rules
void hash_column(const Column &c, const uint64_t *seed, uint64_t *out);
Buf hash_rows(const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint64_t *seed);

}  // namespace b200

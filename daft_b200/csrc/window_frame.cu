// Window aggregates over ROWS and RANGE frames on the device: the reference's WindowPartitionAndDynamicFrameSink
// (src/daft-local-execution/src/sinks/window_partition_and_dynamic_frame.rs) calls RecordBatch::window_agg_dynamic_frame
// (src/daft-recordbatch/src/ops/window.rs:299-385) once per partition; here every partition of the sorted batch is walked by its own
// thread, all of them in one launch, through the state machine of window_frame_core.h (one thread per partition because a sliding
// float sum is defined by the order of its adds and removes).  Inputs: the value column already cast by the host to the type the
// reference accumulates in (Int64 / UInt64 / Float32 / Float64, window_states/sum.rs:137-158, mean.rs:92-100) and the
// partition_begin column of b200_window_ranks.
#include "common.h"
#include "ops.h"
#include "window_frame_core.h"

namespace b200 {

using namespace wf;

template <typename T>
__global__ void window_frame_kernel(FrameSpec f, const T *__restrict__ v, const uint8_t *__restrict__ bits, const int64_t *__restrict__ pbegin,
                                    int64_t n, T *out_t, uint64_t *out_u64, double *out_f64, uint8_t *ok, int *negative_frame) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || pbegin[i] != i) return;   // the first row of a partition works for all of it
    int64_t e = i + 1;
    while (e < n && pbegin[e] == i) e++;
    if (!frame_partition<T>(f, v, bits, i, e - i, out_t, out_u64, out_f64, ok)) *negative_frame = 1;
}
// one byte per row -> LSB-first validity bitmap, one output byte per thread
__global__ void window_frame_pack_kernel(const uint8_t *__restrict__ ok, int64_t n, uint8_t *__restrict__ bitmap) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= (n + 7) / 8) return;
    uint32_t b = 0;
    for (int k = 0; k < 8; k++)
        if (w * 8 + k < n && ok[w * 8 + k]) b |= 1u << k;
    bitmap[w] = (uint8_t)b;
}

static inline int wf_grid(int64_t n) { return (int)std::max<int64_t>(1, (n + 255) / 256); }

Batch *op_window_frame(const Column *values, const Column &partition_begin, const Column *order, int32_t agg, int32_t start_kind, int64_t start,
                       int32_t end_kind, int64_t end, int32_t descending, int64_t min_periods, const char *name) {
    require_device();
    require_plain(partition_begin);
    const int64_t n = partition_begin.length;
    if (partition_begin.dtype != B200_I64 || partition_begin.null_count > 0) fail("window_frame: partition_begin must be a non-null Int64 column");
    if (agg < WF_SUM || agg > WF_MAX) fail("window_frame: unknown aggregate");
    if (start_kind < WF_UNBOUNDED || start_kind > WF_RANGE || end_kind < WF_UNBOUNDED || end_kind > WF_RANGE) fail("window_frame: unknown boundary kind");
    if (start_kind == WF_ROWS && end_kind == WF_ROWS && start > end && start - end > 1) fail("Negative frame size is not allowed");   // window.rs:156-160
    const bool counts = agg == WF_COUNT_ALL || agg == WF_COUNT_VALID || agg == WF_COUNT_NULL;
    FrameSpec f{agg, WF_I64, start_kind, end_kind, start, end, min_periods, descending ? 1 : 0, nullptr, nullptr};
    if (start_kind == WF_RANGE || end_kind == WF_RANGE) {
        if (!order) fail("window_frame: a RANGE frame needs the ORDER BY column");
        require_plain(*order);
        if (order->dtype != B200_I64 || order->length != n) fail("window_frame: the ORDER BY column of a RANGE frame must be Int64 (the host casts), one value per row");
        f.order = as<int64_t>(order->data);
        f.order_bits = order->null_count > 0 ? as<uint8_t>(order->validity) : nullptr;
    }
    const uint8_t *bits = nullptr;
    const void *v = nullptr;
    int32_t out_dtype = B200_U64;
    if (values) {
        require_plain(*values);
        if (values->length != n) fail("window_frame: the value column and partition_begin differ in length");
        if (values->null_count > 0) bits = as<uint8_t>(values->validity);
    }
    if (!counts) {
        if (!values) fail("window_frame: this aggregate needs a value column");
        switch (values->dtype) {
            case B200_I64: f.vtype = WF_I64; break;
            case B200_U64: f.vtype = WF_U64; break;
            case B200_F32: f.vtype = WF_F32; break;
            case B200_F64: f.vtype = WF_F64; break;
            default: fail("window_frame: values must be Int64, UInt64, Float32 or Float64 (the host casts)");
        }
        if (agg == WF_MEAN && f.vtype != WF_F64) fail("window_frame: mean accumulates Float64 values");
        if ((agg == WF_MIN || agg == WF_MAX) && f.vtype != WF_I64 && f.vtype != WF_U64) fail("window_frame: min / max over frames are built for integer-like values");
        v = values->data->ptr;
        out_dtype = agg == WF_MEAN ? B200_F64 : values->dtype;
    } else if (values && values->dtype == B200_NULL) {
        fail("window_frame: count over a Null column is not built");
    }
    const int width = dtype_width(out_dtype);
    Buf out = dev_alloc((size_t)std::max<int64_t>(n, 1) * width);
    Buf ok = dev_alloc((size_t)std::max<int64_t>(n, 1));
    Buf bitmap = dev_alloc_zero((size_t)(std::max<int64_t>(n, 1) + 7) / 8 + 8);
    Buf negative = dev_alloc_zero(8);
    if (n > 0) {
        cudaStream_t st = ctx().stream;
        const int64_t *pb = as<int64_t>(partition_begin.data);
        KernelTimer t("window_frame");
        if (counts || f.vtype == WF_I64)
            window_frame_kernel<int64_t><<<wf_grid(n), 256, 0, st>>>(f, counts ? nullptr : (const int64_t *)v, bits, pb, n, as<int64_t>(out), as<uint64_t>(out), nullptr, as<uint8_t>(ok), as<int>(negative));
        else if (f.vtype == WF_U64)
            window_frame_kernel<uint64_t><<<wf_grid(n), 256, 0, st>>>(f, (const uint64_t *)v, bits, pb, n, as<uint64_t>(out), nullptr, nullptr, as<uint8_t>(ok), as<int>(negative));
        else if (f.vtype == WF_F32)
            window_frame_kernel<float><<<wf_grid(n), 256, 0, st>>>(f, (const float *)v, bits, pb, n, as<float>(out), nullptr, nullptr, as<uint8_t>(ok), as<int>(negative));
        else
            window_frame_kernel<double><<<wf_grid(n), 256, 0, st>>>(f, (const double *)v, bits, pb, n, as<double>(out), nullptr, as<double>(out), as<uint8_t>(ok), as<int>(negative));
        count_launch();
        B200_CUDA(cudaGetLastError());
        window_frame_pack_kernel<<<wf_grid((n + 7) / 8), 256, 0, st>>>(as<uint8_t>(ok), n, as<uint8_t>(bitmap));
        count_launch();
        B200_CUDA(cudaGetLastError());
        if (read_i64(as<int64_t>(negative)) != 0) fail("Negative frame size is not allowed");   // window.rs:264-268, found by some row's RANGE search
    }
    Column c;
    c.name = name ? name : "window_frame";
    c.dtype = out_dtype;
    c.length = n;
    c.data = out;
    c.data_bytes = n * width;
    c.validity = bitmap;
    c.null_count = n - count_set_bits(as<uint8_t>(bitmap), n);
    if (c.null_count == 0) c.validity.reset();
    auto b = std::make_unique<Batch>();
    b->num_rows = n;
    b->cols.push_back(std::move(c));
    op_done();
    return b.release();
}

}  // namespace b200

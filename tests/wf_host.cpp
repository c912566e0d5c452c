// Host build of daft_b200/csrc/window_frame_core.h (TEST INFRASTRUCTURE): the state machine the device kernel runs per partition,
// driven the way window_frame_kernel drives it -- the row that begins a partition walks all of it.
#include "../daft_b200/csrc/window_frame_core.h"

using namespace b200::wf;

template <typename T>
static bool run(const FrameSpec &f, const void *v, const uint8_t *bits, const int64_t *pbegin, int64_t n, void *out, uint8_t *ok) {
    bool fine = true;
    for (int64_t i = 0; i < n; i++) {
        if (pbegin[i] != i) continue;
        int64_t e = i + 1;
        while (e < n && pbegin[e] == i) e++;
        fine = frame_partition<T>(f, (const T *)v, bits, i, e - i, (T *)out, (uint64_t *)out, (double *)out, ok) && fine;
    }
    return fine;
}

// -> 0 ok, 1 bad value type, 2 some frame had a negative size
extern "C" int wf_host_run(int agg, int vtype, int start_kind, long long start, int end_kind, long long end, int descending, long long min_periods,
                           const void *v, const uint8_t *bits, const long long *order, const uint8_t *order_bits, const long long *pbegin, long long n,
                           void *out, uint8_t *ok) {
    FrameSpec f{agg, vtype, start_kind, end_kind, start, end, min_periods, descending, (const int64_t *)order, order_bits};
    const bool counts = agg == WF_COUNT_ALL || agg == WF_COUNT_VALID || agg == WF_COUNT_NULL;
    const int64_t *pb = (const int64_t *)pbegin;
    bool fine;
    if (counts || vtype == WF_I64) fine = run<int64_t>(f, counts ? nullptr : v, bits, pb, n, out, ok);
    else if (vtype == WF_U64) fine = run<uint64_t>(f, v, bits, pb, n, out, ok);
    else if (vtype == WF_F32) fine = run<float>(f, v, bits, pb, n, out, ok);
    else if (vtype == WF_F64) fine = run<double>(f, v, bits, pb, n, out, ok);
    else return 1;
    return fine ? 0 : 2;
}

// Host build of daft_b200/csrc/window_frame_core.h (TEST INFRASTRUCTURE): the state machine the device kernel runs per partition,
// driven the way window_frame_kernel drives it -- the row that begins a partition walks all of it.
#include "../daft_b200/csrc/window_frame_core.h"

using namespace b200::wf;

template <typename T>
static void run(const FrameSpec &f, const void *v, const uint8_t *bits, const int64_t *pbegin, int64_t n, void *out, uint8_t *ok) {
    for (int64_t i = 0; i < n; i++) {
        if (pbegin[i] != i) continue;
        int64_t e = i + 1;
        while (e < n && pbegin[e] == i) e++;
        frame_partition<T>(f, (const T *)v, bits, i, e - i, (T *)out, (uint64_t *)out, (double *)out, ok);
    }
}

extern "C" int wf_host_run(int agg, int vtype, int has_start, long long start, int has_end, long long end, long long min_periods, const void *v,
                           const uint8_t *bits, const long long *pbegin, long long n, void *out, uint8_t *ok) {
    FrameSpec f{agg, vtype, has_start, has_end, start, end, min_periods};
    const bool counts = agg == WF_COUNT_ALL || agg == WF_COUNT_VALID || agg == WF_COUNT_NULL;
    const int64_t *pb = (const int64_t *)pbegin;
    if (counts || vtype == WF_I64) run<int64_t>(f, counts ? nullptr : v, bits, pb, n, out, ok);
    else if (vtype == WF_U64) run<uint64_t>(f, v, bits, pb, n, out, ok);
    else if (vtype == WF_F32) run<float>(f, v, bits, pb, n, out, ok);
    else if (vtype == WF_F64) run<double>(f, v, bits, pb, n, out, ok);
    else return 1;
    return 0;
}

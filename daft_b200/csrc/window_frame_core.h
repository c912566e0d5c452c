// Window aggregates over ROWS and RANGE frames: RecordBatch::window_agg_dynamic_frame
// (src/daft-recordbatch/src/ops/window.rs:299-385 -> window_agg_rows_incremental :388-438, window_agg_rows :441-483,
// window_agg_range_incremental :486-566, window_agg_range :569-633), one partition per call.  The reference slides an aggregate state over the partition -- rows that left the frame are REMOVED from the state, rows
// that entered it are ADDED (update_incremental_state, window.rs:274-296; states in ops/window_states/{sum,count,mean}.rs) -- so a
// float sum is the result of exactly that sequence of `+=` and `-=`; the same sequence runs here, which is why a partition is walked
// by one thread (partitions run in parallel).  min / max have no incremental state in the reference (window_states/mod.rs:47-49):
// every row re-aggregates its frame (window_agg_rows); here for integer-like values only.
//
// Host and device build of the same text (tests/wf_host.cpp compiles it with g++).
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define WF_HD __host__ __device__ __forceinline__
#else
#define WF_HD inline
#endif

namespace b200 {
namespace wf {

enum Agg { WF_SUM = 0, WF_COUNT_ALL = 1, WF_COUNT_VALID = 2, WF_COUNT_NULL = 3, WF_MEAN = 4, WF_MIN = 5, WF_MAX = 6 };
enum VType { WF_I64 = 0, WF_U64 = 1, WF_F32 = 2, WF_F64 = 3 };

enum Bound { WF_UNBOUNDED = 0, WF_ROWS = 1, WF_RANGE = 2 };

struct FrameSpec {
    int32_t agg, vtype;
    int32_t has_start, has_end;   // Bound: 0 UNBOUNDED PRECEDING / FOLLOWING, 1 row offset, 2 offset on the ORDER BY value (RANGE)
    int64_t start, end;           // offsets relative to the current row / the current row's order value (negative = preceding)
    int64_t min_periods;          // frames holding fewer rows give null
    int32_t descending;           // RANGE frames: the direction the partition is sorted in
    const int64_t *order;         // RANGE frames: the ORDER BY column as Int64 (sorted batch), its validity bits (null = all valid)
    const uint8_t *order_bits;
};

// calculate_row_frame_bounds (window.rs:136-163): usize saturating_add_signed, then min(total_rows)
WF_HD int64_t clamp_add(int64_t base, int64_t off, int64_t total) {
    if (off < 0) return off < -base ? 0 : base + off;
    return off > total - base ? total : base + off;
}
WF_HD bool bit_set(const uint8_t *bits, int64_t i) { return !bits || ((bits[i >> 3] >> (i & 7)) & 1); }

// calculate_range_frame_bounds (window.rs:167-272).  A RANGE boundary is searched FORWARD from where the previous frame's boundary
// was (`prev`): the first row whose order value v satisfies, ascending, v >= cur + start (frame start) / v > cur + end (frame end);
// descending, v <= cur + start / v < cur + end.  A comparison with a null on either side is false (`.get(0).unwrap_or(false)`), so
// null order values are stepped over and a null current value sends both boundaries to the end of the partition.  cur + offset wraps.
WF_HD int64_t range_search(const FrameSpec &f, int64_t base, int64_t row, int64_t total, int64_t prev, int64_t off, bool is_end) {
    int64_t idx = prev;
    if (!bit_set(f.order_bits, base + row)) return idx > total ? idx : total;
    const int64_t bound = (int64_t)((uint64_t)f.order[base + row] + (uint64_t)off);
    for (; idx < total; idx++) {
        if (!bit_set(f.order_bits, base + idx)) continue;
        const int64_t v = f.order[base + idx];
        const bool hit = f.descending ? (is_end ? v < bound : v <= bound) : (is_end ? v > bound : v >= bound);
        if (hit) break;
    }
    return idx;
}
// -> false when the frame would have a negative size (the reference raises "Negative frame size is not allowed")
WF_HD bool frame_bounds(const FrameSpec &f, int64_t base, int64_t row, int64_t total, int64_t prev_s, int64_t prev_e, int64_t &fs, int64_t &fe) {
    // descending + both boundaries RANGE: the reference swaps them (window.rs:507-513)
    const bool swap = f.descending && f.has_start == WF_RANGE && f.has_end == WF_RANGE;
    const int64_t so = swap ? f.end : f.start, eo = swap ? f.start : f.end;
    fs = f.has_start == WF_UNBOUNDED ? 0 : f.has_start == WF_ROWS ? clamp_add(row, f.start, total) : range_search(f, base, row, total, prev_s, so, false);
    fe = f.has_end == WF_UNBOUNDED ? total : f.has_end == WF_ROWS ? clamp_add(row + 1, f.end, total) : range_search(f, base, row, total, prev_e, eo, true);
    if (fe < fs) {
        fe = fs;
        return false;
    }
    return true;
}

template <typename T> struct Wrap { typedef T U; };
template <> struct Wrap<int64_t> { typedef uint64_t U; };   // `+=` / `-=` on i64 wrap in the reference's release build
template <typename T> WF_HD bool value_is_nan(T) { return false; }
template <> WF_HD bool value_is_nan<float>(float v) { return v != v; }
template <> WF_HD bool value_is_nan<double>(double v) { return v != v; }
template <typename T> WF_HD T nan_of() { return (T)0; }   // T::Native::from_f64(NAN).unwrap_or_else(zero): integers never count a NaN
template <> WF_HD float nan_of<float>() {   // f32::NAN / f64::NAN: the positive quiet NaN with an empty payload
    const uint32_t b = 0x7fc00000u;
    float x;
    memcpy(&x, &b, 4);
    return x;
}
template <> WF_HD double nan_of<double>() {
    const uint64_t b = 0x7ff8000000000000ULL;
    double x;
    memcpy(&x, &b, 8);
    return x;
}

// SumWindowState + CountWindowState (window_states/sum.rs:66-123, count.rs:29-72): NaNs are counted, not summed; the counters are
// usize and may pass through "negative" values when a frame starts after row 0 (the first update removes rows it never added,
// then adds them: window.rs:282-289) -- unsigned wrap-around, as in the reference's release build.
template <typename T> struct State {
    T sum;
    uint64_t valid_not_nan, nan, total, valid;
};
template <typename T> WF_HD void state_add(State<T> &s, const T *v, const uint8_t *bits, int64_t base, int64_t a, int64_t b) {
    typedef typename Wrap<T>::U U;
    s.total += (uint64_t)(b - a);
    for (int64_t i = a; i < b; i++) {
        if (!bit_set(bits, base + i)) continue;
        s.valid++;
        if (v) {
            const T x = v[base + i];
            if (value_is_nan(x)) s.nan++;
            else { s.sum = (T)((U)s.sum + (U)x); s.valid_not_nan++; }
        }
    }
}
template <typename T> WF_HD void state_remove(State<T> &s, const T *v, const uint8_t *bits, int64_t base, int64_t a, int64_t b) {
    typedef typename Wrap<T>::U U;
    s.total -= (uint64_t)(b - a);
    for (int64_t i = a; i < b; i++) {
        if (!bit_set(bits, base + i)) continue;
        s.valid--;
        if (v) {
            const T x = v[base + i];
            if (value_is_nan(x)) s.nan--;
            else { s.sum = (T)((U)s.sum - (U)x); s.valid_not_nan--; }
        }
    }
}

// One partition: rows [base, base + total) of the sorted batch.  `v` may be null for the counts.  Outputs are indexed like the
// inputs: out_t (sum / min / max, type T), out_u64 (counts), out_f64 (mean), ok = 1 where the row's result is not null.
// Returns false when some row's frame had a negative size.
template <typename T>
WF_HD bool frame_partition(const FrameSpec &f, const T *v, const uint8_t *bits, int64_t base, int64_t total, T *out_t, uint64_t *out_u64,
                           double *out_f64, uint8_t *ok) {
    const uint64_t min_periods = f.min_periods < 0 ? 0 : (uint64_t)f.min_periods;
    bool fine = true;
    if (f.agg == WF_MIN || f.agg == WF_MAX) {
        int64_t ps = 0, pe = 0;   // window_agg_range: the search resumes from EVERY previous row's frame (window.rs:609-610)
        for (int64_t row = 0; row < total; row++) {
            int64_t fs, fe;
            fine = frame_bounds(f, base, row, total, ps, pe, fs, fe) && fine;
            ps = fs;
            pe = fe;
            bool any = false;
            T best = (T)0;
            if (fe >= fs && (uint64_t)(fe - fs) >= min_periods)
                for (int64_t i = fs; i < fe; i++) {
                    if (!bit_set(bits, base + i)) continue;
                    const T x = v[base + i];
                    if (!any || (f.agg == WF_MIN ? x < best : x > best)) best = x;
                    any = true;
                }
            out_t[base + row] = any ? best : (T)0;
            ok[base + row] = any ? 1 : 0;
        }
        return fine;
    }
    State<T> s;
    s.sum = (T)0;
    s.valid_not_nan = s.nan = s.total = s.valid = 0;
    int64_t prev_s = 0, prev_e = 0;
    for (int64_t row = 0; row < total; row++) {
        int64_t fs, fe;   // the incremental path resumes a RANGE search from the last frame that was large enough (window.rs:525-552)
        fine = frame_bounds(f, base, row, total, prev_s, prev_e, fs, fe) && fine;
        const bool enough = fe >= fs && (uint64_t)(fe - fs) >= min_periods;
        if (enough) {
            if (fs > prev_s) state_remove(s, v, bits, base, prev_s, fs);
            if (fe > prev_e) state_add(s, v, bits, base, prev_e, fe);
            prev_s = fs;
            prev_e = fe;
        }
        // evaluate() runs for every row, on the state the last sufficient frame left behind; `enough` masks the result
        const int64_t o = base + row;
        switch (f.agg) {
            case WF_SUM:
                out_t[o] = s.nan > 0 ? nan_of<T>() : s.sum;
                ok[o] = enough && s.valid_not_nan > 0;
                break;
            case WF_MEAN: {   // sum series / count series, both as Float64 (mean.rs:75-80); the count is of VALID rows, NaNs included
                const double sum = s.nan > 0 ? nan_of<double>() : (double)s.sum;
                const bool k = enough && s.valid_not_nan > 0;
                out_f64[o] = k ? sum / (double)s.valid : 0.0;
                ok[o] = k;
                break;
            }
            case WF_COUNT_ALL: out_u64[o] = s.total; ok[o] = enough; break;
            case WF_COUNT_VALID: out_u64[o] = s.valid; ok[o] = enough; break;
            default: out_u64[o] = s.total - s.valid; ok[o] = enough; break;
        }
    }
    return fine;
}

}  // namespace wf
}  // namespace b200

/*
 * cpu_engine.cpp -- morsel-parallel CPU restatement of the reference's execution of TPC-H Q1 and Q5.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY (bench.py `cpu_baseline` and `--impl reference`, tests/).  Nothing under
 * daft_b200/ links or calls this file.  It is "CPU restatement of the reference algorithm, N cores" -- never "Daft":
 * the reference engine (Rust) cannot be built in this image.
 *
 * What is restated, and from where (paths relative to /root/reference):
 *   - the Swordfish threading scheme: the source is cut into morsels of 131 072 rows
 *     (src/common/daft-config: default_morsel_size), one worker per host thread pulls morsels and runs the whole
 *     streaming part of the pipeline on each (intermediate ops and join probes are streaming operators,
 *     src/daft-local-execution/src/intermediate_ops/intermediate_op.rs, src/daft-local-execution/src/join/join_node.rs);
 *   - Filter = evaluate the predicate into a boolean mask, then copy every column through it
 *     (src/daft-recordbatch/src/lib.rs:520-570); Project = one full temporary column per expression node
 *     (lib.rs:1280-1400);
 *   - grouped aggregation per morsel = agg_groupby_inline (src/daft-recordbatch/src/ops/inline_agg.rs:748-928):
 *     XXH3 hash_rows over the key columns, hash-table probe to dense group ids in first-occurrence order, then one
 *     scatter pass per accumulator -- the functions of daft_oracle.c; partials are hash-partitioned into N buckets and
 *     every bucket is finalised on its own (grouped_aggregate.rs:270-488: agg-then-partition, per-bucket final
 *     aggregate + projection; populate_aggregation_stages, src/daft-local-plan/src/agg.rs:13-416: Mean -> Sum + Count);
 *   - hash join = ProbeTable (src/daft-recordbatch/src/probeable/probe_table.rs:56-185): hash_rows of the build keys,
 *     SEQUENTIAL insertion into a hash-keyed table of insertion-ordered row lists (one build task at a time per join,
 *     src/daft-local-execution/src/join/build.rs:99-124; the builds of different joins overlap), probes run per
 *     morsel on all workers: hash_rows of the probe keys, table lookup + value comparison, (probe row, build row)
 *     pairs, then `take` of every output column (inner_join.rs:11-85);
 *   - join order and build sides of Q5: benchmarking/tpch/answers.py:124-143 with the build on the smaller side
 *     (src/daft-local-execution/src/pipeline.rs:1106-1194).
 *
 * The tables come from the same counter-based generator as the device path (daft_b200/csrc/tpch_gen_core.h is a
 * header of pure functions; including it does not load the product library).
 */
#include <algorithm>
#include <malloc.h>
#include <sched.h>
#include <cstring>
#include <cstdio>
#include <climits>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../daft_b200/csrc/tpch_gen_core.h"

extern "C" {
/* daft_oracle.c (same shared object) */
typedef struct {
    int32_t dtype;
    int32_t _pad;
    int64_t length;
    const uint8_t *validity;
    const int64_t *offsets;
    const void *data;
} orc_col;
uint64_t orc_xxh3_64(const uint8_t *in, size_t len, uint64_t seed);
int orc_hash_column(const orc_col *c, const uint64_t *seed, uint64_t *out);
int orc_hash_rows(const orc_col *cols, int ncols, uint64_t *out);
int64_t orc_group_ids(const orc_col *keys, int nkeys, uint32_t *group_ids, uint64_t *first_rows, uint64_t *group_sizes);
void orc_sum_i64(const int64_t *v, const uint8_t *validity, const uint32_t *gid, int64_t n, int64_t ng, int64_t *out, uint8_t *out_valid);
void orc_sum_f64(const double *v, const uint8_t *validity, const uint32_t *gid, int64_t n, int64_t ng, double *out, uint8_t *out_valid);
void orc_count(const uint8_t *validity, int is_null_type, int mode, const uint32_t *gid, int64_t n, int64_t ng, uint64_t *out);
}

#define CPU_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

using namespace b200::tpchgen;
constexpr int64_t MORSEL = 131072;
enum { T_I64 = 5, T_F64 = 11, T_DATE32 = 12, T_UTF8 = 14 };

// Workers the process may actually run: the CPUs of its affinity mask, capped by the cgroup's CFS quota (a container that
// sees 128 CPUs but is allowed 12 of them spends most of 128 workers' time throttled).
int usable_threads() {
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = std::min(n > 0 ? n : CPU_COUNT(&set), CPU_COUNT(&set));
    if (n < 1) n = 1;
    long long quota = -1, period = 100000;
    if (FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r")) {  // cgroup v2: "<quota|max> <period>"
        char q[64] = {0};
        if (fscanf(f, "%63s %lld", q, &period) >= 1 && strcmp(q, "max") != 0) quota = atoll(q);
        fclose(f);
    } else if (FILE *g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {  // cgroup v1
        if (fscanf(g, "%lld", &quota) != 1) quota = -1;
        fclose(g);
        if (FILE *h = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) {
            if (fscanf(h, "%lld", &period) != 1) period = 100000;
            fclose(h);
        }
    }
    if (quota > 0 && period > 0) n = (int)std::min<long long>(n, std::max<long long>(1, (quota + period - 1) / period));
    return n;
}
int hw_threads(int threads) { return threads > 0 ? threads : usable_threads(); }

// Per-morsel temporaries are megabyte-sized vectors.  glibc serves those with mmap / munmap by default, so every morsel
// would fault its pages in afresh under the process-wide mm lock -- at 128 workers that lock, not the query, is what
// gets measured.  The reference runs on a caching allocator (jemalloc); keep freed blocks in the per-thread arenas.
void tune_allocator() {
    static std::once_flag once;
    std::call_once(once, [] {
        mallopt(M_MMAP_THRESHOLD, 1 << 30);
        mallopt(M_TRIM_THRESHOLD, INT_MAX);   // never hand arena memory back between morsels
        mallopt(M_TOP_PAD, 64 << 20);
    });
}

// morsel-parallel driver: `fn(worker, morsel_begin, morsel_end)`; every worker pulls the next morsel
void run_morsels(int64_t n, int threads, const std::function<void(int, int64_t, int64_t)> &fn) {
    tune_allocator();
    const int64_t nm = (n + MORSEL - 1) / MORSEL;
    std::atomic<int64_t> next{0};
    auto work = [&](int w) {
        for (;;) {
            const int64_t m = next.fetch_add(1);
            if (m >= nm) break;
            fn(w, m * MORSEL, std::min(n, (m + 1) * MORSEL));
        }
    };
    const int nt = (int)std::min<int64_t>(threads, std::max<int64_t>(nm, 1));
    std::vector<std::thread> th;
    for (int w = 1; w < nt; w++) th.emplace_back(work, w);
    work(0);
    for (auto &t : th) t.join();
}

// a LargeUtf8 column owned by this engine
struct StrCol {
    std::vector<int64_t> off{0};
    std::vector<uint8_t> bytes;
    int64_t size() const { return (int64_t)off.size() - 1; }
    void push(const uint8_t *p, int64_t len) {
        bytes.insert(bytes.end(), p, p + len);
        off.push_back((int64_t)bytes.size());
    }
};
orc_col str_col(const int64_t *off, const uint8_t *bytes, int64_t n) { return orc_col{T_UTF8, 0, n, nullptr, off, bytes}; }
orc_col i64_col(const int64_t *v, int64_t n) { return orc_col{T_I64, 0, n, nullptr, nullptr, v}; }

template <typename T>
void take(const T *src, const int64_t *idx, int64_t n, std::vector<T> &out) {
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; i++) out[(size_t)i] = src[idx[i]];
}
void take_str(const int64_t *off, const uint8_t *bytes, const int64_t *idx, int64_t n, StrCol &out) {
    out.off.assign(1, 0);
    out.bytes.clear();
    out.off.reserve((size_t)n + 1);
    for (int64_t i = 0; i < n; i++) {
        const int64_t r = idx[i];
        out.push(bytes + off[r], off[r + 1] - off[r]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// ProbeTable (probe_table.rs:56-185): keyed by the row hash, insertion-ordered row lists, value comparison at probe.
// Built by ONE thread (sequential inserts), probed read-only by all workers.
// ---------------------------------------------------------------------------------------------------------------
struct ProbeTable {
    std::vector<uint64_t> hash;   // slot -> hash of its key
    std::vector<int64_t> head;    // slot -> first build row of the key, -1 = empty
    std::vector<int64_t> next, tail;
    uint64_t mask = 0;
    int nkeys = 0;
    const int64_t *key[2] = {nullptr, nullptr};  // Int64 key columns of the build side (no nulls in TPC-H)
    int64_t rows = 0;

    void build(const int64_t *k0, const int64_t *k1, int64_t n) {
        nkeys = k1 ? 2 : 1;
        key[0] = k0;
        key[1] = k1;
        rows = n;
        std::vector<uint64_t> h((size_t)std::max<int64_t>(n, 1));
        orc_col cols[2] = {i64_col(k0, n), i64_col(k1 ? k1 : k0, n)};
        orc_hash_rows(cols, nkeys, h.data());
        uint64_t cap = 16;
        while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
        mask = cap - 1;
        hash.assign(cap, 0);
        head.assign(cap, -1);
        next.assign((size_t)std::max<int64_t>(n, 1), -1);
        tail.assign((size_t)std::max<int64_t>(n, 1), -1);
        for (int64_t i = 0; i < n; i++) {
            uint64_t s = h[(size_t)i] & mask;
            for (;;) {
                const int64_t hd = head[s];
                if (hd < 0) {
                    head[s] = i;
                    hash[s] = h[(size_t)i];
                    tail[(size_t)i] = i;
                    break;
                }
                if (hash[s] == h[(size_t)i] && k0[hd] == k0[i] && (!k1 || k1[hd] == k1[i])) {
                    next[(size_t)tail[(size_t)hd]] = i;
                    tail[(size_t)hd] = i;
                    break;
                }
                s = (s + 1) & mask;
            }
        }
    }
    // probe one morsel: pairs in probe-row order, matches in insertion order
    void probe(const int64_t *p0, const int64_t *p1, int64_t n, std::vector<int64_t> &pidx, std::vector<int64_t> &bidx) const {
        std::vector<uint64_t> h((size_t)std::max<int64_t>(n, 1));
        orc_col cols[2] = {i64_col(p0, n), i64_col(p1 ? p1 : p0, n)};
        orc_hash_rows(cols, nkeys, h.data());
        pidx.clear();
        bidx.clear();
        for (int64_t i = 0; i < n; i++) {
            uint64_t s = h[(size_t)i] & mask;
            for (;;) {
                const int64_t hd = head[s];
                if (hd < 0) break;
                if (hash[s] == h[(size_t)i] && key[0][hd] == p0[i] && (nkeys == 1 || key[1][hd] == p1[i])) {
                    for (int64_t r = hd; r >= 0; r = next[(size_t)r]) {
                        pidx.push_back(i);
                        bidx.push_back(r);
                    }
                    break;
                }
                s = (s + 1) & mask;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// a partial aggregate: string keys (one or two columns, joined with a 0x1f separator for the merge map) -> 8 values
// ---------------------------------------------------------------------------------------------------------------
struct Partial {
    std::string k0, k1;
    int64_t sum_qty = 0;
    double f[4] = {0, 0, 0, 0};  // Q1: price, disc_price, charge, disc;  Q5: revenue in f[0]
    uint64_t cnt[3] = {0, 0, 0};
    uint64_t hash = 0;
};

}  // namespace

// =====================================================================================================================
// host-side table generation (same values as the device generator)
// =====================================================================================================================
CPU_EXPORT int64_t cpu_tpch_num_rows(int table, double sf) {
    const Scale s = scale_of(sf);
    switch (table) {
        case 0:
            if (sf == 1.0) return 6001215;
            if (sf == 10.0) return 59986052;
            if (sf == 100.0) return 600037902;
            return std::max<int64_t>(1, (int64_t)(6001215.0 * sf + 0.5));
        case 1: return s.n_orders;
        case 2: return s.n_customers;
        case 3: return s.n_suppliers;
        case 4: return 25;
        case 5: return 5;
    }
    return -1;
}
// fills `data` (8 / 4 / 1 bytes per row by dtype) and, for the flag string columns, `offsets[n + 1]`;
// returns the dtype code (b200_dtype), -1 for an unknown column.  N_NAME / R_NAME: call cpu_tpch_name.
CPU_EXPORT int cpu_tpch_fill(const char *column, double sf, int64_t row_begin, int64_t row_end, uint64_t seed, void *data, int64_t *offsets,
                             int threads) {
    int col = -1;
    for (int c = 0; c < NUM_COLS; c++)
        if (strcasecmp(kCols[c].name, column) == 0) col = c;
    if (col < 0 || col == N_NAME || col == R_NAME) return -1;
    const int64_t n = row_end - row_begin;
    const Scale sc = scale_of(sf);
    if (col == N_REGIONKEY) {
        for (int64_t r = row_begin; r < row_end; r++) ((int64_t *)data)[r - row_begin] = kNationRegion[r];
        return kCols[col].dtype;
    }
    const int kind = kCols[col].dtype == T_DATE32 ? 1 : (kCols[col].dtype == T_UTF8 ? 2 : 0);
    const int nt = (int)std::max<int64_t>(1, std::min<int64_t>(hw_threads(threads), n / 65536 + 1));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
        const int64_t b = n * t / nt, e = n * (t + 1) / nt;
        th.emplace_back([=] {
            for (int64_t i = b; i < e; i++) gen_store(col, kind, seed, row_begin + i, i, sc, data, offsets);
        });
    }
    for (auto &x : th) x.join();
    if (kind == 2) offsets[n] = n;
    return kCols[col].dtype;
}
CPU_EXPORT const char *cpu_tpch_name(int is_region, int64_t row) { return is_region ? kRegions[row] : kNations[row]; }

// =====================================================================================================================
// Q1 (benchmarking/tpch/answers.py:12-32)
// out_*: up to `cap` groups, sorted by (L_RETURNFLAG, L_LINESTATUS); returns the number of groups
// =====================================================================================================================
CPU_EXPORT int cpu_usable_threads() { return usable_threads(); }

CPU_EXPORT int64_t cpu_q1(int64_t n, const int64_t *qty, const double *price, const double *disc, const double *tax, const int64_t *rf_off,
                          const uint8_t *rf, const int64_t *ls_off, const uint8_t *ls, const int32_t *ship, int32_t ship_cutoff, int threads,
                          int64_t cap, char *out_rf, char *out_ls, int64_t *sum_qty, double *sum_base, double *sum_disc_price, double *sum_charge,
                          double *avg_qty, double *avg_price, double *avg_disc, uint64_t *count_order, int64_t *tasks_out) {
    const int nt = hw_threads(threads);
    std::vector<std::vector<Partial>> per_worker((size_t)nt);  // the sink's per-worker state
    std::atomic<int64_t> tasks{0};
    run_morsels(n, nt, [&](int w, int64_t b, int64_t e) {
        tasks.fetch_add(1);
        const int64_t m = e - b;
        // ---- Filter: predicate -> mask -> every column copied through it (lib.rs:520-570)
        std::vector<uint8_t> mask((size_t)m);
        for (int64_t i = 0; i < m; i++) mask[(size_t)i] = ship[b + i] <= ship_cutoff;
        std::vector<int64_t> idx;
        idx.reserve((size_t)m);
        for (int64_t i = 0; i < m; i++)
            if (mask[(size_t)i]) idx.push_back(b + i);
        const int64_t k = (int64_t)idx.size();
        std::vector<int64_t> f_qty;
        std::vector<double> f_price, f_disc, f_tax;
        std::vector<int32_t> f_ship;
        StrCol f_rf, f_ls;
        take(qty, idx.data(), k, f_qty);
        take(price, idx.data(), k, f_price);
        take(disc, idx.data(), k, f_disc);
        take(tax, idx.data(), k, f_tax);
        take(ship, idx.data(), k, f_ship);
        take_str(rf_off, rf, idx.data(), k, f_rf);
        take_str(ls_off, ls, idx.data(), k, f_ls);
        // ---- Project: one temporary per expression node
        std::vector<double> t1((size_t)k), dp((size_t)k), t3((size_t)k), ch((size_t)k);
        for (int64_t i = 0; i < k; i++) t1[(size_t)i] = 1.0 - f_disc[(size_t)i];
        for (int64_t i = 0; i < k; i++) dp[(size_t)i] = f_price[(size_t)i] * t1[(size_t)i];
        for (int64_t i = 0; i < k; i++) t3[(size_t)i] = 1.0 + f_tax[(size_t)i];
        for (int64_t i = 0; i < k; i++) ch[(size_t)i] = dp[(size_t)i] * t3[(size_t)i];
        // ---- partial aggregate (agg_groupby_inline): hash_rows -> group ids -> one scatter pass per accumulator
        if (k == 0) return;
        orc_col keys[2] = {str_col(f_rf.off.data(), f_rf.bytes.data(), k), str_col(f_ls.off.data(), f_ls.bytes.data(), k)};
        std::vector<uint32_t> gid((size_t)k);
        std::vector<uint64_t> first((size_t)k), sizes((size_t)k);
        const int64_t ng = orc_group_ids(keys, 2, gid.data(), first.data(), sizes.data());
        std::vector<int64_t> s_qty((size_t)ng);
        std::vector<double> s_price((size_t)ng), s_dp((size_t)ng), s_ch((size_t)ng), s_disc((size_t)ng);
        std::vector<uint64_t> c_qty((size_t)ng), c_price((size_t)ng), c_disc((size_t)ng);
        std::vector<uint8_t> valid((size_t)ng);
        orc_sum_i64(f_qty.data(), nullptr, gid.data(), k, ng, s_qty.data(), valid.data());
        orc_sum_f64(f_price.data(), nullptr, gid.data(), k, ng, s_price.data(), valid.data());
        orc_sum_f64(dp.data(), nullptr, gid.data(), k, ng, s_dp.data(), valid.data());
        orc_sum_f64(ch.data(), nullptr, gid.data(), k, ng, s_ch.data(), valid.data());
        orc_sum_f64(f_disc.data(), nullptr, gid.data(), k, ng, s_disc.data(), valid.data());
        orc_count(nullptr, 0, 1, gid.data(), k, ng, c_qty.data());
        orc_count(nullptr, 0, 1, gid.data(), k, ng, c_price.data());
        orc_count(nullptr, 0, 1, gid.data(), k, ng, c_disc.data());
        // the partial's rows (keys taken at the groups' first rows), hash-partitioned for the per-bucket finalize
        std::vector<uint64_t> kh((size_t)k);
        orc_hash_rows(keys, 2, kh.data());
        for (int64_t g = 0; g < ng; g++) {
            const int64_t r = (int64_t)first[(size_t)g];
            Partial p;
            p.k0.assign((const char *)f_rf.bytes.data() + f_rf.off[(size_t)r], (size_t)(f_rf.off[(size_t)r + 1] - f_rf.off[(size_t)r]));
            p.k1.assign((const char *)f_ls.bytes.data() + f_ls.off[(size_t)r], (size_t)(f_ls.off[(size_t)r + 1] - f_ls.off[(size_t)r]));
            p.sum_qty = s_qty[(size_t)g];
            p.f[0] = s_price[(size_t)g];
            p.f[1] = s_dp[(size_t)g];
            p.f[2] = s_ch[(size_t)g];
            p.f[3] = s_disc[(size_t)g];
            p.cnt[0] = c_qty[(size_t)g];
            p.cnt[1] = c_price[(size_t)g];
            p.cnt[2] = c_disc[(size_t)g];
            p.hash = kh[(size_t)r];
            per_worker[(size_t)w].push_back(std::move(p));
        }
    });
    // ---- finalize: bucket b = partial rows with hash % N == b, final aggregate + projection per bucket, in parallel
    std::vector<std::vector<Partial>> finals((size_t)nt);
    {
        std::vector<std::thread> th;
        for (int b = 0; b < nt; b++)
            th.emplace_back([&, b] {
                std::vector<Partial> &out = finals[(size_t)b];
                for (auto &wv : per_worker)
                    for (auto &p : wv) {
                        if ((int)(p.hash % (uint64_t)nt) != b) continue;
                        Partial *hit = nullptr;
                        for (auto &q : out)
                            if (q.hash == p.hash && q.k0 == p.k0 && q.k1 == p.k1) hit = &q;
                        if (!hit) {
                            out.push_back(p);
                            continue;
                        }
                        hit->sum_qty += p.sum_qty;
                        for (int j = 0; j < 4; j++) hit->f[j] += p.f[j];
                        for (int j = 0; j < 3; j++) hit->cnt[j] += p.cnt[j];
                    }
            });
        for (auto &t : th) t.join();
    }
    std::vector<Partial> all;
    for (auto &v : finals)
        for (auto &p : v) all.push_back(p);
    std::sort(all.begin(), all.end(), [](const Partial &a, const Partial &b) { return a.k0 != b.k0 ? a.k0 < b.k0 : a.k1 < b.k1; });
    const int64_t ng = std::min<int64_t>((int64_t)all.size(), cap);
    for (int64_t g = 0; g < ng; g++) {
        const Partial &p = all[(size_t)g];
        out_rf[g] = p.k0.empty() ? 0 : p.k0[0];
        out_ls[g] = p.k1.empty() ? 0 : p.k1[0];
        sum_qty[g] = p.sum_qty;
        sum_base[g] = p.f[0];
        sum_disc_price[g] = p.f[1];
        sum_charge[g] = p.f[2];
        avg_qty[g] = (double)p.sum_qty / (double)p.cnt[0];  // merge_mean = sum / count (functions/agg/mod.rs:29-53)
        avg_price[g] = p.f[0] / (double)p.cnt[1];
        avg_disc[g] = p.f[3] / (double)p.cnt[2];
        count_order[g] = p.cnt[0];
    }
    if (tasks_out) *tasks_out = tasks.load();
    return (int64_t)all.size();
}

// =====================================================================================================================
// Q5 (answers.py:124-143).  Tables are passed as plain column arrays; nation / region come from the generator.
// out_name: 32-byte slots; returns the number of groups, sorted by revenue descending.
// =====================================================================================================================
CPU_EXPORT int64_t cpu_q5(int64_t n_li, const int64_t *l_orderkey, const int64_t *l_suppkey, const double *l_price, const double *l_disc,
                          int64_t n_ord, const int64_t *o_orderkey, const int64_t *o_custkey, const int32_t *o_date, int64_t n_cust,
                          const int64_t *c_custkey, const int64_t *c_nationkey, int64_t n_supp, const int64_t *s_suppkey,
                          const int64_t *s_nationkey, int32_t date_lo, int32_t date_hi, int threads, int64_t cap, char *out_name,
                          double *out_revenue, int64_t *tasks_out, int64_t *probe_rows_out) {
    const int nt = hw_threads(threads);
    std::atomic<int64_t> tasks{0}, probe_rows{0};
    // ---- region (R_NAME == 'ASIA') x nation: tiny
    std::vector<int64_t> asia_nation;  // N_NATIONKEY of the nations in ASIA
    for (int r = 0; r < 5; r++)
        if (strcmp(kRegions[r], "ASIA") == 0)
            for (int nk = 0; nk < 25; nk++)
                if (kNationRegion[nk] == r) asia_nation.push_back(nk);
    // ---- x supplier: build on the (5-row) nation side, probe the supplier morsels -> (N_NAME, S_SUPPKEY, N_NATIONKEY)
    ProbeTable nat_tbl;
    nat_tbl.build(asia_nation.data(), nullptr, (int64_t)asia_nation.size());
    std::vector<std::vector<int64_t>> w_supp((size_t)nt), w_nat((size_t)nt);
    run_morsels(n_supp, nt, [&](int w, int64_t b, int64_t e) {
        tasks.fetch_add(1);
        std::vector<int64_t> pi, bi;
        nat_tbl.probe(s_nationkey + b, nullptr, e - b, pi, bi);
        for (size_t i = 0; i < pi.size(); i++) {
            w_supp[(size_t)w].push_back(s_suppkey[b + pi[i]]);
            w_nat[(size_t)w].push_back(asia_nation[(size_t)bi[i]]);
        }
    });
    std::vector<int64_t> a_supp, a_nat;
    for (int w = 0; w < nt; w++) {
        a_supp.insert(a_supp.end(), w_supp[(size_t)w].begin(), w_supp[(size_t)w].end());
        a_nat.insert(a_nat.end(), w_nat[(size_t)w].begin(), w_nat[(size_t)w].end());
    }
    // ---- orders filter (morsel-parallel: mask, then both kept columns copied through it)
    std::vector<std::vector<int64_t>> w_ok((size_t)nt), w_oc((size_t)nt);
    run_morsels(n_ord, nt, [&](int w, int64_t b, int64_t e) {
        tasks.fetch_add(1);
        const int64_t m = e - b;
        std::vector<uint8_t> m1((size_t)m), m2((size_t)m), mk((size_t)m);
        for (int64_t i = 0; i < m; i++) m1[(size_t)i] = o_date[b + i] >= date_lo;
        for (int64_t i = 0; i < m; i++) m2[(size_t)i] = o_date[b + i] < date_hi;
        for (int64_t i = 0; i < m; i++) mk[(size_t)i] = m1[(size_t)i] & m2[(size_t)i];
        for (int64_t i = 0; i < m; i++)
            if (mk[(size_t)i]) {
                w_ok[(size_t)w].push_back(o_orderkey[b + i]);
                w_oc[(size_t)w].push_back(o_custkey[b + i]);
            }
    });
    std::vector<int64_t> f_ok, f_oc;
    for (int w = 0; w < nt; w++) {
        f_ok.insert(f_ok.end(), w_ok[(size_t)w].begin(), w_ok[(size_t)w].end());
        f_oc.insert(f_oc.end(), w_oc[(size_t)w].begin(), w_oc[(size_t)w].end());
    }
    // ---- the three probe tables: each built by ONE task (build.rs:99-124), the three builds overlap
    ProbeTable t_supp, t_ord, t_cust;
    {
        std::thread b1([&] { t_supp.build(a_supp.data(), nullptr, (int64_t)a_supp.size()); });
        std::thread b2([&] { t_ord.build(f_ok.data(), nullptr, (int64_t)f_ok.size()); });
        std::thread b3([&] { t_cust.build(c_custkey, c_nationkey, n_cust); });
        b1.join();
        b2.join();
        b3.join();
    }
    // ---- lineitem morsels through the three probes, the projection and the partial aggregate
    std::vector<std::vector<Partial>> per_worker((size_t)nt);
    run_morsels(n_li, nt, [&](int w, int64_t b, int64_t e) {
        tasks.fetch_add(1);
        const int64_t m = e - b;
        probe_rows.fetch_add(m);
        std::vector<int64_t> pi, bi;
        // join 3: build = ASIA suppliers, probe = lineitem; output select: N_NAME, price, disc, L_ORDERKEY, N_NATIONKEY
        t_supp.probe(l_suppkey + b, nullptr, m, pi, bi);
        int64_t k = (int64_t)pi.size();
        std::vector<int64_t> j_ok((size_t)k), j_nat((size_t)k);
        std::vector<double> j_price((size_t)k), j_disc((size_t)k);
        StrCol j_name;
        for (int64_t i = 0; i < k; i++) {
            const int64_t pr = b + pi[(size_t)i], br = bi[(size_t)i];
            j_ok[(size_t)i] = l_orderkey[pr];
            j_price[(size_t)i] = l_price[pr];
            j_disc[(size_t)i] = l_disc[pr];
            j_nat[(size_t)i] = a_nat[(size_t)br];
        }
        for (int64_t i = 0; i < k; i++) {
            const char *nm = kNations[a_nat[(size_t)bi[(size_t)i]]];
            j_name.push((const uint8_t *)nm, (int64_t)strlen(nm));
        }
        // join 4: build = orders of 1994, probe = the running intermediate on L_ORDERKEY
        probe_rows.fetch_add(k);
        t_ord.probe(j_ok.data(), nullptr, k, pi, bi);
        const int64_t k2 = (int64_t)pi.size();
        std::vector<int64_t> q_cust((size_t)k2), q_nat;
        std::vector<double> q_price, q_disc;
        StrCol q_name;
        take(j_nat.data(), pi.data(), k2, q_nat);
        take(j_price.data(), pi.data(), k2, q_price);
        take(j_disc.data(), pi.data(), k2, q_disc);
        take_str(j_name.off.data(), j_name.bytes.data(), pi.data(), k2, q_name);
        for (int64_t i = 0; i < k2; i++) q_cust[(size_t)i] = f_oc[(size_t)bi[(size_t)i]];
        // join 5: build = customer on (C_CUSTKEY, C_NATIONKEY), probe on (O_CUSTKEY, N_NATIONKEY)
        probe_rows.fetch_add(k2);
        t_cust.probe(q_cust.data(), q_nat.data(), k2, pi, bi);
        const int64_t k3 = (int64_t)pi.size();
        std::vector<double> r_price, r_disc;
        StrCol r_name;
        take(q_price.data(), pi.data(), k3, r_price);
        take(q_disc.data(), pi.data(), k3, r_disc);
        take_str(q_name.off.data(), q_name.bytes.data(), pi.data(), k3, r_name);
        if (k3 == 0) return;
        // select N_NAME, price * (1 - disc) as value; partial aggregate by N_NAME
        std::vector<double> t1((size_t)k3), val((size_t)k3);
        for (int64_t i = 0; i < k3; i++) t1[(size_t)i] = 1.0 - r_disc[(size_t)i];
        for (int64_t i = 0; i < k3; i++) val[(size_t)i] = r_price[(size_t)i] * t1[(size_t)i];
        orc_col key = str_col(r_name.off.data(), r_name.bytes.data(), k3);
        std::vector<uint32_t> gid((size_t)k3);
        std::vector<uint64_t> first((size_t)k3), sizes((size_t)k3), kh((size_t)k3);
        const int64_t ng = orc_group_ids(&key, 1, gid.data(), first.data(), sizes.data());
        std::vector<double> s((size_t)ng);
        std::vector<uint8_t> valid((size_t)ng);
        orc_sum_f64(val.data(), nullptr, gid.data(), k3, ng, s.data(), valid.data());
        orc_hash_rows(&key, 1, kh.data());
        for (int64_t g = 0; g < ng; g++) {
            const int64_t r = (int64_t)first[(size_t)g];
            Partial p;
            p.k0.assign((const char *)r_name.bytes.data() + r_name.off[(size_t)r], (size_t)(r_name.off[(size_t)r + 1] - r_name.off[(size_t)r]));
            p.f[0] = s[(size_t)g];
            p.hash = kh[(size_t)r];
            per_worker[(size_t)w].push_back(std::move(p));
        }
    });
    // ---- finalize (5 groups: the per-bucket scheme as in Q1, run inline)
    std::vector<Partial> all;
    for (auto &wv : per_worker)
        for (auto &p : wv) {
            Partial *hit = nullptr;
            for (auto &q : all)
                if (q.hash == p.hash && q.k0 == p.k0) hit = &q;
            if (!hit) all.push_back(p);
            else hit->f[0] += p.f[0];
        }
    std::sort(all.begin(), all.end(), [](const Partial &a, const Partial &b) { return a.f[0] > b.f[0]; });
    const int64_t ng = std::min<int64_t>((int64_t)all.size(), cap);
    for (int64_t g = 0; g < ng; g++) {
        memset(out_name + g * 32, 0, 32);
        memcpy(out_name + g * 32, all[(size_t)g].k0.data(), std::min<size_t>(31, all[(size_t)g].k0.size()));
        out_revenue[g] = all[(size_t)g].f[0];
    }
    if (tasks_out) *tasks_out = tasks.load();
    if (probe_rows_out) *probe_rows_out = probe_rows.load();
    return (int64_t)all.size();
}

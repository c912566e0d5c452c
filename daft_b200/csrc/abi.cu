// extern "C" boundary of libdaftb200.so (declarations + reference citations: include/daft_b200.h).
// Every entry point converts C++ exceptions into the daft-ext error convention
// (int status + malloc'd message, src/daft-ext/src/abi/mod.rs:84-107).
#include "common.h"
#include "jit.h"
#include "ops.h"

using namespace b200;

namespace {
int set_err(char **errmsg, const std::string &msg) {
    if (errmsg) {
        char *m = static_cast<char *>(malloc(msg.size() + 1));
        if (m) memcpy(m, msg.c_str(), msg.size() + 1);
        *errmsg = m;
    }
    return 1;
}
template <class F>
int guard(char **errmsg, F &&f) {
    try {
        // Entry points may be called from any host thread: a new thread's current device is 0, so every call selects the
        // library's device first (launches and stream-ordered allocations on ctx().stream need it current).
        if (ctx().ready) cudaSetDevice(ctx().device);
        f();
        return 0;
    } catch (const std::exception &e) {
        return set_err(errmsg, e.what());
    } catch (...) {
        return set_err(errmsg, "unknown error");
    }
}
const Column &col_at(const Batch *b, int32_t i) {
    if (!b) fail("null batch");
    if (i < 0 || i >= (int32_t)b->cols.size()) fail("column index out of range");
    return b->cols[i];
}
// Operators work on plain columns: views (join outputs) are gathered here, on first use, and only the columns asked for.
Batch plain(const b200_batch *b, const std::vector<int32_t> *cols = nullptr) {
    if (!b) fail("null batch");
    return materialize_batch(*b, cols);
}
// operators that move whole rows between buffers (partition, exchange, filter compaction) handle flat columns only
Batch flat(const b200_batch *b, const char *what) {
    Batch p = plain(b);
    for (const Column &c : p.cols)
        if (c.dtype == B200_LIST) fail(std::string(what) + ": list columns are not supported (column '" + c.name + "')");
    return p;
}
Batch plain_keys(const b200_batch *b, const int32_t *key_cols, int32_t nkeys) {
    if (!b) fail("null batch");
    std::vector<int32_t> k(key_cols, key_cols + (nkeys > 0 ? nkeys : 0));
    return materialize_batch(*b, &k);
}
Batch plain_for_plan(const b200_batch *b, int op, const uint8_t *plan, size_t plan_len) {
    if (!b) fail("null batch");
    const std::vector<int32_t> cols = jit::referenced_columns(op, plan, plan_len);
    for (int32_t c : cols)
        if (c < 0 || c >= (int32_t)b->cols.size()) fail("plan: column index out of range");
    return materialize_batch(*b, &cols);
}
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int b200_abi_version(void) { return DAFT_B200_ABI_VERSION; }

int b200_init(int device_ordinal, char **errmsg) {
    return guard(errmsg, [&] {
        Context &c = ctx();
        if (c.ready && c.device == device_ordinal) return;
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            fail(std::string("daft_b200: no CUDA device available (") + cudaGetErrorString(e) + "); there is no CPU fallback");
        if (device_ordinal < 0 || device_ordinal >= ndev) fail("daft_b200: device ordinal out of range");
        B200_CUDA(cudaSetDevice(device_ordinal));
        cudaDeviceProp prop;
        B200_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
        if (prop.major < 10) fail("daft_b200: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
        c.sm_count = prop.multiProcessorCount;
        if (!c.stream) B200_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        cudaMemPool_t pool;
        B200_CUDA(cudaDeviceGetDefaultMemPool(&pool, device_ordinal));
        uint64_t thr = UINT64_MAX;
        B200_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        c.device = device_ordinal;
        c.ready = true;
    });
}

void b200_free_string(char *s) { free(s); }
void *b200_stream(void) { return ctx().stream; }
int b200_synchronize(char **errmsg) {
    return guard(errmsg, [&] {
        require_device();
        B200_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}
int64_t b200_launch_count(void) { return g_launches.load(); }

// device time of the dominant kernel (the fused aggregate) measured with CUDA events on the launching stream
int b200_kernel_timing(int enable) {
    jit::enable_kernel_timing(enable != 0);
    kernel_timing_enable(enable != 0);
    return 0;
}
int b200_kernel_timing_report(char **out_json, char **errmsg) {
    return guard(errmsg, [&] {
        const std::string t = kernel_timing_report();
        char *m = static_cast<char *>(malloc(t.size() + 1));
        if (!m) fail("out of memory");
        memcpy(m, t.c_str(), t.size() + 1);
        *out_json = m;
    });
}
double b200_last_agg_kernel_ms(void) {
    try {
        return jit::last_timed_kernel_ms();
    } catch (...) {
        return -1.0;
    }
}

int b200_set_jit_cache_dir(const char *path, char **errmsg) {
    return guard(errmsg, [&] { jit::set_cache_dir(path ? path : ""); });
}
int b200_precompile(int32_t op, const uint8_t *plan, size_t plan_len, const int32_t *col_dtypes, const uint8_t *col_nullable,
                    int32_t ncols, char **errmsg) {
    return guard(errmsg, [&] {
        std::vector<int32_t> dt(col_dtypes, col_dtypes + ncols);
        std::vector<uint8_t> nl(col_nullable, col_nullable + ncols);
        jit::precompile(op, plan, plan_len, dt, nl);
    });
}

int b200_jit_describe(int32_t op, const uint8_t *plan, size_t plan_len, const int32_t *col_dtypes, const uint8_t *col_nullable,
                      int32_t ncols, char **out_text, char **errmsg) {
    return guard(errmsg, [&] {
        std::vector<int32_t> dt(col_dtypes, col_dtypes + ncols);
        std::vector<uint8_t> nl(col_nullable, col_nullable + ncols);
        jit::Compiled meta;
        const int variant = (op >> 8) & 0xff;  // aggregate build: 0 direct loads, 1 global table, 2 TMA-staged (jit.h)
        const bool with_take = ((op >> 16) & 1) != 0;  // projection evaluated at gathered rows (negative index -> null row)
        op &= 0xff;
        const std::string src = jit::generate(op, plan, plan_len, dt, nl, with_take, meta, variant);
        std::string t = "ptrs=";
        for (size_t i = 0; i < meta.ptrs.size(); i++)
            t += (i ? "," : "") + std::to_string(meta.ptrs[i].kind) + ":" + std::to_string(meta.ptrs[i].index);
        t += "\nints=";
        for (size_t i = 0; i < meta.ints.size(); i++) t += (i ? "," : "") + std::to_string(meta.ints[i]);
        t += "\nouts=";
        for (size_t i = 0; i < meta.outs.size(); i++)
            t += (i ? "," : "") + std::to_string(meta.outs[i].dtype) + ":" + (meta.outs[i].nullable ? "1" : "0");
        if (op == B200_OP_AGG) {  // what the host needs to read the group table back (ops_jit.cu op_agg_first_rows)
            t += "\nnkeys=" + std::to_string(meta.nkeys) + "\nnlane=" + std::to_string(meta.NLANE) + "\nnwarp=" + std::to_string(meta.NWARP) + "\naggs=";
            for (size_t i = 0; i < meta.aggs.size(); i++) {
                const jit::AggOut &a = meta.aggs[i];
                t += (i ? "," : "") + std::to_string(a.kind) + ":" + std::to_string(a.count_mode) + ":" + std::to_string(a.val_dtype) + ":" +
                     std::to_string(a.out_dtype) + ":" + std::to_string(a.lane_cell) + ":" + std::to_string(a.cnt_cell) + ":" +
                     (a.child_nullable ? "1" : "0") + ":" + (a.const_zero ? "1" : "0") + ":" + a.name;
            }
            t += "\nnancells=";  // per aggregate: nan_cell:last_cell (float min / max: the NaN-forgetting fold, jit.h AggOut)
            for (size_t i = 0; i < meta.aggs.size(); i++)
                t += (i ? "," : "") + std::to_string(meta.aggs[i].nan_cell) + ":" + std::to_string(meta.aggs[i].last_cell);
            t += "\nfirstvalid=";  // per aggregate: any_value(ignore_nulls) over a nullable child keeps the first valid row here
            for (size_t i = 0; i < meta.aggs.size(); i++) t += (i ? "," : "") + std::to_string(meta.aggs[i].first_valid_cell);
            t += "\ncellops=";
            for (size_t i = 0; i < meta.lane_cell_ops.size(); i++) t += (i ? "," : "") + std::to_string(meta.lane_cell_ops[i]);
            static const char *hex = "0123456789abcdef";
            t += "\nkeyplan=";
            for (uint8_t b : meta.key_project_plan) { t += hex[b >> 4]; t += hex[b & 15]; }
        }
        t += "\n---\n" + src;
        char *m = static_cast<char *>(malloc(t.size() + 1));
        if (!m) fail("out of memory");
        memcpy(m, t.c_str(), t.size() + 1);
        *out_text = m;
    });
}

// ---- batches -------------------------------------------------------------------------------------
int b200_batch_import(struct ArrowArray *array, struct ArrowSchema *schema, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = import_batch(array, schema); });
}
int b200_batch_export(const b200_batch *batch, struct ArrowArray *out_array, struct ArrowSchema *out_schema, char **errmsg) {
    return guard(errmsg, [&] {
        export_batch(plain(batch), out_array, out_schema);
    });
}
int b200_batch_from_device(const b200_column_desc *cols, int32_t ncols, int64_t num_rows, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = batch_from_device(cols, ncols, num_rows); });
}
int b200_batch_describe(const b200_batch *batch, int32_t col, b200_column_desc *out, char **errmsg) {
    return guard(errmsg, [&] {
        const Column &c = col_at(batch, col);
        out->dtype = c.dtype;
        out->_reserved = 0;
        out->length = c.length;
        out->name = c.name.c_str();
        if (c.is_view()) {  // a join output that nobody has read yet: schema only (b200_batch_materialize gives buffers)
            out->null_count = -1;
            out->validity = out->offsets = out->data = nullptr;
            out->data_bytes = 0;
            return;
        }
        out->null_count = c.null_count;
        out->validity = c.validity ? c.validity->ptr : nullptr;
        out->offsets = c.offsets ? c.offsets->ptr : nullptr;
        out->data = c.data ? c.data->ptr : nullptr;
        out->data_bytes = c.data_bytes;
    });
}
int b200_batch_materialize(const b200_batch *batch, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        require_device();
        *out = new Batch(plain(batch));
        op_done();
    });
}
int b200_batch_dtypes(const b200_batch *batch, int32_t *out_dtypes, int32_t n, char **errmsg) {
    return guard(errmsg, [&] {
        if (!batch || !out_dtypes) fail("null argument");
        if (n < 0 || n > (int32_t)batch->cols.size()) fail("dtypes: column count out of range");
        for (int32_t i = 0; i < n; i++) {
            const Column &c = batch->cols[i];
            out_dtypes[i] = c.dtype == B200_LIST && c.child ? (B200_LIST | (c.child->dtype << 8)) : c.dtype;
        }
    });
}
int64_t b200_batch_num_rows(const b200_batch *batch) { return batch ? batch->num_rows : -1; }
int32_t b200_batch_num_columns(const b200_batch *batch) { return batch ? (int32_t)batch->cols.size() : -1; }
int64_t b200_batch_size_bytes(const b200_batch *batch) {
    if (!batch) return -1;
    int64_t s = 0;
    for (auto &c0 : batch->cols) {
        const Column &c = c0.is_view() ? *c0.base : c0;  // a view is sized by what it selects from
        s += c.data_bytes;
        if (c.offsets) s += (c.length + 1) * 8;
        if (c.validity) s += (c.length + 7) / 8;
    }
    return s;
}
void b200_batch_free(b200_batch *batch) { delete batch; }

int b200_batch_select(const b200_batch *batch, const int32_t *cols, int32_t ncols, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        auto o = std::make_unique<Batch>();
        o->num_rows = batch->num_rows;
        for (int i = 0; i < ncols; i++) o->cols.push_back(col_at(batch, cols[i]));
        *out = o.release();
    });
}
int b200_batch_rename(b200_batch *batch, int32_t col, const char *name, char **errmsg) {
    return guard(errmsg, [&] {
        col_at(batch, col);
        batch->cols[col].name = name ? name : "";
    });
}
int b200_batch_union(const b200_batch *a, const b200_batch *b, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!a || !b) fail("null batch");
        if (a->num_rows != b->num_rows && !a->cols.empty() && !b->cols.empty()) fail("union: row counts differ");
        auto o = std::make_unique<Batch>(*a);
        if (a->cols.empty()) o->num_rows = b->num_rows;
        for (auto &c : b->cols) o->cols.push_back(c);
        *out = o.release();
    });
}
int b200_batch_slice(const b200_batch *batch, int64_t start, int64_t end, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!batch) fail("null batch");
        if (start < 0 || end < start || end > batch->num_rows) fail("slice: bounds out of range");
        auto o = std::make_unique<Batch>();
        o->num_rows = end - start;
        for (auto &c : batch->cols) o->cols.push_back(slice_column(c, start, end));
        *out = o.release();
    });
}

// ---- hashing / partition -----------------------------------------------------------------------------
int b200_hash_rows(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const b200_batch *seed, b200_batch **out,
                   char **errmsg) {
    return guard(errmsg, [&] {
        require_device();
        const Batch in = plain_keys(batch, key_cols, nkeys);
        const uint64_t *sd = nullptr;
        if (seed) {
            const Column s = materialize(col_at(seed, 0));
            if (s.dtype != B200_U64) fail(std::string("seed data type expected to be uint64, got ") + dtype_name(s.dtype));
            if (s.length != batch->num_rows) fail("seed length does not match array length");
            if (s.null_count) fail("seed must not contain nulls");
            sd = as<uint64_t>(s.data);
        }
        Buf h = hash_rows(in, key_cols, nkeys, sd);
        auto o = std::make_unique<Batch>();
        o->num_rows = batch->num_rows;
        o->cols.push_back(make_u64_column("hash", h, batch->num_rows));
        *out = o.release();
    });
}
int b200_partition_by_hash(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, int32_t num_partitions,
                           b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        std::vector<Batch *> parts = op_partition_by_hash(flat(batch, "partition_by_hash"), key_cols, nkeys, num_partitions);
        for (int i = 0; i < num_partitions; i++) out[i] = parts[i];
    });
}

int b200_partition_by_random(const b200_batch *batch, int32_t num_partitions, uint64_t seed, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        std::vector<Batch *> parts = op_partition_by_random(flat(batch, "partition_by_random"), num_partitions, seed);
        for (int i = 0; i < num_partitions; i++) out[i] = parts[i];
    });
}
int b200_search_sorted(const b200_batch *sorted, const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending,
                       b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!sorted || !batch) fail("null batch");
        Buf pos = op_search_sorted(plain(sorted), plain_keys(batch, key_cols, nkeys), key_cols, nkeys, descending);
        auto o = std::make_unique<Batch>();
        o->num_rows = batch->num_rows;
        o->cols.push_back(make_u64_column("idx", pos, batch->num_rows));
        *out = o.release();
    });
}
int b200_partition_by_range(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const b200_batch *boundaries,
                            const uint8_t *descending, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!batch || !boundaries) fail("null batch");
        std::vector<Batch *> parts = op_partition_by_range(flat(batch, "partition_by_range"), key_cols, nkeys, plain(boundaries), descending);
        for (size_t i = 0; i < parts.size(); i++) out[i] = parts[i];
    });
}

// ---- filter / project / agg -----------------------------------------------------------------------
int b200_filter(const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = op_filter(flat(batch, "filter"), plan, plan_len); });
}
int b200_eval_expression_list(const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = op_project(plain_for_plan(batch, B200_OP_PROJECT, plan, plan_len), plan, plan_len, nullptr, 0); });
}
int b200_agg(const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = op_agg(plain_for_plan(batch, B200_OP_AGG, plan, plan_len), plan, plan_len); });
}
int b200_groupby_create(const uint8_t *partial_plan, size_t partial_len, const uint8_t *final_plan, size_t final_len,
                        const uint8_t *project_plan, size_t project_len, b200_groupby_state **out, char **errmsg) {
    return guard(errmsg, [&] { *out = groupby_create(partial_plan, partial_len, final_plan, final_len, project_plan, project_len); });
}
int b200_groupby_push(b200_groupby_state *st, const b200_batch *batch, char **errmsg) {
    return guard(errmsg, [&] {
        if (!batch) fail("null batch");
        groupby_push(st, *batch);
    });
}
int b200_groupby_finalize(b200_groupby_state *st, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = groupby_finalize(st); });
}
void b200_groupby_free(b200_groupby_state *st) { groupby_free(st); }
int b200_dedup(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = op_dedup(plain(batch), key_cols, nkeys); });
}

// ---- take / concat ----------------------------------------------------------------------------------
int b200_take(const b200_batch *batch, const b200_batch *indices, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        require_device();
        const Column ic = materialize(col_at(indices, 0));
        Buf idx = indices_to_i64(ic);
        check_take_indices(as<int64_t>(idx), ic.length, batch->num_rows, ic.null_count > 0);
        *out = gather_batch(plain(batch), as<int64_t>(idx), ic.length, ic.null_count > 0);
        op_done();
    });
}
// ---- joins over sorted inputs, cross join ---------------------------------------------------------
int b200_merge_join_sorted(const b200_batch *left, const int32_t *left_keys, const b200_batch *right, const int32_t *right_keys, int32_t nkeys,
                           b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!left || !right) fail("null batch");
        // only the key columns are read here: payload columns stay views and are gathered once, through the output's selection vectors
        const Batch l = plain_keys(left, left_keys, nkeys), r = plain_keys(right, right_keys, nkeys);
        *out = op_merge_join_sorted(l, left_keys, r, right_keys, nkeys);
    });
}
int b200_asof_join_sorted(const b200_batch *left, const int32_t *left_keys, const b200_batch *right, const int32_t *right_keys, int32_t nkeys,
                          b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!left || !right) fail("null batch");
        const Batch l = plain_keys(left, left_keys, nkeys), r = plain_keys(right, right_keys, nkeys);
        *out = op_asof_join_sorted(l, left_keys, r, right_keys, nkeys);
    });
}
int b200_cross_join(const b200_batch *left, const b200_batch *right, int32_t outer_loop_left, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!left || !right) fail("null batch");
        *out = op_cross_join(*left, *right, outer_loop_left != 0);
    });
}
// ---- window functions over sorted partitions -------------------------------------------------------
int b200_window_ranks(const b200_batch *sorted, const int32_t *partition_cols, int32_t npartition, const int32_t *order_cols, int32_t norder,
                      b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!sorted) fail("null batch");
        std::vector<int32_t> k(partition_cols, partition_cols + (npartition > 0 ? npartition : 0));
        k.insert(k.end(), order_cols, order_cols + (norder > 0 ? norder : 0));
        for (int32_t c : k)
            if (c < 0 || c >= (int32_t)sorted->cols.size()) fail("window: key column index out of range");
        *out = op_window_ranks(materialize_batch(*sorted, &k), partition_cols, npartition, order_cols, norder);
    });
}
int b200_window_offset(const b200_batch *values, const b200_batch *partition_ids, int64_t offset, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!values || !partition_ids) fail("null batch");
        *out = op_window_offset(*values, materialize(col_at(partition_ids, 0)), offset);
    });
}
int b200_window_frame(const b200_batch *values, const b200_batch *partition_begin, const b200_batch *order, int32_t agg, int32_t start_kind,
                      int64_t start, int32_t end_kind, int64_t end, int32_t descending, int64_t min_periods, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        if (!partition_begin) fail("null batch");
        Column v, o;
        if (values) v = materialize(col_at(values, 0));
        if (order) o = materialize(col_at(order, 0));
        *out = op_window_frame(values ? &v : nullptr, materialize(col_at(partition_begin, 0)), order ? &o : nullptr, agg, start_kind, start, end_kind, end,
                               descending, min_periods, values ? v.name.c_str() : "count");
    });
}
int b200_concat(const b200_batch *const *batches, int32_t n, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        require_device();
        if (n < 1) fail("concat: need at least one batch");
        auto o = std::make_unique<Batch>();
        const size_t nc = batches[0]->cols.size();
        std::vector<Batch> ins;
        for (int i = 0; i < n; i++) {
            if (batches[i]->cols.size() != nc) fail("concat: schema mismatch");
            o->num_rows += batches[i]->num_rows;
            ins.push_back(plain(batches[i]));
        }
        for (size_t c = 0; c < nc; c++) {
            std::vector<const Column *> cs;
            for (int i = 0; i < n; i++) cs.push_back(&ins[i].cols[c]);
            o->cols.push_back(concat_columns(cs));
        }
        op_done();
        *out = o.release();
    });
}

// ---- sort --------------------------------------------------------------------------------------------
int b200_argsort(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending,
                 const uint8_t *nulls_first, b200_batch **out_indices, char **errmsg) {
    return guard(errmsg, [&] {
        Buf idx = op_argsort(plain_keys(batch, key_cols, nkeys), key_cols, nkeys, descending, nulls_first);
        auto o = std::make_unique<Batch>();
        o->num_rows = batch->num_rows;
        o->cols.push_back(make_u64_column("indices", idx, batch->num_rows));
        *out_indices = o.release();
    });
}
int b200_sort(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending,
              const uint8_t *nulls_first, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        const Batch in = plain(batch);
        Buf idx = op_argsort(in, key_cols, nkeys, descending, nulls_first);
        *out = gather_batch(in, as<int64_t>(idx), in.num_rows, false, true);
        op_done();
    });
}
int b200_top_n(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending,
               const uint8_t *nulls_first, int64_t limit, int64_t offset, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] {
        // RecordBatch::top_n = argsort then slice (src/daft-recordbatch/src/ops/sort.rs:47-61)
        const int64_t n = batch->num_rows;
        const int64_t start = std::min<int64_t>(std::max<int64_t>(offset, 0), n);
        const int64_t end = std::min<int64_t>(n, start + std::max<int64_t>(limit, 0));
        const Batch in = plain(batch);
        Buf idx = op_top_indices(in, key_cols, nkeys, descending, nulls_first, end);
        *out = gather_batch(in, as<int64_t>(idx) + start, end - start, false);
        op_done();
    });
}

// ---- join --------------------------------------------------------------------------------------------
int b200_probe_table_create(int32_t nkeys, const uint8_t *null_equals_null, int32_t track_indices, b200_probe_table **out,
                            char **errmsg) {
    return guard(errmsg, [&] { *out = probe_table_create(nkeys, null_equals_null, track_indices != 0); });
}
int b200_probe_table_add(b200_probe_table *pt, const b200_batch *build_batch, const int32_t *key_cols, char **errmsg) {
    return guard(errmsg, [&] { probe_table_add(pt, *build_batch, key_cols); });
}
int b200_probe_table_build(b200_probe_table *pt, char **errmsg) {
    return guard(errmsg, [&] { probe_table_build(pt); });
}
int64_t b200_probe_table_num_rows(const b200_probe_table *pt) { return probe_table_num_rows(pt); }
void b200_probe_table_free(b200_probe_table *pt) { probe_table_free(pt); }
int b200_probe_exists(const b200_probe_table *pt, const b200_batch *probe, const int32_t *key_cols, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = probe_exists(pt, *probe, key_cols); });
}
int b200_probe_indices(const b200_probe_table *pt, const b200_batch *probe, const int32_t *key_cols, b200_batch **out,
                       char **errmsg) {
    return guard(errmsg, [&] { *out = probe_indices(pt, *probe, key_cols); });
}
int b200_join_state_create(const b200_probe_table *pt, b200_join_state **out, char **errmsg) {
    return guard(errmsg, [&] { *out = join_state_create(pt); });
}
void b200_join_state_free(b200_join_state *st) { join_state_free(st); }
int b200_hash_join_probe(const b200_probe_table *pt, b200_join_state *st, const b200_batch *probe, const int32_t *key_cols,
                         int32_t join_type, int32_t build_on_left, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = hash_join_probe(pt, st, *probe, key_cols, join_type, build_on_left != 0); });
}
int b200_hash_join_finalize(const b200_probe_table *pt, b200_join_state *st, const b200_batch *probe_schema_like,
                            int32_t build_on_left, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = hash_join_finalize(pt, st, plain(probe_schema_like), build_on_left != 0); });
}

// ---- exchange ----------------------------------------------------------------------------------------
int b200_comm_unique_id(uint8_t *out128, char **errmsg) {
    return guard(errmsg, [&] { comm_unique_id(out128); });
}
int b200_comm_init(const uint8_t *id128, int32_t rank, int32_t world, b200_comm **out, char **errmsg) {
    return guard(errmsg, [&] { *out = comm_init(id128, rank, world); });
}
void b200_comm_free(b200_comm *c) { comm_free(c); }
int b200_exchange_by_hash(b200_comm *c, const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, b200_batch **out,
                          char **errmsg) {
    return guard(errmsg, [&] { *out = exchange_by_hash(c, flat(batch, "exchange_by_hash"), key_cols, nkeys); });
}
int b200_exchange_random(b200_comm *c, const b200_batch *batch, uint64_t seed, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = exchange_random(c, flat(batch, "exchange_random"), seed); });
}
int b200_exchange_by_range(b200_comm *c, const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const b200_batch *boundaries,
                           const uint8_t *descending, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = exchange_by_range(c, flat(batch, "exchange_by_range"), key_cols, nkeys, plain(boundaries), descending); });
}
int b200_agg_allreduce(b200_comm *c, const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out, int32_t *merged,
                       char **errmsg) {
    return guard(errmsg, [&] {
        Batch *r = op_agg_allreduce(c, plain_for_plan(batch, B200_OP_AGG, plan, plan_len), plan, plan_len);
        *out = r;
        *merged = r ? 1 : 0;
    });
}
int b200_all_gather(b200_comm *c, const b200_batch *batch, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = all_gather_batches(c, flat(batch, "all_gather")); });
}

// ---- TPC-H generator -----------------------------------------------------------------------------------
// ---- Parquet -> HBM scan ----------------------------------------------------------------------------
int b200_parquet_read(const char *path, const char *const *columns, int32_t ncolumns, int64_t row_group_begin, int64_t row_group_end,
                      b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = parquet_read(path, columns, ncolumns, row_group_begin, row_group_end); });
}
int b200_parquet_file_info(const char *path, int64_t *row_groups, int64_t *rows, char **errmsg) {
    return guard(errmsg, [&] { parquet_file_info(path, row_groups, rows); });
}
int b200_tpch_generate(int32_t table, double scale_factor, int64_t row_begin, int64_t row_end, uint64_t seed,
                       const char *const *columns, int32_t ncolumns, b200_batch **out, char **errmsg) {
    return guard(errmsg, [&] { *out = tpch_generate(table, scale_factor, row_begin, row_end, seed, columns, ncolumns); });
}
int64_t b200_tpch_num_rows(int32_t table, double scale_factor) { return tpch_num_rows(table, scale_factor); }
int b200_tpch_generate_host(int32_t table, double scale_factor, int64_t row_begin, int64_t row_end, uint64_t seed,
                            const char *const *columns, int32_t ncolumns, struct ArrowArray *out_array,
                            struct ArrowSchema *out_schema, char **errmsg) {
    return guard(errmsg, [&] { tpch_generate_host(table, scale_factor, row_begin, row_end, seed, columns, ncolumns, out_array, out_schema); });
}

}  // extern "C"
#pragma GCC visibility pop

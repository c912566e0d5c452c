/*
 * daft_b200.h -- C ABI of libdaftb200.so: Daft's columnar hot path (hash / partition /
 * group-by aggregate / hash join / sort / filter / project / exchange) as sm_100a CUDA
 * kernels over GPU-resident Arrow buffers.
 *
 * Conventions are the reference's own stable C ABI (src/daft-ext/src/abi/mod.rs:20-113):
 *   - every call returns int, 0 = ok; on error a malloc'd NUL-terminated message is
 *     written to *errmsg and must be released with b200_free_string (abi/mod.rs:50-53);
 *   - data crosses the boundary as Arrow C Data Interface ArrowArray/ArrowSchema pairs
 *     (abi/arrow.rs); inputs are MOVED into the callee (their release callback is called
 *     by this library), outputs carry this library's release callbacks;
 *   - stateful objects are opaque handles with explicit *_free (ctx + fini in the
 *     reference's vtables, abi/mod.rs:66,107);
 *   - threading: entry points may be called from any host thread (each call makes the library's
 *     device current).  Concurrent calls are fine for read-only use of handles -- probing one built
 *     b200_probe_table from many threads (Probeable is Send+Sync,
 *     src/daft-recordbatch/src/probeable/mod.rs:33-85), reading the same batches, pushing to one
 *     b200_groupby_state; mutating calls on ONE handle (add/build, rename, free) must not race
 *     with other calls on it.  All kernels queue on the one library stream.
 *
 * There is NO CPU fallback behind any entry point: without a CUDA device every operator
 * returns an error.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   b200_hash_rows            RecordBatch::hash_rows          src/daft-recordbatch/src/ops/hash.rs:14-26
 *                             kernels::hashing::hash          src/daft-core/src/kernels/hashing.rs:304-445
 *   b200_partition_by_hash    RecordBatch::partition_by_hash  src/daft-recordbatch/src/ops/partition.rs:51-70
 *   b200_agg                  RecordBatch::agg / agg_groupby_inline
 *                                                             src/daft-recordbatch/src/ops/agg.rs:35-88,
 *                                                             src/daft-recordbatch/src/ops/inline_agg.rs:862-928
 *                             (+ the Filter/Project feeding it, fused)
 *   b200_filter               RecordBatch::filter             src/daft-recordbatch/src/lib.rs:520-570
 *   b200_eval_expression_list RecordBatch::eval_expression_list src/daft-recordbatch/src/lib.rs:1603-1631
 *   b200_take / b200_concat   RecordBatch::take / concat      src/daft-recordbatch/src/lib.rs
 *   b200_argsort/sort/top_n   RecordBatch::argsort/sort/top_n src/daft-recordbatch/src/ops/sort.rs:7-61
 *   b200_probe_table_*        ProbeableBuilder::{add_table,build}, Probeable::{probe_indices,probe_exists}
 *                                                             src/daft-recordbatch/src/probeable/mod.rs:33-126,
 *                                                             probe_table.rs:56-213
 *   b200_hash_join_probe/finalize  probe_inner / probe_left_right / probe_outer / probe_anti_semi + finalize_*
 *                                                             src/daft-local-execution/src/join/ (all files)
 *   b200_exchange_*           RepartitionSink + Flight shuffle src/daft-local-execution/src/sinks/repartition.rs:100-370,
 *                                                             src/daft-shuffles/src/ (all files)
 *   b200_merge_join_sorted / b200_asof_join_sorted / b200_cross_join
 *                             RecordBatch::sort_merge_join / asof_join / cross_join
 *                                                             src/daft-recordbatch/src/ops/joins/mod.rs:81-111,154-284
 *   b200_window_ranks / b200_window_offset
 *                             RecordBatch::window_row_number / window_rank / window_offset
 *                                                             src/daft-recordbatch/src/ops/window.rs:636-790
 *   b200_window_frame         RecordBatch::window_agg_dynamic_frame (ROWS and RANGE frames)
 *                                                             src/daft-recordbatch/src/ops/window.rs:299-633
 *   b200_parquet_read         stream_parquet behind the scan-task source
 *                                                             src/daft-parquet/src/read.rs:553,
 *                                                             src/daft-local-execution/src/sources/scan_task.rs:33
 */
#ifndef DAFT_B200_H
#define DAFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAFT_B200_ABI_VERSION 1

/* ---- Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) ---- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
    const char *format;
    const char *name;
    const char *metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema **children;
    struct ArrowSchema *dictionary;
    void (*release)(struct ArrowSchema *);
    void *private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void **buffers;
    struct ArrowArray **children;
    struct ArrowArray *dictionary;
    void (*release)(struct ArrowArray *);
    void *private_data;
};
#endif

/* ---- physical dtypes held on the device (Daft `as_physical` view) ---- */
typedef enum {
    B200_NULL = 0,
    B200_BOOL = 1, /* Arrow bitmap */
    B200_I8 = 2, B200_I16 = 3, B200_I32 = 4, B200_I64 = 5,
    B200_U8 = 6, B200_U16 = 7, B200_U32 = 8, B200_U64 = 9,
    B200_F32 = 10, B200_F64 = 11,
    B200_DATE32 = 12,       /* days since epoch, physical i32 */
    B200_TIMESTAMP = 13,    /* no tz, physical i64 */
    B200_LARGE_UTF8 = 14,   /* i64 offsets -- Daft always holds strings as Large */
    B200_LARGE_BINARY = 15,
    B200_LIST = 16          /* i64 offsets + one child column (Daft List): produced by the list / set aggregates.  Where a
                             * dtype is reported as one int32 (b200_batch_dtypes) the child's code sits in bits 8..15.
                             * List columns can be exported, selected, sliced by the host; operators do not read them. */
} b200_dtype;

/* One column as raw DEVICE pointers (Arrow layout).  Used to wrap buffers that already
 * live in HBM (b200_batch_from_device) and to look inside a batch (b200_batch_describe). */
typedef struct {
    int32_t dtype;           /* b200_dtype */
    int32_t _reserved;
    int64_t length;
    int64_t null_count;
    const void *validity;    /* device; Arrow LSB bitmap, NULL = no nulls */
    const void *offsets;     /* device; int64[length+1] for LARGE_UTF8/LARGE_BINARY, else NULL */
    const void *data;        /* device; values / bitmap (BOOL) / bytes */
    int64_t data_bytes;      /* size of `data` in bytes */
    const char *name;        /* UTF-8, borrowed */
} b200_column_desc;

typedef struct b200_batch b200_batch;             /* device-resident RecordBatch */
typedef struct b200_probe_table b200_probe_table; /* join build side: hash table + build batches */
typedef struct b200_join_state b200_join_state;   /* per-join matched-build-row bitmap (outer/left/right) */

/* ---- library ---- */
int b200_abi_version(void);
int b200_init(int device_ordinal, char **errmsg);  /* selects the device, creates the library stream + pool */
void b200_free_string(char *s);
void *b200_stream(void);                           /* cudaStream_t all operators are enqueued on */
/* Results are stream-ordered: operator calls return when their work is queued on b200_stream(); a host that reads
 * device buffers itself (b200_batch_describe) or hands them to another stream waits here first.  b200_batch_export
 * waits by itself.  (The reference's operators return materialised batches; the equivalent point is the export.) */
int b200_synchronize(char **errmsg);
/* number of kernel launches issued by this library since load (bench.py "gpu_launches") */
int64_t b200_launch_count(void);
/* bench support: when enabled, the fused-aggregate kernel launch is bracketed by CUDA events on the library
 * stream; b200_last_agg_kernel_ms returns the duration of the most recent one (-1 if none). */
int b200_kernel_timing(int enable);
double b200_last_agg_kernel_ms(void);
/* per-kernel device time since the last report, as JSON text {"<kernel>": {"launches": n, "ms": total}, ...} (waits for the
 * stream; *out_json is malloc'd: release with b200_free_string).  Covers the join probe / build kernels, the radix sort
 * passes, partition and exchange kernels; the fused aggregate is reported by b200_last_agg_kernel_ms. */
int b200_kernel_timing_report(char **out_json, char **errmsg);
/* JIT cache: directory holding generated .cu/.cubin (default <lib dir>/_jitcache) */
int b200_set_jit_cache_dir(const char *path, char **errmsg);
/* compile (no launch) the kernels a plan needs -- usable without a GPU (build check) */
int b200_precompile(int32_t op, const uint8_t *plan, size_t plan_len, const int32_t *col_dtypes,
                    const uint8_t *col_nullable, int32_t ncols, char **errmsg);
/* Code generation only (no CUDA, no compile): the kernel parameter layout and the generated source of a plan, as text
 *   "ptrs=<kind>:<index>,...\nints=<kind>,...\nouts=<dtype>:<nullable>,...\n---\n<CUDA source>"
 * (ptr kinds: 0 column data, 1 column validity, 2 column offsets, 3 output data, 4 output validity, 5 output lengths,
 * 6 output offsets, 7 take indices, ...; see csrc/jit.h).  Bit 16 of `op` asks for the projection build that evaluates at
 * gathered rows (take indices, negative = null row).  For aggregate plans bits 8-15 of `op` select the build
 * (0 direct loads, 1 global table, 2 TMA-staged) and the text also carries the group-table cell layout
 * (nkeys / nlane / nwarp / aggs / cellops).  Lets tests/test_jit_host_emulation.py run the generated per-row code on the
 * host against the oracle.  *out_text is malloc'd: release with b200_free_string. */
int b200_jit_describe(int32_t op, const uint8_t *plan, size_t plan_len, const int32_t *col_dtypes,
                      const uint8_t *col_nullable, int32_t ncols, char **out_text, char **errmsg);

/* ---- batches ---- */
/* HOST -> device.  `array` must be a struct array (one child per column), as produced by
 * RecordBatch export; it is moved (released by this call).  Utf8/Binary are widened to Large. */
int b200_batch_import(struct ArrowArray *array, struct ArrowSchema *schema, b200_batch **out, char **errmsg);
/* device -> HOST; out_array/out_schema get this library's release callbacks. */
int b200_batch_export(const b200_batch *batch, struct ArrowArray *out_array, struct ArrowSchema *out_schema,
                      char **errmsg);
/* wrap existing DEVICE buffers without copying (borrowed; caller keeps them alive).  Requirements on borrowed buffers:
 * every buffer must be followed by 16 readable bytes (string values are fetched with aligned word loads, validity /
 * Boolean bitmaps with 32-bit word loads: csrc/common.h) and fixed-width `data` must be aligned to its element size.
 * String offsets may start anywhere (a sliced Arrow array): they are re-based to 0 here (a shifted copy of the offsets;
 * the value bytes stay borrowed); offsets that point outside [0, data_bytes] are rejected. */
int b200_batch_from_device(const b200_column_desc *cols, int32_t ncols, int64_t num_rows, b200_batch **out,
                           char **errmsg);
/* Columns of a join output are VIEWS (row-id selection vectors over the join inputs) until an operator reads them: for
 * such a column describe reports dtype, length and name only (null_count = -1, pointers NULL).  b200_batch_materialize
 * returns the same batch with every column gathered into plain Arrow buffers (shared with `batch`: the gather happens once). */
int b200_batch_describe(const b200_batch *batch, int32_t col, b200_column_desc *out, char **errmsg);
int b200_batch_materialize(const b200_batch *batch, b200_batch **out, char **errmsg);
/* dtype codes of the first n columns in one call (hosts that already know the column names skip b200_batch_describe) */
int b200_batch_dtypes(const b200_batch *batch, int32_t *out_dtypes, int32_t n, char **errmsg);
int64_t b200_batch_num_rows(const b200_batch *batch);
int32_t b200_batch_num_columns(const b200_batch *batch);
int64_t b200_batch_size_bytes(const b200_batch *batch);
void b200_batch_free(b200_batch *batch);
/* columns by index -> new batch sharing the same device buffers */
int b200_batch_select(const b200_batch *batch, const int32_t *cols, int32_t ncols, b200_batch **out, char **errmsg);
int b200_batch_rename(b200_batch *batch, int32_t col, const char *name, char **errmsg);
/* horizontal union of columns (RecordBatch::union) */
int b200_batch_union(const b200_batch *a, const b200_batch *b, b200_batch **out, char **errmsg);
int b200_batch_slice(const b200_batch *batch, int64_t start, int64_t end, b200_batch **out, char **errmsg);

/* ---- a1/a2: hashing ---- */
/* out: one UInt64 column "hash", no nulls.  seed: optional batch whose column 0 is UInt64[n]. */
int b200_hash_rows(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const b200_batch *seed,
                   b200_batch **out, char **errmsg);

/* ---- a3: hash partition ---- */
/* out[0..P): rows with hash_rows(keys) % P == p, input order preserved inside a partition. */
int b200_partition_by_hash(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys,
                           int32_t num_partitions, b200_batch **out, char **errmsg);

/* RepartitionSpec::Random (RecordBatch::partition_by_random, src/daft-recordbatch/src/ops/partition.rs:72-86): out[0..P), rows
 * placed by a seeded counter-based generator; input order kept inside a partition.  Deterministic per (seed, input); the
 * placement is NOT rand::StdRng's (the reference's tests pin conservation and determinism only). */
int b200_partition_by_random(const b200_batch *batch, int32_t num_partitions, uint64_t seed, b200_batch **out, char **errmsg);
/* a12 RepartitionSpec::Range: positions[i] = rows of `sorted` (exactly the key columns, ascending / descending per flag,
 * nulls first where descending) that compare <= key row i (RecordBatch::search_sorted,
 * src/daft-recordbatch/src/ops/search_sorted.rs:13-36; src/daft-core/src/kernels/search_sorted.rs:363-433) -> UInt64 column "idx" */
int b200_search_sorted(const b200_batch *sorted, const b200_batch *batch, const int32_t *key_cols, int32_t nkeys,
                       const uint8_t *descending, b200_batch **out, char **errmsg);
/* out[0 .. boundaries.num_rows]: rows whose search_sorted position is p, input order preserved
 * (RecordBatch::partition_by_range, src/daft-recordbatch/src/ops/partition.rs:88-101); empty boundaries -> out[0] = the batch */
int b200_partition_by_range(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const b200_batch *boundaries,
                            const uint8_t *descending, b200_batch **out, char **errmsg);

/* ---- a10/a11: filter / project.  `plan` = serialised expression program (see below). ---- */
int b200_filter(const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out, char **errmsg);
int b200_eval_expression_list(const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out,
                              char **errmsg);

/* ---- a4/a5/a6: grouped / ungrouped aggregate with the filter + projections fused in ---- */
int b200_agg(const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out, char **errmsg);
/* GroupedAggregateSink / AggregateSink state (src/daft-local-execution/src/sinks/grouped_aggregate.rs:146-488,
 * sinks/aggregate.rs:54-123): make_state -> sink(morsel)* -> finalize.  The three AGG / AGG / PROJECT plans are what the
 * reference's planner hands the sink (populate_aggregation_stages, src/daft-local-plan/src/agg.rs:13-416): the partial
 * aggregation run on every pushed batch, the final aggregation over the concatenated partials (its input schema = the
 * partial plan's output: key columns, then aggregates) and the final projection.  Partials stay in HBM.  push may be
 * called from several host threads (the reference gives every worker its own state; here one state serialises its
 * bookkeeping and the kernels queue on the library stream). */
typedef struct b200_groupby_state b200_groupby_state;
int b200_groupby_create(const uint8_t *partial_plan, size_t partial_len, const uint8_t *final_plan, size_t final_len,
                        const uint8_t *project_plan, size_t project_len, b200_groupby_state **out, char **errmsg);
int b200_groupby_push(b200_groupby_state *st, const b200_batch *batch, char **errmsg);
int b200_groupby_finalize(b200_groupby_state *st, b200_batch **out, char **errmsg);
void b200_groupby_free(b200_groupby_state *st);
/* a13: distinct rows over key columns (first occurrence kept, first-occurrence order) */
int b200_dedup(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, b200_batch **out, char **errmsg);

/* ---- take / concat ---- */
/* indices: column 0 of `indices`, any integer dtype; null index -> null row.  An index outside [0, rows) fails with
 * "index out of bounds" (tests/recordbatch/test_take.py:31-34): the indices are validated on the device before the gather,
 * which costs one 4-byte read back. */
int b200_take(const b200_batch *batch, const b200_batch *indices, b200_batch **out, char **errmsg);
int b200_concat(const b200_batch *const *batches, int32_t n, b200_batch **out, char **errmsg);

/* ---- a9: sort ---- */
int b200_argsort(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending,
                 const uint8_t *nulls_first, b200_batch **out_indices /* UInt64 */, char **errmsg);
int b200_sort(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending,
              const uint8_t *nulls_first, b200_batch **out, char **errmsg);
int b200_top_n(const b200_batch *batch, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending,
               const uint8_t *nulls_first, int64_t limit, int64_t offset, b200_batch **out, char **errmsg);

/* ---- a7/a8: hash join ---- */
typedef enum { B200_JOIN_INNER = 0, B200_JOIN_LEFT = 1, B200_JOIN_RIGHT = 2, B200_JOIN_OUTER = 3,
               B200_JOIN_SEMI = 4, B200_JOIN_ANTI = 5 } b200_join_type;

/* ProbeableBuilder: create -> add_table* -> build.  `key_cols` index into each added batch. */
int b200_probe_table_create(int32_t nkeys, const uint8_t *null_equals_null /* nkeys or NULL */,
                            int32_t track_indices, b200_probe_table **out, char **errmsg);
int b200_probe_table_add(b200_probe_table *pt, const b200_batch *build_batch, const int32_t *key_cols,
                         char **errmsg);
int b200_probe_table_build(b200_probe_table *pt, char **errmsg);
int64_t b200_probe_table_num_rows(const b200_probe_table *pt);
void b200_probe_table_free(b200_probe_table *pt);
/* Probeable::probe_exists -> Boolean column; Probeable::probe_indices -> (probe_idx, build_idx) UInt64 pairs
 * in probe-row order, matches of one probe row in build insertion order. */
int b200_probe_exists(const b200_probe_table *pt, const b200_batch *probe, const int32_t *key_cols,
                      b200_batch **out, char **errmsg);
int b200_probe_indices(const b200_probe_table *pt, const b200_batch *probe, const int32_t *key_cols,
                       b200_batch **out, char **errmsg);
/* One probe batch through the join operator.  Output columns: probe-side columns `probe_out_cols`
 * and build-side columns `build_out_cols` gathered per match, build side first iff build_on_left. */
int b200_join_state_create(const b200_probe_table *pt, b200_join_state **out, char **errmsg);
void b200_join_state_free(b200_join_state *st);
int b200_hash_join_probe(const b200_probe_table *pt, b200_join_state *st /* NULL for inner/semi/anti */,
                         const b200_batch *probe, const int32_t *key_cols, int32_t join_type,
                         int32_t build_on_left, b200_batch **out, char **errmsg);
/* unmatched build rows (left/right/outer builds), null-extended on the probe side whose schema is taken
 * from `probe_schema_like` (a batch with the probe-side columns; may have 0 rows). */
int b200_hash_join_finalize(const b200_probe_table *pt, b200_join_state *st, const b200_batch *probe_schema_like,
                            int32_t build_on_left, b200_batch **out, char **errmsg);

/* ---- joins over SORTED inputs and the cross join ----
 * RecordBatch::sort_merge_join / asof_join / cross_join (src/daft-recordbatch/src/ops/joins/mod.rs:81-111,154-229,231-284) behind
 * SortMergeJoin / AsofJoinOperator / CrossJoinOperator (src/daft-local-execution/src/join/{sort_merge_join,asof_join,cross_join}.rs).
 * Output of all three: [left columns..., right columns...] as views (selection vectors), names unchanged -- the caller applies the
 * join's output schema.  Key dtypes must match pairwise (cast to the supertype first, ops/joins/mod.rs:22-46).
 *  merge_join_sorted: inner join; both sides sorted ascending, nulls last, on their keys; a null in any key matches nothing.
 *  asof_join_sorted : keys = by columns..., then the on column; `right` sorted ascending (nulls last) on them, `left` in any
 *                     order; every left row appears once, with the last right row of its by-group whose on-key is <= its own
 *                     (strategy "backward"), or a null row.
 *  cross_join       : every pair; outer_loop_left != 0 repeats each left row |right| times in a row (JoinSide::Left). */
int b200_merge_join_sorted(const b200_batch *left, const int32_t *left_keys, const b200_batch *right, const int32_t *right_keys,
                           int32_t nkeys, b200_batch **out, char **errmsg);
int b200_asof_join_sorted(const b200_batch *left, const int32_t *left_keys, const b200_batch *right, const int32_t *right_keys,
                          int32_t nkeys, b200_batch **out, char **errmsg);
int b200_cross_join(const b200_batch *left, const b200_batch *right, int32_t outer_loop_left, b200_batch **out, char **errmsg);

/* ---- window functions ----
 * RecordBatch::window_row_number / window_rank / window_offset (src/daft-recordbatch/src/ops/window.rs:636-790) for ALL partitions
 * of a batch at once: `sorted` is ordered by its partition key columns, then its order key columns (b200_sort), where the
 * reference hands its kernels one partition at a time (src/daft-local-execution/src/sinks/window_partition_and_order_by.rs).
 *  window_ranks : five columns, one row per input row: row_number, rank, dense_rank (UInt64), partition_begin (Int64: index of
 *                 the partition's first row), partition_id (Int64, dense, in sorted order).  Rows are peers when their order
 *                 keys are equal with null == null and NaN == NaN (build_multi_array_is_equal, window.rs:661-668).
 *  window_offset: the columns of `values` read at row + offset when that row lies in the same partition (lead: offset > 0,
 *                 lag: offset < 0), null rows otherwise, plus an Int64 column "__src" = the source row or -1 (the caller puts
 *                 the default value there, window.rs:706-789). */
int b200_window_ranks(const b200_batch *sorted, const int32_t *partition_cols, int32_t npartition, const int32_t *order_cols,
                      int32_t norder, b200_batch **out, char **errmsg);
int b200_window_offset(const b200_batch *values, const b200_batch *partition_ids, int64_t offset, b200_batch **out, char **errmsg);
/*  window_frame : one aggregate over a ROWS or RANGE frame -- RecordBatch::window_agg_dynamic_frame (window.rs:299-633;
 *                 WindowPartitionAndDynamicFrameSink, sinks/window_partition_and_dynamic_frame.rs) -- for every partition of the sorted
 *                 batch.  `values`: one column, already cast to the type the reference accumulates in (Int64 / UInt64 / Float32 /
 *                 Float64: window_states/sum.rs:137-158; Float64 for mean), or NULL for count(*); `partition_begin`: that column of
 *                 b200_window_ranks.  agg: 0 sum, 1 count(all), 2 count(valid), 3 count(null), 4 mean, 5 min, 6 max (min / max: Int64 /
 *                 UInt64 values).  start_kind / end_kind: 0 UNBOUNDED PRECEDING / FOLLOWING, 1 `start` / `end` are row offsets from the
 *                 current row, 2 they are offsets on the ORDER BY value (RANGE: `order` = that one column as Int64, `descending` its
 *                 direction; NULL otherwise).  Frames with fewer than min_periods rows give null.  Sums slide exactly as the
 *                 reference's state does (rows leaving the frame are subtracted, rows entering it added), so floats agree bit for
 *                 bit.  A frame of negative size fails with the reference's message. */
int b200_window_frame(const b200_batch *values, const b200_batch *partition_begin, const b200_batch *order, int32_t agg, int32_t start_kind,
                      int64_t start, int32_t end_kind, int64_t end, int32_t descending, int64_t min_periods, b200_batch **out, char **errmsg);

/* ---- a12: exchange (hash repartition over NCCL; one process per GPU) ---- */
typedef struct b200_comm b200_comm;
int b200_comm_unique_id(uint8_t *out128, char **errmsg); /* rank 0: ncclGetUniqueId, 128 bytes */
int b200_comm_init(const uint8_t *id128, int32_t rank, int32_t world, b200_comm **out, char **errmsg);
void b200_comm_free(b200_comm *c);
/* hash-partition `batch` into `world` pieces and all-to-all them; out = this rank's piece. */
int b200_exchange_by_hash(b200_comm *c, const b200_batch *batch, const int32_t *key_cols, int32_t nkeys,
                          b200_batch **out, char **errmsg);
/* RepartitionSpec::Random / ::Range over the communicator (src/daft-local-execution/src/sinks/repartition.rs:166-184).  Range:
 * `boundaries` holds world - 1 rows of the key columns (RecordBatch::quantiles of the gathered samples,
 * src/daft-distributed/src/pipeline_node/sort.rs:153-188); rank r receives the rows whose search_sorted position is r. */
int b200_exchange_random(b200_comm *c, const b200_batch *batch, uint64_t seed, b200_batch **out, char **errmsg);
int b200_exchange_by_range(b200_comm *c, const b200_batch *batch, const int32_t *key_cols, int32_t nkeys,
                           const b200_batch *boundaries, const uint8_t *descending, b200_batch **out, char **errmsg);
/* Low-cardinality HashAggregate over one scan partition per rank (the reference's partial aggregate -> gather -> final
 * aggregate, src/daft-distributed/src/pipeline_node/aggregate.rs:281-340, as ONE stage): every rank runs b200_agg's
 * kernel, the group tables are all-reduced over NVLink (one grouped NCCL launch), and every rank finishes with the global
 * result.  `plan` is the single-stage AGG plan (means included).  *merged = 0 (and *out = NULL) on EVERY rank when the
 * tables cannot be merged slot by slot (more than ~1000 groups, key codes that are not injective, any_value / Float32
 * sums, a group without a local row): the caller then runs the two-stage plan through b200_all_gather. */
int b200_agg_allreduce(b200_comm *c, const b200_batch *batch, const uint8_t *plan, size_t plan_len, b200_batch **out,
                       int32_t *merged, char **errmsg);
/* every rank receives the concatenation of all ranks' batches in rank order (broadcast join / partial-agg merge) */
int b200_all_gather(b200_comm *c, const b200_batch *batch, b200_batch **out, char **errmsg);

/* ---- Parquet -> HBM scan ----
 * The reference's scan source (src/daft-local-execution/src/sources/scan_task.rs:33 -> src/daft-parquet/src/read.rs:553
 * stream_parquet) for a local file: row groups [row_group_begin, row_group_end) (end < 0 = all) of the named columns
 * (NULL / 0 = every column) as one batch.  The host reads the footer, the column chunks' bytes and the page headers; the
 * compressed chunks are copied to HBM as they are and every page is decompressed and decoded on the device.
 * Supported: flat schemas; BOOLEAN, INT32 (incl. DATE, INT(8 | 16 | 32) signed and unsigned), INT64 (incl. TIMESTAMP millis / micros / nanos -> the
 * library's microsecond Timestamp, UINT_64), FLOAT, DOUBLE, BYTE_ARRAY
 * (String -> LargeUtf8, otherwise LargeBinary); PLAIN and PLAIN_DICTIONARY / RLE_DICTIONARY pages, data pages V1 and V2 (RLE
 * Booleans); UNCOMPRESSED and SNAPPY.  Anything else is an error naming the column and the feature (no CPU fallback). */
int b200_parquet_read(const char *path, const char *const *columns, int32_t ncolumns, int64_t row_group_begin,
                      int64_t row_group_end, b200_batch **out, char **errmsg);
/* footer only: number of row groups and rows */
int b200_parquet_file_info(const char *path, int64_t *row_groups, int64_t *rows, char **errmsg);

/* ---- synthetic TPC-H tables generated straight into HBM (bench / tests) ---- */
typedef enum { B200_TPCH_LINEITEM = 0, B200_TPCH_ORDERS = 1, B200_TPCH_CUSTOMER = 2, B200_TPCH_SUPPLIER = 3,
               B200_TPCH_NATION = 4, B200_TPCH_REGION = 5, B200_TPCH_PART = 6, B200_TPCH_PARTSUPP = 7 } b200_tpch_table;
/* rows [row_begin,row_end) of `table` at `scale_factor`; columns selected by name list (NULL = all implemented). */
int b200_tpch_generate(int32_t table, double scale_factor, int64_t row_begin, int64_t row_end, uint64_t seed,
                       const char *const *columns, int32_t ncolumns, b200_batch **out, char **errmsg);
int64_t b200_tpch_num_rows(int32_t table, double scale_factor);
/* same generator on the host (fills caller-provided host buffers through Arrow export) */
int b200_tpch_generate_host(int32_t table, double scale_factor, int64_t row_begin, int64_t row_end, uint64_t seed,
                            const char *const *columns, int32_t ncolumns, struct ArrowArray *out_array,
                            struct ArrowSchema *out_schema, char **errmsg);

/*
 * ---- expression / plan encoding (little-endian byte stream) ----
 *
 * This is the serialised form of the plan-node parameters the reference hands its operators
 * (src/daft-local-plan/src/plan.rs: Project:1884, Filter:1921, UnGroupedAggregate:2030,
 * HashAggregate:2040-2047) with BoundExpr column references already resolved to indices.
 *
 *   str    := u32 len, bytes
 *   node   := u8 tag, payload
 *     0 COL      u32 column_index
 *     1 LIT      u8 dtype, u8 is_null, (BOOL/ints: i64 | F32/F64: f64 | DATE32/TIMESTAMP: i64 | UTF8/BINARY: str)
 *     2 BINARY   u8 op, node lhs, node rhs
 *     3 UNARY    u8 op, node child
 *     4 CAST     u8 dtype, node child   (numeric / bool / date casts; a value the target integer type cannot hold, or a
 *                NaN / infinite float cast to an integer, is NULL: arrow-rs safe cast, src/daft-core/src/array/ops/cast.rs:194-196)
 *     5 IF_ELSE  node predicate, node if_true, node if_false
 *     6 IS_IN    node child, u32 n, n x node(LIT)
 *     7 BETWEEN  node child, node lower, node upper
 *     8 FILL_NULL node child, node fill
 *     9 FUNCTION u8 fn, u8 nargs, nargs x node
 *   binary ops (src/daft-dsl Operator, evaluated at src/daft-recordbatch/src/lib.rs:1376-1400):
 *     0 add 1 sub 2 mul 3 true_div 4 floor_div 5 mod 6 lt 7 le 8 eq 9 ne 10 ge 11 gt 12 and 13 or 14 xor
 *     15 eq_null_safe 16 shift_left 17 shift_right
 *     and / or / xor: Kleene logic on booleans; bitwise on integers (null if either side is null,
 *     src/daft-core/src/array/ops/bitwise.rs, tests/recordbatch/test_bitwise.py).
 *     lt..gt / eq / ne / eq_null_safe / BETWEEN on floats use IEEE 754 totalOrder like the arrow-rs `cmp` kernels the reference calls
 *     (src/daft-core/src/array/ops/comparison.rs:541-603): NaN == NaN, +NaN > +inf, -0.0 < +0.0.  IS_IN on floats follows the
 *     reference's BTreeSet<FloatWrapper> (is_in.rs:44-66): any NaN matches a NaN item, -0.0 matches 0.0.
 *     floor_div / mod follow the reference's Rust arithmetic (src/daft-core/src/array/ops/arithmetic.rs:245-300,
 *     series/ops/arithmetic.rs:456-466): on integers the quotient truncates toward zero and the remainder takes the sign
 *     of the dividend; on floats floor_div = floor(l / r) and mod = fmod(l, r).  An integer divisor of 0 yields null
 *     (the reference panics); MIN / -1 wraps.
 *   unary ops: 0 not 1 negate 2 is_null 3 not_null 4 abs
 *   functions: 0 year 1 month 2 day 3 utf8_length 4 utf8_startswith 5 utf8_endswith 6 utf8_contains
 *              7 utf8_left 8 day_of_week 9 sqrt (Float32 / Float64 operand, same type; IEEE correctly rounded)
 *              10 is_nan 11 not_nan 12 is_inf (float operand -> Boolean; nulls stay null, src/daft-core/src/array/ops/float.rs)
 *              13 ceil 14 floor (float operand, same type) 15 sign (numeric operand, same type: signum; 0 / 1 for unsigned)
 *              16 round (float operand, Int32 literal decimals: (v * 10^d).round() / 10^d, half away from zero,
 *                 src/daft-core/src/array/ops/round.rs:14-35)
 *              17 match_seq (string operand, Int32 literal flags, then string literals): the operand contains the literals in
 *                 this order without overlap; flags bit 0 anchors the first at the start, bit 1 the last at the end -- utf8
 *                 `match` / `like` for patterns made of literal segments and `.*` / `%` (TPC-H Q13 / Q16, ClickBench)
 *   Operands of a binary op must already have identical dtypes (the host inserts the supertype
 *   casts the reference decides at plan time); true_div requires F64 operands.
 *
 *   PROJECT plan := u32 nexprs, nexprs x (str name, node)
 *   FILTER  plan := node predicate               (conjuncts already AND-ed, lib.rs:520-541)
 *   AGG     plan := u8 has_filter, [node filter], u32 nkeys, nkeys x (str name, node),
 *                   u32 naggs, naggs x (u8 kind, u8 count_mode, str name, node child)
 *     agg kinds: 0 count 1 sum 2 min 3 max 4 mean 5 any_value 6 list 7 set 8 concat;  count modes: 0 all 1 valid 2 null;
 *     any_value mode: 0 first row of the group, 1 ignore_nulls (first non-null row, null for an all-null group);
 *     list: the group's values in row order, nulls kept (B200_LIST);  set: its distinct non-null values in first-occurrence
 *     order (B200_LIST);  concat: its non-null strings back to back (string children only);  mode byte 0 for all three;
 *     min / max over Boolean are bool_and / bool_or
 */
#define B200_OP_PROJECT 0
#define B200_OP_FILTER 1
#define B200_OP_AGG 2

#ifdef __cplusplus
}
#endif
#endif /* DAFT_B200_H */

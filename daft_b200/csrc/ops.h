// Operators implemented ahead of time (fixed-function kernels): sort, partition, join, exchange, generator.
#pragma once
#include "common.h"

struct b200_probe_table;
struct b200_join_state;
struct b200_comm;

namespace b200 {

// batch.cu
struct HostCol {
    std::string name;
    int32_t dtype;
    int64_t length;
    void *data;        // malloc'd, ownership moves into the exported ArrowArray
    int64_t *offsets;  // malloc'd or null
    int64_t data_bytes;
};
void export_host_columns(std::vector<HostCol> &cols, int64_t num_rows, ArrowArray *out_array, ArrowSchema *out_schema);
Batch *import_batch(ArrowArray *array, ArrowSchema *schema);
void export_batch(const Batch &b, ArrowArray *out_array, ArrowSchema *out_schema);
Batch *batch_from_device(const b200_column_desc *cols, int32_t ncols, int64_t num_rows);

// sort.cu -- stable LSD radix sort of (u64 key, i64 payload) pairs and everything built on it
// sorts in place logically: on return keys/vals hold the sorted sequence (buffers may be swapped)
void radix_sort_pairs(Buf &keys, Buf &vals, int64_t n, int begin_bit, int end_bit, bool descending);
void fill_iota(int64_t *out, int64_t n);
Buf op_argsort(const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending, const uint8_t *nulls_first);
Buf op_search_sorted(const Batch &sorted, const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending);
std::vector<Batch *> op_partition_by_range(const Batch &b, const int32_t *key_cols, int32_t nkeys, const Batch &boundaries,
                                           const uint8_t *descending);
// the first k entries of that argsort, found by selection when k is small (sort.cu)
Buf op_top_indices(const Batch &b, const int32_t *key_cols, int32_t nkeys, const uint8_t *descending, const uint8_t *nulls_first, int64_t k);
Buf sort_slots_by_minrow(Buf slots, Buf minrow, int64_t ng);
// the joins over sorted inputs and the cross join (sort.cu): outputs are [left columns..., right columns...] as views
Batch *op_merge_join_sorted(const Batch &left, const int32_t *left_keys, const Batch &right, const int32_t *right_keys, int32_t nkeys);
Batch *op_asof_join_sorted(const Batch &left, const int32_t *left_keys, const Batch &right, const int32_t *right_keys, int32_t nkeys);
Batch *op_cross_join(const Batch &left, const Batch &right, bool outer_loop_left);
// window functions over a batch sorted by (partition keys, order keys) (sort.cu)
Batch *op_window_ranks(const Batch &sorted, const int32_t *part_cols, int32_t npart, const int32_t *order_cols, int32_t norder);
Batch *op_window_offset(const Batch &values, const Column &partition_id, int64_t offset);
// window_frame.cu: an aggregate over a ROWS / RANGE frame of every partition (values: null for count(*); order: the ORDER BY column, RANGE only)
Batch *op_window_frame(const Column *values, const Column &partition_begin, const Column *order, int32_t agg, int32_t start_kind, int64_t start,
                       int32_t end_kind, int64_t end, int32_t descending, int64_t min_periods, const char *name);

// ops_jit.cu: ascending indices of the set (or, inverted, clear) bits among the first n bits
Buf bitmap_select(const uint32_t *words, int64_t n, bool invert, int64_t *count_out);

// partition.cu
// stable partition of rows by pid = hash % P: returns row permutation and fills host offsets[P+1]
Buf partition_permutation(const uint64_t *hashes, int64_t n, int32_t P, std::vector<int64_t> &offsets);
// every partition of `b` back to back in one batch, P <= 256 (direct scatter: no row-id permutation, no per-column gather)
// rowid_out != null: columns the scatter cannot place (nullable / Boolean / strings) are left empty and *rowid_out
// receives the partition-ordered source row ids; otherwise they are gathered into the staged batch
// Where a row's partition comes from: a u64 per row (hash_rows output or explicit targets; partition = value % P), or ONE
// never-null 4- / 8-byte integer key column hashed in registers (XXH3-64 of the value, the same number hash_rows writes) --
// the hash column is then neither written nor read.
struct PartSrc {
    const uint64_t *hashes = nullptr;
    const void *key = nullptr;
    int width = 0;  // 4 or 8 with `key`
    Buf keep;       // owns `hashes` when this object computed them
};
// key mode when `key_cols` is one suitable column (and the direct scatter will be used), hash_rows otherwise
PartSrc part_src_for_keys(const Batch &b, const int32_t *key_cols, int32_t nkeys, bool allow_fused);
Batch *partition_staged(const Batch &b, const PartSrc &src, int32_t P, std::vector<int64_t> &offsets, Buf *rowid_out = nullptr);
// the direct scatter in two steps (peer-store exchange): per-partition counts, then rows written straight to per-destination
// pointers (peer memory)
struct PartitionPlan {
    int64_t n = 0, nblocks = 0, chunk = 0;
    int32_t P = 0;
    Buf base;                      // scanned (partition, CTA chunk) histogram
    Buf pid;                       // one byte per row: its partition (written by the count pass, read by the scatter)
    std::vector<int64_t> offsets;  // rows of partition p = [offsets[p], offsets[p + 1]) in partition order
};
PartitionPlan partition_count(const PartSrc &src, int64_t n, int32_t P);
void partition_scatter_to_peers(const PartitionPlan &pl, const Batch &b, const std::vector<int> &cols,
                                const std::vector<std::vector<void *>> &dest_ptrs, Buf *rowid_out);
Buf random_targets(int64_t n, int32_t P, uint64_t seed, uint64_t salt);  // u64 target per row (RepartitionSpec::Random)
std::vector<Batch *> op_partition_by_random(const Batch &b, int32_t P, uint64_t seed);
std::vector<Batch *> op_partition_by_hash(const Batch &b, const int32_t *key_cols, int32_t nkeys, int32_t P);
Batch *op_dedup(const Batch &b, const int32_t *key_cols, int32_t nkeys);

// join.cu
b200_probe_table *probe_table_create(int32_t nkeys, const uint8_t *null_equals_null, bool track_indices);
void probe_table_add(b200_probe_table *pt, const Batch &build, const int32_t *key_cols);
void probe_table_build(b200_probe_table *pt);
int64_t probe_table_num_rows(const b200_probe_table *pt);
void probe_table_free(b200_probe_table *pt);
Batch *probe_exists(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols);
Batch *probe_indices(const b200_probe_table *pt, const Batch &probe, const int32_t *key_cols);
b200_join_state *join_state_create(const b200_probe_table *pt);
void join_state_free(b200_join_state *st);
Batch *hash_join_probe(const b200_probe_table *pt, b200_join_state *st, const Batch &probe, const int32_t *key_cols,
                       int32_t join_type, bool build_on_left);
Batch *hash_join_finalize(const b200_probe_table *pt, b200_join_state *st, const Batch &probe_like, bool build_on_left);

// exchange.cu
void comm_unique_id(uint8_t *out128);
b200_comm *comm_init(const uint8_t *id128, int32_t rank, int32_t world);
void comm_free(b200_comm *c);
Batch *exchange_by_hash(b200_comm *c, const Batch &b, const int32_t *key_cols, int32_t nkeys);
Batch *exchange_random(b200_comm *c, const Batch &b, uint64_t seed);
Batch *exchange_by_range(b200_comm *c, const Batch &b, const int32_t *key_cols, int32_t nkeys, const Batch &boundaries, const uint8_t *descending);
Batch *all_gather_batches(b200_comm *c, const Batch &b);
void allreduce_agg_table(b200_comm *c, unsigned long long *acc, size_t cap, const std::vector<int> &cell_ops, const std::vector<char> &keep_local,
                         unsigned long long *tagmin, int *status_flag);

// parquet.cu: Parquet column chunks decoded on the device into a batch (flat columns; PLAIN / dictionary; Snappy / uncompressed)
Batch *parquet_read(const char *path, const char *const *columns, int32_t ncolumns, int64_t rg_begin, int64_t rg_end);
void parquet_file_info(const char *path, int64_t *row_groups, int64_t *rows);

// tpch_gen.cu
Batch *tpch_generate(int32_t table, double sf, int64_t row_begin, int64_t row_end, uint64_t seed, const char *const *columns,
                     int32_t ncolumns);
int64_t tpch_num_rows(int32_t table, double sf);
void tpch_generate_host(int32_t table, double sf, int64_t row_begin, int64_t row_end, uint64_t seed, const char *const *columns,
                        int32_t ncolumns, ArrowArray *out_array, ArrowSchema *out_schema);

}  // namespace b200

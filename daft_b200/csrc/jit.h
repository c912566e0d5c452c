// Host side of the JIT: plan parsing, typed code generation, NVRTC compile + cubin cache.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "common.h"

struct b200_comm;
struct b200_groupby_state;

namespace b200 {
namespace jit {

enum PtrKind {
    P_COL_DATA = 0, P_COL_VALID, P_COL_OFFSETS,
    P_OUT_DATA, P_OUT_VALID, P_OUT_LEN, P_OUT_OFFSETS,
    P_TAKE, P_SLOT_HASH, P_SLOT_ROW, P_ACC, P_OVERFLOW
};
struct PtrSlot { int kind; int index; };
enum IntKind { I_N = 0, I_MASK, I_RSTRIDE /* words between group-table records: 1 = one array per cell */ };

struct OutCol {
    std::string name;
    int32_t dtype;
    bool nullable;        // a validity bitmap is produced
    bool expr_nullable;   // the expression itself can yield nulls (ignoring a take's negative indices)
};

enum CellOp { C_ADD_F64 = 0, C_ADD_F32, C_ADD_I64, C_MIN, C_MAX };
enum AggKind { AGG_COUNT = 0, AGG_SUM, AGG_MIN, AGG_MAX, AGG_MEAN, AGG_ANY_VALUE, AGG_LIST, AGG_SET, AGG_CONCAT };
inline bool is_list_agg(int kind) { return kind == AGG_LIST || kind == AGG_SET || kind == AGG_CONCAT; }

struct AggOut {
    std::string name;
    int kind = 0, count_mode = 0;
    int32_t val_dtype = 0;   // dtype of the (widened) value the cell accumulates
    int32_t out_dtype = 0;
    int lane_cell = -1;      // index into lane cells, -1 if none
    int cnt_cell = -1;       // warp count cell with #valid (or #null for count-null), -1 => use nrows
    bool child_nullable = false;
    bool const_zero = false; // count(null) over a non-nullable child
    int proj_index = -1;     // AGG_ANY_VALUE: column of the any-value projection (child evaluated at the group's first row);
                             // AGG_LIST / SET / CONCAT: column of the list projection (child evaluated at every row)
    // Float min / max reproduce the reference's left fold `if a < b {a} else {b}` (inline_agg.rs:271-298 == compare_agg.rs
    // :75-95), which FORGETS everything before a NaN: the result is the min / max of the values after the group's last NaN,
    // and NaN when the last valid value is one.  lane_cell then folds the non-NaN values only, nan_cell = 1 + row of the
    // group's last NaN (0: none), last_cell = 1 + row of its last valid value.
    int nan_cell = -1, last_cell = -1;
    // any_value(ignore_nulls) over a nullable child: the cell holds the group's first row whose child is valid (all ones: none)
    int first_valid_cell = -1;
    int child_col = -1;      // AGG_LIST / SET / CONCAT over a bare column: the input column itself is gathered (no projection)
};

struct Compiled {
    cudaKernel_t fn = nullptr;
    cudaLibrary_t lib = nullptr;
    std::vector<PtrSlot> ptrs;
    std::vector<int> ints;
    // project
    std::vector<OutCol> outs;
    bool has_string_out = false;
    bool with_take = false;
    // agg
    int nkeys = 0;
    int G1 = 0, NLANE = 0, NWARP = 0;
    int astride = 1;  // 64-bit words per global-table record (code + cells, padded to 32-byte sectors)
    std::vector<int> lane_cell_ops;
    std::vector<AggOut> aggs;
    bool has_nan_fold = false;  // some aggregate carries nan_cell / last_cell (see AggOut)
    size_t smem_bytes = 0;
    int block_threads = 256;
    int stage_rows = 0, stage_bytes = 0;  // staged agg variant: bytes per shared-memory stage (0 = nothing to stage)
    std::vector<uint8_t> key_project_plan;
    std::vector<uint8_t> anyval_project_plan;  // PROJECT plan of the any_value children (evaluated at the first rows)
    std::vector<uint8_t> listagg_project_plan; // PROJECT plan of the list / set / concat children (evaluated at every row)
    int regs = 0;
    // (AGG, first-choice build) what the previous run of this plan found: the reference samples every query's first
    // morsel to choose its strategy (grouped_aggregate.rs:205-245); a plan that needed the global-table kernel last
    // time starts with it, sized from the group count it saw
    mutable int64_t last_groups = -1;
    std::string source_path;
};

// Codegen only (no CUDA needed): returns the CUDA source for the plan; fills `meta` (fn stays null).
// variant (AGG only): 0 = low-cardinality kernel (tier-1 shared-memory cells), 1 = global-table kernel,
// 2 = the low-cardinality kernel fed by per-warp TMA bulk copies (cp.async.bulk + mbarrier stages),
// 3 = the NaN-fold fix-up pass over a finished group table (re-folds float min / max over the rows after each group's last NaN)
std::string generate(int op, const uint8_t *plan, size_t plan_len, const std::vector<int32_t> &col_dtypes,
                     const std::vector<uint8_t> &col_nullable, bool with_take, Compiled &meta, int variant = 0);
// column indices the plan reads
// the expressions of a PROJECT plan: name, byte range of [str name, node] inside the plan, column index when the node is a bare column
struct ProjectEntry {
    std::string name;
    size_t begin = 0, end = 0;
    int32_t bare_col = -1;
};
std::vector<ProjectEntry> project_entries(const uint8_t *plan, size_t plan_len);
std::vector<int32_t> referenced_columns(int op, const uint8_t *plan, size_t plan_len);
// NVRTC -> cubin (cached on disk); works without a GPU.
std::vector<char> compile_to_cubin(const std::string &source, std::string *path_out);
// generate + compile + load; cached per (op, plan, dtypes, nullability, take)
const Compiled &get(int op, const uint8_t *plan, size_t plan_len, const Batch &input, bool with_take, int variant = 0);
void set_cache_dir(const std::string &dir);
void precompile(int op, const uint8_t *plan, size_t plan_len, const std::vector<int32_t> &col_dtypes,
                const std::vector<uint8_t> &col_nullable);

// launch helper: fills the parameter block from resolver callbacks
struct LaunchArgs {
    std::vector<void *> ptr;
    std::vector<long long> ival;
};
// `timed`: bracket the launch with CUDA events on the library stream when kernel timing is enabled
void launch(const Compiled &k, const LaunchArgs &a, int grid, int block, size_t smem, int pass, bool timed);
void enable_kernel_timing(bool on);
double last_timed_kernel_ms();

}  // namespace jit

// operators built on the JIT (ops_jit.cu)
// take_idx != null: evaluate at rows take_idx[i]; take_nonneg promises that no index is negative
Batch *op_project(const Batch &in, const uint8_t *plan, size_t len, const int64_t *take_idx, int64_t take_n, bool take_nonneg = false);
Batch *op_filter(const Batch &in, const uint8_t *plan, size_t len);
Batch *op_agg(const Batch &in, const uint8_t *plan, size_t len);
// also returns the groups' first rows; `coll`: all-reduce the group table across the communicator's ranks (ops_jit.cu)
Batch *op_agg_first_rows(const Batch &in, const uint8_t *plan, size_t len, Buf *first_rows, int64_t *ngroups, b200_comm *coll = nullptr);
Batch *op_agg_allreduce(b200_comm *comm, const Batch &in, const uint8_t *plan, size_t len);  // nullptr: use the two-stage plan
b200_groupby_state *groupby_create(const uint8_t *p1, size_t n1, const uint8_t *p2, size_t n2, const uint8_t *p3, size_t n3);
void groupby_free(b200_groupby_state *s);
void groupby_push(b200_groupby_state *s, const Batch &in);
Batch *groupby_finalize(b200_groupby_state *s);

}  // namespace b200

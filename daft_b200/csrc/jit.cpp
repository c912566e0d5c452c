// Plan parser + typed CUDA code generator + NVRTC driver for the fused pipeline kernels.
//
// What is generated is only the *expression* part of a kernel -- straight-line typed code for the
// Filter / Project / group-key / aggregate-input expressions of one plan node (the reference's
// tree-walking interpreter at src/daft-recordbatch/src/lib.rs:1280-1400 materialises one full column
// per AST node instead).  The kernel skeletons themselves are hand-written in jit_prelude.cuh.
#include "jit.h"

#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <fstream>
#include <functional>
#include <mutex>
#include <sstream>

namespace b200 {
namespace jit {

extern const char *const kPreludeXxh3_chunks[];  // xxh3.cuh        (embedded at build time, see embed.py)
extern const char *const kPreludeJit_chunks[];   // jit_prelude.cuh
static std::string join_chunks(const char *const *c) {
    std::string s;
    for (; *c; ++c) s += *c;
    return s;
}

// ------------------------------------------------------------------------------------------------
// plan reader
// ------------------------------------------------------------------------------------------------
struct Reader {
    const uint8_t *p;
    size_t n, pos = 0;
    void need(size_t k) const {
        if (pos + k > n) fail("plan: truncated");
    }
    uint8_t u8() { need(1); return p[pos++]; }
    uint32_t u32() { need(4); uint32_t v; memcpy(&v, p + pos, 4); pos += 4; return v; }
    int64_t i64() { need(8); int64_t v; memcpy(&v, p + pos, 8); pos += 8; return v; }
    std::string str() {
        uint32_t l = u32();
        need(l);
        std::string s(reinterpret_cast<const char *>(p + pos), l);
        pos += l;
        return s;
    }
};

enum Tag { T_COL = 0, T_LIT, T_BINARY, T_UNARY, T_CAST, T_IF_ELSE, T_IS_IN, T_BETWEEN, T_FILL_NULL, T_FUNCTION };
enum BinOp { B_ADD = 0, B_SUB, B_MUL, B_TRUE_DIV, B_FLOOR_DIV, B_MOD, B_LT, B_LE, B_EQ, B_NE, B_GE, B_GT, B_AND, B_OR,
             B_XOR, B_EQ_NULL_SAFE, B_SHL, B_SHR };
enum UnOp { U_NOT = 0, U_NEG, U_IS_NULL, U_NOT_NULL, U_ABS };
enum Fn { F_YEAR = 0, F_MONTH, F_DAY, F_UTF8_LENGTH, F_STARTSWITH, F_ENDSWITH, F_CONTAINS, F_LEFT, F_DOW, F_SQRT,
          F_IS_NAN, F_NOT_NAN, F_IS_INF, F_CEIL, F_FLOOR, F_SIGN, F_ROUND, F_MATCH_SEQ, F_HOUR, F_MINUTE, F_SECOND, F_TS_TRUNC, F_SUBSTR, F_RX_MATCH, F_RX_EXTRACT, F_RX_REPLACE_GROUP, F_TS_DATE, F_EXP, F_LN, F_LOG2, F_LOG10, F_SIN, F_COS, F_TAN, F_CBRT, F_LSTRIP, F_RSTRIP, F_RIGHT, F_FIND };

struct Node {
    int tag = 0, op = 0;
    int32_t dtype = B200_NULL;
    bool nullable = false;
    int col = -1;
    bool lit_null = false;
    int64_t lit_i = 0;     // ints / bool / date / timestamp / raw f64 bits
    std::string lit_s;
    std::vector<int> kids;
    std::string key;       // canonical structural key (for de-duplication)
};

struct Ast {
    std::vector<Node> nodes;
    const std::vector<int32_t> *col_dtypes;
    const std::vector<uint8_t> *col_nullable;  // per column: bit 0 = has nulls, bits 1-3 = fixed string length hint
    int list_col_refs = 0;  // references to list columns: only the bare child of a concat aggregate may be one
    bool nullable(int c) const { return ((*col_nullable)[c] & 1) != 0; }
    // String columns whose bytes divide evenly over the rows at 1..4 B per value (flags, codes): the aggregate kernel is
    // specialised for that length and fetches the value speculatively; every row still verifies it against the offsets.
    int fixed_len(int c) const { return ((*col_nullable)[c] >> 1) & 7; }
};

static bool is_numeric(int32_t dt) { return is_integer(dt) || is_float(dt); }

// Casts go through arrow-rs `cast` with its default CastOptions { safe: true } (src/daft-core/src/array/ops/cast.rs:194-196;
// arrow-cast 57.3.0, Cargo.lock:298): a value the target integer type cannot hold becomes NULL.  Integer sources use
// num-traits 0.2.19 NumCast (exact range check), float sources succeed iff the value truncated toward zero is in range
// (NaN fails).  Date32 / Timestamp cast as their physical Int32 / Int64.
static bool int_like(int32_t dt) { return is_integer(dt) || dt == B200_DATE32 || dt == B200_TIMESTAMP; }
static bool int_like_signed(int32_t dt) { return is_signed_int(dt) || dt == B200_DATE32 || dt == B200_TIMESTAMP; }
static bool cast_can_fail(int32_t src, int32_t dst) {
    if (is_string(src)) return !is_string(dst);  // parsing
    if (!int_like(dst)) return false;
    if (is_float(src)) return true;
    if (!int_like(src)) return false;  // bool -> 0 / 1
    const int bs = 8 * dtype_width(src), bd = 8 * dtype_width(dst);
    if (int_like_signed(src)) return !int_like_signed(dst) || bs > bd;
    return int_like_signed(dst) ? bs >= bd : bs > bd;
}

static int parse_node(Reader &r, Ast &a) {
    Node nd;
    nd.tag = r.u8();
    switch (nd.tag) {
        case T_COL: {
            nd.col = (int)r.u32();
            if (nd.col < 0 || nd.col >= (int)a.col_dtypes->size()) fail("plan: column index out of range");
            nd.dtype = (*a.col_dtypes)[nd.col];
            if (nd.dtype == B200_LIST) a.list_col_refs++;  // legal only as the bare child of a concat aggregate (gen_agg checks)
            nd.nullable = a.nullable(nd.col) || nd.dtype == B200_NULL;
            nd.key = "c" + std::to_string(nd.col);
            break;
        }
        case T_LIT: {
            nd.dtype = r.u8();
            nd.lit_null = r.u8() != 0;
            if (is_string(nd.dtype)) nd.lit_s = r.str();
            else nd.lit_i = r.i64();
            nd.nullable = nd.lit_null || nd.dtype == B200_NULL;
            nd.key = "l" + std::to_string(nd.dtype) + (nd.lit_null ? "n" : "") + ":" +
                     (is_string(nd.dtype) ? nd.lit_s : std::to_string(nd.lit_i));
            break;
        }
        case T_BINARY: {
            nd.op = r.u8();
            int l = parse_node(r, a), rr = parse_node(r, a);
            nd.kids = {l, rr};
            const Node &L = a.nodes[l], &R = a.nodes[rr];
            if (L.dtype != R.dtype && L.dtype != B200_NULL && R.dtype != B200_NULL)
                fail(std::string("plan: binary op operands must have identical dtypes (host inserts casts): ") +
                     dtype_name(L.dtype) + " vs " + dtype_name(R.dtype));
            int32_t t = L.dtype != B200_NULL ? L.dtype : R.dtype;
            nd.nullable = L.nullable || R.nullable;
            switch (nd.op) {
                case B_ADD: case B_SUB: case B_MUL: case B_FLOOR_DIV: case B_MOD:
                    if (!is_numeric(t)) fail(std::string("plan: arithmetic on non-numeric dtype ") + dtype_name(t));
                    nd.dtype = t;
                    if ((nd.op == B_FLOOR_DIV || nd.op == B_MOD) && is_integer(t)) nd.nullable = true;  // x / 0 -> null
                    break;
                case B_TRUE_DIV:
                    if (t != B200_F64) fail("plan: true_div requires float64 operands");
                    nd.dtype = B200_F64;
                    break;
                case B_LT: case B_LE: case B_EQ: case B_NE: case B_GE: case B_GT:
                    nd.dtype = B200_BOOL;
                    break;
                case B_EQ_NULL_SAFE:
                    nd.dtype = B200_BOOL;
                    nd.nullable = false;
                    break;
                case B_AND: case B_OR: case B_XOR:
                    // booleans: Kleene and / or, xor; integers: bitwise, null if either side is null
                    // (src/daft-core/src/array/ops/bitwise.rs; tests/recordbatch/test_bitwise.py)
                    if (t != B200_BOOL && t != B200_NULL && !is_integer(t)) fail("plan: and/or/xor need boolean or integer operands");
                    nd.dtype = is_integer(t) ? t : B200_BOOL;
                    break;
                case B_SHL: case B_SHR:
                    if (!is_integer(t)) fail("plan: shift needs integer operands");
                    nd.dtype = t;
                    break;
                default: fail("plan: unknown binary op");
            }
            nd.key = "b" + std::to_string(nd.op) + "(" + L.key + "," + R.key + ")";
            break;
        }
        case T_UNARY: {
            nd.op = r.u8();
            int c = parse_node(r, a);
            nd.kids = {c};
            const Node &C = a.nodes[c];
            switch (nd.op) {
                case U_NOT: nd.dtype = B200_BOOL; nd.nullable = C.nullable; break;
                case U_NEG: case U_ABS: nd.dtype = C.dtype; nd.nullable = C.nullable; break;
                case U_IS_NULL: case U_NOT_NULL: nd.dtype = B200_BOOL; nd.nullable = false; break;
                default: fail("plan: unknown unary op");
            }
            nd.key = "u" + std::to_string(nd.op) + "(" + C.key + ")";
            break;
        }
        case T_CAST: {
            nd.dtype = r.u8();
            int c = parse_node(r, a);
            nd.kids = {c};
            nd.nullable = a.nodes[c].nullable || cast_can_fail(a.nodes[c].dtype, nd.dtype);
            nd.key = "k" + std::to_string(nd.dtype) + "(" + a.nodes[c].key + ")";
            break;
        }
        case T_IF_ELSE: {
            int p = parse_node(r, a), t = parse_node(r, a), e = parse_node(r, a);
            nd.kids = {p, t, e};
            const Node &Tn = a.nodes[t], &En = a.nodes[e];
            if (Tn.dtype != En.dtype && Tn.dtype != B200_NULL && En.dtype != B200_NULL) fail("plan: if_else branch dtypes differ");
            nd.dtype = Tn.dtype != B200_NULL ? Tn.dtype : En.dtype;
            nd.nullable = a.nodes[p].nullable || Tn.nullable || En.nullable;
            nd.key = "i(" + a.nodes[p].key + "," + Tn.key + "," + En.key + ")";
            break;
        }
        case T_IS_IN: {
            int c = parse_node(r, a);
            nd.kids = {c};
            uint32_t n = r.u32();
            nd.key = "n(" + a.nodes[c].key;
            for (uint32_t i = 0; i < n; i++) {
                int k = parse_node(r, a);
                if (a.nodes[k].tag != T_LIT) fail("plan: is_in items must be literals");
                nd.kids.push_back(k);
                nd.key += "," + a.nodes[k].key;
            }
            nd.key += ")";
            nd.dtype = B200_BOOL;
            nd.nullable = a.nodes[c].nullable;
            break;
        }
        case T_BETWEEN: {
            int c = parse_node(r, a), lo = parse_node(r, a), hi = parse_node(r, a);
            nd.kids = {c, lo, hi};
            nd.dtype = B200_BOOL;
            nd.nullable = a.nodes[c].nullable || a.nodes[lo].nullable || a.nodes[hi].nullable;
            nd.key = "w(" + a.nodes[c].key + "," + a.nodes[lo].key + "," + a.nodes[hi].key + ")";
            break;
        }
        case T_FILL_NULL: {
            int c = parse_node(r, a), f = parse_node(r, a);
            nd.kids = {c, f};
            nd.dtype = a.nodes[c].dtype != B200_NULL ? a.nodes[c].dtype : a.nodes[f].dtype;
            nd.nullable = a.nodes[c].nullable && a.nodes[f].nullable;
            nd.key = "f(" + a.nodes[c].key + "," + a.nodes[f].key + ")";
            break;
        }
        case T_FUNCTION: {
            nd.op = r.u8();
            int na = r.u8();
            nd.key = "F" + std::to_string(nd.op) + "(";
            for (int i = 0; i < na; i++) {
                int k = parse_node(r, a);
                nd.kids.push_back(k);
                nd.nullable = nd.nullable || a.nodes[k].nullable;
                nd.key += a.nodes[k].key + ",";
            }
            nd.key += ")";
            switch (nd.op) {
                case F_YEAR: nd.dtype = B200_I32; break;
                case F_MONTH: case F_DAY: case F_DOW: nd.dtype = B200_U32; break;
                case F_UTF8_LENGTH: nd.dtype = B200_U64; break;
                case F_STARTSWITH: case F_ENDSWITH: case F_CONTAINS: nd.dtype = B200_BOOL; break;
                case F_LEFT: nd.dtype = B200_LARGE_UTF8; break;
                case F_SQRT: {  // operand cast by the host to its floating representation (Float32 for 8/16-bit integers and Float32,
                                // Float64 otherwise: DataType::to_floating_representation, src/daft-schema/src/dtype.rs:756-772)
                    if (nd.kids.size() != 1 || !is_float(a.nodes[nd.kids[0]].dtype)) fail("plan: sqrt needs one float operand");
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    break;
                }
                case F_IS_NAN: case F_NOT_NAN: case F_IS_INF: {  // src/daft-core/src/array/ops/float.rs:13-110: nulls stay null
                    if (nd.kids.size() != 1 || !is_float(a.nodes[nd.kids[0]].dtype)) fail("plan: is_nan / not_nan / is_inf need one float operand");
                    nd.dtype = B200_BOOL;
                    break;
                }
                case F_CEIL: case F_FLOOR: {  // src/daft-functions/src/numeric/ceil.rs:25-38, floor.rs (integers are returned unchanged by the host)
                    if (nd.kids.size() != 1 || !is_float(a.nodes[nd.kids[0]].dtype)) fail("plan: ceil / floor need one float operand");
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    break;
                }
                case F_SIGN: {  // src/daft-functions/src/numeric/sign.rs:26-42 -> array/ops/sign.rs: signum, 0 / 1 for unsigned integers
                    if (nd.kids.size() != 1 || !is_numeric(a.nodes[nd.kids[0]].dtype)) fail("plan: sign needs one numeric operand");
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    break;
                }
                case F_ROUND: {  // src/daft-core/src/array/ops/round.rs:14-35: (v * 10^d).round() / 10^d, half away from zero
                    if (nd.kids.size() != 2 || !is_float(a.nodes[nd.kids[0]].dtype)) fail("plan: round needs a float operand and a decimals literal");
                    const Node &d = a.nodes[nd.kids[1]];
                    if (d.tag != T_LIT || d.dtype != B200_I32 || d.lit_null) fail("plan: round decimals must be an Int32 literal");
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    nd.nullable = a.nodes[nd.kids[0]].nullable;
                    break;
                }
                case F_MATCH_SEQ: {  // utf8 match / like restricted to literal segments in order (flags: 1 = anchored start, 2 = anchored end)
                    if (nd.kids.size() < 2 || !is_string(a.nodes[nd.kids[0]].dtype)) fail("plan: match needs a string operand");
                    for (size_t i = 1; i < nd.kids.size(); i++) {
                        const Node &l = a.nodes[nd.kids[i]];
                        if (l.tag != T_LIT || l.lit_null || (i == 1 ? l.dtype != B200_I32 : !is_string(l.dtype)))
                            fail("plan: match takes an Int32 flags literal followed by string literals");
                    }
                    nd.dtype = B200_BOOL;
                    nd.nullable = a.nodes[nd.kids[0]].nullable;
                    break;
                }
                case F_HOUR: case F_MINUTE: case F_SECOND: {  // dt.hour() / minute() / second() of a Timestamp(us) (src/daft-functions-temporal): UInt32
                    if (nd.kids.size() != 1 || a.nodes[nd.kids[0]].dtype != B200_TIMESTAMP) fail("plan: hour / minute / second need one timestamp operand");
                    nd.dtype = B200_U32;
                    break;
                }
                case F_TS_TRUNC: {  // dt.truncate(interval) relative to the epoch; the interval arrives as an Int64 literal in microseconds
                    if (nd.kids.size() != 2 || a.nodes[nd.kids[0]].dtype != B200_TIMESTAMP) fail("plan: truncate needs a timestamp operand and an interval literal");
                    const Node &u = a.nodes[nd.kids[1]];
                    if (u.tag != T_LIT || u.dtype != B200_I64 || u.lit_null || u.lit_i <= 0) fail("plan: the truncate interval must be a positive Int64 literal (microseconds)");
                    nd.dtype = B200_TIMESTAMP;
                    nd.nullable = a.nodes[nd.kids[0]].nullable;
                    break;
                }
                case F_SUBSTR: {  // utf8 substr(start, length) in characters (src/daft-functions-utf8/src/substr.rs:269-294): null when the value has
                                  // no character at `start` or the length is 0; a length literal of -1 stands for "to the end"
                    if (nd.kids.size() != 3 || !is_string(a.nodes[nd.kids[0]].dtype)) fail("plan: substr needs a string operand, a start and a length");
                    for (int i = 1; i <= 2; i++) {
                        const Node &l = a.nodes[nd.kids[i]];
                        if (l.tag != T_LIT || l.dtype != B200_I64 || l.lit_null || l.lit_i < (i == 1 ? 0 : -1)) fail("plan: substr start / length must be Int64 literals (start >= 0, length >= -1)");
                    }
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    nd.nullable = true;
                    break;
                }
                case F_EXP: case F_LN: case F_LOG2: case F_LOG10: case F_SIN: case F_COS: case F_TAN: case F_CBRT: {
                    // unary float math (src/daft-functions/src/numeric/{exp,log,trigonometry,cbrt}.rs) in the operand's floating representation,
                    // like sqrt; CUDA's libdevice results are within 1-2 ulp of the host's libm: a floating-point tolerance applies
                    if (nd.kids.size() != 1 || !is_float(a.nodes[nd.kids[0]].dtype)) fail("plan: exp / ln / log2 / log10 / sin / cos / tan / cbrt need one float operand");
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    break;
                }
                case F_LSTRIP: case F_RSTRIP: {
                    if (nd.kids.size() != 1 || !is_string(a.nodes[nd.kids[0]].dtype)) fail("plan: lstrip / rstrip need one string operand");
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    break;
                }
                case F_RIGHT: {
                    if (nd.kids.size() != 2 || !is_string(a.nodes[nd.kids[0]].dtype) || a.nodes[nd.kids[1]].dtype != B200_I64) fail("plan: right needs a string and an Int64 count");
                    nd.dtype = a.nodes[nd.kids[0]].dtype;
                    break;
                }
                case F_FIND: {  // str::find: BYTE index of the first occurrence, -1 when absent (src/daft-functions-utf8/src/find.rs:85)
                    if (nd.kids.size() != 2 || !is_string(a.nodes[nd.kids[0]].dtype) || !is_string(a.nodes[nd.kids[1]].dtype)) fail("plan: find needs two string operands");
                    nd.dtype = B200_I64;
                    break;
                }
                case F_TS_DATE: {  // dt.date() of a Timestamp(us): days since the epoch, floor semantics
                    if (nd.kids.size() != 1 || a.nodes[nd.kids[0]].dtype != B200_TIMESTAMP) fail("plan: date() needs one timestamp operand");
                    nd.dtype = B200_DATE32;
                    break;
                }
                case F_RX_MATCH: case F_RX_EXTRACT: case F_RX_REPLACE_GROUP: {
                    // regular expressions (daft_b200/regex.py compiles the pattern on the host): operand, the VM program as a binary literal,
                    // the pattern text (kept for diagnostics and the oracle).  match -> Boolean; extract -> the selected group or null
                    // (src/daft-functions-utf8/src/extract.rs); replace_group -> the group when the (fully anchored) pattern matches,
                    // the value unchanged when it does not (regexp_replace(x, '^...(g)...$', '\\1'), replace.rs)
                    if (nd.kids.size() != 3 || !is_string(a.nodes[nd.kids[0]].dtype)) fail("plan: a regex function needs a string operand, a program and a pattern");
                    const Node &prog = a.nodes[nd.kids[1]];
                    if (prog.tag != T_LIT || !is_string(prog.dtype) || prog.lit_null || prog.lit_s.size() < 5) fail("plan: the regex program must be a binary literal");
                    const size_t ninst = (unsigned char)prog.lit_s[0];
                    if (ninst == 0 || ninst > 64 || prog.lit_s.size() < 1 + 4 * ninst) fail("plan: malformed regex program");
                    for (size_t i = 0; i < ninst; i++) {   // every jump stays inside the program, every class table inside the literal
                        const unsigned char *in = (const unsigned char *)prog.lit_s.data() + 1 + 4 * i;
                        if (in[0] > 9) fail("plan: unknown regex instruction");
                        if ((in[0] == 4 && (in[1] >= ninst || in[2] >= ninst)) || (in[0] == 5 && in[1] >= ninst)) fail("plan: regex jump out of range");
                        if (in[0] == 3) {
                            const size_t off = (size_t)in[1] | ((size_t)in[3] << 8);
                            if (off >= prog.lit_s.size() || off + 1 + 2 * (size_t)(unsigned char)prog.lit_s[off] > prog.lit_s.size()) fail("plan: regex class table out of range");
                        }
                        if (in[0] != 0 && in[0] != 4 && in[0] != 5 && i + 1 >= ninst) fail("plan: regex program runs off its end");
                    }
                    nd.dtype = nd.op == F_RX_MATCH ? B200_BOOL : a.nodes[nd.kids[0]].dtype;
                    nd.nullable = nd.op == F_RX_EXTRACT ? true : a.nodes[nd.kids[0]].nullable;
                    break;
                }
                default: fail("plan: unknown function");
            }
            break;
        }
        default: fail("plan: unknown node tag " + std::to_string(nd.tag));
    }
    a.nodes.push_back(std::move(nd));
    return (int)a.nodes.size() - 1;
}

// ------------------------------------------------------------------------------------------------
// code generation
// ------------------------------------------------------------------------------------------------
static const char *ctype(int32_t dt) {
    switch (dt) {
        case B200_NULL: case B200_BOOL: return "bool";
        case B200_I8: return "int8_t";
        case B200_I16: return "int16_t";
        case B200_I32: case B200_DATE32: return "int32_t";
        case B200_I64: case B200_TIMESTAMP: return "int64_t";
        case B200_U8: return "uint8_t";
        case B200_U16: return "uint16_t";
        case B200_U32: return "uint32_t";
        case B200_U64: return "uint64_t";
        case B200_F32: return "float";
        case B200_F64: return "double";
        case B200_LARGE_UTF8: case B200_LARGE_BINARY: return "Str";
    }
    return "bool";
}
static const char *utype(int32_t dt) {  // same-width unsigned type for wrapping integer arithmetic
    switch (dtype_width(dt)) {
        case 1: return "uint8_t";
        case 2: return "uint16_t";
        case 4: return "uint32_t";
        default: return "uint64_t";
    }
}
static std::string zero_of(int32_t dt) {
    if (is_string(dt)) return "Str{(const uint8_t*)0, 0}";
    if (dt == B200_BOOL || dt == B200_NULL) return "false";
    return std::string("(") + ctype(dt) + ")0";
}

struct CG {
    const Ast &ast;
    Compiled &meta;
    std::ostringstream globals;
    int nlit = 0;
    std::map<std::pair<int, int>, int> ptr_index;
    std::map<int, int> int_index;

    int ptr_slot(int kind, int index) {
        auto key = std::make_pair(kind, index);
        auto it = ptr_index.find(key);
        if (it != ptr_index.end()) return it->second;
        int s = (int)meta.ptrs.size();
        meta.ptrs.push_back({kind, index});
        ptr_index[key] = s;
        return s;
    }
    int int_slot(int kind) {
        auto it = int_index.find(kind);
        if (it != int_index.end()) return it->second;
        int s = (int)meta.ints.size();
        meta.ints.push_back(kind);
        int_index[kind] = s;
        return s;
    }
    std::string P(int kind, int index) { return "p.ptr[" + std::to_string(ptr_slot(kind, index)) + "]"; }
};

// How a column is fetched: directly from global memory at `row` (guarded by `rowok`), or from a
// pre-loaded Raw struct `r` (the aggregate kernel issues all loads of U rows before any compute).
enum Access { DIRECT, RAW };

struct ExprGen {
    CG &cg;
    Access access;
    std::ostream &os;
    std::string pfx;                 // unique prefix for the SSA names of this function body
    std::map<std::string, int> done; // canonical key -> node id already emitted (CSE)

    std::string v(int id) { return pfx + "v" + std::to_string(id); }
    std::string k(int id) { return pfx + "k" + std::to_string(id); }

    int emit(int id) {
        const Node &n = cg.ast.nodes[id];
        auto it = done.find(n.key);
        if (it != done.end()) return it->second;
        std::vector<int> kid;
        for (int c : n.kids) kid.push_back(emit(c));
        const std::string V = v(id), K = k(id), T = ctype(n.dtype);
        switch (n.tag) {
            case T_COL: {
                const int c = n.col;
                const bool nullable = cg.ast.nullable(c);
                if (access == RAW) {
                    if (n.dtype == B200_NULL) {
                        os << "    const bool " << V << " = false; const bool " << K << " = false;\n";
                    } else if (is_string(n.dtype)) {
                        os << "    const Str " << V << "{(const uint8_t*)" << cg.P(P_COL_DATA, c) << " + r.o" << c << "a, r.o" << c
                           << "b - r.o" << c << "a};\n";
                        os << "    const bool " << K << " = " << (nullable ? "r.k" + std::to_string(c) : std::string("true")) << ";\n";
                    } else {
                        os << "    const " << T << " " << V << " = r.c" << c << "; const bool " << K << " = "
                           << (nullable ? "r.k" + std::to_string(c) : std::string("true")) << ";\n";
                    }
                } else {
                    if (n.dtype == B200_NULL) {
                        os << "    const bool " << V << " = false; const bool " << K << " = false;\n";
                    } else {
                        if (nullable)
                            os << "    const bool " << K << " = rowok && bit_at((const uint8_t*)" << cg.P(P_COL_VALID, c) << ", row);\n";
                        else
                            os << "    const bool " << K << " = rowok;\n";
                        if (is_string(n.dtype)) {
                            os << "    Str " << V << "{(const uint8_t*)0, 0};\n";
                            os << "    if (rowok) { const int64_t* of_ = (const int64_t*)" << cg.P(P_COL_OFFSETS, c) << "; const int64_t a_ = of_[row]; "
                               << V << ".p = (const uint8_t*)" << cg.P(P_COL_DATA, c) << " + a_; " << V << ".len = of_[row + 1] - a_; }\n";
                        } else if (n.dtype == B200_BOOL) {
                            os << "    const bool " << V << " = rowok && bit_at((const uint8_t*)" << cg.P(P_COL_DATA, c) << ", row);\n";
                        } else {
                            os << "    const " << T << " " << V << " = rowok ? ((const " << T << "*)" << cg.P(P_COL_DATA, c) << ")[row] : (" << T << ")0;\n";
                        }
                    }
                }
                break;
            }
            case T_LIT: {
                if (n.dtype == B200_NULL || n.lit_null) {
                    os << "    const " << T << " " << V << " = " << zero_of(n.dtype) << "; const bool " << K << " = false;\n";
                } else if (is_string(n.dtype)) {
                    int li = cg.nlit++;
                    cg.globals << "static __device__ const uint8_t b200_lit" << li << "[" << (n.lit_s.size() ? n.lit_s.size() : 1) << "] = {";
                    for (size_t i = 0; i < n.lit_s.size(); i++) cg.globals << (i ? "," : "") << (unsigned)(uint8_t)n.lit_s[i];
                    if (n.lit_s.empty()) cg.globals << "0";
                    cg.globals << "};\n";
                    os << "    const Str " << V << "{b200_lit" << li << ", " << n.lit_s.size() << "}; const bool " << K << " = true;\n";
                } else if (n.dtype == B200_F64) {
                    os << "    const double " << V << " = __longlong_as_double(" << n.lit_i << "LL); const bool " << K << " = true;\n";
                } else if (n.dtype == B200_F32) {
                    double d;
                    memcpy(&d, &n.lit_i, 8);
                    float f = (float)d;
                    int32_t bits;
                    memcpy(&bits, &f, 4);
                    os << "    const float " << V << " = __int_as_float(" << bits << "); const bool " << K << " = true;\n";
                } else if (n.dtype == B200_BOOL) {
                    os << "    const bool " << V << " = " << (n.lit_i ? "true" : "false") << "; const bool " << K << " = true;\n";
                } else if (n.dtype == B200_U64) {
                    os << "    const uint64_t " << V << " = " << (uint64_t)n.lit_i << "ULL; const bool " << K << " = true;\n";
                } else {
                    os << "    const " << T << " " << V << " = (" << T << ")(" << n.lit_i << "LL); const bool " << K << " = true;\n";
                }
                break;
            }
            case T_BINARY: {
                const Node &L = cg.ast.nodes[n.kids[0]];
                const int32_t t = L.dtype != B200_NULL ? L.dtype : cg.ast.nodes[n.kids[1]].dtype;
                const std::string a = v(kid[0]), b = v(kid[1]), ka = k(kid[0]), kb = k(kid[1]);
                // Floats compare by IEEE totalOrder (f_key).  Against a finite non-zero literal the IEEE comparison differs from it
                // only for a NaN operand, so the hot predicates (`x >= 0.05`) keep the plain compare plus a NaN-sign fix-up and need
                // no key registers.
                auto plain_lit = [&](int kidx) {
                    const Node &q = cg.ast.nodes[n.kids[kidx]];
                    if (q.tag != T_LIT || q.lit_null || !is_float(q.dtype)) return false;
                    double d;
                    memcpy(&d, &q.lit_i, 8);
                    return d == d && (q.dtype == B200_F32 ? (float)d != 0.0f : d != 0.0);
                };
                auto cmp = [&](const char *op) {
                    if (is_string(t)) os << "    const bool " << V << " = str_cmp(" << a << ", " << b << ") " << op << " 0;\n";
                    else if (is_float(t) && (plain_lit(1) || plain_lit(0))) {
                        const bool less = op[0] == '<';                       // operator as written: a op b
                        const bool var_left = plain_lit(1);                   // the non-literal side
                        const std::string &x = var_left ? a : b;
                        // a < c holds for a = -NaN; c < b holds for b = +NaN (and symmetrically for > / >=)
                        const char *fix = (less == var_left) ? "f_negnan(" : "f_posnan(";
                        os << "    const bool " << V << " = (" << a << " " << op << " " << b << ") || " << fix << x << ");\n";
                    }
                    else if (is_float(t)) os << "    const bool " << V << " = f_key(" << a << ") " << op << " f_key(" << b << ");\n";  // IEEE totalOrder
                    else os << "    const bool " << V << " = " << a << " " << op << " " << b << ";\n";
                    os << "    const bool " << K << " = " << ka << " && " << kb << ";\n";
                };
                switch (n.op) {
                    case B_ADD: case B_SUB: case B_MUL: {
                        const char *op = n.op == B_ADD ? "+" : (n.op == B_SUB ? "-" : "*");
                        if (is_float(t)) os << "    const " << T << " " << V << " = " << a << " " << op << " " << b << ";\n";
                        else
                            os << "    const " << T << " " << V << " = (" << T << ")((" << utype(t) << ")" << a << " " << op << " (" << utype(t) << ")" << b << ");\n";
                        os << "    const bool " << K << " = " << ka << " && " << kb << ";\n";
                        break;
                    }
                    case B_TRUE_DIV:
                        os << "    const double " << V << " = " << a << " / " << b << "; const bool " << K << " = " << ka << " && " << kb << ";\n";
                        break;
                    case B_FLOOR_DIV: case B_MOD: {
                        const bool div = n.op == B_FLOOR_DIV;
                        if (is_float(t)) {
                            if (div) os << "    const " << T << " " << V << " = floor(" << a << " / " << b << ");\n";
                            else os << "    const " << T << " " << V << " = fmod" << (t == B200_F32 ? "f" : "") << "(" << a << ", " << b << ");\n";  // Rust `%` on floats
                            os << "    const bool " << K << " = " << ka << " && " << kb << ";\n";
                        } else {
                            os << "    const bool " << K << " = " << ka << " && " << kb << " && (" << b << " != 0);\n";
                            if (is_signed_int(t))
                                os << "    const " << T << " " << V << " = " << K << " ? " << (div ? "trunc_div_s" : "trunc_rem_s") << "<" << T << ">(" << a << ", " << b << ") : (" << T << ")0;\n";
                            else
                                os << "    const " << T << " " << V << " = " << K << " ? (" << T << ")(" << a << (div ? " / " : " % ") << b << ") : (" << T << ")0;\n";
                        }
                        break;
                    }
                    case B_LT: cmp("<"); break;
                    case B_LE: cmp("<="); break;
                    case B_GE: cmp(">="); break;
                    case B_GT: cmp(">"); break;
                    case B_EQ: case B_NE: case B_EQ_NULL_SAFE: {
                        std::string eq;
                        if (is_string(t)) eq = "str_eq(" + a + ", " + b + ")";
                        else if (is_float(t) && (plain_lit(0) || plain_lit(1))) eq = "(" + a + " == " + b + ")";  // unique bit pattern
                        else if (is_float(t)) eq = "(f_key(" + a + ") == f_key(" + b + "))";  // bit equality: NaN == NaN, -0.0 != 0.0
                        else eq = "(" + a + " == " + b + ")";
                        if (n.op == B_EQ_NULL_SAFE) {
                            os << "    const bool " << V << " = (" << ka << " && " << kb << ") ? " << eq << " : (" << ka << " == " << kb << "); const bool " << K << " = true;\n";
                        } else {
                            os << "    const bool " << V << " = " << (n.op == B_NE ? "!" : "") << eq << "; const bool " << K << " = " << ka << " && " << kb << ";\n";
                        }
                        break;
                    }
                    case B_AND: case B_OR: case B_XOR:
                        if (is_integer(t)) {  // bitwise; null if either side is null
                            os << "    const " << T << " " << V << " = (" << T << ")(" << a << (n.op == B_AND ? " & " : (n.op == B_OR ? " | " : " ^ ")) << b
                               << "); const bool " << K << " = " << ka << " && " << kb << ";\n";
                        } else if (n.op == B_AND) {  // Kleene (src/daft-core/src/array/ops/comparison.rs:725-786)
                            os << "    const bool " << V << " = (" << ka << " && " << a << ") && (" << kb << " && " << b << ");\n";
                            os << "    const bool " << K << " = (" << ka << " && " << kb << ") || (" << ka << " && !" << a << ") || (" << kb << " && !" << b << ");\n";
                        } else if (n.op == B_OR) {   // comparison.rs:788-841
                            os << "    const bool " << V << " = (" << ka << " && " << a << ") || (" << kb << " && " << b << ");\n";
                            os << "    const bool " << K << " = (" << ka << " && " << kb << ") || (" << ka << " && " << a << ") || (" << kb << " && " << b << ");\n";
                        } else {
                            os << "    const bool " << V << " = " << a << " != " << b << "; const bool " << K << " = " << ka << " && " << kb << ";\n";
                        }
                        break;
                    case B_SHL: case B_SHR:
                        os << "    const " << T << " " << V << " = (" << T << ")(" << a << (n.op == B_SHL ? " << " : " >> ") << b << "); const bool " << K << " = " << ka << " && " << kb << ";\n";
                        break;
                }
                break;
            }
            case T_UNARY: {
                const std::string a = v(kid[0]), ka = k(kid[0]);
                switch (n.op) {
                    case U_NOT: os << "    const bool " << V << " = !" << a << "; const bool " << K << " = " << ka << ";\n"; break;
                    case U_NEG:
                        if (is_float(n.dtype)) os << "    const " << T << " " << V << " = -" << a << ";";
                        else os << "    const " << T << " " << V << " = (" << T << ")(0 - (" << utype(n.dtype) << ")" << a << ");";
                        os << " const bool " << K << " = " << ka << ";\n";
                        break;
                    case U_ABS:
                        if (is_float(n.dtype)) os << "    const " << T << " " << V << " = fabs(" << a << ");";
                        else if (is_signed_int(n.dtype)) os << "    const " << T << " " << V << " = " << a << " < 0 ? (" << T << ")(0 - (" << utype(n.dtype) << ")" << a << ") : " << a << ";";
                        else os << "    const " << T << " " << V << " = " << a << ";";
                        os << " const bool " << K << " = " << ka << ";\n";
                        break;
                    case U_IS_NULL: os << "    const bool " << V << " = !" << ka << "; const bool " << K << " = true;\n"; break;
                    case U_NOT_NULL: os << "    const bool " << V << " = " << ka << "; const bool " << K << " = true;\n"; break;
                }
                break;
            }
            case T_CAST: {
                const Node &C = cg.ast.nodes[n.kids[0]];
                const std::string a = v(kid[0]), ka = k(kid[0]);
                if (C.dtype == n.dtype) {
                    os << "    const " << T << " " << V << " = " << a << "; const bool " << K << " = " << ka << ";\n";
                } else if (C.dtype == B200_NULL) {
                    os << "    const " << T << " " << V << " = " << zero_of(n.dtype) << "; const bool " << K << " = false;\n";
                } else if (is_string(C.dtype) && is_string(n.dtype)) {
                    os << "    const Str " << V << " = " << a << "; const bool " << K << " = " << ka << ";\n";
                } else if (is_string(C.dtype) && is_integer(n.dtype)) {
                    // trimmed, then [+-]?[0-9]+ in range of the target, else null (str_parse_int in the prelude)
                    const std::string tv = pfx + "t" + std::to_string(id);
                    os << "    unsigned long long " << tv << " = 0; const bool " << K << " = " << ka << " && str_parse_int(" << a << ", "
                       << 8 * dtype_width(n.dtype) << ", " << (is_signed_int(n.dtype) ? "true" : "false") << ", " << tv << ");\n";
                    os << "    const " << T << " " << V << " = " << K << " ? (" << T << ")" << tv << " : (" << T << ")0;\n";
                } else if (is_string(C.dtype) || is_string(n.dtype)) {
                    fail("plan: only Utf8 -> integer among the casts between strings and other types is built on the device");
                } else if (n.dtype == B200_BOOL) {
                    os << "    const bool " << V << " = " << a << " != 0; const bool " << K << " = " << ka << ";\n";
                } else if (!cast_can_fail(C.dtype, n.dtype)) {
                    os << "    const " << T << " " << V << " = (" << T << ")" << a << "; const bool " << K << " = " << ka << ";\n";
                } else if (is_float(C.dtype)) {
                    // out-of-range / NaN -> null: the value truncated toward zero must lie in [MIN, MAX] of the target
                    const int bits = 8 * dtype_width(n.dtype);
                    const bool sg = int_like_signed(n.dtype);
                    char lo[64], hi[64];
                    snprintf(lo, sizeof lo, "%.1f", sg ? -std::ldexp(1.0, bits - 1) : 0.0);
                    snprintf(hi, sizeof hi, "%.1f", std::ldexp(1.0, sg ? bits - 1 : bits));
                    const std::string tv = pfx + "t" + std::to_string(id);
                    os << "    const double " << tv << " = trunc((double)" << a << ");\n";
                    os << "    const bool " << K << " = " << ka << " && " << tv << " >= " << lo << " && " << tv << " < " << hi << ";\n";
                    os << "    const " << T << " " << V << " = " << K << " ? (" << T << ")" << a << " : (" << T << ")0;\n";
                } else {
                    // narrowing / sign-changing integer cast: null unless the value survives the round trip with its sign
                    const std::string S = ctype(C.dtype), tv = pfx + "t" + std::to_string(id);
                    os << "    const " << T << " " << tv << " = (" << T << ")" << a << ";\n";
                    os << "    const bool " << K << " = " << ka << " && ((" << S << ")" << tv << " == " << a << ") && ("
                       << (int_like_signed(n.dtype) ? "(" + tv + " < 0)" : std::string("false")) << " == "
                       << (int_like_signed(C.dtype) ? "(" + a + " < 0)" : std::string("false")) << ");\n";
                    os << "    const " << T << " " << V << " = " << K << " ? " << tv << " : (" << T << ")0;\n";
                }
                break;
            }
            case T_IF_ELSE: {
                const std::string p_ = v(kid[0]), kp = k(kid[0]);
                os << "    const bool " << pfx << "s" << id << " = " << kp << " && " << p_ << ";\n";
                os << "    const " << T << " " << V << " = " << pfx << "s" << id << " ? " << v(kid[1]) << " : " << v(kid[2]) << ";\n";
                os << "    const bool " << K << " = " << kp << " && (" << pfx << "s" << id << " ? " << k(kid[1]) << " : " << k(kid[2]) << ");\n";
                break;
            }
            case T_IS_IN: {
                const Node &C = cg.ast.nodes[n.kids[0]];
                const std::string a = v(kid[0]);
                os << "    bool " << V << " = false;\n";
                for (size_t i = 1; i < kid.size(); i++) {
                    const Node &it2 = cg.ast.nodes[n.kids[i]];
                    if (it2.lit_null || it2.dtype == B200_NULL) continue;
                    if (it2.dtype != C.dtype) fail("plan: is_in item dtype differs from the tested expression");
                    if (is_string(C.dtype)) os << "    " << V << " = " << V << " || str_eq(" << a << ", " << v(kid[i]) << ");\n";
                    else if (is_float(C.dtype))  // BTreeSet<FloatWrapper>: all NaNs are one element, -0.0 == 0.0 (is_in.rs:44-66)
                        os << "    " << V << " = " << V << " || (" << a << " == " << v(kid[i]) << ") || (isnan(" << a << ") && isnan(" << v(kid[i]) << "));\n";
                    else os << "    " << V << " = " << V << " || (" << a << " == " << v(kid[i]) << ");\n";
                }
                os << "    const bool " << K << " = " << k(kid[0]) << ";\n";
                break;
            }
            case T_BETWEEN: {
                const Node &C = cg.ast.nodes[n.kids[0]];
                const bool fk = is_float(C.dtype);  // floats compare by totalOrder, like the comparison operators
                const std::string a = fk ? "f_key(" + v(kid[0]) + ")" : v(kid[0]), lo = fk ? "f_key(" + v(kid[1]) + ")" : v(kid[1]),
                                  hi = fk ? "f_key(" + v(kid[2]) + ")" : v(kid[2]);
                if (is_string(C.dtype))
                    os << "    const bool " << V << " = str_cmp(" << a << ", " << lo << ") >= 0 && str_cmp(" << a << ", " << hi << ") <= 0;\n";
                else
                    os << "    const bool " << V << " = " << a << " >= " << lo << " && " << a << " <= " << hi << ";\n";
                // value.gte(lower).and(value.lte(upper)) with the Kleene `and` (src/daft-core/src/array/ops/between.rs:30-32):
                // null unless both comparisons are known, or one of them is known to be false
                {
                    const std::string kg = "(" + k(kid[0]) + " && " + k(kid[1]) + ")", kl = "(" + k(kid[0]) + " && " + k(kid[2]) + ")";
                    const std::string ge = is_string(C.dtype) ? "(str_cmp(" + a + ", " + lo + ") >= 0)" : "(" + a + " >= " + lo + ")";
                    const std::string le = is_string(C.dtype) ? "(str_cmp(" + a + ", " + hi + ") <= 0)" : "(" + a + " <= " + hi + ")";
                    os << "    const bool " << K << " = (" << kg << " && " << kl << ") || (" << kg << " && !" << ge << ") || (" << kl << " && !" << le << ");\n";
                }
                break;
            }
            case T_FILL_NULL:
                os << "    const " << T << " " << V << " = " << k(kid[0]) << " ? " << v(kid[0]) << " : " << v(kid[1]) << ";\n";
                os << "    const bool " << K << " = " << k(kid[0]) << " || " << k(kid[1]) << ";\n";
                break;
            case T_FUNCTION: {
                std::string allk = "true";
                for (int c : kid) allk += " && " + k(c);
                const std::string a = v(kid[0]);
                switch (n.op) {
                    case F_YEAR: os << "    const int32_t " << V << " = date_year((int32_t)" << a << ");\n"; break;
                    case F_MONTH: os << "    const uint32_t " << V << " = date_month((int32_t)" << a << ");\n"; break;
                    case F_DAY: os << "    const uint32_t " << V << " = date_day((int32_t)" << a << ");\n"; break;
                    case F_DOW: os << "    const uint32_t " << V << " = date_dow((int32_t)" << a << ");\n"; break;
                    case F_UTF8_LENGTH: os << "    const uint64_t " << V << " = (uint64_t)str_char_len(" << a << ");\n"; break;
                    case F_STARTSWITH: os << "    const bool " << V << " = str_startswith(" << a << ", " << v(kid[1]) << ");\n"; break;
                    case F_ENDSWITH: os << "    const bool " << V << " = str_endswith(" << a << ", " << v(kid[1]) << ");\n"; break;
                    case F_CONTAINS: os << "    const bool " << V << " = str_contains(" << a << ", " << v(kid[1]) << ");\n"; break;
                    case F_LEFT: os << "    const Str " << V << " = str_left(" << a << ", (int64_t)" << v(kid[1]) << ");\n"; break;
                    case F_SQRT: os << "    const " << T << " " << V << " = " << (n.dtype == B200_F32 ? "sqrtf(" : "sqrt(") << a << ");\n"; break;  // IEEE correctly rounded
                    case F_IS_NAN: os << "    const bool " << V << " = (" << a << " != " << a << ");\n"; break;
                    case F_NOT_NAN: os << "    const bool " << V << " = (" << a << " == " << a << ");\n"; break;
                    case F_IS_INF: os << "    const bool " << V << " = (" << a << " - " << a << " != 0) && (" << a << " == " << a << ");\n"; break;
                    case F_CEIL: os << "    const " << T << " " << V << " = " << (n.dtype == B200_F32 ? "ceilf(" : "ceil(") << a << ");\n"; break;
                    case F_FLOOR: os << "    const " << T << " " << V << " = " << (n.dtype == B200_F32 ? "floorf(" : "floor(") << a << ");\n"; break;
                    case F_ROUND: {
                        const int64_t dec = cg.ast.nodes[kid[1]].lit_i;
                        if (dec == 0) {
                            os << "    const " << T << " " << V << " = " << (n.dtype == B200_F32 ? "roundf(" : "round(") << a << ");\n";
                        } else {
                            char mbuf[64];
                            snprintf(mbuf, sizeof mbuf, "%.17g", std::pow(10.0, (double)dec));
                            os << "    const " << T << " " << V << " = (" << T << ")(round((double)" << a << " * " << mbuf << ") / " << mbuf << ");\n";
                        }
                        allk = k(kid[0]);
                        break;
                    }
                    case F_MATCH_SEQ: {
                        const int64_t flags = cg.ast.nodes[kid[1]].lit_i;
                        const size_t nseg = kid.size() - 2;
                        os << "    bool " << V << " = true;\n    {\n      int64_t pos_ = 0;\n";
                        for (size_t j = 0; j < nseg; j++) {
                            const std::string sg = v(kid[2 + j]);
                            const bool first = j == 0, last = j + 1 == nseg;
                            if (first && (flags & 1)) {
                                os << "      if (" << V << ") { " << V << " = str_startswith(" << a << ", " << sg << "); pos_ = " << sg << ".len; }\n";
                            } else if (last && (flags & 2)) {
                                os << "      if (" << V << ") { const int64_t f_ = " << a << ".len - " << sg << ".len; " << V << " = f_ >= pos_ && str_endswith(" << a
                                   << ", " << sg << "); pos_ = f_ + " << sg << ".len; }\n";
                            } else {
                                os << "      if (" << V << ") { const int64_t f_ = str_find_from(" << a << ", " << sg << ", pos_); " << V << " = f_ >= 0; pos_ = f_ + " << sg
                                   << ".len; }\n";
                            }
                        }
                        if (nseg == 1 && (flags & 3) == 3) os << "      " << V << " = " << V << " && " << a << ".len == " << v(kid[2]) << ".len;\n";
                        if (nseg == 0 && (flags & 3) == 3) os << "      " << V << " = " << a << ".len == 0;\n";
                        os << "    }\n";
                        allk = k(kid[0]);
                        break;
                    }
                    case F_EXP: case F_LN: case F_LOG2: case F_LOG10: case F_SIN: case F_COS: case F_TAN: case F_CBRT: {
                        static const char *names[] = {"exp", "log", "log2", "log10", "sin", "cos", "tan", "cbrt"};
                        os << "    const " << T << " " << V << " = " << names[n.op - F_EXP] << (n.dtype == B200_F32 ? "f(" : "(") << a << ");\n";
                        break;
                    }
                    case F_LSTRIP: os << "    const Str " << V << " = str_lstrip(" << a << ");\n"; break;
                    case F_RSTRIP: os << "    const Str " << V << " = str_rstrip(" << a << ");\n"; break;
                    case F_RIGHT: os << "    const Str " << V << " = str_right(" << a << ", (int64_t)" << v(kid[1]) << ");\n"; break;
                    case F_FIND: os << "    const int64_t " << V << " = str_find_from(" << a << ", " << v(kid[1]) << ", 0);\n"; break;
                    case F_TS_DATE: os << "    const int32_t " << V << " = (int32_t)((" << a << " - ts_floor_mod((int64_t)" << a << ", 86400000000LL)) / 86400000000LL);\n"; break;
                    case F_RX_MATCH:
                        os << "    const bool " << V << " = rx_exec(" << a << ", " << v(kid[1]) << ").matched;\n";
                        allk = k(kid[0]);
                        break;
                    case F_RX_EXTRACT: case F_RX_REPLACE_GROUP: {
                        const std::string r_ = pfx + "rx" + std::to_string(id);
                        os << "    const RxResult " << r_ << " = rx_exec(" << a << ", " << v(kid[1]) << ");\n";
                        if (n.op == F_RX_EXTRACT) {
                            os << "    const Str " << V << " = (" << r_ << ".matched && " << r_ << ".c0 >= 0) ? Str{" << a << ".p + " << r_ << ".c0, (int64_t)(" << r_ << ".c1 - " << r_
                               << ".c0)} : Str{" << a << ".p, 0};\n";
                            allk = k(kid[0]) + " && " + r_ + ".matched && " + r_ + ".c0 >= 0";
                        } else {
                            os << "    const Str " << V << " = !" << r_ << ".matched ? " << a << " : (" << r_ << ".c0 >= 0 ? Str{" << a << ".p + " << r_ << ".c0, (int64_t)(" << r_
                               << ".c1 - " << r_ << ".c0)} : Str{" << a << ".p, 0});\n";
                            allk = k(kid[0]);
                        }
                        break;
                    }
                    case F_SUBSTR:
                        os << "    bool " << pfx << "ok" << id << " = false;\n";
                        os << "    const Str " << V << " = str_substr(" << a << ", (int64_t)" << cg.ast.nodes[kid[1]].lit_i << "LL, (int64_t)" << cg.ast.nodes[kid[2]].lit_i
                           << "LL, " << pfx << "ok" << id << ");\n";
                        allk = k(kid[0]) + " && " + pfx + "ok" + std::to_string(id);
                        break;
                    case F_HOUR: os << "    const uint32_t " << V << " = ts_hour((int64_t)" << a << ");\n"; break;
                    case F_MINUTE: os << "    const uint32_t " << V << " = ts_minute((int64_t)" << a << ");\n"; break;
                    case F_SECOND: os << "    const uint32_t " << V << " = ts_second((int64_t)" << a << ");\n"; break;
                    case F_TS_TRUNC:
                        os << "    const int64_t " << V << " = ts_trunc((int64_t)" << a << ", (int64_t)" << cg.ast.nodes[kid[1]].lit_i << "LL);\n";
                        allk = k(kid[0]);
                        break;
                    case F_SIGN:
                        if (is_float(n.dtype))  // f64::signum: NaN stays NaN, otherwise +-1 with the operand's sign (signum(+-0.0) = +-1.0)
                            os << "    const " << T << " " << V << " = (" << a << " != " << a << ") ? " << a << " : "
                               << (n.dtype == B200_F32 ? "copysignf(1.0f, " : "copysign(1.0, ") << a << ");\n";
                        else if (is_signed_int(n.dtype))
                            os << "    const " << T << " " << V << " = (" << T << ")((" << a << " > 0) - (" << a << " < 0));\n";
                        else
                            os << "    const " << T << " " << V << " = (" << T << ")(" << a << " != 0);\n";
                        break;
                }
                os << "    const bool " << K << " = " << allk << ";\n";
                break;
            }
        }
        done[n.key] = id;
        return id;
    }
};

static void collect_cols(const Ast &a, int id, std::vector<int> &cols) {
    const Node &n = a.nodes[id];
    if (n.tag == T_COL) {
        for (int c : cols)
            if (c == n.col) return;
        cols.push_back(n.col);
    }
    for (int c : n.kids) collect_cols(a, c, cols);
}

static std::string file_header(CG &cg) {
    std::ostringstream o;
    o << "using namespace b200;\nusing namespace b200::jit;\n";
    o << cg.globals.str();
    return o.str();
}

// ---- PROJECT -------------------------------------------------------------------------------------
static std::string gen_project(Reader &r, Ast &ast, Compiled &meta, bool with_take) {
    CG cg{ast, meta};
    uint32_t nexpr = r.u32();
    std::vector<int> roots;
    std::vector<std::string> names;
    for (uint32_t i = 0; i < nexpr; i++) {
        names.push_back(r.str());
        roots.push_back(parse_node(r, ast));
    }
    if (ast.list_col_refs) fail("plan: expressions cannot read list columns");
    const int iv_n = cg.int_slot(I_N);
    const int take_slot = with_take ? cg.ptr_slot(P_TAKE, 0) : -1;
    std::ostringstream body;
    ExprGen eg{cg, DIRECT, body, "", {}};
    body << "    const bool rowok = active && row >= 0;\n";
    for (uint32_t i = 0; i < nexpr; i++) {
        int id = eg.emit(roots[i]);
        const Node &n = ast.nodes[roots[i]];
        OutCol oc{names[i], n.dtype, n.nullable || with_take, n.nullable || n.dtype == B200_NULL};
        if (n.dtype == B200_NULL) oc.nullable = true;
        meta.outs.push_back(oc);
        const std::string V = eg.v(id), K = eg.k(id);
        body << "    // ---- output " << i << " (" << dtype_name(n.dtype) << ")\n";
        if (n.dtype == B200_NULL) continue;
        if (oc.nullable) {
            body << "    { const unsigned m_ = __ballot_sync(FULL_MASK, active && " << K << "); if (pass == 0 && (threadIdx.x & 31) == 0) ((uint32_t*)"
                 << cg.P(P_OUT_VALID, (int)i) << ")[out_row >> 5] = m_; }\n";
        }
        if (is_string(n.dtype)) {
            meta.has_string_out = true;
            body << "    if (pass == 0) { if (active) ((int64_t*)" << cg.P(P_OUT_LEN, (int)i) << ")[out_row] = " << K << " ? " << V << ".len : 0; }\n";
            body << "    else if (active && " << K << ") { uint8_t* d_ = (uint8_t*)" << cg.P(P_OUT_DATA, (int)i) << " + ((const int64_t*)"
                 << cg.P(P_OUT_OFFSETS, (int)i) << ")[out_row]; for (int64_t j_ = 0; j_ < " << V << ".len; j_++) d_[j_] = " << V << ".p[j_]; }\n";
        } else if (n.dtype == B200_BOOL) {
            body << "    { const unsigned m_ = __ballot_sync(FULL_MASK, active && " << K << " && " << V << "); if (pass == 0 && (threadIdx.x & 31) == 0) ((uint32_t*)"
                 << cg.P(P_OUT_DATA, (int)i) << ")[out_row >> 5] = m_; }\n";
        } else {
            body << "    if (pass == 0 && active) ((" << ctype(n.dtype) << "*)" << cg.P(P_OUT_DATA, (int)i) << ")[out_row] = " << K << " ? " << V << " : (" << ctype(n.dtype) << ")0;\n";
        }
    }
    std::ostringstream o;
    o << file_header(cg);
    o << "struct Prog {\n";
    o << "  typedef Params<" << meta.ptrs.size() << ", " << meta.ints.size() << "> P;\n";
    o << "  static constexpr int IV_N = " << iv_n << ";\n  static constexpr int PTR_TAKE = " << take_slot << ";\n";
    o << "  __device__ __forceinline__ static void row(const P& p, const int64_t out_row, const int64_t row, const bool active, const int pass) {\n";
    o << body.str();
    o << "  }\n};\n";
    o << "extern \"C\" __global__ void __launch_bounds__(256) b200_project_k(const __grid_constant__ Prog::P p, int pass) { project_body<Prog>(p, pass); }\n";
    return o.str();
}

// ---- FILTER (fused compaction) ---------------------------------------------------------------------------
static std::string gen_filter(Reader &r, Ast &ast, Compiled &meta) {
    CG cg{ast, meta};
    const int root = parse_node(r, ast);
    if (ast.list_col_refs) fail("plan: expressions cannot read list columns");
    if (ast.nodes[root].dtype != B200_BOOL && ast.nodes[root].dtype != B200_NULL) fail("filter: predicate must be boolean");
    const int iv_n = cg.int_slot(I_N);
    const int p_counts = cg.ptr_slot(P_OUT_LEN, 0);
    std::ostringstream kb, cp;
    ExprGen eg{cg, DIRECT, kb, "", {}};
    kb << "    const bool rowok = true;\n";
    const int id = eg.emit(root);
    kb << "    return " << eg.k(id) << " && " << eg.v(id) << ";\n";  // a null predicate drops the row (lib.rs:542-563)
    // fixed-width columns without nulls are compacted in the kernel; the rest are gathered by the host via row ids
    bool need_idx = false;
    const int ncols = (int)ast.col_dtypes->size();
    for (int c = 0; c < ncols; c++) {
        const int32_t dt = (*ast.col_dtypes)[c];
        const int w = dtype_width(dt);
        OutCol oc{"", dt, false, false};
        if (w && !ast.nullable(c)) {
            // column-major over the tile's FILTER_SUB rows: all loads of a column in flight, then its stores
            cp << "    { " << ctype(dt) << " v_[FILTER_SUB];\n"
               << "#pragma unroll\n      for (int j = 0; j < FILTER_SUB; j++) v_[j] = dst[j] >= 0 ? ((const " << ctype(dt) << "*)" << cg.P(P_COL_DATA, c) << ")[src[j]] : " << zero_of(dt) << ";\n"
               << "#pragma unroll\n      for (int j = 0; j < FILTER_SUB; j++) if (dst[j] >= 0) ((" << ctype(dt) << "*)" << cg.P(P_OUT_DATA, c) << ")[dst[j]] = v_[j]; }\n";
            oc.nullable = false;
        } else {
            oc.nullable = true;  // marks "host gathers this column"
            if (dt != B200_NULL) need_idx = true;
        }
        meta.outs.push_back(oc);
    }
    const int p_idx = need_idx ? cg.ptr_slot(P_TAKE, 0) : -1;
    std::ostringstream o;
    o << file_header(cg);
    o << "struct Prog {\n  typedef Params<" << meta.ptrs.size() << ", " << meta.ints.size() << "> P;\n";
    o << "  static constexpr int IV_N = " << iv_n << ", PTR_COUNTS = " << p_counts << ", PTR_IDX = " << p_idx << ";\n";
    o << "  __device__ __forceinline__ static bool keep(const P& p, const int64_t row) {\n" << kb.str() << "  }\n";
    o << "  __device__ __forceinline__ static void copy(const P& p, const int64_t (&src)[FILTER_SUB], const int64_t (&dst)[FILTER_SUB]) {\n" << cp.str() << "  }\n";
    o << "};\n";
    o << "extern \"C\" __global__ void __launch_bounds__(256) b200_filter_k(const __grid_constant__ Prog::P p, int pass) { filter_body<Prog>(p, pass); }\n";
    meta.with_take = need_idx;
    return o.str();
}

// ---- AGG -----------------------------------------------------------------------------------------
static std::string hash_expr(int32_t dt, const std::string &v, const std::string &k, const std::string &seed) {
    // Daft null rules, see hash.cu header
    const std::string nullh = "0x2D06800538D394C2ULL";
    std::string h;
    switch (dt) {
        case B200_NULL: return "xxh::hash_empty(" + seed + ")";
        case B200_BOOL: h = "xxh::hash_u8(" + v + " ? (uint8_t)'1' : (uint8_t)'0', " + seed + ")"; break;
        case B200_F32: h = "xxh::hash_u32((uint32_t)__float_as_int(" + v + "), " + seed + ")"; break;
        case B200_F64: h = "xxh::hash_u64((uint64_t)__double_as_longlong(" + v + "), " + seed + ")"; break;
        case B200_LARGE_UTF8: case B200_LARGE_BINARY:
            return "(" + k + " ? xxh::hash_bytes(" + v + ".p, " + v + ".len, " + seed + ") : xxh::hash_empty(" + seed + "))";
        default:
            switch (dtype_width(dt)) {
                case 1: h = "xxh::hash_u8((uint8_t)" + v + ", " + seed + ")"; break;
                case 2: h = "xxh::hash_u16((uint16_t)" + v + ", " + seed + ")"; break;
                case 4: h = "xxh::hash_u32((uint32_t)" + v + ", " + seed + ")"; break;
                default: h = "xxh::hash_u64((uint64_t)" + v + ", " + seed + ")"; break;
            }
    }
    return "(" + k + " ? " + h + " : " + nullh + ")";
}
static std::string bits_equal(int32_t dt, const std::string &a, const std::string &b) {
    if (is_string(dt)) return "str_eq(" + a + ", " + b + ")";
    if (dt == B200_F32) return "(__float_as_int(" + a + ") == __float_as_int(" + b + "))";
    if (dt == B200_F64) return "(__double_as_longlong(" + a + ") == __double_as_longlong(" + b + "))";
    if (dt == B200_NULL) return "true";
    return "(" + a + " == " + b + ")";
}
static int32_t sum_dtype(int32_t dt) {  // inline_agg.rs:424-460
    if (is_signed_int(dt)) return B200_I64;
    if (is_unsigned_int(dt)) return B200_U64;
    if (dt == B200_F32 || dt == B200_F64) return dt;
    fail(std::string("agg: sum not supported for dtype ") + dtype_name(dt));
}
static std::string ord_of(int32_t dt, const std::string &v) {
    if (dt == B200_F64) return "ord_f64(" + v + ")";
    if (dt == B200_F32) return "ord_f32(" + v + ")";
    if (is_unsigned_int(dt) || dt == B200_BOOL) return "ord_u64((uint64_t)" + v + ")";
    return "ord_i64((int64_t)" + v + ")";
}

// flavor: 0 = the value itself; 1 = float min / max candidate (NaN values are skipped); 2 = 1 + row of the last NaN value;
// 3 = 1 + row of the last valid value (2 and 3 are C_MAX cells over the row number); 4 = first row with a valid value (C_MIN
// over the row number: any_value(ignore_nulls))
struct LaneCell { int op; std::string child_key; int val_slot; int32_t val_dtype; int flavor; };
struct WarpCell { std::string cond_key; int val_slot; bool count_nulls; };

static std::string gen_agg(Reader &r, Ast &ast, Compiled &meta, int variant) {
    const bool nanfix_variant = variant == 3;   // generated like the global-table build, different kernel body
    const bool rowslot_variant = variant == 4;  // likewise: writes every row's table slot (list / set / concat aggregates)
    const bool global_variant = variant == 1 || nanfix_variant || rowslot_variant, staged_variant = variant == 2;
    CG cg{ast, meta};
    const bool has_filter = r.u8() != 0;
    int filter_root = -1;
    if (has_filter) filter_root = parse_node(r, ast);
    const uint32_t nkeys = r.u32();
    std::vector<int> key_roots;
    std::vector<std::string> key_names;
    {
        // remember the raw bytes of the key list to rebuild a PROJECT plan for the output key columns
        size_t start = r.pos;
        for (uint32_t i = 0; i < nkeys; i++) {
            key_names.push_back(r.str());
            key_roots.push_back(parse_node(r, ast));
        }
        meta.key_project_plan.resize(4 + (r.pos - start));
        memcpy(meta.key_project_plan.data(), &nkeys, 4);
        memcpy(meta.key_project_plan.data() + 4, r.p + start, r.pos - start);
    }
    meta.nkeys = (int)nkeys;
    const uint32_t naggs = r.u32();
    struct AggIn { int kind, mode; std::string name; int root; };
    std::vector<AggIn> aggs;
    uint32_t nanyval = 0, nlist = 0;
    std::vector<uint8_t> anyval_bytes;  // [name, node] of every any_value child = the entries of a PROJECT plan
    std::vector<uint8_t> list_bytes;    // likewise for the list / set / concat children
    for (uint32_t i = 0; i < naggs; i++) {
        AggIn a;
        a.kind = r.u8();
        a.mode = r.u8();
        const size_t start = r.pos;
        a.name = r.str();
        a.root = parse_node(r, ast);
        if (a.kind == AGG_ANY_VALUE) {
            nanyval++;
            anyval_bytes.insert(anyval_bytes.end(), r.p + start, r.p + r.pos);
        }
        if (is_list_agg(a.kind) && ast.nodes[a.root].tag != T_COL) {
            nlist++;
            list_bytes.insert(list_bytes.end(), r.p + start, r.p + r.pos);
        }
        aggs.push_back(a);
    }
    if (nanyval) {
        meta.anyval_project_plan.resize(4);
        memcpy(meta.anyval_project_plan.data(), &nanyval, 4);
        meta.anyval_project_plan.insert(meta.anyval_project_plan.end(), anyval_bytes.begin(), anyval_bytes.end());
    }
    if (nlist) {
        meta.listagg_project_plan.resize(4);
        memcpy(meta.listagg_project_plan.data(), &nlist, 4);
        meta.listagg_project_plan.insert(meta.listagg_project_plan.end(), list_bytes.begin(), list_bytes.end());
    }
    if (r.pos != r.n) fail("plan: trailing bytes");

    // ---- value slots (distinct aggregate-input expressions) and cells
    std::vector<int> val_roots;            // node id per value slot
    std::vector<int32_t> val_dtypes;       // dtype after widening
    std::vector<int32_t> val_src_dtypes;
    std::map<std::string, int> val_index;
    auto val_slot = [&](int root, int32_t widened) {
        const Node &n = ast.nodes[root];
        std::string key = n.key + "#" + std::to_string(widened);
        auto it = val_index.find(key);
        if (it != val_index.end()) return it->second;
        int s = (int)val_roots.size();
        val_roots.push_back(root);
        val_dtypes.push_back(widened);
        val_src_dtypes.push_back(n.dtype);
        val_index[key] = s;
        return s;
    };
    std::vector<LaneCell> lane_cells;
    std::vector<WarpCell> warp_cells;
    auto lane_cell = [&](int op, int vslot, int flavor = 0) {
        for (size_t i = 0; i < lane_cells.size(); i++)
            if (lane_cells[i].op == op && lane_cells[i].val_slot == vslot && lane_cells[i].flavor == flavor) return (int)i;
        lane_cells.push_back({op, "", vslot, val_dtypes[vslot], flavor});
        return (int)lane_cells.size() - 1;
    };
    auto warp_cell = [&](int vslot, bool count_nulls) {
        for (size_t i = 0; i < warp_cells.size(); i++)
            if (warp_cells[i].val_slot == vslot && warp_cells[i].count_nulls == count_nulls) return (int)i;
        warp_cells.push_back({"", vslot, count_nulls});
        return (int)warp_cells.size() - 1;
    };
    int anyval_seen = 0, list_seen = 0, legit_list_refs = 0;
    for (auto &a : aggs) {
        const Node &c = ast.nodes[a.root];
        AggOut o;
        o.name = a.name;
        o.kind = a.kind;
        o.count_mode = a.mode;
        o.child_nullable = c.nullable;
        switch (a.kind) {
            case AGG_COUNT: {
                o.out_dtype = B200_U64;
                if (a.mode == 0) break;  // all -> nrows
                if (!c.nullable) { o.const_zero = (a.mode == 2); break; }
                int vs = val_slot(a.root, c.dtype);
                o.cnt_cell = warp_cell(vs, a.mode == 2);
                break;
            }
            case AGG_SUM: case AGG_MEAN: {
                int32_t w = sum_dtype(c.dtype);
                int vs = val_slot(a.root, w);
                o.val_dtype = w;
                o.out_dtype = a.kind == AGG_MEAN ? B200_F64 : w;
                o.lane_cell = lane_cell(w == B200_F64 ? C_ADD_F64 : (w == B200_F32 ? C_ADD_F32 : C_ADD_I64), vs);
                if (c.nullable) o.cnt_cell = warp_cell(vs, false);
                break;
            }
            case AGG_MIN: case AGG_MAX: {
                if (!(is_numeric(c.dtype) || c.dtype == B200_DATE32 || c.dtype == B200_TIMESTAMP || c.dtype == B200_BOOL))
                    fail(std::string("agg: min/max not supported on the device for dtype ") + dtype_name(c.dtype));
                int vs = val_slot(a.root, c.dtype);
                o.val_dtype = c.dtype;
                o.out_dtype = c.dtype;
                if (is_float(c.dtype)) {  // the reference's NaN-forgetting left fold (AggOut::nan_cell)
                    o.lane_cell = lane_cell(a.kind == AGG_MIN ? C_MIN : C_MAX, vs, 1);
                    o.nan_cell = lane_cell(C_MAX, vs, 2);
                    o.last_cell = lane_cell(C_MAX, vs, 3);
                    meta.has_nan_fold = true;
                } else {
                    o.lane_cell = lane_cell(a.kind == AGG_MIN ? C_MIN : C_MAX, vs);
                }
                if (c.nullable) o.cnt_cell = warp_cell(vs, false);
                break;
            }
            case AGG_ANY_VALUE: {
                // Series::any_value (src/daft-core/src/series/ops/agg.rs:353-382): the child at the group's first row; with
                // ignore_nulls the first NON-NULL row, which needs its own per-group row cell -- only offered where it is the
                // same thing (children that cannot be null)
                o.out_dtype = c.dtype;
                o.proj_index = anyval_seen++;
                if (a.mode == 1 && c.nullable && c.dtype != B200_NULL) {  // the first NON-NULL row of the group: its own row cell
                    const int vs = val_slot(a.root, c.dtype);
                    o.first_valid_cell = lane_cell(C_MIN, vs, 4);
                }
                break;
            }
            case AGG_LIST: case AGG_SET: case AGG_CONCAT: {
                // grouped_list / grouped_set / grouped_concat (src/daft-core/src/array/ops/list_agg.rs:30-57, series/ops/agg.rs:
                // 531-575, array/ops/concat_agg.rs:138-161): no accumulator cell -- the host orders the rows by group after the
                // table is built (ops_jit.cu build_list_aggs) and gathers the child, evaluated at every row, in that order
                if (a.kind == AGG_CONCAT && !is_string(c.dtype) && c.dtype != B200_LIST)
                    fail(std::string("agg: concat is offered for string and list columns, got ") + dtype_name(c.dtype));
                if (a.kind != AGG_CONCAT && c.dtype == B200_LIST) fail("agg: list / set of a list column is not supported");
                o.val_dtype = c.dtype;
                o.out_dtype = a.kind == AGG_CONCAT ? c.dtype : B200_LIST;
                if (c.tag == T_COL) {
                    o.child_col = c.col;
                    if (c.dtype == B200_LIST) legit_list_refs++;
                } else {
                    o.proj_index = list_seen++;
                }
                break;
            }
            default: fail("agg: unsupported aggregation kind " + std::to_string(a.kind));
        }
        meta.aggs.push_back(o);
    }
    if (ast.list_col_refs != legit_list_refs) fail("plan: expressions cannot read list columns");
    // lane cells = value cells, then the group's row count and its first row; warp cells = ballot counts
    const int NL = (int)lane_cells.size(), NW = (int)warp_cells.size(), NLT = NL + 2;
    meta.NLANE = NLT;
    meta.NWARP = NW;
    const int astride = (1 + NLT + NW + 3) / 4 * 4;  // record = code + cells, whole 32-byte sectors
    meta.astride = astride;
    for (auto &c : lane_cells) meta.lane_cell_ops.push_back(c.op);
    meta.lane_cell_ops.push_back(C_ADD_I64);  // row count
    meta.lane_cell_ops.push_back(C_MIN);      // first row
    int G1 = (60 * 1024) / (8 * NLT * 256);
    G1 = std::max(1, std::min(8, G1));
    if (nkeys == 0) G1 = 1;
    meta.G1 = G1;
    auto env_int = [](const char *name, int dflt) {
        const char *v = getenv(name);
        return v && *v ? atoi(v) : dflt;
    };
    const int MINB = std::max(1, std::min(8, env_int("B200_AGG_MINB", 2)));
    // tier-2 of the global kernel: a shared-memory table of (code, cells) absorbing mid-cardinality inputs (budget 96 KB)
    int t2_slots = 64;
    while ((size_t)(t2_slots * 2) * 8 * (size_t)(1 + NLT + NW) <= 96 * 1024) t2_slots *= 2;
    if (env_int("B200_AGG_T2", 1) == 0) t2_slots = 2;

    const int iv_n = cg.int_slot(I_N), iv_mask = cg.int_slot(I_MASK), iv_rstride = cg.int_slot(I_RSTRIDE);
    const int p_hash = cg.ptr_slot(P_SLOT_HASH, 0), p_ovf = cg.ptr_slot(P_OVERFLOW, 0);  // slot tags, status flag
    const int p_rowslot = rowslot_variant ? cg.ptr_slot(P_SLOT_ROW, 0) : 0;                // int64 per input row
    // global cells: [0,NL) value cells, NL = row count, NL+1 = first row, [NL+2, NL+2+NW) ballot count cells
    const int cell_nrows = NL, cell_minrow = NL + 1;
    std::vector<int> p_acc(NLT + NW);
    for (int i = 0; i < NLT + NW; i++) p_acc[i] = cg.ptr_slot(P_ACC, i);

    // ---- exact key code budget (see `code()` below): bytes per string key, 0 = no exact code
    int exact_B = 0;
    bool exact_possible = nkeys > 0;
    int wide_key = -1, wide_avail = 0;  // the one 8-byte integer key whose VALUE decides exactness row by row
    {
        int bits_fixed = 0, nstr = 0, n64 = 0, idx64 = -1;
        for (uint32_t i = 0; i < nkeys; i++) {
            const Node &n = ast.nodes[key_roots[i]];
            if (is_string(n.dtype)) nstr++;
            else if (n.dtype == B200_BOOL) bits_fixed += 1;
            else if (n.dtype != B200_NULL) bits_fixed += 8 * dtype_width(n.dtype);
            if (n.nullable) bits_fixed += 1;
            if (n.dtype == B200_I64 || n.dtype == B200_U64 || n.dtype == B200_TIMESTAMP) { n64++; idx64 = (int)i; }
        }
        if (bits_fixed > 62) {
            exact_possible = false;
            // Int64 keys (every TPC-H key column) rarely use their top bits: pack the value when it fits the bits
            // left over by the other keys, otherwise that row takes the fingerprint + verification path
            if (n64 == 1 && nstr == 0 && bits_fixed - 64 <= 30) {
                exact_possible = true;
                wide_key = idx64;
                wide_avail = 62 - (bits_fixed - 64);
            }
        }
        if (exact_possible && nstr) {
            exact_B = std::min(7, ((62 - bits_fixed) / nstr - 3) / 8);
            if (exact_B < 1) exact_possible = false;
        }
    }
    // string columns that appear ONLY as bare group keys can be prefetched with the cheap 4-byte pack
    std::vector<int> key_only_str_cols;
    if (exact_possible && exact_B >= 1 && exact_B <= 4) {
        for (uint32_t i = 0; i < nkeys; i++) {
            const Node &n = ast.nodes[key_roots[i]];
            if (n.tag == T_COL && is_string(n.dtype)) key_only_str_cols.push_back(n.col);
        }
    }
    auto is_pack4_col = [&](int c) {
        for (int x : key_only_str_cols)
            if (x == c) return true;
        return false;
    };

    // ---- columns referenced anywhere -> Raw struct
    std::vector<int> cols;
    if (filter_root >= 0) collect_cols(ast, filter_root, cols);
    for (int k : key_roots) collect_cols(ast, k, cols);
    for (int v : val_roots) collect_cols(ast, v, cols);
    std::ostringstream raw, load, load2, pf;
    const int chunk_rows = 32 * std::max(1, std::min(8, env_int("B200_AGG_U", 2)));
    // L2 prefetch of the warp's next chunk (jit_prelude.cuh agg_chunk).  Measured on B200 (SF10): plans that read string
    // columns gain (Q1 0.636 -> 0.571 ms, count-by-flags 0.490 -> 0.449 ms: their 9-11 narrow loads per row are latency
    // bound), numeric-only plans that already stream at ~6 TB/s lose 10-15% -- so it is on exactly when a string
    // column is read.  B200_AGG_PREFETCH = 0 off / 1 all columns / 2 string columns only; default: auto.
    bool reads_strings = false;
    for (int c : cols) reads_strings = reads_strings || is_string((*ast.col_dtypes)[c]);
    const int pf_env = env_int("B200_AGG_PREFETCH", -1);
    const int pf_mode = pf_env >= 0 ? pf_env : (reads_strings ? 1 : 0);
    bool pf_is_string_col = false;
    int pf_lines = 0;
    auto pf_span = [&](const std::string &ptr, int bytes_per_row, int extra_bytes) {  // lines of one chunk of a column
        if (pf_mode == 0 || (pf_mode == 2 && !pf_is_string_col)) return;
        const int lines = (chunk_rows * bytes_per_row + extra_bytes + 127) / 128;  // chunks start on 64-row boundaries of 256-byte aligned buffers
        // lanes [pf_lines, pf_lines + lines) own this column's lines (second slot for footprints beyond 32 lines)
        for (int l = 0; l < lines; l++, pf_lines++) {
            if (pf_lines >= 64) break;
            const int slot = pf_lines / 32, ln = pf_lines % 32;
            pf << "    if (lane == " << ln << ") { f.b" << slot << " = (const char*)" << ptr << " + " << l * 128 << "; f.w" << slot << " = " << bytes_per_row << "; }\n";
        }
    };
    size_t raw_bytes = 0;
    // staged variant: R rows of every fixed-width column / R+2 string offsets / R*L bytes of a fixed-length string
    // column are bulk-copied into a per-warp shared-memory stage; ld_* = the loader reading that stage
    const int SW = staged_variant ? (env_int("B200_AGG_STAGE_WARPS", 8) >= 16 ? 16 : 8) : 8;  // warps per CTA
    const int SRPL = std::max(1, std::min(4, env_int("B200_AGG_STAGE_RPL", 3)));                   // rows per thread per stage
    const int SR = SW * 32 * SRPL;
    int SS = std::max(2, std::min(8, env_int("B200_AGG_STAGES", 2)));
    int stage_off = 0, stage_fixed_tx = 0;
    std::ostringstream st_issue, ld_a, ld_b;
    auto seg = [&](int bytes) {
        const int o = stage_off;
        stage_off += (bytes + 15) / 16 * 16;
        return o;
    };
    for (int c : cols) {
        const int32_t dt = (*ast.col_dtypes)[c];
        const bool nullable = ast.nullable(c);
        if (dt == B200_NULL) continue;
        if (staged_variant) {
            if (nullable) ld_a << "      r.k" << c << " = bit_at((const uint8_t*)" << cg.P(P_COL_VALID, c) << ", row);\n";
            if (is_string(dt)) {
                const int bytes = (SR + 2) * 8, o = seg(bytes);
                stage_fixed_tx += bytes;
                st_issue << "    bulk_g2s(dst + " << o << ", (const char*)" << cg.P(P_COL_OFFSETS, c) << " + base * 8, " << bytes << ", bar);\n";
                ld_a << "      { const int64_t* of_ = (const int64_t*)(st + " << o << "); r.o" << c << "a = of_[rr]; r.o" << c << "b = of_[rr + 1]; }\n";
                if (is_pack4_col(c) && ast.fixed_len(c) > 0) {
                    const int L = ast.fixed_len(c);
                    const int ob = seg(SR * L);
                    stage_fixed_tx += SR * L;
                    st_issue << "    bulk_g2s(dst + " << ob << ", (const char*)" << cg.P(P_COL_DATA, c) << " + base * " << L << ", " << SR * L << ", bar);\n";
                    ld_a << "      ok_ = ok_ && (r.o" << c << "b - r.o" << c << "a == " << L << ") && (r.o" << c << "a == row * " << L << ");\n";
                    ld_b << "      if (fast_) r.pk" << c << " = (uint64_t)smem_pack4(st + " << ob << " + rr * " << L << ", " << L << ");\n"
                         << "      else r.pk" << c << " = str_pack4(Str{(const uint8_t*)" << cg.P(P_COL_DATA, c) << " + r.o" << c << "a, r.o" << c << "b - r.o" << c << "a});\n";
                } else if (is_pack4_col(c)) {
                    ld_b << "      r.pk" << c << " = str_pack4(Str{(const uint8_t*)" << cg.P(P_COL_DATA, c) << " + r.o" << c << "a, r.o" << c << "b - r.o" << c << "a});\n";
                } else {
                    ld_b << "      r.pk" << c << " = str_pack8(Str{(const uint8_t*)" << cg.P(P_COL_DATA, c) << " + r.o" << c << "a, r.o" << c << "b - r.o" << c << "a});\n";
                }
            } else if (dt == B200_BOOL) {
                ld_a << "      r.c" << c << " = bit_at((const uint8_t*)" << cg.P(P_COL_DATA, c) << ", row);\n";
            } else {
                const int w = dtype_width(dt), bytes = SR * w, o = seg(bytes);
                stage_fixed_tx += bytes;
                st_issue << "    bulk_g2s(dst + " << o << ", (const char*)" << cg.P(P_COL_DATA, c) << " + base * " << w << ", " << bytes << ", bar);\n";
                ld_a << "      r.c" << c << " = ((const " << ctype(dt) << "*)(st + " << o << "))[rr];\n";
            }
        }
        if (nullable) {
            raw << "    bool k" << c << ";\n";
            load << "    r.k" << c << " = bit_at((const uint8_t*)" << cg.P(P_COL_VALID, c) << ", row);\n";
            raw_bytes += 4;
        }
        if (is_string(dt)) {
            raw << "    int64_t o" << c << "a, o" << c << "b; uint64_t pk" << c << ";\n";
            load << "    { const int64_t* of_ = (const int64_t*)" << cg.P(P_COL_OFFSETS, c) << "; r.o" << c << "a = of_[row]; r.o" << c << "b = of_[row + 1]; }\n";
            pf_is_string_col = true;
            pf_span(cg.P(P_COL_OFFSETS, c), 8, 8);
            if (is_pack4_col(c) && ast.fixed_len(c) > 0) pf_span(cg.P(P_COL_DATA, c), ast.fixed_len(c), 0);
            pf_is_string_col = false;
            // second load phase: the first <= 8 bytes of the value (depends on the offsets of phase one)
            if (is_pack4_col(c) && ast.fixed_len(c) > 0) {
                // Fixed-length value columns (flags, codes): the kernel is specialised for L = data_bytes / rows.  The
                // value's bytes are fetched SPECULATIVELY at data + row * L together with the offsets (one memory
                // round trip instead of two) and used only if the row's offsets confirm the guess.
                const int L = ast.fixed_len(c);
                raw << "    unsigned int sw" << c << ";\n";
                if (L == 1 || L == 2 || L == 4) {
                    const char *ty = L == 1 ? "uint8_t" : (L == 2 ? "uint16_t" : "uint32_t");
                    load << "    r.sw" << c << " = ((const " << ty << "*)" << cg.P(P_COL_DATA, c) << ")[row];\n";
                } else {
                    load << "    { const uint8_t* q_ = (const uint8_t*)" << cg.P(P_COL_DATA, c) << " + row * 3; r.sw" << c
                         << " = (unsigned)q_[0] | ((unsigned)q_[1] << 8) | ((unsigned)q_[2] << 16); }\n";
                }
                load2 << "    { const long long len_ = r.o" << c << "b - r.o" << c << "a;\n"
                      << "      if (r.o" << c << "a == row * " << L << " && len_ == " << L << ") r.pk" << c << " = (uint64_t)r.sw" << c << ";\n"
                      << "      else r.pk" << c << " = str_pack4(Str{(const uint8_t*)" << cg.P(P_COL_DATA, c) << " + r.o" << c << "a, len_}); }\n";
                raw_bytes += 28;
            } else if (is_pack4_col(c)) {
                load2 << "    r.pk" << c << " = str_pack4(Str{(const uint8_t*)" << cg.P(P_COL_DATA, c) << " + r.o" << c << "a, r.o" << c << "b - r.o" << c << "a});\n";
                raw_bytes += 24;
            } else {
                load2 << "    r.pk" << c << " = str_pack8(Str{(const uint8_t*)" << cg.P(P_COL_DATA, c) << " + r.o" << c << "a, r.o" << c << "b - r.o" << c << "a});\n";
                raw_bytes += 24;
            }
        } else if (dt == B200_BOOL) {
            raw << "    bool c" << c << ";\n";
            load << "    r.c" << c << " = bit_at((const uint8_t*)" << cg.P(P_COL_DATA, c) << ", row);\n";
            raw_bytes += 4;
        } else {
            raw << "    " << ctype(dt) << " c" << c << ";\n";
            load << "    r.c" << c << " = ((const " << ctype(dt) << "*)" << cg.P(P_COL_DATA, c) << ")[row];\n";
            pf_span(cg.P(P_COL_DATA, c), dtype_width(dt), 0);
            raw_bytes += std::max(4, dtype_width(dt));
        }
    }
    const int U = std::max(1, std::min(8, env_int("B200_AGG_U", 2)));
    // shared memory of the tier-1 kernels: lane cells, ballot cells, slots / codes / key values (generous sizeof(Keys))
    const size_t tier1_smem = 8 * (size_t)SW * G1 * NLT * 32 + 8 * (size_t)SW * G1 * std::max(NW, 1) + 16 * (size_t)G1 +
                              (size_t)G1 * (40 * (size_t)nkeys + 16);
    bool stage_fits = true;
    if (staged_variant) {
        // Measured on B200 (profiles/r01_agg_variants.md): the TMA build wins when two CTAs with 768-row stages fit on
        // an SM (ungrouped and narrow plans: +7..35%); with smaller stages or one CTA per SM the direct-load build
        // is faster (TPC-H Q1), so the plan is then declared "not staged" and the host runs the direct build.
        const size_t budget = (SW == 16 ? 227 * 1024 : (227 * 1024) / 2) - 1024;
        if (env_int("B200_AGG_STAGE_FORCE", 0)) {
            while (SS > 2 && tier1_smem + (size_t)SS * (stage_off + 16) > 226 * 1024) SS--;
            stage_fits = stage_off > 0 && tier1_smem + (size_t)SS * (stage_off + 16) <= 226 * 1024;
        } else {
            stage_fits = stage_off > 0 && tier1_smem + (size_t)SS * (stage_off + 16) <= budget;
        }
    }

    // ---- Keys / Vals structs
    std::ostringstream keys_s, vals_s;
    for (uint32_t i = 0; i < nkeys; i++) {
        const Node &n = ast.nodes[key_roots[i]];
        keys_s << "    " << ctype(n.dtype) << " v" << i << "; bool k" << i << ";\n";
        if (is_string(n.dtype)) keys_s << "    uint64_t p" << i << ";\n";
    }
    for (size_t i = 0; i < val_roots.size(); i++) vals_s << "    " << ctype(val_dtypes[i]) << " a" << i << "; bool k" << i << ";\n";

    // ---- eval (RAW access): filter, keys, vals
    std::ostringstream ev;
    {
        ExprGen eg{cg, RAW, ev, "", {}};
        if (filter_root >= 0) {
            int id = eg.emit(filter_root);
            ev << "    if (!(" << eg.k(id) << " && " << eg.v(id) << ")) return false;\n";  // null predicate drops the row (lib.rs:542-563)
        }
        for (uint32_t i = 0; i < nkeys; i++) {
            int id = eg.emit(key_roots[i]);
            ev << "    keys.v" << i << " = " << eg.v(id) << "; keys.k" << i << " = " << eg.k(id) << ";\n";
            if (is_string(ast.nodes[key_roots[i]].dtype)) {
                const Node &kn = ast.nodes[key_roots[i]];
                if (kn.tag == T_COL)  // prefix already fetched by load2()
                    ev << "    keys.p" << i << " = keys.k" << i << " ? r.pk" << kn.col << " : 0ULL;\n";
                else
                    ev << "    keys.p" << i << " = keys.k" << i << " ? str_pack8(keys.v" << i << ") : 0ULL;\n";
            }
        }
        for (size_t i = 0; i < val_roots.size(); i++) {
            int id = eg.emit(val_roots[i]);
            const Node &n = ast.nodes[val_roots[i]];
            if (n.dtype == B200_NULL) ev << "    vals.a" << i << " = " << zero_of(val_dtypes[i]) << "; vals.k" << i << " = false;\n";
            else ev << "    vals.a" << i << " = (" << ctype(val_dtypes[i]) << ")" << eg.v(id) << "; vals.k" << i << " = " << eg.k(id) << ";\n";
        }
        ev << "    return true;\n";
    }
    // ---- eval_keys (DIRECT access at an arbitrary row)
    std::ostringstream ek;
    {
        ExprGen eg{cg, DIRECT, ek, "", {}};
        ek << "    const bool rowok = true;\n";
        for (uint32_t i = 0; i < nkeys; i++) {
            int id = eg.emit(key_roots[i]);
            ek << "    keys.v" << i << " = " << eg.v(id) << "; keys.k" << i << " = " << eg.k(id) << ";\n";
            if (is_string(ast.nodes[key_roots[i]].dtype))
                ek << "    keys.p" << i << " = keys.k" << i << " ? str_pack8(keys.v" << i << ") : 0ULL;\n";
        }
    }
    // ---- hash / equality
    std::ostringstream hs, eq;
    hs << "    uint64_t h = 0;\n";
    for (uint32_t i = 0; i < nkeys; i++) {
        const Node &n = ast.nodes[key_roots[i]];
        if (is_string(n.dtype))
            hs << "    h = (k.k" << i << " ? hash_str_packed(k.v" << i << ", k.p" << i << ", " << (i == 0 ? "0ULL" : "h") << ") : xxh::hash_empty("
               << (i == 0 ? "0ULL" : "h") << "));\n";
        else
            hs << "    h = " << hash_expr(n.dtype, "k.v" + std::to_string(i), "k.k" + std::to_string(i), i == 0 ? "0ULL" : "h") << ";\n";
    }
    hs << "    return h;\n";
    eq << "    return true";
    for (uint32_t i = 0; i < nkeys; i++) {
        const Node &n = ast.nodes[key_roots[i]];
        const std::string si = std::to_string(i);
        if (is_string(n.dtype))
            eq << " && (a.k" << si << " == b.k" << si << ") && (!a.k" << si << " || str_eq_packed(a.v" << si << ", a.p" << si << ", b.v" << si << ", b.p" << si << "))";
        else
            eq << " && (a.k" << si << " == b.k" << si << ") && (!a.k" << si << " || " << bits_equal(n.dtype, "a.v" + si, "b.v" + si) << ")";
    }
    eq << ";\n";
    // cheap 64-bit fingerprint of the key values for the tier-1 lookup (a pre-filter only: keys_equal decides)
    std::ostringstream fps;
    fps << "    uint64_t f = 0x9E3779B97F4A7C15ULL;\n";
    for (uint32_t i = 0; i < nkeys; i++) {
        const Node &n = ast.nodes[key_roots[i]];
        const std::string si = std::to_string(i);
        std::string bits;
        if (is_string(n.dtype)) {
            // values longer than the packed prefix contribute a hash of ALL their bytes (shared prefixes are common)
            const bool p4 = n.tag == T_COL && is_pack4_col(n.col);
            bits = "(((uint64_t)k.v" + si + ".len <= " + (p4 ? "4" : "8") + "ULL ? k.p" + si + " : xxh::hash_bytes(k.v" + si + ".p, k.v" + si +
                   ".len, 0ULL)) ^ ((uint64_t)k.v" + si + ".len << 56))";
        }
        else if (n.dtype == B200_F64) bits = "(uint64_t)__double_as_longlong(k.v" + si + ")";
        else if (n.dtype == B200_F32) bits = "(uint64_t)(uint32_t)__float_as_int(k.v" + si + ")";
        else if (n.dtype == B200_NULL) bits = "0ULL";
        else bits = "(uint64_t)k.v" + si;
        fps << "    f = (f ^ (k.k" << si << " ? " << bits << " : 0xA5A5A5A5A5A5A5A5ULL)) * 0xD6E8FEB86659FD93ULL; f ^= f >> 32;\n";
    }
    fps << "    return f;\n";
    // Exact key code: when every key is narrow (ints <= 32 bit, bools, short strings) the whole key tuple is packed
    // injectively into 62 bits -- a tier-1 match on the code then needs no verification at all.  Rows whose strings
    // exceed their byte budget (and plans with wide keys) fall back to fingerprint | bit63 + one verification.
    std::ostringstream cds;
    {
        const bool possible = exact_possible;
        const int B = exact_B;
        if (!possible) {
            cds << "    return fingerprint(k) | 0x8000000000000000ULL;\n";
        } else {
            cds << "    uint64_t c = 0x4000000000000000ULL; bool ex = true;\n";
            int sh = 0;
            for (uint32_t i = 0; i < nkeys; i++) {
                const Node &n = ast.nodes[key_roots[i]];
                const std::string si = std::to_string(i);
                if (is_string(n.dtype)) {
                    const unsigned long long mask = (1ULL << (8 * B)) - 1ULL;
                    cds << "    { const uint64_t l_ = k.k" << si << " ? (uint64_t)k.v" << si << ".len : 0ULL; ex = ex && (l_ <= " << B << "ULL); c |= (k.p" << si
                        << " & " << mask << "ULL) << " << sh << "; c |= (l_ & 7ULL) << " << (sh + 8 * B) << "; }\n";
                    sh += 8 * B + 3;
                } else if ((int)i == wide_key) {
                    cds << "    { const uint64_t v_ = k.k" << si << " ? (uint64_t)k.v" << si << " : 0ULL; ex = ex && ((v_ >> " << wide_avail
                        << ") == 0ULL); c |= (v_ & " << ((1ULL << wide_avail) - 1ULL) << "ULL) << " << sh << "; }\n";
                    sh += wide_avail;
                } else if (n.dtype != B200_NULL) {
                    std::string bits;
                    int w = n.dtype == B200_BOOL ? 1 : 8 * dtype_width(n.dtype);
                    if (n.dtype == B200_BOOL) bits = "(uint64_t)k.v" + si;
                    else if (n.dtype == B200_F32) bits = "(uint64_t)(uint32_t)__float_as_int(k.v" + si + ")";
                    else bits = std::string("(uint64_t)(") + utype(n.dtype) + ")k.v" + si;
                    cds << "    c |= (k.k" << si << " ? " << bits << " : 0ULL) << " << sh << ";\n";
                    sh += w;
                }
                if (n.nullable) {
                    cds << "    c |= (uint64_t)(!k.k" << si << ") << " << sh << ";\n";
                    sh += 1;
                }
            }
            cds << "    return ex ? c : (fingerprint(k) | 0x8000000000000000ULL);\n";
        }
    }

    // ---- lane cells: init / update / global update / flush
    std::ostringstream li, lu, gu, wc, fl;
    std::ostringstream su, sfl, sinit;  // tier-2 (shared-memory table of the global kernel): update / flush / init of slot cells
    for (int c = 0; c < NL; c++) {
        const LaneCell &lc = lane_cells[c];
        const std::string cell = "cells[" + std::to_string(c * 32) + "]";
        const std::string val = "v.a" + std::to_string(lc.val_slot), vk0 = "v.k" + std::to_string(lc.val_slot);
        const std::string is_nan = "(" + val + " != " + val + ")";
        const std::string vk = lc.flavor == 1 ? "(" + vk0 + " && !" + is_nan + ")" : lc.flavor == 2 ? "(" + vk0 + " && " + is_nan + ")" : vk0;
        const std::string row_ord = "(unsigned long long)(row + 1)";
        const std::string gp = "((unsigned long long*)" + std::string("p.ptr[") + std::to_string(p_acc[c]) + "]) + slot";
        const std::string ordv = lc.flavor == 4 ? "(unsigned long long)row" : lc.flavor >= 2 ? row_ord : ord_of(lc.val_dtype, val);
        {   // tier-2 cell c of slot t lives at t2c[c * T2 + t]
            const std::string sp = "(t2c + " + std::to_string(c) + " * T2 + t)";
            const std::string sv = "t2c[" + std::to_string(c) + " * T2 + t]";
            switch (lc.op) {
                case C_ADD_F64:
                    sinit << "    " << sv << " = 0ULL;\n";
                    su << "    if (" << vk << ") atomicAdd((double*)" << sp << ", " << val << ");\n";
                    sfl << "    atomicAdd((double*)(" << gp << "), __longlong_as_double((long long)" << sv << "));\n";
                    break;
                case C_ADD_F32:
                    sinit << "    " << sv << " = 0ULL;\n";
                    su << "    if (" << vk << ") atomicAdd((float*)" << sp << ", " << val << ");\n";
                    sfl << "    atomicAdd((float*)(" << gp << "), __int_as_float((int)" << sv << "));\n";
                    break;
                case C_ADD_I64:
                    sinit << "    " << sv << " = 0ULL;\n";
                    su << "    if (" << vk << ") atomicAdd(" << sp << ", (unsigned long long)" << val << ");\n";
                    sfl << "    atomicAdd(" << gp << ", " << sv << ");\n";
                    break;
                case C_MIN:
                    sinit << "    " << sv << " = 0xFFFFFFFFFFFFFFFFULL;\n";
                    su << "    if (" << vk << ") atomicMin(" << sp << ", (unsigned long long)" << ordv << ");\n";
                    sfl << "    atomicMin(" << gp << ", " << sv << ");\n";
                    break;
                case C_MAX:
                    sinit << "    " << sv << " = 0ULL;\n";
                    su << "    if (" << vk << ") atomicMax(" << sp << ", (unsigned long long)" << ordv << ");\n";
                    sfl << "    atomicMax(" << gp << ", " << sv << ");\n";
                    break;
            }
        }
        switch (lc.op) {
            case C_ADD_F64:
                li << "    " << cell << " = 0ULL;\n";
                lu << "    if (" << vk << ") " << cell << " = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)" << cell << ") + " << val << ");\n";
                gu << "    if (" << vk << ") atomicAdd((double*)(" << gp << "), " << val << ");\n";
                break;
            case C_ADD_F32:
                li << "    " << cell << " = 0ULL;\n";
                lu << "    if (" << vk << ") " << cell << " = (unsigned long long)(unsigned)__float_as_int(__int_as_float((int)" << cell << ") + " << val << ");\n";
                gu << "    if (" << vk << ") atomicAdd((float*)(" << gp << "), " << val << ");\n";
                break;
            case C_ADD_I64:
                li << "    " << cell << " = 0ULL;\n";
                lu << "    if (" << vk << ") " << cell << " += (unsigned long long)" << val << ";\n";
                gu << "    if (" << vk << ") atomicAdd(" << gp << ", (unsigned long long)" << val << ");\n";
                break;
            case C_MIN:
                li << "    " << cell << " = 0xFFFFFFFFFFFFFFFFULL;\n";
                lu << "    if (" << vk << ") { const unsigned long long o_ = " << ordv << "; if (o_ < " << cell << ") " << cell << " = o_; }\n";
                gu << "    if (" << vk << ") atomicMin(" << gp << ", (unsigned long long)" << ordv << ");\n";
                break;
            case C_MAX:
                li << "    " << cell << " = 0ULL;\n";
                lu << "    if (" << vk << ") { const unsigned long long o_ = " << ordv << "; if (o_ > " << cell << ") " << cell << " = o_; }\n";
                gu << "    if (" << vk << ") atomicMax(" << gp << ", (unsigned long long)" << ordv << ");\n";
                break;
        }
        // flush: combine the 8 warps' copies for this lane, butterfly over lanes, one atomic per CTA
        fl << "    {\n";
        const std::string rd = "lcell[(size_t)w_ * lstride + " + std::to_string(c * 32) + " + lane]";
        switch (lc.op) {
            case C_ADD_F64:
                fl << "      double a_ = 0.0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += __longlong_as_double((long long)" << rd << ");\n";
                fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);\n";
                fl << "      if (lane == 0) atomicAdd((double*)(" << gp << "), a_);\n";
                break;
            case C_ADD_F32:
                fl << "      float a_ = 0.f; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += __int_as_float((int)" << rd << ");\n";
                fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);\n";
                fl << "      if (lane == 0) atomicAdd((float*)(" << gp << "), a_);\n";
                break;
            case C_ADD_I64:
                fl << "      unsigned long long a_ = 0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += " << rd << ";\n";
                fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);\n";
                fl << "      if (lane == 0) atomicAdd(" << gp << ", a_);\n";
                break;
            case C_MIN:
                fl << "      unsigned long long a_ = 0xFFFFFFFFFFFFFFFFULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = " << rd << "; a_ = t_ < a_ ? t_ : a_; }\n";
                fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ < a_ ? t_ : a_; }\n";
                fl << "      if (lane == 0) atomicMin(" << gp << ", a_);\n";
                break;
            case C_MAX:
                fl << "      unsigned long long a_ = 0ULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = " << rd << "; a_ = t_ > a_ ? t_ : a_; }\n";
                fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ > a_ ? t_ : a_; }\n";
                fl << "      if (lane == 0) atomicMax(" << gp << ", a_);\n";
                break;
        }
        fl << "    }\n";
    }
    // row count and first row: two more lane cells (plain read-modify-write, no ballots)
    {
        const std::string cn = "cells[" + std::to_string(NL * 32) + "]", cm = "cells[" + std::to_string((NL + 1) * 32) + "]";
        li << "    " << cn << " = 0ULL;\n    " << cm << " = 0xFFFFFFFFFFFFFFFFULL;\n";
        lu << "    " << cn << " += 1ULL;\n    if ((unsigned long long)row < " << cm << ") " << cm << " = (unsigned long long)row;\n";
        gu << "    atomicAdd(((unsigned long long*)p.ptr[" << p_acc[cell_nrows] << "]) + slot, 1ULL);\n";
        gu << "    atomicMin(((unsigned long long*)p.ptr[" << p_acc[cell_minrow] << "]) + slot, (unsigned long long)row);\n";
        sinit << "    t2c[" << NL << " * T2 + t] = 0ULL; t2c[" << (NL + 1) << " * T2 + t] = 0xFFFFFFFFFFFFFFFFULL;\n";
        su << "    atomicAdd(t2c + " << NL << " * T2 + t, 1ULL);\n    atomicMin(t2c + " << (NL + 1) << " * T2 + t, (unsigned long long)row);\n";
        sfl << "    atomicAdd(((unsigned long long*)p.ptr[" << p_acc[cell_nrows] << "]) + slot, t2c[" << NL << " * T2 + t]);\n";
        sfl << "    atomicMin(((unsigned long long*)p.ptr[" << p_acc[cell_minrow] << "]) + slot, t2c[" << (NL + 1) << " * T2 + t]);\n";
        fl << "    { unsigned long long a_ = 0; for (int w_ = 0; w_ < AGG_WARPS; w_++) a_ += lcell[(size_t)w_ * lstride + " << NL * 32 << " + lane];\n";
        fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);\n";
        fl << "      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[" << p_acc[cell_nrows] << "]) + slot, a_); }\n";
        fl << "    { unsigned long long a_ = 0xFFFFFFFFFFFFFFFFULL; for (int w_ = 0; w_ < AGG_WARPS; w_++) { unsigned long long t_ = lcell[(size_t)w_ * lstride + "
           << (NL + 1) * 32 << " + lane]; a_ = t_ < a_ ? t_ : a_; }\n";
        fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) { unsigned long long t_ = __shfl_xor_sync(FULL_MASK, a_, d_); a_ = t_ < a_ ? t_ : a_; }\n";
        fl << "      if (lane == 0) atomicMin(((unsigned long long*)p.ptr[" << p_acc[cell_minrow] << "]) + slot, a_); }\n";
    }
    // ballot count cells (valid / null counts of nullable aggregate inputs)
    for (int c = 0; c < NW; c++) {
        const WarpCell &w = warp_cells[c];
        const std::string vk = "v.k" + std::to_string(w.val_slot);
        const std::string cond = w.count_nulls ? "!" + vk : vk;
        wc << "    { const unsigned m_ = __ballot_sync(FULL_MASK, in_g && " << cond << "); if (lane == 0) wcell[" << c << "] += __popc(m_); }\n";
        gu << "    if (" << cond << ") atomicAdd(((unsigned long long*)p.ptr[" << p_acc[NLT + c] << "]) + slot, 1ULL);\n";
        sinit << "    t2c[" << (NLT + c) << " * T2 + t] = 0ULL;\n";
        su << "    if (" << cond << ") atomicAdd(t2c + " << (NLT + c) << " * T2 + t, 1ULL);\n";
        sfl << "    if (t2c[" << (NLT + c) << " * T2 + t]) atomicAdd(((unsigned long long*)p.ptr[" << p_acc[NLT + c] << "]) + slot, t2c[" << (NLT + c) << " * T2 + t]);\n";
        fl << "    { unsigned long long a_ = lane < AGG_WARPS ? wcell[(size_t)lane * wstride + " << c << "] : 0ULL;\n";
        fl << "      for (int d_ = 16; d_ > 0; d_ >>= 1) a_ += __shfl_xor_sync(FULL_MASK, a_, d_);\n";
        fl << "      if (lane == 0 && a_) atomicAdd(((unsigned long long*)p.ptr[" << p_acc[NLT + c] << "]) + slot, a_); }\n";
    }

    std::ostringstream o;
    if (SW != 8) o << "#define B200_AGG_WARPS " << SW << "\n";
    o << file_header(cg);
    o << "struct Prog {\n";
    o << "  typedef Params<" << meta.ptrs.size() << ", " << meta.ints.size() << "> P;\n";
    o << "  static constexpr int IV_N = " << iv_n << ", IV_MASK = " << iv_mask << ", PTR_TAG = " << p_hash << ", PTR_MINROW = " << p_acc[cell_minrow]
      << ", PTR_FLAG = " << p_ovf << ", PTR_ROWSLOT = " << p_rowslot << ", MODE = " << (exact_possible ? 0 : 2) << ";\n";
    o << "  static constexpr int ASTRIDE = " << astride << ", IV_RSTRIDE = " << iv_rstride << ";\n";
    o << "  static constexpr int NKEYS = " << nkeys << ", G1 = " << G1 << ", NLANE = " << NLT << ", NWARP = " << NW << ", U = " << U << ";\n";
    o << "  struct Raw {\n" << raw.str() << "    int pad_;\n  };\n";
    o << "  struct Keys {\n" << keys_s.str() << "    int pad_;\n  };\n";
    o << "  struct Vals {\n" << vals_s.str() << "    int pad_;\n  };\n";
    o << "  __device__ __forceinline__ static void load(const P& p, const int64_t row, Raw& r) {\n" << load.str() << "  }\n";
    o << "  __device__ __forceinline__ static void load2(const P& p, const int64_t row, Raw& r) {\n" << load2.str() << "  }\n";
    o << "  static constexpr bool PREFETCH = " << (pf_mode ? "true" : "false") << ";\n";
    o << "  static constexpr int PFDIST = " << std::max(1, std::min(8, env_int("B200_AGG_PFDIST", 1))) << ";\n";
    o << "  static constexpr int PF_LINES = " << pf_lines << ";\n";
    o << "  __device__ __forceinline__ static void prefetch_setup(const P& p, const int lane, PfLane& f) {\n"
      << "    f.b0 = nullptr; f.b1 = nullptr; f.w0 = 0; f.w1 = 0;\n" << pf.str() << "  }\n";
    o << "  __device__ __forceinline__ static bool eval(const P& p, const Raw& r, Keys& keys, Vals& vals) {\n" << ev.str() << "  }\n";
    o << "  __device__ __forceinline__ static void eval_keys(const P& p, const int64_t row, Keys& keys) {\n" << ek.str() << "  }\n";
    o << "  __device__ __forceinline__ static bool keys_equal(const Keys& a, const Keys& b) {\n" << eq.str() << "  }\n";
    {
        std::ostringstream fk;
        for (uint32_t i = 0; i < nkeys; i++)
            if (is_string(ast.nodes[key_roots[i]].dtype))
                fk << "    k.p" << i << " = k.k" << i << " ? str_pack8(k.v" << i << ") : 0ULL;\n";
        o << "  __device__ __forceinline__ static void finish_keys(Keys& k) {\n" << fk.str() << "  }\n";
    }
    o << "  __device__ __forceinline__ static uint64_t fingerprint(const Keys& k) {\n" << fps.str() << "  }\n";
    o << "  __device__ __forceinline__ static uint64_t code(const Keys& k) {\n" << cds.str() << "  }\n";
    o << "  __device__ __forceinline__ static void lane_init(unsigned long long* cells) {\n" << li.str() << "  }\n";
    o << "  __device__ __forceinline__ static void lane_update(unsigned long long* cells, const Vals& v, const int64_t row) {\n" << lu.str() << "  }\n";
    o << "  __device__ __forceinline__ static void global_update(const P& p, const long long slot, const Vals& v, const int64_t row) {\n" << gu.str() << "  }\n";
    {   // fix-up pass: re-fold the float min / max candidates over the rows after the group's last NaN
        std::ostringstream nf;
        for (const auto &ag : meta.aggs) {
            if (ag.nan_cell < 0) continue;
            const LaneCell &lc = lane_cells[ag.lane_cell];
            const std::string val = "v.a" + std::to_string(lc.val_slot), vk0 = "v.k" + std::to_string(lc.val_slot);
            nf << "    if (" << vk0 << " && !(" << val << " != " << val << ") && (unsigned long long)(row + 1) > ((const unsigned long long*)p.ptr["
               << p_acc[ag.nan_cell] << "])[slot]) " << (lc.op == C_MIN ? "atomicMin" : "atomicMax") << "(((unsigned long long*)p.ptr[" << p_acc[ag.lane_cell]
               << "]) + slot, (unsigned long long)" << ord_of(lc.val_dtype, val) << ");\n";
        }
        o << "  __device__ __forceinline__ static void nanfix_update(const P& p, const long long slot, const Vals& v, const int64_t row) {\n" << nf.str() << "  }\n";
    }
    o << "  static constexpr int T2 = " << t2_slots << ", NCELLS = " << (NLT + NW) << ";\n";
    o << "  __device__ __forceinline__ static void t2_init(unsigned long long* t2c, const int t) {\n" << sinit.str() << "  }\n";
    o << "  __device__ __forceinline__ static void t2_update(unsigned long long* t2c, const int t, const Vals& v, const int64_t row) {\n" << su.str() << "  }\n";
    o << "  __device__ __forceinline__ static void t2_flush(const P& p, const long long slot, const unsigned long long* t2c, const int t) {\n" << sfl.str() << "  }\n";
    o << "  __device__ __forceinline__ static void warp_cells(const Vals& v, const bool in_g, unsigned long long* wcell, const int lane) {\n" << wc.str() << "  }\n";
    o << "  __device__ __forceinline__ static void flush(const P& p, const long long slot, const unsigned long long* lcell, const size_t lstride,\n"
         "                                               const unsigned long long* wcell, const size_t wstride, const int lane) {\n" << fl.str() << "  }\n";
    if (staged_variant) {
        o << "  static constexpr int STAGE_ROWS = " << SR << ", STAGES = " << SS << ", STAGE_BYTES = " << stage_off << ", RPL = " << SRPL << ";\n";
        o << "  __device__ __forceinline__ static void stage_issue(const P& p, const uint32_t dst, const uint32_t bar, const int64_t base) {\n"
          << "    mbar_expect_tx(bar, " << stage_fixed_tx << "u);\n" << st_issue.str() << "  }\n";
        o << "  static constexpr bool LATE = " << (env_int("B200_AGG_STAGE_LATE", 1) ? "true" : "false") << ";\n";
        // one row (rr = its index inside the stage); the fixed-length string check is a vote of the row's warp
        o << "  __device__ __forceinline__ static void load_staged(const P& p, const unsigned char* st, const int64_t row, const int rr, Raw& r) {\n"
          << "    bool ok_ = true;\n"
          << ld_a.str()
          << "    const bool fast_ = __all_sync(FULL_MASK, ok_);\n"
          << ld_b.str() << "  }\n";
    }
    o << "};\n";
    // Ungrouped staged plans (Q6-shaped scans) whose tier-1 cells and stages fit three times into an SM are compiled for three
    // resident CTAs: the extra bytes in flight are what the staged build is for, and at 256 threads that needs <= 80 registers
    // (ptxas otherwise settles a few registers above that and silently drops to two CTAs).
    int staged_minb = MINB;
    if (staged_variant && SW == 8 && nkeys == 0 &&
        3 * (tier1_smem + (size_t)SS * (size_t)stage_off + (size_t)16 * SS + 1024) <= (size_t)227 * 1024)
        staged_minb = std::max(MINB, 3);
    if (staged_variant)
        o << "extern \"C\" __global__ void __launch_bounds__(" << SW * 32 << ", " << (SW == 16 ? 1 : staged_minb) << ") b200_agg_k(const __grid_constant__ Prog::P p) { agg_body_staged<Prog>(p); }\n";
    else if (nanfix_variant)
        o << "extern \"C\" __global__ void __launch_bounds__(256, 2) b200_agg_k(const __grid_constant__ Prog::P p) { agg_body_nanfix<Prog>(p); }\n";
    else if (rowslot_variant)
        o << "extern \"C\" __global__ void __launch_bounds__(256, 2) b200_agg_k(const __grid_constant__ Prog::P p) { agg_body_rowslot<Prog>(p); }\n";
    else if (global_variant)
        o << "extern \"C\" __global__ void __launch_bounds__(256, 2) b200_agg_k(const __grid_constant__ Prog::P p) { agg_body_global<Prog>(p); }\n";
    else
        o << "extern \"C\" __global__ void __launch_bounds__(256, " << MINB << ") b200_agg_k(const __grid_constant__ Prog::P p) { agg_body<Prog>(p); }\n";
    // shared memory: lane cells, ballot cells, tier-1 slots/ready flags/key values (generous bound for sizeof(Keys))
    meta.smem_bytes = (nanfix_variant || rowslot_variant) ? 0 : global_variant ? 8 * (size_t)t2_slots * (size_t)(1 + NLT + NW)
                                     : tier1_smem;
    meta.block_threads = SW * 32;
    if (staged_variant) {
        meta.smem_bytes += (size_t)SS * (size_t)stage_off + (size_t)16 * SS;
        meta.stage_bytes = stage_fits ? stage_off : 0;  // 0: the host runs the direct-load build instead
        meta.stage_rows = SR;
        if (!stage_fits) meta.smem_bytes = 0;
    }
    return o.str();
}

// Column indices a plan reads (structure walk only, no typing): operators gather just these when the input holds views.
static void scan_node(Reader &r, std::vector<int32_t> &cols) {
    const int tag = r.u8();
    switch (tag) {
        case T_COL: cols.push_back((int32_t)r.u32()); break;
        case T_LIT: {
            const int dt = r.u8();
            r.u8();
            if (is_string(dt)) r.str();
            else r.i64();
            break;
        }
        case T_BINARY: r.u8(); scan_node(r, cols); scan_node(r, cols); break;
        case T_UNARY: case T_CAST: r.u8(); scan_node(r, cols); break;
        case T_IF_ELSE: case T_BETWEEN: scan_node(r, cols); scan_node(r, cols); scan_node(r, cols); break;
        case T_IS_IN: {
            scan_node(r, cols);
            const uint32_t n = r.u32();
            for (uint32_t i = 0; i < n; i++) scan_node(r, cols);
            break;
        }
        case T_FILL_NULL: scan_node(r, cols); scan_node(r, cols); break;
        case T_FUNCTION: {
            r.u8();
            const int n = r.u8();
            for (int i = 0; i < n; i++) scan_node(r, cols);
            break;
        }
        default: fail("plan: unknown node tag");
    }
}
std::vector<int32_t> referenced_columns(int op, const uint8_t *plan, size_t plan_len) {
    Reader r{plan, plan_len};
    std::vector<int32_t> cols;
    if (op == B200_OP_PROJECT) {
        const uint32_t n = r.u32();
        for (uint32_t i = 0; i < n; i++) { r.str(); scan_node(r, cols); }
    } else if (op == B200_OP_FILTER) {
        scan_node(r, cols);
    } else if (op == B200_OP_AGG) {
        if (r.u8()) scan_node(r, cols);
        const uint32_t nk = r.u32();
        for (uint32_t i = 0; i < nk; i++) { r.str(); scan_node(r, cols); }
        const uint32_t na = r.u32();
        for (uint32_t i = 0; i < na; i++) { r.u8(); r.u8(); r.str(); scan_node(r, cols); }
    } else {
        fail("jit: unknown op");
    }
    std::sort(cols.begin(), cols.end());
    cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
    return cols;
}

std::vector<ProjectEntry> project_entries(const uint8_t *plan, size_t plan_len) {
    Reader r{plan, plan_len};
    std::vector<ProjectEntry> out;
    const uint32_t n = r.u32();
    for (uint32_t i = 0; i < n; i++) {
        ProjectEntry e;
        e.begin = r.pos;
        e.name = r.str();
        const size_t node = r.pos;
        std::vector<int32_t> cols;
        scan_node(r, cols);
        e.end = r.pos;
        if (plan[node] == T_COL && e.end - node == 5) memcpy(&e.bare_col, plan + node + 1, 4);
        out.push_back(std::move(e));
    }
    if (r.pos != r.n) fail("plan: trailing bytes");
    return out;
}

std::string generate(int op, const uint8_t *plan, size_t plan_len, const std::vector<int32_t> &col_dtypes,
                     const std::vector<uint8_t> &col_nullable, bool with_take, Compiled &meta, int variant) {
    Reader r{plan, plan_len};
    Ast ast;
    ast.col_dtypes = &col_dtypes;
    ast.col_nullable = &col_nullable;
    meta.with_take = with_take;
    std::string body;
    if (op == B200_OP_PROJECT) {
        body = gen_project(r, ast, meta, with_take);
        if (r.pos != r.n) fail("plan: trailing bytes");
    } else if (op == B200_OP_FILTER) {
        body = gen_filter(r, ast, meta);
        if (r.pos != r.n) fail("plan: trailing bytes");
    } else if (op == B200_OP_AGG) {
        body = gen_agg(r, ast, meta, variant);
    } else {
        fail("jit: unknown op");
    }
    return body;
}

// ------------------------------------------------------------------------------------------------
// NVRTC + cache
// ------------------------------------------------------------------------------------------------
static std::string g_cache_dir;
static std::mutex g_mu;

static uint64_t fnv1a(const std::string &s) {
    uint64_t h = 1469598103934665603ULL;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
    return h;
}
static std::string default_cache_dir() {
    Dl_info info;
    if (dladdr((void *)&default_cache_dir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t s = p.rfind('/');
        if (s != std::string::npos) return p.substr(0, s) + "/_jitcache";
    }
    return "/tmp/daft_b200_jitcache";
}
void set_cache_dir(const std::string &dir) {
    std::lock_guard<std::mutex> g(g_mu);
    g_cache_dir = dir;
}
static std::string cache_dir() {
    if (g_cache_dir.empty()) g_cache_dir = default_cache_dir();
    mkdir(g_cache_dir.c_str(), 0755);
    return g_cache_dir;
}

static std::atomic<unsigned long long> &prog_counter() {
    static std::atomic<unsigned long long> c{0};
    return c;
}
std::vector<char> compile_to_cubin(const std::string &body, std::string *path_out) {
    // leading "#define ..." lines of the generated code configure the prelude and go in front of it
    size_t defs_end = 0;
    while (body.compare(defs_end, 8, "#define ") == 0) defs_end = body.find('\n', defs_end) + 1;
    const std::string full = body.substr(0, defs_end) + join_chunks(kPreludeXxh3_chunks) + "\n" + join_chunks(kPreludeJit_chunks) + "\n" + body.substr(defs_end);
    char name[64];
    snprintf(name, sizeof name, "k_%016llx", (unsigned long long)fnv1a(full));
    const std::string base = cache_dir() + "/" + name;
    if (path_out) *path_out = base + ".cu";
    // Cache file = cubin bytes followed by a 16-byte trailer {magic, FNV-1a of the bytes}: a file another process is
    // still writing, or one cut short, fails the check and is recompiled instead of being handed to the loader.
    static const uint64_t kCubinMagic = 0x4232303043554231ULL;  // "B200CUB1"
    {
        std::ifstream f(base + ".cubin", std::ios::binary);
        if (f) {
            std::vector<char> buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            if (buf.size() > 16) {
                uint64_t tr[2];
                memcpy(tr, buf.data() + buf.size() - 16, 16);
                buf.resize(buf.size() - 16);
                if (tr[0] == kCubinMagic && tr[1] == fnv1a(std::string(buf.begin(), buf.end()))) return buf;
            }
        }
    }
    const std::string uniq = "." + std::to_string((long long)getpid()) + "." + std::to_string(prog_counter()++) + ".tmp";
    {   // several ranks may generate the same plan at once: private temporary names, atomic renames
        const std::string tmp = base + ".cu" + uniq;
        {
            std::ofstream f(tmp);
            f << full;
        }
        if (rename(tmp.c_str(), (base + ".cu").c_str()) != 0) remove(tmp.c_str());
    }
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, full.c_str(), (base + ".cu").c_str(), 0, nullptr, nullptr) != NVRTC_SUCCESS)  // full path: ncu --import-source
        fail("nvrtcCreateProgram failed");
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--fmad=false", "-default-device"};
    nvrtcResult res = nvrtcCompileProgram(prog, 5, opts);
    if (res != NVRTC_SUCCESS) {
        size_t ls = 0;
        nvrtcGetProgramLogSize(prog, &ls);
        std::string log(ls, '\0');
        nvrtcGetProgramLog(prog, &log[0]);
        nvrtcDestroyProgram(&prog);
        fail("JIT compile failed for " + base + ".cu:\n" + log);
    }
    size_t cs = 0;
    nvrtcGetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    nvrtcGetCUBIN(prog, cubin.data());
    nvrtcDestroyProgram(&prog);
    {
        const std::string tmp = base + ".cubin" + uniq;
        bool ok;
        {
            std::ofstream f(tmp, std::ios::binary);
            const uint64_t tr[2] = {kCubinMagic, fnv1a(std::string(cubin.begin(), cubin.end()))};
            f.write(cubin.data(), (std::streamsize)cubin.size());
            f.write(reinterpret_cast<const char *>(tr), 16);
            ok = (bool)f;
        }
        if (!ok || rename(tmp.c_str(), (base + ".cubin").c_str()) != 0) remove(tmp.c_str());  // the cache is best effort
    }
    return cubin;
}

static std::map<std::string, std::unique_ptr<Compiled>> g_kernels;

static std::string cache_key(int op, const uint8_t *plan, size_t plan_len, const std::vector<int32_t> &dt,
                             const std::vector<uint8_t> &nl, bool with_take) {
    std::string k;
    k.push_back((char)op);
    k.push_back((char)with_take);
    k.append(reinterpret_cast<const char *>(plan), plan_len);
    k.push_back('|');
    for (size_t i = 0; i < dt.size(); i++) { k.push_back((char)dt[i]); k.push_back((char)nl[i]); }
    return k;
}

const Compiled &get(int op, const uint8_t *plan, size_t plan_len, const Batch &input, bool with_take, int variant) {
    std::vector<int32_t> dt;
    std::vector<uint8_t> nl;
    for (auto &c : input.cols) {
        dt.push_back(c.dtype);
        uint8_t f = c.nullable() ? 1 : 0;
        if (op == B200_OP_AGG && is_string(c.dtype) && c.length > 0 && c.data_bytes % c.length == 0) {
            const int64_t L = c.data_bytes / c.length;  // fixed-length hint (see Ast::fixed_len)
            if (L >= 1 && L <= 4) f |= (uint8_t)(L << 1);
        }
        nl.push_back(f);
    }
    std::lock_guard<std::mutex> g(g_mu);
    const std::string key = cache_key(op, plan, plan_len, dt, nl, with_take) + (char)('0' + variant);
    auto it = g_kernels.find(key);
    if (it != g_kernels.end()) return *it->second;
    auto c = std::make_unique<Compiled>();
    const std::string src = generate(op, plan, plan_len, dt, nl, with_take, *c, variant);
    if (op == B200_OP_AGG && variant == 2 && c->stage_bytes == 0) {  // plan not worth staging: nothing to compile
        auto &none = *c;
        g_kernels[key] = std::move(c);
        return none;
    }
    std::vector<char> cubin = compile_to_cubin(src, &c->source_path);
    require_device();
    // runtime "library" API (cudart dlopens the driver itself; the .so has no link-time libcuda dependency)
    B200_CUDA(cudaLibraryLoadData(&c->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    B200_CUDA(cudaLibraryGetKernel(&c->fn, c->lib, op == B200_OP_AGG ? "b200_agg_k" : (op == B200_OP_FILTER ? "b200_filter_k" : "b200_project_k")));
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, (const void *)c->fn) == cudaSuccess) c->regs = fa.numRegs;
    if (c->smem_bytes > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute((const void *)c->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
    auto &ref = *c;
    g_kernels[key] = std::move(c);
    return ref;
}

static void precompile_one(int op, const uint8_t *plan, size_t plan_len, const std::vector<int32_t> &col_dtypes,
                           const std::vector<uint8_t> &col_nullable);
void precompile(int op, const uint8_t *plan, size_t plan_len, const std::vector<int32_t> &col_dtypes,
                const std::vector<uint8_t> &col_nullable) {
    precompile_one(op, plan, plan_len, col_dtypes, col_nullable);
    if (op == B200_OP_AGG) {  // and the build specialised for 1-byte string columns (TPC-H flags and codes)
        std::vector<uint8_t> f = col_nullable;
        bool any = false;
        for (size_t i = 0; i < f.size(); i++)
            if (is_string(col_dtypes[i])) { f[i] |= 1 << 1; any = true; }
        if (any) precompile_one(op, plan, plan_len, col_dtypes, f);
    }
}
static void precompile_one(int op, const uint8_t *plan, size_t plan_len, const std::vector<int32_t> &col_dtypes,
                           const std::vector<uint8_t> &col_nullable) {
    std::lock_guard<std::mutex> g(g_mu);
    Compiled c;
    const std::string src = generate(op, plan, plan_len, col_dtypes, col_nullable, false, c, 0);
    compile_to_cubin(src, nullptr);
    if (op == B200_OP_AGG) {  // the high-cardinality variant of the same plan
        Compiled cg2;
        compile_to_cubin(generate(op, plan, plan_len, col_dtypes, col_nullable, false, cg2, 1), nullptr);
        Compiled cg3;  // and the TMA-staged build of the low-cardinality kernel (when the plan qualifies)
        const std::string src3 = generate(op, plan, plan_len, col_dtypes, col_nullable, false, cg3, 2);
        if (cg3.stage_bytes > 0) compile_to_cubin(src3, nullptr);
        if (c.has_nan_fold) {  // the float min / max fix-up pass (variant 3)
            Compiled cg4;
            compile_to_cubin(generate(op, plan, plan_len, col_dtypes, col_nullable, false, cg4, 3), nullptr);
        }
    }
    if (op == B200_OP_AGG && !c.anyval_project_plan.empty()) {
        Compiled c3;
        compile_to_cubin(generate(B200_OP_PROJECT, c.anyval_project_plan.data(), c.anyval_project_plan.size(), col_dtypes, col_nullable, true, c3, 0), nullptr);
    }
    if (op == B200_OP_AGG && !c.listagg_project_plan.empty()) {
        Compiled c5, c6;
        compile_to_cubin(generate(op, plan, plan_len, col_dtypes, col_nullable, false, c5, 4), nullptr);
        compile_to_cubin(generate(B200_OP_PROJECT, c.listagg_project_plan.data(), c.listagg_project_plan.size(), col_dtypes, col_nullable, false, c6, 0), nullptr);
    }
    if (op == B200_OP_AGG && !c.key_project_plan.empty() && c.nkeys > 0) {
        Compiled c2;
        const std::string src2 = generate(B200_OP_PROJECT, c.key_project_plan.data(), c.key_project_plan.size(), col_dtypes, col_nullable, true, c2, 0);
        compile_to_cubin(src2, nullptr);
    }
}

static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static bool g_timing = false, g_have_sample = false;
void enable_kernel_timing(bool on) { g_timing = on; }
double last_timed_kernel_ms() {
    if (!g_have_sample) return -1.0;
    float ms = 0.f;
    B200_CUDA(cudaEventSynchronize(g_ev1));
    B200_CUDA(cudaEventElapsedTime(&ms, g_ev0, g_ev1));
    return (double)ms;
}

void launch(const Compiled &k, const LaunchArgs &a, int grid, int block, size_t smem, int pass, bool timed) {
    // parameter block layout == jit_prelude.cuh Params<NPTR,NINT>
    const size_t np = std::max<size_t>(k.ptrs.size(), 1), ni = std::max<size_t>(k.ints.size(), 1);
    std::vector<uint64_t> block_buf(np + ni, 0);
    for (size_t i = 0; i < a.ptr.size(); i++) block_buf[i] = (uint64_t)(uintptr_t)a.ptr[i];
    for (size_t i = 0; i < a.ival.size(); i++) block_buf[np + i] = (uint64_t)a.ival[i];
    void *args[2] = {block_buf.data(), &pass};
    const bool t = timed && g_timing;
    if (t) {
        if (!g_ev0) {
            B200_CUDA(cudaEventCreate(&g_ev0));
            B200_CUDA(cudaEventCreate(&g_ev1));
        }
        B200_CUDA(cudaEventRecord(g_ev0, ctx().stream));
    }
    B200_CUDA(cudaLaunchKernel((const void *)k.fn, dim3((unsigned)grid), dim3((unsigned)block), args, smem, ctx().stream));
    if (t) {
        B200_CUDA(cudaEventRecord(g_ev1, ctx().stream));
        g_have_sample = true;
    }
    count_launch();
}

}  // namespace jit
}  // namespace b200

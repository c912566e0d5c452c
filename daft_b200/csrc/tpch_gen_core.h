// Counter-based, dbgen-shaped TPC-H value generator: pure functions of (column, seed, row), shared by the device
// generator (tpch_gen.cu), the host generator behind b200_tpch_generate_host, and the standalone CPU engine under
// oracle/ (which includes this header so that the CPU arm sees the same tables without loading libdaftb200.so).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <strings.h>

#ifndef B200_I64  // standalone use: the physical dtype codes of include/daft_b200.h
#define B200_I64 5
#define B200_F64 11
#define B200_DATE32 12
#define B200_LARGE_UTF8 14
#endif

namespace b200 {
namespace tpchgen {

#ifdef __CUDACC__
#define GEN_HD __host__ __device__ __forceinline__
#else
#define GEN_HD inline
#endif

GEN_HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
GEN_HD uint64_t rnd(uint64_t seed, uint32_t stream, uint64_t idx) { return mix64(mix64(seed ^ ((uint64_t)stream << 40)) ^ idx); }
GEN_HD int64_t rnd_range(uint64_t seed, uint32_t stream, uint64_t idx, int64_t lo, int64_t hi) {
    return lo + (int64_t)(rnd(seed, stream, idx) % (uint64_t)(hi - lo + 1));
}

constexpr int32_t DAY_1992_01_01 = 8035;
constexpr int32_t DAY_1998_08_02 = 10440;  // last order date (ENDDATE - 151 days)
constexpr int32_t DAY_1995_06_17 = 9298;   // dbgen CURRENTDATE

struct Scale {
    int64_t n_orders, n_customers, n_suppliers, n_parts;
};
inline Scale scale_of(double sf) {
    Scale s;
    s.n_orders = std::max<int64_t>(1, (int64_t)(1500000.0 * sf + 0.5));
    s.n_customers = std::max<int64_t>(1, (int64_t)(150000.0 * sf + 0.5));
    s.n_suppliers = std::max<int64_t>(1, (int64_t)(10000.0 * sf + 0.5));
    s.n_parts = std::max<int64_t>(1, (int64_t)(200000.0 * sf + 0.5));
    return s;
}

// lineitem row -> (order index, line number); 7 consecutive orders carry 4,1,7,2,6,3,5 lines = 28
GEN_HD void line_of(int64_t row, int64_t n_orders, int64_t &oidx, int32_t &lno) {
    const int64_t blk = row / 28;
    const int r = (int)(row % 28);
    int o, l;
    if (r < 4) { o = 0; l = r; }
    else if (r < 5) { o = 1; l = r - 4; }
    else if (r < 12) { o = 2; l = r - 5; }
    else if (r < 14) { o = 3; l = r - 12; }
    else if (r < 20) { o = 4; l = r - 14; }
    else if (r < 23) { o = 5; l = r - 20; }
    else { o = 6; l = r - 23; }
    oidx = (blk * 7 + o) % n_orders;
    lno = l + 1;
}
GEN_HD int64_t order_key(int64_t oidx) { return (oidx / 8) * 32 + (oidx % 8) + 1; }  // dbgen's sparse order keys
GEN_HD int32_t order_date(uint64_t seed, int64_t oidx) {
    return (int32_t)rnd_range(seed, 11, (uint64_t)oidx, DAY_1992_01_01, DAY_1998_08_02);
}
// the i-th (0..3) supplier of a part: dbgen's PART_SUPP_BRIDGE (specification 4.2.3, PS_SUPPKEY), S = number of suppliers
GEN_HD int64_t part_supplier(int64_t partkey, int i, int64_t S) { return (partkey + (int64_t)i * (S / 4 + (partkey - 1) / S)) % S + 1; }
// 'F' / 'O' / 'P' from the ship dates of the order's lines in the 28-line pattern (all shipped / none / some)
GEN_HD char order_status(uint64_t seed, int64_t oidx) {
    const int o = (int)(oidx % 7);
    const int start = o == 0 ? 0 : o == 1 ? 4 : o == 2 ? 5 : o == 3 ? 12 : o == 4 ? 14 : o == 5 ? 20 : 23;
    const int count = o == 0 ? 4 : o == 1 ? 1 : o == 2 ? 7 : o == 3 ? 2 : o == 4 ? 6 : o == 5 ? 3 : 5;
    const int32_t odate = order_date(seed, oidx);
    int open = 0;
    for (int l = 0; l < count; l++) {
        const int64_t row = (oidx / 7) * 28 + start + l;
        open += odate + (int32_t)rnd_range(seed, 21, (uint64_t)row, 1, 121) > DAY_1995_06_17;
    }
    return open == 0 ? 'F' : open == count ? 'O' : 'P';
}

enum ColId {
    L_ORDERKEY = 0, L_PARTKEY, L_SUPPKEY, L_LINENUMBER, L_QUANTITY, L_EXTENDEDPRICE, L_DISCOUNT, L_TAX, L_RETURNFLAG,
    L_LINESTATUS, L_SHIPDATE, L_COMMITDATE, L_RECEIPTDATE,
    O_ORDERKEY, O_CUSTKEY, O_ORDERDATE, O_TOTALPRICE, O_SHIPPRIORITY,
    C_CUSTKEY, C_NATIONKEY, C_ACCTBAL,
    S_SUPPKEY, S_NATIONKEY, S_ACCTBAL,
    N_NATIONKEY, N_REGIONKEY, N_NAME,
    R_REGIONKEY, R_NAME,
    // dictionary-valued strings (round 2): the generator draws an index, the column holds the dictionary word
    L_SHIPMODE, O_ORDERPRIORITY, C_MKTSEGMENT,
    P_PARTKEY, P_SIZE, P_TYPE,
    // the columns the rest of benchmarking/tpch/answers.py reads (Q2, Q7-Q11, Q13, Q15-Q17, Q19-Q22)
    L_SHIPINSTRUCT, O_ORDERSTATUS, O_COMMENT,
    C_NAME, C_ADDRESS, C_PHONE, C_COMMENT,
    S_NAME, S_ADDRESS, S_PHONE, S_COMMENT,
    P_NAME, P_MFGR, P_BRAND, P_CONTAINER, P_RETAILPRICE,
    PS_PARTKEY, PS_SUPPKEY, PS_AVAILQTY, PS_SUPPLYCOST,
    NUM_COLS
};
struct ColInfo { const char *name; int32_t table; int32_t dtype; };
static const ColInfo kCols[NUM_COLS] = {
    {"L_ORDERKEY", 0, B200_I64}, {"L_PARTKEY", 0, B200_I64}, {"L_SUPPKEY", 0, B200_I64}, {"L_LINENUMBER", 0, B200_I64},
    {"L_QUANTITY", 0, B200_I64}, {"L_EXTENDEDPRICE", 0, B200_F64}, {"L_DISCOUNT", 0, B200_F64}, {"L_TAX", 0, B200_F64},
    {"L_RETURNFLAG", 0, B200_LARGE_UTF8}, {"L_LINESTATUS", 0, B200_LARGE_UTF8}, {"L_SHIPDATE", 0, B200_DATE32},
    {"L_COMMITDATE", 0, B200_DATE32}, {"L_RECEIPTDATE", 0, B200_DATE32},
    {"O_ORDERKEY", 1, B200_I64}, {"O_CUSTKEY", 1, B200_I64}, {"O_ORDERDATE", 1, B200_DATE32}, {"O_TOTALPRICE", 1, B200_F64},
    {"O_SHIPPRIORITY", 1, B200_I64},
    {"C_CUSTKEY", 2, B200_I64}, {"C_NATIONKEY", 2, B200_I64}, {"C_ACCTBAL", 2, B200_F64},
    {"S_SUPPKEY", 3, B200_I64}, {"S_NATIONKEY", 3, B200_I64}, {"S_ACCTBAL", 3, B200_F64},
    {"N_NATIONKEY", 4, B200_I64}, {"N_REGIONKEY", 4, B200_I64}, {"N_NAME", 4, B200_LARGE_UTF8},
    {"R_REGIONKEY", 5, B200_I64}, {"R_NAME", 5, B200_LARGE_UTF8},
    {"L_SHIPMODE", 0, B200_LARGE_UTF8}, {"O_ORDERPRIORITY", 1, B200_LARGE_UTF8}, {"C_MKTSEGMENT", 2, B200_LARGE_UTF8},
    {"P_PARTKEY", 6, B200_I64}, {"P_SIZE", 6, B200_I64}, {"P_TYPE", 6, B200_LARGE_UTF8},
    {"L_SHIPINSTRUCT", 0, B200_LARGE_UTF8}, {"O_ORDERSTATUS", 1, B200_LARGE_UTF8}, {"O_COMMENT", 1, B200_LARGE_UTF8},
    {"C_NAME", 2, B200_LARGE_UTF8}, {"C_ADDRESS", 2, B200_LARGE_UTF8}, {"C_PHONE", 2, B200_LARGE_UTF8}, {"C_COMMENT", 2, B200_LARGE_UTF8},
    {"S_NAME", 3, B200_LARGE_UTF8}, {"S_ADDRESS", 3, B200_LARGE_UTF8}, {"S_PHONE", 3, B200_LARGE_UTF8}, {"S_COMMENT", 3, B200_LARGE_UTF8},
    {"P_NAME", 6, B200_LARGE_UTF8}, {"P_MFGR", 6, B200_LARGE_UTF8}, {"P_BRAND", 6, B200_LARGE_UTF8}, {"P_CONTAINER", 6, B200_LARGE_UTF8},
    {"P_RETAILPRICE", 6, B200_F64},
    {"PS_PARTKEY", 7, B200_I64}, {"PS_SUPPKEY", 7, B200_I64}, {"PS_AVAILQTY", 7, B200_I64}, {"PS_SUPPLYCOST", 7, B200_F64},
};
// dictionary columns: number of words (TPC-H specification 4.2.2.13: ship modes, order priorities, market segments;
// P_TYPE = one syllable of each of three lists, 6 x 5 x 5 = 150 combinations)
static const char *kShipModes[7] = {"REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"};
static const char *kPriorities[5] = {"1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"};
static const char *kSegments[5] = {"AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"};
static const char *kTypeS1[6] = {"STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO"};
static const char *kTypeS2[5] = {"ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED"};
static const char *kTypeS3[5] = {"TIN", "NICKEL", "BRASS", "STEEL", "COPPER"};
static const char *kInstructs[4] = {"DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"};
// P_CONTAINER = one syllable of each of two lists (5 x 8), P_BRAND = Brand#MN with M the manufacturer (specification 4.2.3)
static const char *kContS1[5] = {"SM", "LG", "MED", "JUMBO", "WRAP"};
static const char *kContS2[8] = {"CASE", "BOX", "BAG", "JAR", "PKG", "PACK", "CAN", "DRUM"};
inline bool is_dict_col(int col) {
    return col == L_SHIPMODE || col == O_ORDERPRIORITY || col == C_MKTSEGMENT || col == P_TYPE || col == L_SHIPINSTRUCT || col == P_MFGR ||
           col == P_BRAND || col == P_CONTAINER;
}
inline int dict_size(int col) {
    switch (col) {
        case L_SHIPMODE: return 7;
        case P_TYPE: return 150;
        case L_SHIPINSTRUCT: return 4;
        case P_BRAND: return 25;
        case P_CONTAINER: return 40;
        default: return 5;
    }
}
inline std::string dict_word(int col, int k) {
    switch (col) {
        case L_SHIPMODE: return kShipModes[k];
        case O_ORDERPRIORITY: return kPriorities[k];
        case C_MKTSEGMENT: return kSegments[k];
        case L_SHIPINSTRUCT: return kInstructs[k];
        case P_MFGR: return std::string("Manufacturer#") + (char)('1' + k);
        case P_BRAND: return std::string("Brand#") + (char)('1' + k / 5) + (char)('1' + k % 5);
        case P_CONTAINER: return std::string(kContS1[k / 8]) + " " + kContS2[k % 8];
        default: return std::string(kTypeS1[k / 25]) + " " + kTypeS2[(k / 5) % 5] + " " + kTypeS3[k % 5];
    }
}
// the dictionary index of a dictionary column's row
GEN_HD int gen_dict_index(int col, uint64_t seed, int64_t row) {
    switch (col) {
        case L_SHIPMODE: return (int)(rnd(seed, 61, (uint64_t)row) % 7);
        case O_ORDERPRIORITY: return (int)(rnd(seed, 62, (uint64_t)row) % 5);
        case C_MKTSEGMENT: return (int)(rnd(seed, 63, (uint64_t)row) % 5);
        case L_SHIPINSTRUCT: return (int)(rnd(seed, 66, (uint64_t)row) % 4);
        case P_BRAND: return (int)(rnd(seed, 67, (uint64_t)row) % 25);
        case P_MFGR: return (int)(rnd(seed, 67, (uint64_t)row) % 25) / 5;  // the brand's manufacturer
        case P_CONTAINER: return (int)(rnd(seed, 68, (uint64_t)row) % 40);
        default: return (int)(rnd(seed, 64, (uint64_t)row) % 150);
    }
}
static const char *kNations[25] = {"ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE", "GERMANY", "INDIA",
                                   "INDONESIA", "IRAN", "IRAQ", "JAPAN", "JORDAN", "KENYA", "MOROCCO", "MOZAMBIQUE", "PERU", "CHINA",
                                   "ROMANIA", "SAUDI ARABIA", "VIETNAM", "RUSSIA", "UNITED KINGDOM", "UNITED STATES"};
static const int kNationRegion[25] = {0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1};
static const char *kRegions[5] = {"AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"};

// one fixed-width value (returned as 8 raw bytes); flag columns return the ASCII byte
GEN_HD uint64_t gen_value(int col, uint64_t seed, int64_t row, Scale sc) {
    union { double d; uint64_t u; int64_t i; } v;
    v.u = 0;
    if (col <= L_RECEIPTDATE) {
        int64_t oidx;
        int32_t lno;
        line_of(row, sc.n_orders, oidx, lno);
        const int32_t odate = order_date(seed, oidx);
        const int32_t ship = odate + (int32_t)rnd_range(seed, 21, (uint64_t)row, 1, 121);
        switch (col) {
            case L_ORDERKEY: v.i = order_key(oidx); break;
            case L_PARTKEY: v.i = rnd_range(seed, 22, (uint64_t)row, 1, sc.n_parts); break;
            case L_SUPPKEY:  // one of the part's four suppliers, as dbgen draws it: (L_PARTKEY, L_SUPPKEY) is a partsupp key
                v.i = part_supplier(rnd_range(seed, 22, (uint64_t)row, 1, sc.n_parts), (int)(rnd(seed, 23, (uint64_t)row) % 4), sc.n_suppliers);
                break;
            case L_LINENUMBER: v.i = lno; break;
            case L_QUANTITY: v.i = rnd_range(seed, 24, (uint64_t)row, 1, 50); break;
            case L_EXTENDEDPRICE: {
                const int64_t pk = rnd_range(seed, 22, (uint64_t)row, 1, sc.n_parts);
                const int64_t cents = 90000 + ((pk / 10) % 20001) + 100 * (pk % 1000);
                const int64_t qty = rnd_range(seed, 24, (uint64_t)row, 1, 50);
                v.d = (double)(qty * cents) / 100.0;
                break;
            }
            case L_DISCOUNT: v.d = (double)rnd_range(seed, 25, (uint64_t)row, 0, 10) / 100.0; break;
            case L_TAX: v.d = (double)rnd_range(seed, 26, (uint64_t)row, 0, 8) / 100.0; break;
            case L_SHIPDATE: v.i = ship; break;
            case L_COMMITDATE: v.i = odate + (int32_t)rnd_range(seed, 27, (uint64_t)row, 30, 90); break;
            case L_RECEIPTDATE: v.i = ship + (int32_t)rnd_range(seed, 28, (uint64_t)row, 1, 30); break;
            case L_RETURNFLAG: {
                const int32_t receipt = ship + (int32_t)rnd_range(seed, 28, (uint64_t)row, 1, 30);
                v.u = receipt <= DAY_1995_06_17 ? ((rnd(seed, 29, (uint64_t)row) & 1) ? 'R' : 'A') : 'N';
                break;
            }
            case L_LINESTATUS: v.u = ship > DAY_1995_06_17 ? 'O' : 'F'; break;
        }
        return v.u;
    }
    switch (col) {
        case O_ORDERKEY: v.i = order_key(row); break;
        case O_CUSTKEY: {
            int64_t ck = rnd_range(seed, 31, (uint64_t)row, 1, sc.n_customers);
            if (ck % 3 == 0) ck = ck > 1 ? ck - 1 : ck + 1;  // dbgen: a third of the customers never order
            v.i = ck;
            break;
        }
        case O_ORDERDATE: v.i = order_date(seed, row); break;
        case O_TOTALPRICE: v.d = (double)rnd_range(seed, 32, (uint64_t)row, 85000, 55000000) / 100.0; break;
        case O_SHIPPRIORITY: v.i = 0; break;
        case C_CUSTKEY: case S_SUPPKEY: v.i = row + 1; break;
        case C_NATIONKEY: v.i = rnd_range(seed, 41, (uint64_t)row, 0, 24); break;
        case C_ACCTBAL: v.d = (double)rnd_range(seed, 42, (uint64_t)row, -99999, 999999) / 100.0; break;
        case S_NATIONKEY: v.i = rnd_range(seed, 51, (uint64_t)row, 0, 24); break;
        case S_ACCTBAL: v.d = (double)rnd_range(seed, 52, (uint64_t)row, -99999, 999999) / 100.0; break;
        case N_NATIONKEY: case R_REGIONKEY: v.i = row; break;
        case P_PARTKEY: v.i = row + 1; break;
        case P_SIZE: v.i = rnd_range(seed, 65, (uint64_t)row, 1, 50); break;
        case P_RETAILPRICE: {
            const int64_t pk = row + 1;
            v.d = (double)(90000 + ((pk / 10) % 20001) + 100 * (pk % 1000)) / 100.0;
            break;
        }
        case O_ORDERSTATUS: v.u = (uint64_t)order_status(seed, row); break;
        case PS_PARTKEY: v.i = row / 4 + 1; break;
        case PS_SUPPKEY: v.i = part_supplier(row / 4 + 1, (int)(row % 4), sc.n_suppliers); break;
        case PS_AVAILQTY: v.i = rnd_range(seed, 81, (uint64_t)row, 1, 9999); break;
        case PS_SUPPLYCOST: v.d = (double)rnd_range(seed, 82, (uint64_t)row, 100, 100000) / 100.0; break;
        default: break;
    }
    return v.u;
}

// kind: 0 = 8-byte value, 1 = 4-byte value, 2 = flag string (offsets = i, one byte per row)
GEN_HD void gen_store(int col, int kind, uint64_t seed, int64_t row, int64_t i, Scale sc, void *data, int64_t *offsets) {
    const uint64_t u = gen_value(col, seed, row, sc);
    if (kind == 0) ((uint64_t *)data)[i] = u;
    else if (kind == 1) ((int32_t *)data)[i] = (int32_t)(int64_t)u;
    else {
        ((uint8_t *)data)[i] = (uint8_t)u;
        offsets[i] = i;
    }
}

// ---- generated strings: names, phones, addresses, comments (a pure function of column, seed and row, like every other value) ----
// out == nullptr counts only.  The longest value is 117 bytes (C_COMMENT).
struct StrOut {
    uint8_t *p;
    int n;
};
GEN_HD void put(StrOut &s, char c) {
    if (s.p) s.p[s.n] = (uint8_t)c;
    s.n++;
}
GEN_HD void put_str(StrOut &s, const char *w) {
    while (*w) put(s, *w++);
}
GEN_HD void put_num(StrOut &s, int64_t v, int width) {  // zero-padded decimal
    int64_t p10 = 1;
    for (int i = 1; i < width; i++) p10 *= 10;
    for (; p10 > 0; p10 /= 10) put(s, (char)('0' + (v / p10) % 10));
}
GEN_HD int c_strlen(const char *w) {
    int n = 0;
    while (w[n]) n++;
    return n;
}
// the text grammar's vocabulary (specification 4.2.2.14, a subset): comments are words of this list separated by blanks
constexpr int kNumWords = 48;
GEN_HD const char *text_word(int k) {
    switch (k) {
        case 0: return "furiously"; case 1: return "special"; case 2: return "requests"; case 3: return "packages";
        case 4: return "deposits"; case 5: return "accounts"; case 6: return "pending"; case 7: return "ironic";
        case 8: return "express"; case 9: return "regular"; case 10: return "final"; case 11: return "bold";
        case 12: return "slyly"; case 13: return "carefully"; case 14: return "blithely"; case 15: return "quickly";
        case 16: return "foxes"; case 17: return "ideas"; case 18: return "theodolites"; case 19: return "pinto";
        case 20: return "beans"; case 21: return "instructions"; case 22: return "dependencies"; case 23: return "excuses";
        case 24: return "platelets"; case 25: return "asymptotes"; case 26: return "courts"; case 27: return "dolphins";
        case 28: return "sleep"; case 29: return "wake"; case 30: return "are"; case 31: return "cajole";
        case 32: return "haggle"; case 33: return "nag"; case 34: return "use"; case 35: return "boost";
        case 36: return "among"; case 37: return "above"; case 38: return "the"; case 39: return "even";
        case 40: return "unusual"; case 41: return "silent"; case 42: return "daring"; case 43: return "according";
        case 44: return "to"; case 45: return "after"; case 46: return "never"; default: return "along";
    }
}
// P_NAME: distinct-looking colour words (specification 4.2.3 lists 92; 32 of them here), five per name
constexpr int kNumColors = 32;
GEN_HD const char *color_word(int k) {
    switch (k) {
        case 0: return "almond"; case 1: return "antique"; case 2: return "aquamarine"; case 3: return "azure";
        case 4: return "beige"; case 5: return "bisque"; case 6: return "black"; case 7: return "blanched";
        case 8: return "blue"; case 9: return "blush"; case 10: return "brown"; case 11: return "burlywood";
        case 12: return "chartreuse"; case 13: return "chocolate"; case 14: return "coral"; case 15: return "cornflower";
        case 16: return "cream"; case 17: return "cyan"; case 18: return "dark"; case 19: return "dodger";
        case 20: return "firebrick"; case 21: return "forest"; case 22: return "frosted"; case 23: return "ghost";
        case 24: return "goldenrod"; case 25: return "green"; case 26: return "honeydew"; case 27: return "hot";
        case 28: return "indian"; case 29: return "ivory"; case 30: return "khaki"; default: return "lace";
    }
}
GEN_HD void put_text(StrOut &s, uint64_t seed, uint32_t stream, int64_t row, int lo, int hi) {
    const int target = lo + (int)(rnd(seed, stream, (uint64_t)row * 32) % (uint64_t)(hi - lo + 1));
    const int begin = s.n;
    for (int j = 1; j < 32; j++) {
        const char *w = text_word((int)(rnd(seed, stream, (uint64_t)row * 32 + (uint64_t)j) % kNumWords));
        if (s.n - begin + c_strlen(w) + 1 > target) break;
        if (s.n > begin) put(s, ' ');
        put_str(s, w);
    }
}
GEN_HD void put_phone(StrOut &s, uint64_t seed, uint32_t stream, int64_t row, int64_t nationkey) {
    const uint64_t r = rnd(seed, stream, (uint64_t)row);
    put_num(s, nationkey + 10, 2);
    put(s, '-');
    put_num(s, 100 + (int64_t)(r % 900), 3);
    put(s, '-');
    put_num(s, 100 + (int64_t)((r >> 16) % 900), 3);
    put(s, '-');
    put_num(s, 1000 + (int64_t)((r >> 32) % 9000), 4);
}
GEN_HD void put_address(StrOut &s, uint64_t seed, uint32_t stream, int64_t row) {
    const int len = 10 + (int)(rnd(seed, stream, (uint64_t)row * 8) % 31);
    for (int j = 0; j < len; j++) {
        const uint64_t r = rnd(seed, stream, (uint64_t)row * 8 + 1 + (uint64_t)(j / 8)) >> (8 * (j % 8));
        const int k = (int)(r & 63);
        put(s, k < 26 ? (char)('a' + k) : k < 52 ? (char)('A' + k - 26) : k < 62 ? (char)('0' + k - 52) : k == 62 ? ' ' : ',');
    }
}
inline bool is_generated_string(int col) {
    return col == O_COMMENT || col == C_NAME || col == C_ADDRESS || col == C_PHONE || col == C_COMMENT || col == S_NAME || col == S_ADDRESS ||
           col == S_PHONE || col == S_COMMENT || col == P_NAME;
}
// writes the value of a generated-string column (out may be nullptr) and returns its length in bytes
GEN_HD int gen_string(int col, uint64_t seed, int64_t row, uint8_t *out) {
    StrOut s{out, 0};
    switch (col) {
        case C_NAME: put_str(s, "Customer#"); put_num(s, row + 1, 9); break;
        case S_NAME: put_str(s, "Supplier#"); put_num(s, row + 1, 9); break;
        case C_PHONE: put_phone(s, seed, 43, row, rnd_range(seed, 41, (uint64_t)row, 0, 24)); break;  // country code = C_NATIONKEY + 10
        case S_PHONE: put_phone(s, seed, 53, row, rnd_range(seed, 51, (uint64_t)row, 0, 24)); break;
        case C_ADDRESS: put_address(s, seed, 44, row); break;
        case S_ADDRESS: put_address(s, seed, 54, row); break;
        case C_COMMENT: put_text(s, seed, 45, row, 29, 116); break;
        case O_COMMENT: put_text(s, seed, 33, row, 19, 78); break;
        case S_COMMENT: {
            // 5 suppliers in 10 000 complain, 5 recommend (specification 4.2.3: "Customer ... Complaints" / "Customer ... Recommends")
            const uint64_t r = rnd(seed, 56, (uint64_t)row) % 2000;
            put_text(s, seed, 55, row, 25, r < 2 ? 60 : 100);
            if (r < 2) {
                put_str(s, " Customer ");
                put_str(s, text_word((int)(rnd(seed, 57, (uint64_t)row) % kNumWords)));
                put_str(s, r == 0 ? " Complaints" : " Recommends");
            }
            break;
        }
        case P_NAME:
            for (int j = 0; j < 5; j++) {
                if (j) put(s, ' ');
                put_str(s, color_word((int)(rnd(seed, 69, (uint64_t)row * 8 + (uint64_t)j) % kNumColors)));
            }
            break;
        default: break;
    }
    return s.n;
}

}  // namespace tpchgen
}  // namespace b200

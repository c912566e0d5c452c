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
};
// dictionary columns: number of words (TPC-H specification 4.2.2.13: ship modes, order priorities, market segments;
// P_TYPE = one syllable of each of three lists, 6 x 5 x 5 = 150 combinations)
static const char *kShipModes[7] = {"REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"};
static const char *kPriorities[5] = {"1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"};
static const char *kSegments[5] = {"AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"};
static const char *kTypeS1[6] = {"STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO"};
static const char *kTypeS2[5] = {"ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED"};
static const char *kTypeS3[5] = {"TIN", "NICKEL", "BRASS", "STEEL", "COPPER"};
inline bool is_dict_col(int col) { return col == L_SHIPMODE || col == O_ORDERPRIORITY || col == C_MKTSEGMENT || col == P_TYPE; }
inline int dict_size(int col) { return col == L_SHIPMODE ? 7 : col == P_TYPE ? 150 : 5; }
inline std::string dict_word(int col, int k) {
    switch (col) {
        case L_SHIPMODE: return kShipModes[k];
        case O_ORDERPRIORITY: return kPriorities[k];
        case C_MKTSEGMENT: return kSegments[k];
        default: return std::string(kTypeS1[k / 25]) + " " + kTypeS2[(k / 5) % 5] + " " + kTypeS3[k % 5];
    }
}
// the dictionary index of a dictionary column's row
GEN_HD int gen_dict_index(int col, uint64_t seed, int64_t row) {
    switch (col) {
        case L_SHIPMODE: return (int)(rnd(seed, 61, (uint64_t)row) % 7);
        case O_ORDERPRIORITY: return (int)(rnd(seed, 62, (uint64_t)row) % 5);
        case C_MKTSEGMENT: return (int)(rnd(seed, 63, (uint64_t)row) % 5);
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
            case L_SUPPKEY: v.i = rnd_range(seed, 23, (uint64_t)row, 1, sc.n_suppliers); break;
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


}  // namespace tpchgen
}  // namespace b200

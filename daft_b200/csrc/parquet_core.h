// Parquet column chunks -> the library's Arrow column layout (the reference's scan path is
// src/daft-parquet/src/read.rs:553 stream_parquet behind src/daft-local-execution/src/sources/scan_task.rs:33).
//
// What lives here is everything that touches the file format, as plain functions of raw bytes:
//   * the Thrift compact protocol reader and the footer / page-header structures it fills (host only),
//   * the decoders a page needs -- Snappy, the RLE / bit-packed hybrid (definition levels, dictionary indices), PLAIN and
//     dictionary values -- as __host__ __device__ functions of (page descriptor, column descriptor),
//   * the per-page stages the scan runs: decompress -> dictionary -> values (pass 1) -> string bytes (pass 2).
// The device runs one page per thread through these stages (daft_b200/csrc/parquet.cu) with the compressed column chunks
// resident in HBM; tests compile the very same header for the host (tests/pq_host.cpp) and hold the result to pyarrow's reader.
//
// Supported: flat (non-nested) columns; BOOLEAN, INT32 (also as 8- / 16-bit integers), INT64 (timestamps of any unit), FLOAT, DOUBLE, BYTE_ARRAY; PLAIN, PLAIN_DICTIONARY / RLE_DICTIONARY;
// data pages V1 and V2; UNCOMPRESSED and SNAPPY.  Everything else is refused with an error that names it.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef __CUDACC__
#define PQ_HD __host__ __device__ __forceinline__
#else
#define PQ_HD inline
#endif

namespace b200 {
namespace pq {

enum PhysicalType { T_BOOLEAN = 0, T_INT32 = 1, T_INT64 = 2, T_INT96 = 3, T_FLOAT = 4, T_DOUBLE = 5, T_BYTE_ARRAY = 6, T_FIXED_LEN_BYTE_ARRAY = 7 };
enum Encoding { E_PLAIN = 0, E_PLAIN_DICTIONARY = 2, E_RLE = 3, E_BIT_PACKED = 4, E_DELTA_BINARY_PACKED = 5, E_DELTA_LENGTH_BYTE_ARRAY = 6,
                E_DELTA_BYTE_ARRAY = 7, E_RLE_DICTIONARY = 8, E_BYTE_STREAM_SPLIT = 9 };
enum Codec { C_UNCOMPRESSED = 0, C_SNAPPY = 1, C_GZIP = 2, C_LZO = 3, C_BROTLI = 4, C_LZ4 = 5, C_ZSTD = 6, C_LZ4_RAW = 7 };
enum PageType { P_DATA = 0, P_INDEX = 1, P_DICTIONARY = 2, P_DATA_V2 = 3 };

// ------------------------------------------------------------------------------------------------------------------
// Thrift compact protocol (host): just enough to walk FileMetaData and PageHeader
// ------------------------------------------------------------------------------------------------------------------
struct TReader {
    const uint8_t *p, *end;
    TReader(const uint8_t *b, size_t n) : p(b), end(b + n) {}
    uint8_t byte() {
        if (p >= end) throw std::runtime_error("parquet: truncated thrift data");
        return *p++;
    }
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            const uint8_t b = byte();
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
        }
        throw std::runtime_error("parquet: varint too long");
    }
    int64_t zigzag() {
        const uint64_t v = varint();
        return (int64_t)(v >> 1) ^ -(int64_t)(v & 1);
    }
    std::string binary() {
        const uint64_t n = varint();
        if ((uint64_t)(end - p) < n) throw std::runtime_error("parquet: truncated thrift string");
        std::string s((const char *)p, (size_t)n);
        p += n;
        return s;
    }
    // field header: returns the wire type (0 = stop) and updates `fid`
    int field(int16_t &fid) {
        const uint8_t b = byte();
        if (b == 0) return 0;
        const int delta = b >> 4;
        if (delta) fid = (int16_t)(fid + delta);
        else fid = (int16_t)zigzag();
        return b & 0x0f;
    }
    void list_header(int &elem_type, uint64_t &n) {
        const uint8_t b = byte();
        elem_type = b & 0x0f;
        n = b >> 4;
        if (n == 15) n = varint();
    }
    void skip(int type) {
        switch (type) {
            case 1: case 2: return;               // bool carried by the field header
            case 3: byte(); return;
            case 4: case 5: case 6: varint(); return;
            case 7: for (int i = 0; i < 8; i++) byte(); return;
            case 8: binary(); return;
            case 9: case 10: {
                int et;
                uint64_t n;
                list_header(et, n);
                for (uint64_t i = 0; i < n; i++) {
                    if (et == 1 || et == 2) byte();   // bools inside a list take a byte each
                    else skip(et);
                }
                return;
            }
            case 11: {
                const uint64_t n = varint();
                if (n) {
                    const uint8_t kv = byte();
                    for (uint64_t i = 0; i < n; i++) {
                        skip(kv >> 4);
                        skip(kv & 0x0f);
                    }
                }
                return;
            }
            case 12: {
                int16_t fid = 0;
                for (int t = field(fid); t; t = field(fid)) skip(t);
                return;
            }
            default: throw std::runtime_error("parquet: unknown thrift type " + std::to_string(type));
        }
    }
};

struct SchemaElement {
    std::string name;
    int32_t type = -1, type_length = 0, repetition = 0 /* 0 required, 1 optional, 2 repeated */, num_children = 0, converted = -1;
    bool is_string = false, is_date = false, is_timestamp = false, is_decimal = false, int_unsigned = false;
    int32_t int_bits = 0, ts_unit = 0 /* 1 ms, 2 us, 3 ns */;
};
struct ColumnChunkMeta {
    int32_t type = -1, codec = 0;
    std::vector<int32_t> encodings;
    std::vector<std::string> path;
    int64_t num_values = 0, total_uncompressed = 0, total_compressed = 0, data_page_offset = 0, dictionary_page_offset = -1;
};
struct RowGroupMeta {
    std::vector<ColumnChunkMeta> columns;
    int64_t num_rows = 0;
};
struct FileMeta {
    int32_t version = 0;
    int64_t num_rows = 0;
    std::vector<SchemaElement> schema;   // flattened, element 0 is the root
    std::vector<RowGroupMeta> row_groups;
    std::string created_by;
};

inline void parse_logical_type(TReader &r, SchemaElement &e) {
    int16_t fid = 0;
    for (int t = r.field(fid); t; t = r.field(fid)) {
        if (t != 12) { r.skip(t); continue; }
        if (fid == 1) e.is_string = true;
        else if (fid == 5) e.is_decimal = true;
        else if (fid == 6) e.is_date = true;
        else if (fid == 8) e.is_timestamp = true;
        if (fid == 8 || fid == 10) {   // TIMESTAMP {1 isAdjustedToUTC, 2 unit}, INTEGER {1 bitWidth, 2 isSigned}
            int16_t f2 = 0;
            for (int t2 = r.field(f2); t2; t2 = r.field(f2)) {
                if (fid == 8 && f2 == 2 && t2 == 12) {
                    int16_t f3 = 0;
                    for (int t3 = r.field(f3); t3; t3 = r.field(f3)) {
                        e.ts_unit = f3;
                        r.skip(t3);
                    }
                } else if (fid == 10 && f2 == 1 && t2 == 3) {
                    e.int_bits = r.byte();
                } else if (fid == 10 && f2 == 2 && (t2 == 1 || t2 == 2)) {
                    e.int_unsigned = (t2 == 2);
                } else {
                    r.skip(t2);
                }
            }
        } else {
            r.skip(12);
        }
    }
}
inline SchemaElement parse_schema_element(TReader &r) {
    SchemaElement e;
    int16_t fid = 0;
    for (int t = r.field(fid); t; t = r.field(fid)) {
        switch (fid) {
            case 1: e.type = (int32_t)r.zigzag(); break;
            case 2: e.type_length = (int32_t)r.zigzag(); break;
            case 3: e.repetition = (int32_t)r.zigzag(); break;
            case 4: e.name = r.binary(); break;
            case 5: e.num_children = (int32_t)r.zigzag(); break;
            case 6: e.converted = (int32_t)r.zigzag(); break;
            case 10: parse_logical_type(r, e); break;
            default: r.skip(t);
        }
    }
    // converted types (the pre-2.4 annotation): UTF8 = 0, DECIMAL = 5, DATE = 6, TIMESTAMP_MILLIS = 9, TIMESTAMP_MICROS = 10, UINT_8..64 = 11..14
    if (e.converted == 0) e.is_string = true;
    if (e.converted == 5) e.is_decimal = true;
    if (e.converted == 6) e.is_date = true;
    if (e.converted == 9 || e.converted == 10) {
        e.is_timestamp = true;
        if (!e.ts_unit) e.ts_unit = e.converted == 9 ? 1 : 2;
    }
    if (e.converted >= 11 && e.converted <= 14 && !e.int_bits) {
        e.int_unsigned = true;
        e.int_bits = 8 << (e.converted - 11);
    }
    if (e.converted >= 15 && e.converted <= 18 && !e.int_bits) e.int_bits = 8 << (e.converted - 15);
    return e;
}
inline ColumnChunkMeta parse_column_meta(TReader &r) {
    ColumnChunkMeta m;
    int16_t fid = 0;
    for (int t = r.field(fid); t; t = r.field(fid)) {
        switch (fid) {
            case 1: m.type = (int32_t)r.zigzag(); break;
            case 2: {
                int et;
                uint64_t n;
                r.list_header(et, n);
                for (uint64_t i = 0; i < n; i++) m.encodings.push_back((int32_t)r.zigzag());
                break;
            }
            case 3: {
                int et;
                uint64_t n;
                r.list_header(et, n);
                for (uint64_t i = 0; i < n; i++) m.path.push_back(r.binary());
                break;
            }
            case 4: m.codec = (int32_t)r.zigzag(); break;
            case 5: m.num_values = r.zigzag(); break;
            case 6: m.total_uncompressed = r.zigzag(); break;
            case 7: m.total_compressed = r.zigzag(); break;
            case 9: m.data_page_offset = r.zigzag(); break;
            case 11: m.dictionary_page_offset = r.zigzag(); break;
            default: r.skip(t);
        }
    }
    return m;
}
inline FileMeta parse_file_meta(const uint8_t *buf, size_t n) {
    TReader r(buf, n);
    FileMeta fm;
    int16_t fid = 0;
    for (int t = r.field(fid); t; t = r.field(fid)) {
        if (fid == 1) fm.version = (int32_t)r.zigzag();
        else if (fid == 2) {
            int et;
            uint64_t cnt;
            r.list_header(et, cnt);
            for (uint64_t i = 0; i < cnt; i++) fm.schema.push_back(parse_schema_element(r));
        } else if (fid == 3) fm.num_rows = r.zigzag();
        else if (fid == 4) {
            int et;
            uint64_t cnt;
            r.list_header(et, cnt);
            for (uint64_t i = 0; i < cnt; i++) {
                RowGroupMeta rg;
                int16_t f2 = 0;
                for (int t2 = r.field(f2); t2; t2 = r.field(f2)) {
                    if (f2 == 1) {
                        int et2;
                        uint64_t c2;
                        r.list_header(et2, c2);
                        for (uint64_t j = 0; j < c2; j++) {
                            ColumnChunkMeta cm;
                            int16_t f3 = 0;
                            for (int t3 = r.field(f3); t3; t3 = r.field(f3)) {
                                if (f3 == 3 && t3 == 12) cm = parse_column_meta(r);
                                else r.skip(t3);
                            }
                            rg.columns.push_back(std::move(cm));
                        }
                    } else if (f2 == 3) rg.num_rows = r.zigzag();
                    else r.skip(t2);
                }
                fm.row_groups.push_back(std::move(rg));
            }
        } else if (fid == 6) fm.created_by = r.binary();
        else r.skip(t);
    }
    return fm;
}

struct PageHeader {
    int32_t type = -1, uncompressed_size = 0, compressed_size = 0, num_values = 0, encoding = 0, def_encoding = E_RLE;
    int32_t v2_num_nulls = 0, v2_num_rows = 0, v2_def_len = 0, v2_rep_len = 0;
    bool v2_compressed = true;
    size_t header_bytes = 0;
};
inline PageHeader parse_page_header(const uint8_t *buf, size_t n) {
    TReader r(buf, n);
    PageHeader h;
    int16_t fid = 0;
    for (int t = r.field(fid); t; t = r.field(fid)) {
        if (fid == 1) h.type = (int32_t)r.zigzag();
        else if (fid == 2) h.uncompressed_size = (int32_t)r.zigzag();
        else if (fid == 3) h.compressed_size = (int32_t)r.zigzag();
        else if ((fid == 5 || fid == 7 || fid == 8) && t == 12) {
            int16_t f2 = 0;
            for (int t2 = r.field(f2); t2; t2 = r.field(f2)) {
                if (fid == 5) {
                    if (f2 == 1) h.num_values = (int32_t)r.zigzag();
                    else if (f2 == 2) h.encoding = (int32_t)r.zigzag();
                    else if (f2 == 3) h.def_encoding = (int32_t)r.zigzag();
                    else r.skip(t2);
                } else if (fid == 7) {
                    if (f2 == 1) h.num_values = (int32_t)r.zigzag();
                    else if (f2 == 2) h.encoding = (int32_t)r.zigzag();
                    else r.skip(t2);
                } else {
                    if (f2 == 1) h.num_values = (int32_t)r.zigzag();
                    else if (f2 == 2) h.v2_num_nulls = (int32_t)r.zigzag();
                    else if (f2 == 3) h.v2_num_rows = (int32_t)r.zigzag();
                    else if (f2 == 4) h.encoding = (int32_t)r.zigzag();
                    else if (f2 == 5) h.v2_def_len = (int32_t)r.zigzag();
                    else if (f2 == 6) h.v2_rep_len = (int32_t)r.zigzag();
                    else if (f2 == 7 && (t2 == 1 || t2 == 2)) h.v2_compressed = (t2 == 1);
                    else r.skip(t2);
                }
            }
        } else r.skip(t);
    }
    h.header_bytes = (size_t)(r.p - buf);
    return h;
}

// ------------------------------------------------------------------------------------------------------------------
// Descriptors shared by the host planner and the per-page stages
// ------------------------------------------------------------------------------------------------------------------
struct ColumnDesc {
    int32_t physical = 0;        // PhysicalType
    int32_t max_def = 0;         // 0 = required (no definition levels), 1 = optional
    int32_t width = 0;           // bytes per output value (0 for strings, 1 for the Boolean byte staging)
    int32_t is_string = 0;
    // outputs (n_rows entries unless stated)
    uint8_t *out_values = nullptr;     // fixed width: n_rows * width bytes; BOOLEAN: one byte per row (packed afterwards)
    uint8_t *out_valid = nullptr;      // one byte per row (packed afterwards); null when max_def == 0
    int64_t *out_lengths = nullptr;    // strings: n_rows (+1) lengths, scanned into offsets between pass 1 and pass 2
    int64_t *out_offsets = nullptr;    // strings: n_rows + 1
    uint8_t *out_bytes = nullptr;      // strings: the value bytes (allocated after the scan)
    int32_t *row_dict_index = nullptr; // strings through a dictionary: the index of every row (-1 = null), pass 1 -> pass 2
    int64_t *row_src = nullptr;        // PLAIN strings: byte position of every row's value inside its page (-1 = null)
};
struct PageDesc {
    int64_t src_off = 0;         // page payload inside the chunk buffer
    int64_t scratch_off = -1;    // where the decompressed payload goes (-1 = the payload is used in place)
    int32_t comp_size = 0, uncomp_size = 0;
    int32_t num_values = 0;
    int32_t encoding = 0;        // E_PLAIN or E_RLE_DICTIONARY (PLAIN_DICTIONARY is mapped to it)
    int32_t page_type = 0;       // P_DATA, P_DATA_V2 or P_DICTIONARY
    int32_t codec = 0;
    int32_t v2_def_len = 0, v2_rep_len = 0, v2_compressed = 1;
    int32_t column = 0;          // ColumnDesc index
    int32_t dict = -1;           // DictDesc index, -1 = none
    int64_t row_begin = 0;       // first output row of the page
    int32_t status = 0;          // written by the stages: 0 ok, otherwise an error code (see pq_error_text)
};
struct DictDesc {
    int32_t page = -1;           // PageDesc index of the dictionary page
    int32_t count = 0;
    int64_t str_off_begin = 0;   // strings: this dictionary's slice of dict_str_offsets (count + 1 entries)
};
struct ScanCtx {
    const uint8_t *chunks = nullptr;   // all selected column chunks back to back
    uint8_t *scratch = nullptr;        // decompressed pages
    PageDesc *pages = nullptr;
    ColumnDesc *columns = nullptr;
    DictDesc *dicts = nullptr;
    int32_t *dict_str_offsets = nullptr;   // strings: offsets of the dictionary entries inside their (decompressed) dictionary page
    int32_t npages = 0, ndicts = 0;
};

enum PqError { PQ_OK = 0, PQ_ERR_SNAPPY = 1, PQ_ERR_LEVELS = 2, PQ_ERR_VALUES = 3, PQ_ERR_DICT_INDEX = 4, PQ_ERR_ENCODING = 5 };
inline const char *pq_error_text(int e) {
    switch (e) {
        case PQ_ERR_SNAPPY: return "corrupt Snappy block";
        case PQ_ERR_LEVELS: return "definition levels do not match the page's value count";
        case PQ_ERR_VALUES: return "page values end before the page's value count";
        case PQ_ERR_DICT_INDEX: return "dictionary index out of range";
        case PQ_ERR_ENCODING: return "unsupported page encoding";
        default: return "unknown error";
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Decoders (host + device)
// ------------------------------------------------------------------------------------------------------------------
PQ_HD uint32_t ld_u32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
// non-overlapping copy: whole words when source and destination share their misalignment (page payloads sit at arbitrary byte
// offsets, so this is the lucky half of the cases), bytes otherwise
PQ_HD void copy_bytes(uint8_t *dst, const uint8_t *src, int64_t n) {
    int64_t i = 0;
    const uintptr_t d = (uintptr_t)dst, s = (uintptr_t)src;
    if (((d ^ s) & 7) == 0) {
        while (i < n && ((d + (uintptr_t)i) & 7)) {
            dst[i] = src[i];
            i++;
        }
        for (; i + 8 <= n; i += 8) *(uint64_t *)(dst + i) = *(const uint64_t *)(src + i);
    } else if (((d ^ s) & 3) == 0) {
        while (i < n && ((d + (uintptr_t)i) & 3)) {
            dst[i] = src[i];
            i++;
        }
        for (; i + 4 <= n; i += 4) *(uint32_t *)(dst + i) = *(const uint32_t *)(src + i);
    }
    for (; i < n; i++) dst[i] = src[i];
}
PQ_HD void fill_bytes(uint8_t *dst, uint8_t v, int64_t n) {
    int64_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 7)) dst[i++] = v;
    const uint64_t w = 0x0101010101010101ULL * v;
    for (; i + 8 <= n; i += 8) *(uint64_t *)(dst + i) = w;
    for (; i < n; i++) dst[i] = v;
}

// Snappy raw format: varint uncompressed length, then literal / copy elements.  Returns the number of bytes produced, -1 when corrupt.
PQ_HD int64_t snappy_decompress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap) {
    int64_t ip = 0;
    uint64_t ulen = 0;
    for (int shift = 0;; shift += 7) {
        if (ip >= n || shift > 35) return -1;
        const uint8_t b = src[ip++];
        ulen |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) break;
    }
    if ((int64_t)ulen > cap) return -1;
    int64_t op = 0;
    while (ip < n) {
        const uint8_t tag = src[ip++];
        int64_t len, off;
        switch (tag & 3) {
            case 0: {
                len = (tag >> 2) + 1;
                if (len > 60) {
                    const int extra = (int)len - 60;
                    if (ip + extra > n) return -1;
                    len = 0;
                    for (int i = 0; i < extra; i++) len |= (int64_t)src[ip + i] << (8 * i);
                    len += 1;
                    ip += extra;
                }
                if (ip + len > n || op + len > (int64_t)ulen) return -1;
                copy_bytes(dst + op, src + ip, len);
                ip += len;
                op += len;
                continue;
            }
            case 1:
                if (ip >= n) return -1;
                len = ((tag >> 2) & 7) + 4;
                off = ((int64_t)(tag >> 5) << 8) | src[ip++];
                break;
            case 2:
                if (ip + 2 > n) return -1;
                len = (tag >> 2) + 1;
                off = (int64_t)src[ip] | ((int64_t)src[ip + 1] << 8);
                ip += 2;
                break;
            default:
                if (ip + 4 > n) return -1;
                len = (tag >> 2) + 1;
                off = (int64_t)ld_u32(src + ip);
                ip += 4;
                break;
        }
        if (off == 0 || off > op || op + len > (int64_t)ulen) return -1;
        if (off >= len) copy_bytes(dst + op, dst + op - off, len);   // disjoint ranges
        else for (int64_t i = 0; i < len; i++) dst[op + i] = dst[op + i - off];   // overlapping copies repeat the pattern byte by byte
        op += len;
    }
    return op == (int64_t)ulen ? op : -1;
}

// RLE / bit-packed hybrid (definition levels, dictionary indices): <header varint> then either a repeated value of
// ceil(width / 8) bytes (header & 1 == 0, count = header >> 1) or (header >> 1) groups of 8 bit-packed values.
struct HybridReader {
    const uint8_t *p, *end;
    int32_t width;
    int64_t run_left = 0;     // values left in the current run
    int32_t run_value = 0;    // RLE run
    bool packed = false;
    int64_t bit_pos = 0;      // bit-packed run: bit offset from p
    PQ_HD HybridReader(const uint8_t *b, const uint8_t *e, int32_t w) : p(b), end(e), width(w) {}
    PQ_HD bool next_run() {
        if (packed) {   // skip what is left of the bit-packed run (always whole groups of 8 values)
            p += (bit_pos + run_left * width + 7) / 8;
            packed = false;
        }
        uint64_t h = 0;
        for (int shift = 0;; shift += 7) {
            if (p >= end || shift > 35) return false;
            const uint8_t b = *p++;
            h |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) break;
        }
        if (h & 1) {
            run_left = (int64_t)(h >> 1) * 8;
            packed = true;
            bit_pos = 0;
            if (p + (run_left * width + 7) / 8 > end) {   // the last group may be cut short by the writer: clamp to what is there
                run_left = ((end - p) * 8) / (width ? width : 1);
            }
        } else {
            run_left = (int64_t)(h >> 1);
            const int nb = (width + 7) / 8;
            if (p + nb > end) return false;
            run_value = 0;
            for (int i = 0; i < nb; i++) run_value |= (int32_t)p[i] << (8 * i);
            p += nb;
        }
        return run_left > 0 || (h >> 1) == 0;
    }
    // the next value; ok = false when the stream ends early
    PQ_HD int32_t next(bool &ok) {
        while (run_left == 0) {
            if (!next_run()) {
                ok = false;
                return 0;
            }
        }
        run_left--;
        if (!packed) return run_value;
        int32_t v = 0;
        for (int b = 0; b < width; b++) {
            const int64_t bit = bit_pos + b;
            v |= (int32_t)((p[bit >> 3] >> (bit & 7)) & 1) << b;
        }
        bit_pos += width;
        if (run_left == 0) {
            p += (bit_pos + 7) / 8;
            packed = false;
        }
        return v;
    }
};

// the decompressed payload of a page
PQ_HD const uint8_t *page_data(const ScanCtx &c, const PageDesc &pg) { return pg.scratch_off >= 0 ? c.scratch + pg.scratch_off : c.chunks + pg.src_off; }

// ---- stage 0: decompress (every page that has a scratch slot) ----
PQ_HD void stage_decompress(const ScanCtx &c, int32_t pi) {
    PageDesc &pg = c.pages[pi];
    if (pg.scratch_off < 0) return;
    const uint8_t *src = c.chunks + pg.src_off;
    uint8_t *dst = c.scratch + pg.scratch_off;
    // data page V2: the levels in front are never compressed
    const int64_t lev = pg.page_type == P_DATA_V2 ? (int64_t)pg.v2_def_len + pg.v2_rep_len : 0;
    copy_bytes(dst, src, lev);
    if (pg.codec == C_SNAPPY && (pg.page_type != P_DATA_V2 || pg.v2_compressed)) {
        const int64_t got = snappy_decompress(src + lev, pg.comp_size - lev, dst + lev, pg.uncomp_size - lev);
        if (got != pg.uncomp_size - lev) pg.status = PQ_ERR_SNAPPY;
    } else {
        copy_bytes(dst + lev, src + lev, pg.comp_size - lev);
    }
}

// ---- stage 1: dictionaries of string columns: where every entry starts inside the dictionary page ----
PQ_HD void stage_dictionary(const ScanCtx &c, int32_t di) {
    const DictDesc &d = c.dicts[di];
    PageDesc &pg = c.pages[d.page];
    const ColumnDesc &col = c.columns[pg.column];
    if (!col.is_string || pg.status) return;
    const uint8_t *data = page_data(c, pg);
    int32_t *offs = c.dict_str_offsets + d.str_off_begin;
    int64_t pos = 0;
    for (int32_t k = 0; k < d.count; k++) {
        if (pos + 4 > pg.uncomp_size) {
            pg.status = PQ_ERR_VALUES;
            return;
        }
        const int64_t len = (int64_t)ld_u32(data + pos);
        offs[k] = (int32_t)(pos + 4);
        pos += 4 + len;
        if (pos > pg.uncomp_size) {
            pg.status = PQ_ERR_VALUES;
            return;
        }
    }
    offs[d.count] = (int32_t)(pos + 4);   // entry k spans [offs[k], offs[k + 1] - 4)
}

// ---- stage 2: data pages, pass 1: validity, fixed-width values, string lengths ----
PQ_HD void stage_values(const ScanCtx &c, int32_t pi) {
    PageDesc &pg = c.pages[pi];
    if (pg.page_type == P_DICTIONARY || pg.status) return;
    const ColumnDesc &col = c.columns[pg.column];
    const uint8_t *data = page_data(c, pg);
    const uint8_t *end = data + pg.uncomp_size;
    const int64_t n = pg.num_values, r0 = pg.row_begin;
    // definition levels
    const uint8_t *lev = nullptr, *lev_end = nullptr;
    const uint8_t *vals = data;
    if (pg.page_type == P_DATA_V2) {
        vals = data + pg.v2_rep_len + pg.v2_def_len;
        if (col.max_def > 0) {
            lev = data + pg.v2_rep_len;
            lev_end = lev + pg.v2_def_len;
        }
    } else if (col.max_def > 0) {
        if (data + 4 > end) {
            pg.status = PQ_ERR_LEVELS;
            return;
        }
        const int64_t nb = (int64_t)ld_u32(data);
        lev = data + 4;
        lev_end = lev + nb;
        vals = lev_end;
        if (vals > end) {
            pg.status = PQ_ERR_LEVELS;
            return;
        }
    }
    HybridReader levels(lev, lev_end, 1);
    // the common page of a nullable column without nulls: ONE run-length run of 1s covering the page -- no per-value level decode
    bool all_valid = col.max_def == 0;
    if (!all_valid && lev != nullptr) {
        HybridReader peek(lev, lev_end, 1);
        if (peek.next_run() && !peek.packed && peek.run_value == col.max_def && peek.run_left >= n) all_valid = true;
    }
    if (all_valid && col.out_valid) fill_bytes(col.out_valid + r0, 1, n);
    // dictionary indices
    const bool by_dict = pg.encoding == E_RLE_DICTIONARY;
    const DictDesc *dict = by_dict && pg.dict >= 0 ? &c.dicts[pg.dict] : nullptr;
    if (by_dict && !dict) {
        pg.status = PQ_ERR_ENCODING;
        return;
    }
    int32_t idx_width = 0;
    if (by_dict) {
        if (vals >= end && n > 0) {
            pg.status = PQ_ERR_VALUES;
            return;
        }
        idx_width = vals < end ? (int32_t)vals[0] : 0;
        vals += 1;
        if (idx_width > 32) {   // a corrupt page: dictionary indices are at most 32 bits wide
            pg.status = PQ_ERR_VALUES;
            return;
        }
    }
    const bool bool_rle = pg.encoding == E_RLE;   // BOOLEAN only (the planner checks): <4-byte length> <hybrid runs of 1-bit values>
    if (bool_rle) {
        if (vals + 4 > end && n > 0) {
            pg.status = PQ_ERR_VALUES;
            return;
        }
        vals += 4;
        idx_width = 1;
    }
    HybridReader indices(vals, end, idx_width);
    const uint8_t *dict_data = dict ? page_data(c, c.pages[dict->page]) : nullptr;
    const int32_t *dict_offs = dict && col.is_string ? c.dict_str_offsets + dict->str_off_begin : nullptr;
    int64_t pos = 0;        // PLAIN: byte position inside the values
    int64_t bool_bit = 0;   // PLAIN BOOLEAN: bit position
    const int64_t vbytes = end - vals;
    if (all_valid && !by_dict && !bool_rle && !col.is_string && col.physical != T_BOOLEAN) {   // PLAIN fixed width, no nulls: one copy
        if (n * col.width > vbytes) {
            pg.status = PQ_ERR_VALUES;
            return;
        }
        copy_bytes(col.out_values + r0 * col.width, vals, n * col.width);
        return;
    }
    for (int64_t i = 0; i < n; i++) {
        const int64_t r = r0 + i;
        bool ok = true;
        const bool valid = all_valid || levels.next(ok) == col.max_def;
        if (!ok) {
            pg.status = PQ_ERR_LEVELS;
            return;
        }
        if (col.out_valid && !all_valid) col.out_valid[r] = valid ? 1 : 0;
        if (!valid) {
            if (col.is_string) {
                col.out_lengths[r] = 0;
                if (col.row_dict_index) col.row_dict_index[r] = -1;
                if (col.row_src) col.row_src[r] = -1;
            } else {
                for (int b = 0; b < col.width; b++) col.out_values[r * col.width + b] = 0;
            }
            continue;
        }
        if (by_dict) {
            const int32_t k = indices.next(ok);
            if (!ok) {
                pg.status = PQ_ERR_VALUES;
                return;
            }
            if (k < 0 || k >= dict->count) {
                pg.status = PQ_ERR_DICT_INDEX;
                return;
            }
            if (col.is_string) {
                col.out_lengths[r] = (int64_t)dict_offs[k + 1] - 4 - dict_offs[k];
                col.row_dict_index[r] = k;
            } else if (col.physical == T_BOOLEAN) {
                col.out_values[r] = (dict_data[k >> 3] >> (k & 7)) & 1;
            } else {
                copy_bytes(col.out_values + r * col.width, dict_data + (int64_t)k * col.width, col.width);
            }
        } else if (bool_rle) {
            const int32_t v = indices.next(ok);
            if (!ok) {
                pg.status = PQ_ERR_VALUES;
                return;
            }
            col.out_values[r] = (uint8_t)(v & 1);
        } else if (col.is_string) {
            if (pos + 4 > vbytes) {
                pg.status = PQ_ERR_VALUES;
                return;
            }
            const int64_t len = (int64_t)ld_u32(vals + pos);
            if (pos + 4 + len > vbytes) {
                pg.status = PQ_ERR_VALUES;
                return;
            }
            col.out_lengths[r] = len;
            col.row_src[r] = (vals - data) + pos + 4;
            pos += 4 + len;
        } else if (col.physical == T_BOOLEAN) {
            if ((bool_bit >> 3) >= vbytes) {
                pg.status = PQ_ERR_VALUES;
                return;
            }
            col.out_values[r] = (vals[bool_bit >> 3] >> (bool_bit & 7)) & 1;
            bool_bit++;
        } else {
            if (pos + col.width > vbytes) {
                pg.status = PQ_ERR_VALUES;
                return;
            }
            copy_bytes(col.out_values + r * col.width, vals + pos, col.width);
            pos += col.width;
        }
    }
}

// ---- stage 3: data pages, pass 2 (strings): the bytes, once the scanned offsets say where they go ----
PQ_HD void stage_string_bytes(const ScanCtx &c, int32_t pi) {
    PageDesc &pg = c.pages[pi];
    if (pg.page_type == P_DICTIONARY || pg.status) return;
    const ColumnDesc &col = c.columns[pg.column];
    if (!col.is_string) return;
    const uint8_t *data = page_data(c, pg);
    const bool by_dict = pg.encoding == E_RLE_DICTIONARY;
    const DictDesc *dict = by_dict ? &c.dicts[pg.dict] : nullptr;
    const uint8_t *dict_data = dict ? page_data(c, c.pages[dict->page]) : nullptr;
    const int32_t *dict_offs = dict ? c.dict_str_offsets + dict->str_off_begin : nullptr;
    for (int64_t i = 0; i < pg.num_values; i++) {
        const int64_t r = pg.row_begin + i;
        const int64_t o = col.out_offsets[r], len = col.out_offsets[r + 1] - o;
        if (len == 0) continue;
        const uint8_t *src = by_dict ? dict_data + dict_offs[col.row_dict_index[r]] : data + col.row_src[r];
        copy_bytes(col.out_bytes + o, src, len);
    }
}

// ---- after the stages: narrow / rescale a column that Parquet stores wider or in another unit than its Arrow type ----
// INT32 annotated INT(8 | 16): the low bytes of the little-endian value; TIMESTAMP millis -> micros (x 1000) and nanos -> micros
// (floor division by 1000, the unit the library's Timestamp has).  One value per call: the executors run it over all rows.
enum ConvKind { CONV_NONE = 0, CONV_LOW8 = 1, CONV_LOW16 = 2, CONV_MUL64 = 3, CONV_FLOORDIV64 = 4 };
PQ_HD void convert_one(int kind, const uint8_t *src, uint8_t *dst, int64_t i, int64_t factor) {
    if (kind == CONV_LOW8) {
        dst[i] = src[4 * i];
    } else if (kind == CONV_LOW16) {
        dst[2 * i] = src[4 * i];
        dst[2 * i + 1] = src[4 * i + 1];
    } else {
        int64_t v = ((const int64_t *)src)[i];
        if (kind == CONV_MUL64) {
            v *= factor;
        } else {
            const int64_t q = v / factor, r = v % factor;
            v = r < 0 ? q - 1 : q;
        }
        ((int64_t *)dst)[i] = v;
    }
}

}  // namespace pq
}  // namespace b200

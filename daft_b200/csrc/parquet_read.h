// The Parquet scan's host side: footer -> which byte ranges to read -> page table -> the order of the per-page stages.
// Written against an executor X so that the same planner drives the device (parquet.cu: HBM buffers, one CUDA thread per page)
// and the host build the tests hold to pyarrow (tests/pq_host.cpp: malloc'd buffers, a plain loop over the pages).
//
//   X::Mem                      a buffer handle with .ptr
//   X::alloc(bytes)             uninitialised buffer
//   X::stage(bytes)             host memory the planner reads the file into (page-locked on the device executor)
//   X::upload(mem, src, bytes)  host -> buffer
//   X::download(dst, ptr, bytes)
//   X::run(stage, ctx, count)   stage 0 decompress / 1 dictionary / 2 values / 3 string bytes over `count` pages (dictionaries for 1)
//   X::scan(lengths, offsets, n)      exclusive scan of n int64 lengths into offsets[0..n], returns offsets[n] (the total)
//   X::pack(bytes01, n, bitmap) -> number of zero bytes; packs one byte per row into an LSB-first bitmap
//   X::convert(kind, src, dst, n, factor)   convert_one over n rows (narrow INT32 to 8 / 16 bits, rescale timestamps)
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <thread>

#include "parquet_core.h"

#ifndef B200_BOOL  // standalone use (tests/pq_host.cpp): the physical dtype codes of include/daft_b200.h
#define B200_BOOL 1
#define B200_I8 2
#define B200_I16 3
#define B200_U8 6
#define B200_U16 7
#define B200_I32 4
#define B200_I64 5
#define B200_U32 8
#define B200_U64 9
#define B200_F32 10
#define B200_F64 11
#define B200_DATE32 12
#define B200_TIMESTAMP 13
#define B200_LARGE_UTF8 14
#define B200_LARGE_BINARY 15
#endif

namespace b200 {
namespace pq {

struct FileBytes {   // a whole-file view the planner reads ranges from: pread on a descriptor, safe from several threads at once
    int fd = -1;
    int64_t size = 0;
    explicit FileBytes(const std::string &path) {
        fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("parquet: cannot open " + path);
        struct stat st;
        if (::fstat(fd, &st) != 0) {
            ::close(fd);
            throw std::runtime_error("parquet: cannot stat " + path);
        }
        size = (int64_t)st.st_size;
    }
    ~FileBytes() {
        if (fd >= 0) ::close(fd);
    }
    FileBytes(const FileBytes &) = delete;
    FileBytes &operator=(const FileBytes &) = delete;
    void read(int64_t off, int64_t n, uint8_t *dst) const {
        if (off < 0 || n < 0 || off + n > size) throw std::runtime_error("parquet: read outside the file");
        while (n > 0) {
            const ssize_t got = ::pread(fd, dst, (size_t)n, (off_t)off);
            if (got <= 0) throw std::runtime_error("parquet: short read");
            dst += got;
            off += got;
            n -= got;
        }
    }
};
// the column chunks' byte ranges, read by a few threads at once (the page cache / NVMe queue wants more than one request in flight)
struct ChunkRead {
    int64_t file_off, bytes, dst_off;
};
inline void read_chunks(const FileBytes &fb, const std::vector<ChunkRead> &reads, uint8_t *dst) {
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nt = std::min<size_t>(std::min<size_t>(reads.size(), 8), hw ? hw : 1);
    if (nt <= 1) {
        for (const ChunkRead &r : reads) fb.read(r.file_off, r.bytes, dst + r.dst_off);
        return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::string> errors(nt);
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; t++)
        th.emplace_back([&, t] {
            try {
                for (size_t i = next++; i < reads.size(); i = next++) fb.read(reads[i].file_off, reads[i].bytes, dst + reads[i].dst_off);
            } catch (const std::exception &e) {
                errors[t] = e.what();
            }
        });
    for (auto &x : th) x.join();
    for (const std::string &e : errors)
        if (!e.empty()) throw std::runtime_error(e);
}

inline FileMeta read_footer(const FileBytes &fb) {
    if (fb.size < 12) throw std::runtime_error("parquet: file too small");
    uint8_t tail[8];
    fb.read(fb.size - 8, 8, tail);
    if (memcmp(tail + 4, "PAR1", 4) != 0) throw std::runtime_error("parquet: missing PAR1 magic (encrypted footers are not supported)");
    const int64_t flen = (int64_t)ld_u32(tail);
    if (flen <= 0 || flen + 8 > fb.size) throw std::runtime_error("parquet: bad footer length");
    std::vector<uint8_t> buf((size_t)flen);
    fb.read(fb.size - 8 - flen, flen, buf.data());
    return parse_file_meta(buf.data(), buf.size());
}

struct LeafInfo {
    std::string name;
    int32_t schema_index = 0, dtype = 0, width = 0, max_def = 0, physical = 0;
    bool is_string = false;
    int32_t conv = CONV_NONE, final_width = 0;   // a conversion pass after the stages (parquet_core.h convert_one)
    int64_t conv_factor = 1;
};
// flat schemas only: the leaves are the root's children, leaf i is column chunk i of every row group
inline std::vector<LeafInfo> leaves_of(const FileMeta &fm) {
    if (fm.schema.empty()) throw std::runtime_error("parquet: empty schema");
    std::vector<LeafInfo> out;
    for (size_t i = 1; i < fm.schema.size(); i++) {
        const SchemaElement &e = fm.schema[i];
        LeafInfo l;
        l.name = e.name;
        l.schema_index = (int32_t)i;
        l.dtype = -1;   // resolved (or refused) when the column is selected
        l.physical = e.type;
        l.max_def = e.repetition == 1 ? 1 : 0;
        out.push_back(l);
    }
    return out;
}
inline void resolve_leaf(const FileMeta &fm, LeafInfo &l) {
    const SchemaElement &e = fm.schema[(size_t)l.schema_index];
    const std::string where = "parquet: column '" + e.name + "': ";
    if (e.num_children > 0) throw std::runtime_error(where + "nested (group) columns are not supported");
    if (e.repetition == 2) throw std::runtime_error(where + "repeated fields are not supported");
    if (e.is_decimal) throw std::runtime_error(where + "DECIMAL is not supported");
    switch (e.type) {
        case T_BOOLEAN: l.dtype = B200_BOOL; l.width = 1; break;
        case T_INT32:
            if (e.is_date) l.dtype = B200_DATE32;
            else if (e.int_bits == 8) {
                l.dtype = e.int_unsigned ? B200_U8 : B200_I8;
                l.conv = CONV_LOW8;
                l.final_width = 1;
            } else if (e.int_bits == 16) {
                l.dtype = e.int_unsigned ? B200_U16 : B200_I16;
                l.conv = CONV_LOW16;
                l.final_width = 2;
            } else l.dtype = e.int_unsigned ? B200_U32 : B200_I32;
            l.width = 4;
            break;
        case T_INT64:
            if (e.is_timestamp) {   // the library's Timestamp counts microseconds
                l.dtype = B200_TIMESTAMP;
                if (e.ts_unit == 1) {
                    l.conv = CONV_MUL64;
                    l.conv_factor = 1000;
                    l.final_width = 8;
                } else if (e.ts_unit == 3) {
                    l.conv = CONV_FLOORDIV64;
                    l.conv_factor = 1000;
                    l.final_width = 8;
                } else if (e.ts_unit != 2) throw std::runtime_error(where + "timestamp without a unit");
            } else l.dtype = e.int_unsigned ? B200_U64 : B200_I64;
            l.width = 8;
            break;
        case T_FLOAT: l.dtype = B200_F32; l.width = 4; break;
        case T_DOUBLE: l.dtype = B200_F64; l.width = 8; break;
        case T_BYTE_ARRAY: l.dtype = e.is_string ? B200_LARGE_UTF8 : B200_LARGE_BINARY; l.width = 0; l.is_string = true; break;
        default: throw std::runtime_error(where + "physical type " + std::to_string(e.type) + " (INT96 / FIXED_LEN_BYTE_ARRAY) is not supported");
    }
}

template <class X>
struct ColumnResult {
    std::string name;
    int32_t dtype = 0;
    int64_t length = 0, null_count = 0, data_bytes = 0;
    typename X::Mem data, validity, offsets;
};

// rows of row groups [rg_begin, rg_end) (rg_end < 0 = all) of the named columns (empty = every column)
template <class X>
std::vector<ColumnResult<X>> scan_file(X &x, const std::string &path, const std::vector<std::string> &want, int64_t rg_begin, int64_t rg_end,
                                       int64_t *rows_out) {
    FileBytes fb(path);
    const FileMeta fm = read_footer(fb);
    std::vector<LeafInfo> leaves = leaves_of(fm);
    for (size_t i = 1; i < fm.schema.size(); i++)
        if (fm.schema[i].num_children > 0) throw std::runtime_error("parquet: nested schemas are not supported (group '" + fm.schema[i].name + "')");
    std::vector<int> sel;
    if (want.empty()) {
        for (size_t i = 0; i < leaves.size(); i++) sel.push_back((int)i);
    } else {
        for (const std::string &w : want) {
            int found = -1;
            for (size_t i = 0; i < leaves.size(); i++)
                if (leaves[i].name == w) found = (int)i;
            if (found < 0) throw std::runtime_error("parquet: no column named '" + w + "'");
            sel.push_back(found);
        }
    }
    for (int li : sel) resolve_leaf(fm, leaves[(size_t)li]);
    const int64_t nrg = (int64_t)fm.row_groups.size();
    if (rg_end < 0 || rg_end > nrg) rg_end = nrg;
    if (rg_begin < 0 || rg_begin > rg_end) throw std::runtime_error("parquet: row group range out of bounds");
    int64_t n_rows = 0;
    for (int64_t g = rg_begin; g < rg_end; g++) n_rows += fm.row_groups[(size_t)g].num_rows;
    *rows_out = n_rows;

    // ---- read the chunks and walk their page headers ----
    int64_t chunk_total = 0;
    for (size_t ci = 0; ci < sel.size(); ci++)
        for (int64_t g = rg_begin; g < rg_end; g++) {
            const RowGroupMeta &rg = fm.row_groups[(size_t)g];
            if ((size_t)sel[ci] >= rg.columns.size()) throw std::runtime_error("parquet: row group has fewer column chunks than the schema has leaves");
            if (rg.columns[(size_t)sel[ci]].total_compressed < 0) throw std::runtime_error("parquet: negative column chunk size");
            chunk_total += rg.columns[(size_t)sel[ci]].total_compressed;
        }
    uint8_t *chunks = x.stage((size_t)chunk_total + 16);
    int64_t chunk_used = 0;
    {
        std::vector<ChunkRead> reads;
        int64_t at = 0;
        for (size_t ci = 0; ci < sel.size(); ci++)
            for (int64_t g = rg_begin; g < rg_end; g++) {
                const ColumnChunkMeta &cm = fm.row_groups[(size_t)g].columns[(size_t)sel[ci]];
                int64_t start = cm.data_page_offset;
                if (cm.dictionary_page_offset > 0 && cm.dictionary_page_offset < start) start = cm.dictionary_page_offset;
                if (start < 0 || start + cm.total_compressed > fb.size) throw std::runtime_error("parquet: column chunk lies outside the file");
                reads.push_back({start, cm.total_compressed, at});
                at += cm.total_compressed;
            }
        read_chunks(fb, reads, chunks);
    }
    std::vector<PageDesc> pages;
    std::vector<DictDesc> dicts;
    std::vector<ColumnDesc> cols(sel.size());
    std::vector<char> has_dict_pages(sel.size(), 0), has_plain_pages(sel.size(), 0);
    int64_t scratch_bytes = 0, dict_str_entries = 0;
    for (size_t ci = 0; ci < sel.size(); ci++) {
        const LeafInfo &leaf = leaves[(size_t)sel[ci]];
        ColumnDesc &cd = cols[ci];
        cd.physical = leaf.physical;
        cd.max_def = leaf.max_def;
        cd.width = leaf.width;
        cd.is_string = leaf.is_string ? 1 : 0;
        int64_t row = 0;
        for (int64_t g = rg_begin; g < rg_end; g++) {
            const RowGroupMeta &rg = fm.row_groups[(size_t)g];
            if ((size_t)sel[ci] >= rg.columns.size()) throw std::runtime_error("parquet: row group has fewer column chunks than the schema has leaves");
            const ColumnChunkMeta &cm = rg.columns[(size_t)sel[ci]];
            const std::string where = "parquet: column '" + leaf.name + "': ";
            if (cm.codec != C_UNCOMPRESSED && cm.codec != C_SNAPPY) throw std::runtime_error(where + "compression codec " + std::to_string(cm.codec) + " is not supported (UNCOMPRESSED and SNAPPY are)");
            if (cm.type != leaf.physical) throw std::runtime_error(where + "chunk type differs from the schema");
            const int64_t base = chunk_used;   // the bytes are already in the staging buffer (read_chunks above, same order)
            chunk_used += cm.total_compressed;
            int64_t pos = 0, seen = 0;
            int32_t cur_dict = -1;
            while (seen < cm.num_values) {
                if (pos >= cm.total_compressed) throw std::runtime_error(where + "column chunk ends before its value count");
                const PageHeader h = parse_page_header(chunks + base + pos, (size_t)(cm.total_compressed - pos));
                pos += (int64_t)h.header_bytes;
                if (h.compressed_size < 0 || pos + h.compressed_size > cm.total_compressed) throw std::runtime_error(where + "page runs past its column chunk");
                if (h.type == P_INDEX) {
                    pos += h.compressed_size;
                    continue;
                }
                PageDesc pg;
                pg.src_off = base + pos;
                pg.comp_size = h.compressed_size;
                pg.uncomp_size = h.uncompressed_size;
                pg.num_values = h.num_values;
                pg.page_type = h.type;
                pg.codec = cm.codec;
                pg.column = (int32_t)ci;
                pg.v2_def_len = h.v2_def_len;
                pg.v2_rep_len = h.v2_rep_len;
                pg.v2_compressed = h.v2_compressed ? 1 : 0;
                if (cm.codec != C_UNCOMPRESSED) {
                    pg.scratch_off = scratch_bytes;
                    scratch_bytes += ((int64_t)h.uncompressed_size + 15) / 16 * 16;
                } else if (h.uncompressed_size != h.compressed_size) {
                    throw std::runtime_error(where + "uncompressed page with differing sizes");
                }
                if (h.type == P_DICTIONARY) {
                    if (h.encoding != E_PLAIN && h.encoding != E_PLAIN_DICTIONARY) throw std::runtime_error(where + "dictionary page encoding " + std::to_string(h.encoding) + " is not supported");
                    pg.encoding = E_PLAIN;
                    DictDesc d;
                    d.page = (int32_t)pages.size();
                    d.count = h.num_values;
                    if (leaf.is_string) {
                        d.str_off_begin = dict_str_entries;
                        dict_str_entries += (int64_t)h.num_values + 1;
                    }
                    cur_dict = (int32_t)dicts.size();
                    dicts.push_back(d);
                } else if (h.type == P_DATA || h.type == P_DATA_V2) {
                    if (h.encoding == E_PLAIN) {
                        pg.encoding = E_PLAIN;
                        has_plain_pages[ci] = 1;
                    } else if (h.encoding == E_RLE_DICTIONARY || h.encoding == E_PLAIN_DICTIONARY) {
                        pg.encoding = E_RLE_DICTIONARY;
                        if (cur_dict < 0) throw std::runtime_error(where + "dictionary-encoded page without a dictionary page");
                        has_dict_pages[ci] = 1;
                    } else if (h.encoding == E_RLE && leaf.physical == T_BOOLEAN) {
                        pg.encoding = E_RLE;   // Booleans of data pages V2: a length-prefixed RLE / bit-packed run of 1-bit values
                    } else {
                        throw std::runtime_error(where + "page encoding " + std::to_string(h.encoding) + " (DELTA_* / BYTE_STREAM_SPLIT) is not supported");
                    }
                    if (h.type == P_DATA && leaf.max_def > 0 && h.def_encoding != E_RLE) throw std::runtime_error(where + "BIT_PACKED definition levels are not supported");
                    if (h.type == P_DATA_V2 && h.v2_rep_len != 0) throw std::runtime_error(where + "repetition levels in a flat column");
                    pg.dict = cur_dict;
                    pg.row_begin = row + seen;
                    seen += h.num_values;
                } else {
                    throw std::runtime_error(where + "unknown page type " + std::to_string(h.type));
                }
                pages.push_back(pg);
                pos += h.compressed_size;
            }
            if (seen != rg.num_rows) throw std::runtime_error(where + "value count differs from the row group's row count (nested data?)");
            row += rg.num_rows;
        }
    }

    // ---- buffers ----
    using Mem = typename X::Mem;
    Mem d_chunks = x.alloc((size_t)chunk_total + 16), d_scratch = x.alloc((size_t)scratch_bytes + 16), d_pages = x.alloc(std::max<size_t>(pages.size(), 1) * sizeof(PageDesc)),
        d_cols = x.alloc(std::max<size_t>(cols.size(), 1) * sizeof(ColumnDesc)), d_dicts = x.alloc(std::max<size_t>(dicts.size(), 1) * sizeof(DictDesc)),
        d_dict_offs = x.alloc((size_t)(dict_str_entries + 1) * 4);
    x.upload(d_chunks, chunks, (size_t)chunk_total);
    std::vector<Mem> values(sel.size()), valid(sel.size()), lengths(sel.size()), offsets(sel.size()), ridx(sel.size()), rsrc(sel.size()), bytes(sel.size());
    const size_t nr = (size_t)std::max<int64_t>(n_rows, 1);
    for (size_t ci = 0; ci < sel.size(); ci++) {
        ColumnDesc &cd = cols[ci];
        if (cd.is_string) {
            lengths[ci] = x.alloc((nr + 1) * 8);
            offsets[ci] = x.alloc((nr + 1) * 8);
            cd.out_lengths = (int64_t *)lengths[ci].ptr;
            cd.out_offsets = (int64_t *)offsets[ci].ptr;
            if (has_dict_pages[ci]) {
                ridx[ci] = x.alloc(nr * 4);
                cd.row_dict_index = (int32_t *)ridx[ci].ptr;
            }
            if (has_plain_pages[ci]) {
                rsrc[ci] = x.alloc(nr * 8);
                cd.row_src = (int64_t *)rsrc[ci].ptr;
            }
        } else {
            values[ci] = x.alloc(nr * (size_t)cd.width);
            cd.out_values = (uint8_t *)values[ci].ptr;
        }
        if (cd.max_def > 0) {
            valid[ci] = x.alloc(nr);
            cd.out_valid = (uint8_t *)valid[ci].ptr;
        }
    }
    x.upload(d_pages, pages.data(), pages.size() * sizeof(PageDesc));
    x.upload(d_cols, cols.data(), cols.size() * sizeof(ColumnDesc));
    x.upload(d_dicts, dicts.data(), dicts.size() * sizeof(DictDesc));
    ScanCtx ctx;
    ctx.chunks = (const uint8_t *)d_chunks.ptr;
    ctx.scratch = (uint8_t *)d_scratch.ptr;
    ctx.pages = (PageDesc *)d_pages.ptr;
    ctx.columns = (ColumnDesc *)d_cols.ptr;
    ctx.dicts = (DictDesc *)d_dicts.ptr;
    ctx.dict_str_offsets = (int32_t *)d_dict_offs.ptr;
    ctx.npages = (int32_t)pages.size();
    ctx.ndicts = (int32_t)dicts.size();

    auto check_pages = [&](const char *stage) {
        std::vector<PageDesc> back(pages.size());
        x.download(back.data(), d_pages.ptr, pages.size() * sizeof(PageDesc));
        for (size_t i = 0; i < back.size(); i++)
            if (back[i].status)
                throw std::runtime_error(std::string("parquet: column '") + leaves[(size_t)sel[(size_t)back[i].column]].name + "', page " + std::to_string(i) + " (" + stage +
                                         "): " + pq_error_text(back[i].status));
    };
    // ---- the stages ----
    if (!pages.empty()) {
        x.run(0, ctx, (int)pages.size());
        if (!dicts.empty()) x.run(1, ctx, (int)dicts.size());
        x.run(2, ctx, (int)pages.size());
        check_pages("values");
    }
    bool any_string = false;
    std::vector<int64_t> total_bytes(sel.size(), 0);
    for (size_t ci = 0; ci < sel.size(); ci++) {
        if (!cols[ci].is_string) continue;
        any_string = true;
        total_bytes[ci] = x.scan((int64_t *)lengths[ci].ptr, (int64_t *)offsets[ci].ptr, n_rows);
        bytes[ci] = x.alloc((size_t)std::max<int64_t>(total_bytes[ci], 1));
        cols[ci].out_bytes = (uint8_t *)bytes[ci].ptr;
    }
    if (any_string && !pages.empty()) {
        x.upload(d_cols, cols.data(), cols.size() * sizeof(ColumnDesc));
        x.run(3, ctx, (int)pages.size());
    }
    // ---- results: validity and Booleans packed to bitmaps ----
    std::vector<ColumnResult<X>> out(sel.size());
    for (size_t ci = 0; ci < sel.size(); ci++) {
        const LeafInfo &leaf = leaves[(size_t)sel[ci]];
        ColumnResult<X> &r = out[ci];
        r.name = leaf.name;
        r.dtype = leaf.dtype;
        r.length = n_rows;
        if (cols[ci].max_def > 0) {
            Mem bm = x.alloc((nr + 7) / 8 + 8);
            r.null_count = x.pack((const uint8_t *)valid[ci].ptr, n_rows, (uint8_t *)bm.ptr);
            if (r.null_count > 0) r.validity = bm;
        }
        if (cols[ci].is_string) {
            r.data = bytes[ci];
            r.offsets = offsets[ci];
            r.data_bytes = total_bytes[ci];
        } else if (leaf.dtype == B200_BOOL) {
            Mem bm = x.alloc((nr + 7) / 8 + 8);
            x.pack((const uint8_t *)values[ci].ptr, n_rows, (uint8_t *)bm.ptr);
            r.data = bm;
            r.data_bytes = (n_rows + 7) / 8;
        } else if (leaf.conv != CONV_NONE) {
            Mem fin = x.alloc(nr * (size_t)leaf.final_width);
            x.convert(leaf.conv, (const uint8_t *)values[ci].ptr, (uint8_t *)fin.ptr, n_rows, leaf.conv_factor);
            r.data = fin;
            r.data_bytes = n_rows * leaf.final_width;
        } else {
            r.data = values[ci];
            r.data_bytes = n_rows * cols[ci].width;
        }
    }
    return out;
}

}  // namespace pq
}  // namespace b200

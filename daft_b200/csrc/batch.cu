// Batches: Arrow C Data Interface import (host -> HBM) / export (HBM -> host), wrapping of
// existing device buffers, and the structural operations that never touch row data.
#include "common.h"
#include "ops.h"

namespace b200 {

static int32_t dtype_from_format(const char *f) {
    std::string s(f ? f : "");
    if (s == "n") return B200_NULL;
    if (s == "b") return B200_BOOL;
    if (s == "c") return B200_I8;
    if (s == "C") return B200_U8;
    if (s == "s") return B200_I16;
    if (s == "S") return B200_U16;
    if (s == "i") return B200_I32;
    if (s == "I") return B200_U32;
    if (s == "l") return B200_I64;
    if (s == "L") return B200_U64;
    if (s == "f") return B200_F32;
    if (s == "g") return B200_F64;
    if (s == "tdD") return B200_DATE32;
    if (s.rfind("ts", 0) == 0 && s.size() >= 4) {
        if (s.size() > 4) fail("import: timezone-aware timestamps are not supported (" + s + ")");
        return B200_TIMESTAMP;
    }
    if (s == "u" || s == "U") return B200_LARGE_UTF8;
    if (s == "z" || s == "Z") return B200_LARGE_BINARY;
    fail("import: unsupported Arrow format '" + s + "'");
}

static const char *format_from_dtype(int32_t dt) {
    switch (dt) {
        case B200_NULL: return "n";
        case B200_BOOL: return "b";
        case B200_I8: return "c";
        case B200_U8: return "C";
        case B200_I16: return "s";
        case B200_U16: return "S";
        case B200_I32: return "i";
        case B200_U32: return "I";
        case B200_I64: return "l";
        case B200_U64: return "L";
        case B200_F32: return "f";
        case B200_F64: return "g";
        case B200_DATE32: return "tdD";
        case B200_TIMESTAMP: return "tsu:";
        case B200_LARGE_UTF8: return "U";
        case B200_LARGE_BINARY: return "Z";
        case B200_LIST: return "+L";
    }
    return "n";
}

static inline bool host_bit(const uint8_t *bm, int64_t i) { return (bm[i >> 3] >> (i & 7)) & 1; }

static Buf upload(const void *host, size_t bytes) {
    Buf b = dev_alloc(bytes);
    if (bytes) B200_CUDA(cudaMemcpyAsync(b->ptr, host, bytes, cudaMemcpyHostToDevice, ctx().stream));
    return b;
}

// bitmap with arbitrary bit offset -> device bitmap at offset 0
static Buf upload_bitmap(const uint8_t *bm, int64_t offset, int64_t n, std::vector<std::vector<uint8_t>> &keep) {
    if ((offset & 7) == 0) return upload(bm + (offset >> 3), (size_t)((n + 7) / 8));
    keep.emplace_back((size_t)((n + 7) / 8), 0);
    auto &t = keep.back();
    for (int64_t i = 0; i < n; i++)
        if (host_bit(bm, offset + i)) t[i >> 3] |= (uint8_t)(1u << (i & 7));
    return upload(t.data(), t.size());
}

static Column import_column(const ArrowArray *a, const ArrowSchema *s, std::vector<std::vector<uint8_t>> &keep) {
    Column c;
    c.name = s->name ? s->name : "";
    c.dtype = dtype_from_format(s->format);
    if (c.dtype == B200_TIMESTAMP) c.name = c.name;  // unit is carried by the host-side schema
    c.length = a->length;
    const int64_t off = a->offset, n = a->length;
    if (a->dictionary) fail("import: dictionary arrays are not supported");
    if (c.dtype == B200_NULL) {
        c.null_count = n;
        return c;
    }
    const uint8_t *validity = a->n_buffers > 0 ? static_cast<const uint8_t *>(a->buffers[0]) : nullptr;
    int64_t nulls = a->null_count;
    if (validity && nulls != 0) {
        if (nulls < 0) {
            nulls = 0;
            for (int64_t i = 0; i < n; i++) nulls += !host_bit(validity, off + i);
        }
        if (nulls > 0) c.validity = upload_bitmap(validity, off, n, keep);
    } else {
        nulls = 0;
    }
    c.null_count = nulls;
    int w = dtype_width(c.dtype);
    if (w) {
        c.data = upload(static_cast<const uint8_t *>(a->buffers[1]) + off * w, (size_t)n * w);
        c.data_bytes = n * w;
    } else if (c.dtype == B200_BOOL) {
        c.data = upload_bitmap(static_cast<const uint8_t *>(a->buffers[1]), off, n, keep);
        c.data_bytes = (n + 7) / 8;
    } else {
        const std::string f(s->format);
        const uint8_t *bytes = static_cast<const uint8_t *>(a->buffers[2]);
        int64_t first, last;
        if (f == "U" || f == "Z") {
            const int64_t *o = static_cast<const int64_t *>(a->buffers[1]) + off;
            first = n >= 0 && o ? o[0] : 0;
            last = o ? o[n] : 0;
            if (first == 0) {
                c.offsets = upload(o, (size_t)(n + 1) * 8);
            } else {
                keep.emplace_back((size_t)(n + 1) * 8);
                int64_t *t = reinterpret_cast<int64_t *>(keep.back().data());
                for (int64_t i = 0; i <= n; i++) t[i] = o[i] - first;
                c.offsets = upload(t, (size_t)(n + 1) * 8);
            }
        } else {  // 32-bit offsets -> widen (Daft always holds Large; src/daft-core/src/kernels/hashing.rs:416-438)
            const int32_t *o = static_cast<const int32_t *>(a->buffers[1]) + off;
            first = o ? o[0] : 0;
            last = o ? o[n] : 0;
            keep.emplace_back((size_t)(n + 1) * 8);
            int64_t *t = reinterpret_cast<int64_t *>(keep.back().data());
            for (int64_t i = 0; i <= n; i++) t[i] = (int64_t)o[i] - first;
            c.offsets = upload(t, (size_t)(n + 1) * 8);
        }
        c.data_bytes = last - first;
        c.data = upload(bytes ? bytes + first : nullptr, (size_t)c.data_bytes);
    }
    return c;
}

Batch *import_batch(ArrowArray *array, ArrowSchema *schema) {
    require_device();
    if (!array || !schema || !schema->format || std::string(schema->format) != "+s")
        fail("import: expected a struct array (one child per column)");
    if (array->n_children != schema->n_children) fail("import: array/schema child count mismatch");
    if (array->offset != 0) fail("import: sliced struct arrays are not supported");
    auto out = std::make_unique<Batch>();
    out->num_rows = array->length;
    std::vector<std::vector<uint8_t>> keep;  // host temporaries must outlive the async copies
    for (int64_t i = 0; i < array->n_children; i++) {
        if (array->children[i]->length != array->length) fail("import: child length mismatch");
        out->cols.push_back(import_column(array->children[i], schema->children[i], keep));
    }
    B200_CUDA(cudaStreamSynchronize(ctx().stream));
    if (array->release) array->release(array);
    if (schema->release) schema->release(schema);
    return out.release();
}

// ---- export -----------------------------------------------------------------------------------
struct ExportPriv {
    std::vector<void *> bufs;          // malloc'd
    std::vector<const void *> bufptrs; // ArrowArray.buffers
    std::vector<ArrowArray *> children;
    std::vector<ArrowSchema *> schildren;
    std::string name, format;
};
static void release_array(ArrowArray *a) {
    if (!a || !a->release) return;
    auto *p = static_cast<ExportPriv *>(a->private_data);
    for (auto *c : p->children) {
        if (c->release) c->release(c);
        free(c);
    }
    for (void *b : p->bufs) free(b);
    delete p;
    a->release = nullptr;
}
static void release_schema(ArrowSchema *s) {
    if (!s || !s->release) return;
    auto *p = static_cast<ExportPriv *>(s->private_data);
    for (auto *c : p->schildren) {
        if (c->release) c->release(c);
        free(c);
    }
    delete p;
    s->release = nullptr;
}
static void *host_copy(const void *dev, size_t bytes, ExportPriv *p) {
    void *h = malloc(bytes ? bytes + 8 : 8);
    if (!h) fail("export: out of host memory");
    if (bytes) B200_CUDA(cudaMemcpyAsync(h, dev, bytes, cudaMemcpyDeviceToHost, ctx().stream));
    p->bufs.push_back(h);
    return h;
}
static void export_column(const Column &c, ArrowArray *a, ArrowSchema *s) {
    auto *ap = new ExportPriv, *sp = new ExportPriv;
    sp->name = c.name;
    sp->format = format_from_dtype(c.dtype);
    memset(s, 0, sizeof(*s));
    s->format = sp->format.c_str();
    s->name = sp->name.c_str();
    s->flags = 2;  // nullable
    s->release = release_schema;
    s->private_data = sp;
    memset(a, 0, sizeof(*a));
    a->length = c.length;
    a->null_count = c.null_count;
    a->release = release_array;
    a->private_data = ap;
    const int64_t n = c.length;
    if (c.dtype == B200_NULL) {
        a->n_buffers = 0;
        a->null_count = n;
        return;
    }
    const void *validity = c.null_count > 0 ? host_copy(c.validity->ptr, (size_t)((n + 7) / 8), ap) : nullptr;
    ap->bufptrs.push_back(validity);
    int w = dtype_width(c.dtype);
    if (w) {
        ap->bufptrs.push_back(host_copy(c.data ? c.data->ptr : nullptr, (size_t)n * w, ap));
    } else if (c.dtype == B200_BOOL) {
        ap->bufptrs.push_back(host_copy(c.data ? c.data->ptr : nullptr, (size_t)((n + 7) / 8), ap));
    } else if (c.dtype == B200_LIST) {  // Arrow large_list: [validity, int64 offsets] + one child ("item")
        if (!c.child || !c.offsets) fail("export: list column without offsets / child");
        ap->bufptrs.push_back(host_copy(c.offsets->ptr, (size_t)(n + 1) * 8, ap));
        auto *ca = static_cast<ArrowArray *>(calloc(1, sizeof(ArrowArray)));
        auto *cs = static_cast<ArrowSchema *>(calloc(1, sizeof(ArrowSchema)));
        Column item = *c.child;
        item.name = "item";
        export_column(item, ca, cs);
        ap->children.push_back(ca);
        sp->schildren.push_back(cs);
        a->n_children = 1;
        a->children = ap->children.data();
        s->n_children = 1;
        s->children = sp->schildren.data();
    } else {
        if (c.offsets) {
            ap->bufptrs.push_back(host_copy(c.offsets->ptr, (size_t)(n + 1) * 8, ap));
        } else {
            int64_t *z = static_cast<int64_t *>(calloc((size_t)n + 1, 8));
            ap->bufs.push_back(z);
            ap->bufptrs.push_back(z);
        }
        ap->bufptrs.push_back(host_copy(c.data ? c.data->ptr : nullptr, (size_t)c.data_bytes, ap));
    }
    a->n_buffers = (int64_t)ap->bufptrs.size();
    a->buffers = ap->bufptrs.data();
}

// ---- small batches: every device buffer is packed into one staging block by ONE kernel and comes back with ONE
// copy (a 4-row aggregate result has ~20 buffers; 20 tiny cudaMemcpyAsync calls cost more than the query's epilogue)
struct PackDesc {
    const uint8_t *src;
    int64_t dst_off, bytes;
};
constexpr int PACK_MAX = 96;
struct PackSet {
    int n;
    PackDesc d[PACK_MAX];
};
__global__ void pack_buffers(PackSet ps, uint8_t *__restrict__ dst) {
    const PackDesc d = ps.d[blockIdx.y];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.bytes; i += (int64_t)gridDim.x * blockDim.x)
        dst[d.dst_off + i] = d.src[i];
}

static bool export_batch_packed(const Batch &b, ArrowArray *out_array, ArrowSchema *out_schema) {
    // plan the layout
    struct Plan { int col, slot; int64_t off, bytes; const void *src; };
    std::vector<Plan> plan;
    int64_t total = 0;
    auto add = [&](int col, int slot, const void *src, int64_t bytes) {
        plan.push_back({col, slot, total, bytes, src});
        total += (bytes + 15) & ~(int64_t)15;
    };
    for (size_t ci = 0; ci < b.cols.size(); ci++) {
        const Column &c = b.cols[ci];
        const int64_t n = c.length;
        if (c.dtype == B200_NULL) continue;
        if (c.dtype == B200_LIST) return false;  // nested: the general path
        if (c.null_count > 0) add((int)ci, 0, c.validity->ptr, (n + 7) / 8);
        const int w = dtype_width(c.dtype);
        if (w) add((int)ci, 1, c.data ? c.data->ptr : nullptr, n * w);
        else if (c.dtype == B200_BOOL) add((int)ci, 1, c.data ? c.data->ptr : nullptr, (n + 7) / 8);
        else {
            if (!c.offsets) return false;
            add((int)ci, 1, c.offsets->ptr, (n + 1) * 8);
            add((int)ci, 2, c.data ? c.data->ptr : nullptr, c.data_bytes);
        }
    }
    if (total > (1 << 20) || plan.size() > PACK_MAX) return false;
    uint8_t *host = static_cast<uint8_t *>(malloc((size_t)total + 64));
    if (!host) fail("export: out of host memory");
    if (total > 0) {
        Buf stage = dev_alloc((size_t)total);
        PackSet ps;
        ps.n = 0;
        int64_t maxb = 1;
        for (auto &p : plan)
            if (p.bytes > 0 && p.src) {
                ps.d[ps.n++] = {static_cast<const uint8_t *>(p.src), p.off, p.bytes};
                maxb = std::max(maxb, p.bytes);
            }
        if (ps.n) {
            dim3 grid((unsigned)std::min<int64_t>((maxb + 255) / 256, 64), (unsigned)ps.n);
            pack_buffers<<<grid, 256, 0, ctx().stream>>>(ps, as<uint8_t>(stage));
            count_launch();
        }
        B200_CUDA(cudaMemcpyAsync(host, stage->ptr, (size_t)total, cudaMemcpyDeviceToHost, ctx().stream));
        B200_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    auto *ap = new ExportPriv, *sp = new ExportPriv;
    sp->format = "+s";
    ap->bufs.push_back(host);  // the parent owns the single block
    memset(out_schema, 0, sizeof(*out_schema));
    memset(out_array, 0, sizeof(*out_array));
    for (size_t ci = 0; ci < b.cols.size(); ci++) {
        const Column &c = b.cols[ci];
        auto *ca = static_cast<ArrowArray *>(calloc(1, sizeof(ArrowArray)));
        auto *cs = static_cast<ArrowSchema *>(calloc(1, sizeof(ArrowSchema)));
        auto *cap = new ExportPriv, *csp = new ExportPriv;
        csp->name = c.name;
        csp->format = format_from_dtype(c.dtype);
        cs->format = csp->format.c_str();
        cs->name = csp->name.c_str();
        cs->flags = 2;
        cs->release = release_schema;
        cs->private_data = csp;
        ca->length = c.length;
        ca->null_count = c.dtype == B200_NULL ? c.length : c.null_count;
        ca->release = release_array;
        ca->private_data = cap;
        if (c.dtype != B200_NULL) {
            const int nslots = is_string(c.dtype) ? 3 : 2;
            cap->bufptrs.assign((size_t)nslots, nullptr);
            for (auto &p : plan)
                if (p.col == (int)ci) cap->bufptrs[(size_t)p.slot] = host + p.off;
            ca->n_buffers = nslots;
            ca->buffers = cap->bufptrs.data();
        }
        ap->children.push_back(ca);
        sp->schildren.push_back(cs);
    }
    out_schema->format = sp->format.c_str();
    out_schema->name = sp->name.c_str();
    out_schema->n_children = (int64_t)sp->schildren.size();
    out_schema->children = sp->schildren.data();
    out_schema->release = release_schema;
    out_schema->private_data = sp;
    out_array->length = b.num_rows;
    out_array->n_buffers = 1;
    ap->bufptrs.push_back(nullptr);
    out_array->buffers = ap->bufptrs.data();
    out_array->n_children = (int64_t)ap->children.size();
    out_array->children = ap->children.data();
    out_array->release = release_array;
    out_array->private_data = ap;
    return true;
}

void export_batch(const Batch &b, ArrowArray *out_array, ArrowSchema *out_schema) {
    require_device();
    if (b.num_rows <= 65536 && export_batch_packed(b, out_array, out_schema)) return;
    auto *ap = new ExportPriv, *sp = new ExportPriv;
    sp->format = "+s";
    sp->name = "";
    memset(out_schema, 0, sizeof(*out_schema));
    memset(out_array, 0, sizeof(*out_array));
    for (auto &c : b.cols) {
        auto *ca = static_cast<ArrowArray *>(calloc(1, sizeof(ArrowArray)));
        auto *cs = static_cast<ArrowSchema *>(calloc(1, sizeof(ArrowSchema)));
        ap->children.push_back(ca);
        sp->schildren.push_back(cs);
        export_column(c, ca, cs);
    }
    B200_CUDA(cudaStreamSynchronize(ctx().stream));
    out_schema->format = sp->format.c_str();
    out_schema->name = sp->name.c_str();
    out_schema->n_children = (int64_t)sp->schildren.size();
    out_schema->children = sp->schildren.data();
    out_schema->release = release_schema;
    out_schema->private_data = sp;
    out_array->length = b.num_rows;
    out_array->null_count = 0;
    out_array->n_buffers = 1;
    ap->bufptrs.push_back(nullptr);
    out_array->buffers = ap->bufptrs.data();
    out_array->n_children = (int64_t)ap->children.size();
    out_array->children = ap->children.data();
    out_array->release = release_array;
    out_array->private_data = ap;
}

// host buffers (malloc'd by the caller, ownership moves) -> Arrow struct array
void export_host_columns(std::vector<HostCol> &cols, int64_t num_rows, ArrowArray *out_array, ArrowSchema *out_schema) {
    auto *ap = new ExportPriv, *sp = new ExportPriv;
    sp->format = "+s";
    memset(out_schema, 0, sizeof(*out_schema));
    memset(out_array, 0, sizeof(*out_array));
    for (auto &h : cols) {
        auto *ca = static_cast<ArrowArray *>(calloc(1, sizeof(ArrowArray)));
        auto *cs = static_cast<ArrowSchema *>(calloc(1, sizeof(ArrowSchema)));
        auto *cap = new ExportPriv, *csp = new ExportPriv;
        csp->name = h.name;
        csp->format = format_from_dtype(h.dtype);
        cs->format = csp->format.c_str();
        cs->name = csp->name.c_str();
        cs->flags = 2;
        cs->release = release_schema;
        cs->private_data = csp;
        ca->length = h.length;
        ca->null_count = 0;
        ca->release = release_array;
        ca->private_data = cap;
        cap->bufptrs.push_back(nullptr);
        if (h.offsets) {
            cap->bufs.push_back(h.offsets);
            cap->bufptrs.push_back(h.offsets);
        }
        cap->bufs.push_back(h.data);
        cap->bufptrs.push_back(h.data);
        ca->n_buffers = (int64_t)cap->bufptrs.size();
        ca->buffers = cap->bufptrs.data();
        ap->children.push_back(ca);
        sp->schildren.push_back(cs);
    }
    out_schema->format = sp->format.c_str();
    out_schema->name = sp->name.c_str();
    out_schema->n_children = (int64_t)sp->schildren.size();
    out_schema->children = sp->schildren.data();
    out_schema->release = release_schema;
    out_schema->private_data = sp;
    out_array->length = num_rows;
    out_array->n_buffers = 1;
    ap->bufptrs.push_back(nullptr);
    out_array->buffers = ap->bufptrs.data();
    out_array->n_children = (int64_t)ap->children.size();
    out_array->children = ap->children.data();
    out_array->release = release_array;
    out_array->private_data = ap;
}

__global__ void rebase_offsets_kernel(const int64_t *__restrict__ src, int64_t n1, int64_t first, int64_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n1; i += stride) out[i] = src[i] - first;
}
static Buf rebase_offsets(const int64_t *src, int64_t n, int64_t first) {
    Buf out = dev_alloc((size_t)(n + 1) * 8);
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 256) / 256, (int64_t)ctx().sm_count * 8));
    rebase_offsets_kernel<<<grid, 256, 0, ctx().stream>>>(src, n + 1, first, as<int64_t>(out));
    count_launch();
    return out;
}

Batch *batch_from_device(const b200_column_desc *cols, int32_t ncols, int64_t num_rows) {
    require_device();
    auto out = std::make_unique<Batch>();
    out->num_rows = num_rows;
    for (int i = 0; i < ncols; i++) {
        const auto &d = cols[i];
        if (d.length != num_rows) fail("from_device: column length != num_rows");
        Column c;
        c.name = d.name ? d.name : "";
        c.dtype = d.dtype;
        c.length = d.length;
        c.null_count = d.validity ? d.null_count : 0;
        if (d.dtype == B200_NULL) c.null_count = d.length;
        if (d.validity && d.null_count > 0) c.validity = dev_borrow(d.validity, (size_t)((d.length + 7) / 8));
        if (d.offsets) c.offsets = dev_borrow(d.offsets, (size_t)(d.length + 1) * 8);
        if (d.data) c.data = dev_borrow(d.data, (size_t)d.data_bytes);
        c.data_bytes = d.data_bytes;
        if (is_string(c.dtype) && !d.offsets) fail("from_device: string column without offsets");
        if (is_string(c.dtype) && d.length > 0) {
            // A sliced Arrow string array has offsets[0] != 0.  The operators that copy value bytes wholesale (concat, the
            // packed all-gather, small permutation gathers) take the bytes from `data` + 0: re-base such a column here by
            // viewing the data buffer from its first referenced byte and shifting the offsets once (one 16-byte read back).
            int64_t ends[2] = {0, 0};
            B200_CUDA(cudaMemcpyAsync(&ends[0], d.offsets, 8, cudaMemcpyDeviceToHost, ctx().stream));
            B200_CUDA(cudaMemcpyAsync(&ends[1], (const int64_t *)d.offsets + d.length, 8, cudaMemcpyDeviceToHost, ctx().stream));
            B200_CUDA(cudaStreamSynchronize(ctx().stream));
            if (ends[0] < 0 || ends[1] < ends[0] || ends[1] > d.data_bytes) fail("from_device: string offsets outside the data buffer");
            if (ends[0] != 0) {
                c.offsets = rebase_offsets((const int64_t *)d.offsets, d.length, ends[0]);
                c.data = dev_borrow((const uint8_t *)d.data + ends[0], (size_t)(ends[1] - ends[0]));
            }
            c.data_bytes = ends[1] - ends[0];
        }
        out->cols.push_back(std::move(c));
    }
    return out.release();
}

}  // namespace b200

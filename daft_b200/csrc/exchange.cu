// a12: the repartition exchange as an NCCL all-to-all over NVLink 5 / NVSwitch.
//
// Replaces RepartitionSink -> Arrow-IPC spill files -> Arrow Flight do_get
// (src/daft-local-execution/src/sinks/repartition.rs:100-370, src/daft-shuffles/src/shuffle_cache.rs:36-434,
// src/daft-shuffles/src/server/flight_server.rs:70-383): no disk, no IPC framing, no driver-side
// transposition.  Map side = the same hash_rows % world placement as a3 (so results agree with the
// reference's partitioning), rows gathered destination-contiguous; the G x G row/byte count matrices are
// all-gathered first, then every column buffer moves with one grouped ncclSend/ncclRecv.
// One process per GPU; the communicator is created from an id the host broadcasts out of band.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>

#include "common.h"
#include "kernels.cuh"
#include "ops.h"

struct b200_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // peer-store exchange: every rank owns a receive arena (cudaMalloc, exported over CUDA IPC) that all peers map
    void *arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<void *> peer_arena;  // [world]: mapped base of every rank's arena (own entry = arena)
    int p2p_state = 0;               // 0 untried, 1 usable, -1 unavailable (no IPC / no peer access): NCCL send/recv instead
};

namespace b200 {

// NCCL is bound at first use with dlopen, not at link time: a process that also hosts PyTorch must share
// PyTorch's libnccl.so.2 (same SONAME, possibly newer than the system copy), and processes that never
// exchange data should not load NCCL at all.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
};
static NcclApi &nccl() {
    static NcclApi api = [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process (e.g. PyTorch's)
        if (!h) {
            const char *p = getenv("B200_NCCL_LIB");
            if (p && *p) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
        }
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) fail(std::string("cannot load libnccl.so.2: ") + dlerror());
        NcclApi a;
        auto sym = [&](const char *n) {
            void *f = dlsym(h, n);
            if (!f) fail(std::string("libnccl.so.2 lacks ") + n);
            return f;
        };
        a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
        a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
        a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
        a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
        a.Send = (decltype(a.Send))sym("ncclSend");
        a.Recv = (decltype(a.Recv))sym("ncclRecv");
        a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
        return a;
    }();
    return api;
}

#define B200_NCCL(expr)                                                                                  \
    do {                                                                                                 \
        ncclResult_t _r = (expr);                                                                        \
        if (_r != ncclSuccess) fail(std::string("NCCL error: ") + nccl().GetErrorString(_r) + " in " #expr); \
    } while (0)

void comm_unique_id(uint8_t *out128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    B200_NCCL(nccl().GetUniqueId(&id));
    memcpy(out128, &id, 128);
}
b200_comm *comm_init(const uint8_t *id128, int32_t rank, int32_t world) {
    require_device();
    auto c = std::make_unique<b200_comm>();
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    B200_NCCL(nccl().CommInitRank(&c->comm, world, id, rank));
    c->rank = rank;
    c->world = world;
    return c.release();
}
static void arena_release(b200_comm *c) {
    for (int r = 0; r < (int)c->peer_arena.size(); r++)
        if (r != c->rank && c->peer_arena[r]) cudaIpcCloseMemHandle(c->peer_arena[r]);
    c->peer_arena.clear();
    if (c->arena) cudaFree(c->arena);
    c->arena = nullptr;
    c->arena_bytes = 0;
}
void comm_free(b200_comm *c) {
    if (c) arena_release(c);
    if (c && c->comm) nccl().CommDestroy(c->comm);
    delete c;
}

static int egrid(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx().sm_count * 8)); }

__global__ void bits_to_bytes(const uint8_t *__restrict__ bits, int64_t n, uint8_t *__restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = bits == nullptr ? 1 : (uint8_t)bit_at(bits, i);
}
__global__ void bytes_to_bits(const uint8_t *__restrict__ bytes, int64_t n, uint32_t *__restrict__ out) {
    const int64_t n_round = (n + 31) & ~(int64_t)31;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n_round; i += stride) {
        const unsigned m = __ballot_sync(0xffffffffu, i < n && bytes[i] != 0);
        if ((threadIdx.x & 31) == 0) out[i >> 5] = m;
    }
}
__global__ void offsets_to_lengths(const int64_t *__restrict__ off, int64_t n, int64_t *__restrict__ len) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) len[i] = off[i + 1] - off[i];
}

// who sends how much to whom; *_off/_cnt in elements of the buffer being moved
struct Route {
    std::vector<int64_t> send_off, send_cnt, recv_off, recv_cnt;
    int64_t recv_total = 0;
};

// all ranks learn everybody's send counts, m[src * (world + nextra) + dst], plus `nextra` extra words per rank (the
// exchange piggybacks each column's null count: validity only travels for columns some rank has nulls in)
static std::vector<int64_t> gather_counts(b200_comm *c, const std::vector<int64_t> &send_cnt, const std::vector<int64_t> &extra) {
    cudaStream_t st = ctx().stream;
    const int W = c->world;
    const size_t per = (size_t)W + extra.size();
    std::vector<int64_t> mine(send_cnt);
    mine.insert(mine.end(), extra.begin(), extra.end());
    Buf d_in = dev_alloc(per * 8), d_out = dev_alloc(per * W * 8);
    B200_CUDA(cudaMemcpyAsync(d_in->ptr, mine.data(), per * 8, cudaMemcpyHostToDevice, st));
    B200_NCCL(nccl().AllGather(d_in->ptr, d_out->ptr, per, ncclInt64, c->comm, st));
    std::vector<int64_t> m(per * W);
    B200_CUDA(cudaMemcpyAsync(m.data(), d_out->ptr, per * W * 8, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return m;
}
static Route make_route(b200_comm *c, const std::vector<int64_t> &send_off, const std::vector<int64_t> &send_cnt,
                        const std::vector<int64_t> &extra = {}, std::vector<int64_t> *extra_sum = nullptr) {
    const int W = c->world;
    Route r;
    r.send_off = send_off;
    r.send_cnt = send_cnt;
    const std::vector<int64_t> m = gather_counts(c, send_cnt, extra);
    const size_t per = (size_t)W + extra.size();
    r.recv_off.assign(W, 0);
    r.recv_cnt.assign(W, 0);
    for (int src = 0; src < W; src++) {
        r.recv_cnt[src] = m[(size_t)src * per + c->rank];
        r.recv_off[src] = r.recv_total;
        r.recv_total += r.recv_cnt[src];
    }
    if (extra_sum) {
        extra_sum->assign(extra.size(), 0);
        for (int src = 0; src < W; src++)
            for (size_t e = 0; e < extra.size(); e++) (*extra_sum)[e] += m[(size_t)src * per + W + e];
    }
    return r;
}
static void move_buffer(b200_comm *c, const Route &r, const void *send, void *recv, size_t elem_bytes) {
    cudaStream_t st = ctx().stream;
    B200_NCCL(nccl().GroupStart());
    for (int p = 0; p < c->world; p++) {
        if (r.send_cnt[p])
            B200_NCCL(nccl().Send((const uint8_t *)send + (size_t)r.send_off[p] * elem_bytes, (size_t)r.send_cnt[p] * elem_bytes, ncclUint8, p,
                               c->comm, st));
        if (r.recv_cnt[p])
            B200_NCCL(nccl().Recv((uint8_t *)recv + (size_t)r.recv_off[p] * elem_bytes, (size_t)r.recv_cnt[p] * elem_bytes, ncclUint8, p, c->comm,
                               st));
    }
    B200_NCCL(nccl().GroupEnd());
    count_launch();
}

// `src` columns are already destination-contiguous (rows [row_route.send_off[p], +send_cnt[p]) go to peer p)
// nulls_anywhere[col] = some rank's piece of that column has nulls (null = unknown: validity always travels)
static Batch *route_batch(b200_comm *c, const Batch &src, const Route &rows, const std::vector<int64_t> *nulls_anywhere) {
    cudaStream_t st = ctx().stream;
    const int W = c->world;
    const int64_t n_out = rows.recv_total, n_in = src.num_rows;
    auto out = std::make_unique<Batch>();
    out->num_rows = n_out;
    for (size_t ci = 0; ci < src.cols.size(); ci++) {
        const Column &col = src.cols[ci];
        Column o;
        o.name = col.name;
        o.dtype = col.dtype;
        o.length = n_out;
        if (col.dtype == B200_NULL) {
            o.null_count = n_out;
            out->cols.push_back(std::move(o));
            continue;
        }
        // validity travels one byte per row (bitmaps cannot be cut at arbitrary bit offsets) -- only for columns that have
        // nulls on some rank
        if (!nulls_anywhere || (*nulls_anywhere)[ci] > 0) {
            Buf sb = dev_alloc((size_t)std::max<int64_t>(n_in, 1)), rb = dev_alloc((size_t)std::max<int64_t>(n_out, 1));
            if (n_in) {
                bits_to_bytes<<<egrid(n_in), 256, 0, st>>>(col.null_count > 0 ? as<uint8_t>(col.validity) : nullptr, n_in, as<uint8_t>(sb));
                count_launch();
            }
            move_buffer(c, rows, sb->ptr, rb->ptr, 1);
            o.validity = dev_alloc((size_t)((n_out + 31) / 32 + 1) * 4);
            if (n_out) {
                bytes_to_bits<<<egrid(n_out), 256, 0, st>>>(as<uint8_t>(rb), n_out, as<uint32_t>(o.validity));
                count_launch();
            }
            o.null_count = n_out - count_set_bits(as<uint8_t>(o.validity), n_out);
            if (o.null_count == 0) o.validity.reset();
        }
        const int w = dtype_width(col.dtype);
        if (w) {
            o.data = dev_alloc((size_t)std::max<int64_t>(n_out, 1) * w);
            o.data_bytes = n_out * w;
            move_buffer(c, rows, col.data ? col.data->ptr : nullptr, o.data->ptr, (size_t)w);
        } else if (col.dtype == B200_BOOL) {
            Buf sb = dev_alloc((size_t)std::max<int64_t>(n_in, 1)), rb = dev_alloc((size_t)std::max<int64_t>(n_out, 1));
            if (n_in) {
                bits_to_bytes<<<egrid(n_in), 256, 0, st>>>(as<uint8_t>(col.data), n_in, as<uint8_t>(sb));
                count_launch();
            }
            move_buffer(c, rows, sb->ptr, rb->ptr, 1);
            o.data = dev_alloc((size_t)((n_out + 31) / 32 + 1) * 4);
            o.data_bytes = (n_out + 7) / 8;
            if (n_out) {
                bytes_to_bits<<<egrid(n_out), 256, 0, st>>>(as<uint8_t>(rb), n_out, as<uint32_t>(o.data));
                count_launch();
            }
        } else {  // strings: lengths by row route, bytes by a byte route
            Buf slen = dev_alloc((size_t)std::max<int64_t>(n_in, 1) * 8), rlen = dev_alloc((size_t)(n_out + 1) * 8);
            if (n_in) {
                offsets_to_lengths<<<egrid(n_in), 256, 0, st>>>(as<int64_t>(col.offsets), n_in, as<int64_t>(slen));
                count_launch();
            }
            move_buffer(c, rows, slen->ptr, rlen->ptr, 8);
            o.offsets = dev_alloc((size_t)(n_out + 1) * 8);
            exclusive_scan_i64(as<int64_t>(rlen), as<int64_t>(o.offsets), n_out, as<int64_t>(o.offsets) + n_out);
            // byte offsets of each destination's slice on the send side
            std::vector<int64_t> cut((size_t)W + 1, 0), boff(W), bcnt(W);
            for (int p = 0; p <= W; p++) {
                const int64_t row = p < W ? rows.send_off[p] : n_in;
                if (n_in) B200_CUDA(cudaMemcpyAsync(&cut[p], as<int64_t>(col.offsets) + row, 8, cudaMemcpyDeviceToHost, st));
            }
            B200_CUDA(cudaStreamSynchronize(st));
            for (int p = 0; p < W; p++) {
                boff[p] = cut[p] - cut[0];
                const int64_t endrow = rows.send_off[p] + rows.send_cnt[p];
                int64_t endb = 0;
                if (n_in) {
                    B200_CUDA(cudaMemcpyAsync(&endb, as<int64_t>(col.offsets) + endrow, 8, cudaMemcpyDeviceToHost, st));
                    B200_CUDA(cudaStreamSynchronize(st));
                }
                bcnt[p] = endb - cut[p];
            }
            const Route br = make_route(c, boff, bcnt);
            o.data = dev_alloc((size_t)br.recv_total);
            o.data_bytes = br.recv_total;
            move_buffer(c, br, col.data ? as<uint8_t>(col.data) + cut[0] : nullptr, o.data->ptr, 1);
        }
        out->cols.push_back(std::move(o));
    }
    B200_CUDA(cudaStreamSynchronize(st));
    return out.release();
}

// ---- peer-store exchange: partition + all-to-all in ONE kernel over NVLink peer memory --------------------------------
// Every rank maps every other rank's receive arena (CUDA IPC).  After the count matrix is known, all ranks derive the
// same receive layout (column c of rank d: a region of recv_total[d] x width bytes, sender s's rows at row offset
// sum of counts of the senders before it), and the partition kernel writes each row of a fixed-width, never-null column
// straight into its destination GPU's arena -- the payload crosses NVLink once, tile by tile, while the kernel is still
// partitioning (partition.cu partition_scatter<PEER>); nothing is staged in local HBM and no send / recv is posted.
// A one-word NCCL all-reduce closes the step (all writes landed), then each rank copies its columns out of the arena
// (device-to-device), which frees the arena for the next exchange: the next exchange's count all-gather orders the
// next writes behind these copies on every rank.
static bool arena_ensure(b200_comm *c, size_t need) {  // collective: `need` is identical on every rank
    if (c->p2p_state < 0) return false;
    if (c->arena && c->arena_bytes >= need) return true;
    cudaStream_t st = ctx().stream;
    const int W = c->world;
    B200_CUDA(cudaStreamSynchronize(st));
    arena_release(c);
    size_t bytes = std::max<size_t>(need + need / 4, (size_t)64 << 20);
    bytes = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&c->arena, bytes) != cudaSuccess) {
        cudaGetLastError();
        c->arena = nullptr;
        ok = 0;
    } else if (cudaIpcGetMemHandle(&mine, c->arena) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    // all-gather {ok, handle}: 8 + 64 bytes per rank
    constexpr size_t REC = 8 + sizeof(cudaIpcMemHandle_t);
    std::vector<uint8_t> rec(REC, 0), all(REC * W);
    memcpy(rec.data(), &ok, sizeof(int));
    memcpy(rec.data() + 8, &mine, sizeof(mine));
    Buf d_in = dev_alloc(REC), d_out = dev_alloc(REC * W);
    B200_CUDA(cudaMemcpyAsync(d_in->ptr, rec.data(), REC, cudaMemcpyHostToDevice, st));
    B200_NCCL(nccl().AllGather(d_in->ptr, d_out->ptr, REC, ncclUint8, c->comm, st));
    B200_CUDA(cudaMemcpyAsync(all.data(), d_out->ptr, REC * W, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    bool all_ok = true;
    for (int r = 0; r < W; r++) {
        int o = 0;
        memcpy(&o, all.data() + r * REC, sizeof(int));
        all_ok = all_ok && o;
    }
    int mapped = all_ok ? 1 : 0;
    if (all_ok) {
        c->peer_arena.assign(W, nullptr);
        for (int r = 0; r < W && mapped; r++) {
            if (r == c->rank) {
                c->peer_arena[r] = c->arena;
                continue;
            }
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + r * REC + 8, sizeof(h));
            if (cudaIpcOpenMemHandle(&c->peer_arena[r], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                c->peer_arena[r] = nullptr;
                mapped = 0;
            }
        }
    }
    // every rank must have mapped every peer: agree with a MIN all-reduce
    Buf flag = dev_alloc(4);
    B200_CUDA(cudaMemcpyAsync(flag->ptr, &mapped, 4, cudaMemcpyHostToDevice, st));
    B200_NCCL(nccl().AllReduce(flag->ptr, flag->ptr, 1, ncclInt32, ncclMin, c->comm, st));
    B200_CUDA(cudaMemcpyAsync(&mapped, flag->ptr, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (!mapped) {
        arena_release(c);
        c->p2p_state = -1;
        return false;
    }
    c->arena_bytes = bytes;
    c->p2p_state = 1;
    return true;
}

static Batch *route_batch(b200_comm *c, const Batch &src, const Route &rows, const std::vector<int64_t> *nulls_anywhere);

// returns nullptr when the peer path is not available (the caller takes the NCCL route); collective decisions only
static Batch *exchange_peer_store(b200_comm *c, const Batch &b, const PartSrc &hashes) {
    const int W = c->world;
    const char *env = getenv("B200_EXCHANGE_P2P");
    if ((env && *env == '0') || c->p2p_state < 0 || W > 8 || W < 2) return nullptr;  // B200_EXCHANGE_P2P=0: NCCL route only
    cudaStream_t st = ctx().stream;
    // columns the kernel can place itself: fixed width, no nulls ON ANY RANK (decided from the gathered null counts)
    PartitionPlan pl = partition_count(hashes, b.num_rows, W);
    std::vector<int64_t> scnt(W), soff(W), nulls, nulls_anywhere;
    for (int p = 0; p < W; p++) {
        soff[p] = pl.offsets[p];
        scnt[p] = pl.offsets[p + 1] - pl.offsets[p];
    }
    for (const Column &col : b.cols) nulls.push_back(col.null_count);
    std::vector<int64_t> m = gather_counts(c, scnt, nulls);
    const size_t per = (size_t)W + nulls.size();
    nulls_anywhere.assign(nulls.size(), 0);
    for (int src = 0; src < W; src++)
        for (size_t e = 0; e < nulls.size(); e++) nulls_anywhere[e] += m[(size_t)src * per + W + e];
    std::vector<int> peer_cols, other_cols;
    for (size_t ci = 0; ci < b.cols.size(); ci++) {
        const Column &col = b.cols[ci];
        if (dtype_width(col.dtype) && nulls_anywhere[ci] == 0 && peer_cols.size() < 16) peer_cols.push_back((int)ci);
        else if (col.dtype != B200_NULL) other_cols.push_back((int)ci);
    }
    if (peer_cols.empty()) return nullptr;
    // receive layout of every rank (identical arithmetic everywhere): column regions back to back, 256-byte aligned
    std::vector<int64_t> recv_total(W, 0);
    for (int d = 0; d < W; d++)
        for (int s2 = 0; s2 < W; s2++) recv_total[d] += m[(size_t)s2 * per + d];
    auto region = [&](int d, size_t k) {  // byte offset of peer column k in rank d's arena
        size_t off = 0;
        for (size_t j = 0; j < k; j++) off += ((size_t)recv_total[d] * dtype_width(b.cols[peer_cols[j]].dtype) + 255) & ~(size_t)255;
        return off;
    };
    size_t need = 0;
    for (int d = 0; d < W; d++) need = std::max(need, region(d, peer_cols.size()));
    if (!arena_ensure(c, need)) return nullptr;
    std::vector<std::vector<void *>> dest(peer_cols.size(), std::vector<void *>(W));
    for (size_t k = 0; k < peer_cols.size(); k++) {
        const int w = dtype_width(b.cols[peer_cols[k]].dtype);
        for (int d = 0; d < W; d++) {
            int64_t before = 0;
            for (int s2 = 0; s2 < c->rank; s2++) before += m[(size_t)s2 * per + d];
            dest[k][d] = (uint8_t *)c->peer_arena[d] + region(d, k) + (size_t)before * w;
        }
    }
    Buf rowid;
    partition_scatter_to_peers(pl, b, peer_cols, dest, other_cols.empty() ? nullptr : &rowid);
    // all writes of all ranks have landed once this collective completes
    Buf bar = dev_alloc_zero(4);
    B200_NCCL(nccl().AllReduce(bar->ptr, bar->ptr, 1, ncclInt32, ncclSum, c->comm, st));
    count_launch();
    const int64_t n_out = recv_total[c->rank];
    auto out = std::make_unique<Batch>();
    out->num_rows = n_out;
    out->cols.resize(b.cols.size());
    for (size_t k = 0; k < peer_cols.size(); k++) {
        const Column &col = b.cols[peer_cols[k]];
        const int w = dtype_width(col.dtype);
        Column o;
        o.name = col.name;
        o.dtype = col.dtype;
        o.length = n_out;
        o.data = dev_alloc((size_t)std::max<int64_t>(n_out, 1) * w);
        o.data_bytes = n_out * w;
        if (n_out)
            B200_CUDA(cudaMemcpyAsync(o.data->ptr, (const uint8_t *)c->arena + region(c->rank, k), (size_t)n_out * w, cudaMemcpyDeviceToDevice, st));
        out->cols[peer_cols[k]] = std::move(o);
    }
    for (size_t ci = 0; ci < b.cols.size(); ci++)
        if (b.cols[ci].dtype == B200_NULL) {
            Column o = b.cols[ci];
            o.length = n_out;
            o.null_count = n_out;
            out->cols[ci] = std::move(o);
        }
    if (!other_cols.empty()) {  // nullable / Boolean / string columns: destination-ordered gather, then the NCCL route
        Batch rest;
        rest.num_rows = b.num_rows;
        for (int ci : other_cols) rest.cols.push_back(gather_column(b.cols[ci], as<int64_t>(rowid), b.num_rows, false));
        Route rows;
        rows.send_off = soff;
        rows.send_cnt = scnt;
        rows.recv_off.assign(W, 0);
        rows.recv_cnt.assign(W, 0);
        for (int src = 0; src < W; src++) {
            rows.recv_cnt[src] = m[(size_t)src * per + c->rank];
            rows.recv_off[src] = rows.recv_total;
            rows.recv_total += rows.recv_cnt[src];
        }
        std::vector<int64_t> na;
        for (int ci : other_cols) na.push_back(nulls_anywhere[ci]);
        std::unique_ptr<Batch> routed(route_batch(c, rest, rows, &na));
        for (size_t k = 0; k < other_cols.size(); k++) out->cols[other_cols[k]] = std::move(routed->cols[k]);
    }
    B200_CUDA(cudaStreamSynchronize(st));
    return out.release();
}

// rows of `b` go to rank targets[i] % world (targets: u64 per row), arriving in (source rank, source row) order
static Batch *exchange_by_targets(b200_comm *c, const Batch &b, const PartSrc &h) {
    const int W = c->world;
    {
        KernelTimer ktp("exchange_peer_store");
        if (Batch *direct = exchange_peer_store(c, b, h)) return direct;
    }
    std::vector<int64_t> offsets;
    std::unique_ptr<Batch> staged;
    if (W <= 256 && b.num_rows >= 4096) {  // direct scatter: the payload moves once, straight into destination order
        staged.reset(partition_staged(b, h, W, offsets));
    } else {
        if (!h.hashes) fail("exchange: the row-id partition needs a hash column");  // key mode is only chosen for the direct scatter
        Buf perm = partition_permutation(h.hashes, b.num_rows, W, offsets);
        staged.reset(gather_batch(b, as<int64_t>(perm), b.num_rows, false));
    }
    std::vector<int64_t> soff(W), scnt(W);
    for (int p = 0; p < W; p++) {
        soff[p] = offsets[p];
        scnt[p] = offsets[p + 1] - offsets[p];
    }
    std::vector<int64_t> nulls, nulls_anywhere;
    for (const Column &col : staged->cols) nulls.push_back(col.null_count);
    const Route rows = make_route(c, soff, scnt, nulls, &nulls_anywhere);
    KernelTimer kt("exchange_route");
    return route_batch(c, *staged, rows, &nulls_anywhere);
}

// RepartitionSpec::Hash: the same hash_rows % world placement as partition_by_hash (a3)
Batch *exchange_by_hash(b200_comm *c, const Batch &b, const int32_t *key_cols, int32_t nkeys) {
    require_device();
    if (!c) fail("null communicator");
    return exchange_by_targets(c, b, part_src_for_keys(b, key_cols, nkeys, c->world <= 256 && b.num_rows >= 4096));
}
// RepartitionSpec::Random (src/daft-local-execution/src/sinks/repartition.rs:169-171): see partition.cu random_targets; the
// salt makes the draw a function of (seed, rank, row) so that ranks do not all draw the same sequence
Batch *exchange_random(b200_comm *c, const Batch &b, uint64_t seed) {
    require_device();
    if (!c) fail("null communicator");
    PartSrc src;
    src.keep = random_targets(b.num_rows, c->world, seed, (uint64_t)c->rank << 40);
    src.hashes = as<uint64_t>(src.keep);
    return exchange_by_targets(c, b, src);
}
// RepartitionSpec::Range (repartition.rs:172-176; the exchange leg of the distributed sort,
// src/daft-distributed/src/pipeline_node/sort.rs:249-300): target = search_sorted(boundaries, keys); world - 1 boundaries
Batch *exchange_by_range(b200_comm *c, const Batch &b, const int32_t *key_cols, int32_t nkeys, const Batch &boundaries, const uint8_t *descending) {
    require_device();
    if (!c) fail("null communicator");
    if (boundaries.num_rows != c->world - 1) fail("exchange_by_range: need exactly world - 1 boundary rows");
    if (boundaries.num_rows == 0) return new Batch(b);
    PartSrc src;
    src.keep = op_search_sorted(boundaries, b, key_cols, nkeys, descending);
    src.hashes = as<uint64_t>(src.keep);
    return exchange_by_targets(c, b, src);
}

// ---- small batches (partial aggregates, top-N candidates): ONE fixed-size all-gather ---------------------------------
// Every rank packs its batch into a slab of SLAB_BYTES: a host-built header (row count, per column null count and
// data bytes -- all known on the host without a device read) followed by the 16-byte aligned column buffers.  One
// ncclAllGather moves the slabs, one strided D2H copy brings the W headers back (the only synchronisation), and one
// kernel assembles the concatenated columns (byte copies, offset rebasing, bit-granular validity concatenation).
// A rank whose batch does not fit writes rows = -1 and ALL ranks fall through to the general route below.
constexpr int64_t SLAB_BYTES = 32 * 1024, SLAB_HDR = 2048;
struct SegDesc {
    int kind;  // 0 bytes, 1 string offsets (rebase), 2 bits (src == nullptr: all ones)
    int _pad;
    const uint8_t *src;
    uint8_t *dst;
    int64_t n;         // bytes / offsets entries - 1 / bits
    int64_t dst_base;  // kind 1: byte base added to the rebased offsets; kind 2: first destination bit; kind 1 dst row via dst pointer
};
__global__ void pack_segments(const SegDesc *__restrict__ d) {
    const SegDesc s = d[blockIdx.x];
    if (s.kind == 0) {
        for (int64_t i = threadIdx.x; i < s.n; i += blockDim.x) s.dst[i] = s.src[i];
    } else if (s.kind == 1) {
        const int64_t *so = reinterpret_cast<const int64_t *>(s.src);
        int64_t *dofs = reinterpret_cast<int64_t *>(s.dst);
        const int64_t first = so[0];
        for (int64_t i = threadIdx.x; i <= s.n; i += blockDim.x) dofs[i] = so[i] - first + s.dst_base;
    } else {
        unsigned int *words = reinterpret_cast<unsigned int *>(s.dst);
        for (int64_t i = threadIdx.x; i < s.n; i += blockDim.x) {
            const bool bit = s.src == nullptr || bit_at(s.src, i);
            const int64_t o = s.dst_base + i;
            if (bit) atomicOr(&words[o >> 5], 1u << (o & 31));
        }
    }
}
static void run_segments(const std::vector<SegDesc> &segs) {
    if (segs.empty()) return;
    cudaStream_t st = ctx().stream;
    Buf d = dev_alloc(segs.size() * sizeof(SegDesc));
    B200_CUDA(cudaMemcpyAsync(d->ptr, segs.data(), segs.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, st));
    pack_segments<<<(unsigned)segs.size(), 256, 0, st>>>(as<SegDesc>(d));
    count_launch();
    // no wait: the pageable-host copy above is staged before cudaMemcpyAsync returns, `d` is freed in stream order
}

// returns nullptr when some rank's batch does not fit a slab (collective decision)
static Batch *all_gather_small(b200_comm *c, const Batch &b) {
    cudaStream_t st = ctx().stream;
    const int W = c->world;
    const int nc = (int)b.cols.size();
    // header: [0] rows (-1 = too big), [1] ncols, then per column {null_count, data_bytes}
    std::vector<int64_t> hdr(SLAB_HDR / 8, 0);
    bool fits = 2 + 2 * (int64_t)nc <= SLAB_HDR / 8;
    int64_t off = SLAB_HDR;
    struct Part { int64_t valid_off = -1, offs_off = -1, data_off = -1; };
    std::vector<Part> parts(nc);
    std::vector<SegDesc> segs;
    Buf slab = dev_alloc((size_t)SLAB_BYTES), all = dev_alloc((size_t)SLAB_BYTES * W);
    auto place = [&](int64_t bytes) {
        const int64_t o = off;
        off += (bytes + 15) & ~(int64_t)15;
        return o;
    };
    for (int ci = 0; ci < nc && fits; ci++) {
        const Column &col = b.cols[ci];
        const int64_t n = col.length;
        hdr[2 + 2 * ci] = col.null_count;
        hdr[3 + 2 * ci] = col.data_bytes;
        if (col.dtype == B200_NULL) continue;
        if (col.null_count > 0) parts[ci].valid_off = place((n + 7) / 8);
        if (is_string(col.dtype)) parts[ci].offs_off = place((n + 1) * 8);
        const int w = dtype_width(col.dtype);
        const int64_t db = w ? n * w : (col.dtype == B200_BOOL ? (n + 7) / 8 : col.data_bytes);
        parts[ci].data_off = place(db);
        if (off > SLAB_BYTES) fits = false;
        if (!fits) break;
        uint8_t *base = as<uint8_t>(slab);
        if (col.null_count > 0) segs.push_back({0, 0, as<uint8_t>(col.validity), base + parts[ci].valid_off, (n + 7) / 8, 0});
        if (is_string(col.dtype) && col.offsets) segs.push_back({0, 0, as<uint8_t>(col.offsets), base + parts[ci].offs_off, (n + 1) * 8, 0});
        if (db > 0 && col.data) segs.push_back({0, 0, as<uint8_t>(col.data), base + parts[ci].data_off, db, 0});
    }
    hdr[0] = fits ? b.num_rows : -1;
    hdr[1] = nc;
    B200_CUDA(cudaMemcpyAsync(slab->ptr, hdr.data(), (size_t)SLAB_HDR, cudaMemcpyHostToDevice, st));
    if (fits && !segs.empty()) {
        Buf d = dev_alloc(segs.size() * sizeof(SegDesc));
        B200_CUDA(cudaMemcpyAsync(d->ptr, segs.data(), segs.size() * sizeof(SegDesc), cudaMemcpyHostToDevice, st));
        pack_segments<<<(unsigned)segs.size(), 256, 0, st>>>(as<SegDesc>(d));
        count_launch();
        B200_NCCL(nccl().AllGather(slab->ptr, all->ptr, (size_t)SLAB_BYTES, ncclUint8, c->comm, st));
        count_launch();
        // (d is released after the stream-ordered free; the synchronise below covers the host vector)
    } else {
        B200_NCCL(nccl().AllGather(slab->ptr, all->ptr, (size_t)SLAB_BYTES, ncclUint8, c->comm, st));
        count_launch();
    }
    std::vector<int64_t> hs((size_t)W * (SLAB_HDR / 8));
    B200_CUDA(cudaMemcpy2DAsync(hs.data(), (size_t)SLAB_HDR, all->ptr, (size_t)SLAB_BYTES, (size_t)SLAB_HDR, (size_t)W, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < W; r++) {
        const int64_t *h = &hs[(size_t)r * (SLAB_HDR / 8)];
        if (h[0] < 0) return nullptr;
        if (h[1] != nc) fail("all_gather: ranks disagree on the number of columns");
    }
    // assemble: every rank walks the same layout rule over each source slab
    auto out = std::make_unique<Batch>();
    int64_t total_rows = 0;
    for (int r = 0; r < W; r++) total_rows += hs[(size_t)r * (SLAB_HDR / 8)];
    out->num_rows = total_rows;
    std::vector<int64_t> cursor(W, SLAB_HDR);  // running offset inside each rank's slab
    std::vector<SegDesc> asm_segs;
    const uint8_t *allp = as<uint8_t>(all);
    for (int ci = 0; ci < nc; ci++) {
        const Column &col = b.cols[ci];
        Column o;
        o.name = col.name;
        o.dtype = col.dtype;
        o.length = total_rows;
        if (col.dtype == B200_NULL) {
            o.null_count = total_rows;
            out->cols.push_back(std::move(o));
            continue;
        }
        int64_t nulls = 0, dbytes = 0;
        for (int r = 0; r < W; r++) {
            nulls += hs[(size_t)r * (SLAB_HDR / 8) + 2 + 2 * ci];
            dbytes += hs[(size_t)r * (SLAB_HDR / 8) + 3 + 2 * ci];
        }
        o.null_count = nulls;
        const int w = dtype_width(col.dtype);
        const bool is_bool = col.dtype == B200_BOOL, is_str = is_string(col.dtype);
        if (nulls > 0) o.validity = dev_alloc_zero((size_t)((total_rows + 31) / 32 + 1) * 4);
        if (is_str) o.offsets = dev_alloc((size_t)(total_rows + 1) * 8);
        if (w) {
            o.data = dev_alloc((size_t)std::max<int64_t>(total_rows * w, 1));
            o.data_bytes = total_rows * w;
        } else if (is_bool) {
            o.data = dev_alloc_zero((size_t)((total_rows + 31) / 32 + 1) * 4);
            o.data_bytes = (total_rows + 7) / 8;
        } else {
            o.data = dev_alloc((size_t)std::max<int64_t>(dbytes, 1));
            o.data_bytes = dbytes;
        }
        int64_t row_base = 0, byte_base = 0;
        for (int r = 0; r < W; r++) {
            const int64_t *h = &hs[(size_t)r * (SLAB_HDR / 8)];
            const int64_t n = h[0], rn = h[2 + 2 * ci], rdb = h[3 + 2 * ci];
            const uint8_t *sl = allp + (size_t)r * SLAB_BYTES;
            auto take = [&](int64_t bytes) {
                const int64_t o2 = cursor[r];
                cursor[r] += (bytes + 15) & ~(int64_t)15;
                return sl + o2;
            };
            const uint8_t *vsrc = rn > 0 ? take((n + 7) / 8) : nullptr;
            const uint8_t *osrc = is_str ? take((n + 1) * 8) : nullptr;
            const int64_t db = w ? n * w : (is_bool ? (n + 7) / 8 : rdb);
            const uint8_t *dsrc = take(db);
            if (nulls > 0 && n > 0) asm_segs.push_back({2, 0, vsrc, as<uint8_t>(o.validity), n, row_base});
            if (is_str) asm_segs.push_back({1, 0, osrc, as<uint8_t>(o.offsets) + (size_t)row_base * 8, n, byte_base});
            if (w && n > 0) asm_segs.push_back({0, 0, dsrc, as<uint8_t>(o.data) + (size_t)row_base * w, n * w, 0});
            else if (is_bool && n > 0) asm_segs.push_back({2, 0, dsrc, as<uint8_t>(o.data), n, row_base});
            else if (is_str && rdb > 0) asm_segs.push_back({0, 0, dsrc, as<uint8_t>(o.data) + (size_t)byte_base, rdb, 0});
            row_base += n;
            byte_base += is_str ? rdb : 0;
        }
        if (nulls == 0) o.validity.reset();
        out->cols.push_back(std::move(o));
    }
    run_segments(asm_segs);
    return out.release();
}

// ---- low-cardinality aggregates: the group tables themselves are all-reduced ------------------------------------------
// Every rank ran the same aggregate kernel into a table of the same capacity; a group's slot is a function of its
// key code (open addressing from the same home slot), so slot s holds the same group on every rank unless two codes
// collided differently -- which the caller detects from the MIN / MAX of the tag words.  ONE collective: the tables
// (status word, tags, cells -- a few hundred KB) are all-gathered, and a merge kernel folds the W copies of every slot in
// rank order with the cell's own operator (f64 / 64-bit sums, minima, maxima), so every rank computes bit-identical
// results.  Cells listed in `keep_local` (the groups' first rows) stay local.  (A grouped ncclAllReduce per operator run
// was six small collectives; at 8 GPUs their latencies added up to more than the table transfer.)
struct MergeOps {
    int ncells;
    signed char op[48];  // jit cell operator, -1 = keep the local value
};
__global__ void merge_agg_tables(const unsigned long long *__restrict__ all, int W, size_t words, size_t cap, const __grid_constant__ MergeOps ops,
                                 unsigned long long *__restrict__ acc, unsigned long long *__restrict__ tagmin, int *__restrict__ status) {
    const size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) {
        int st = 0;
        for (int r = 0; r < W; r++) st = max(st, (int)(all[(size_t)r * words] & 0xffffffffULL));
        if (st) atomicMax(status, st);
    }
    if (s >= cap) return;
    unsigned long long tmax = 0ULL, tmin = ~0ULL;
    for (int r = 0; r < W; r++) {
        const unsigned long long t = all[(size_t)r * words + 2 + s];
        tmax = max(tmax, t);
        if (t) tmin = min(tmin, t);
    }
    acc[s] = tmax;
    tagmin[s] = tmin;
    for (int c = 0; c < ops.ncells; c++) {
        const int op = ops.op[c];
        if (op < 0) continue;
        const size_t at = 2 + (size_t)(1 + c) * cap + s;
        unsigned long long v = all[at];
        for (int r = 1; r < W; r++) {
            const unsigned long long x = all[(size_t)r * words + at];
            switch (op) {
                case 0: v = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)v) + __longlong_as_double((long long)x)); break;  // C_ADD_F64
                case 2: v += x; break;            // C_ADD_I64 (wraps like the kernel)
                case 3: v = min(v, x); break;     // C_MIN
                default: v = max(v, x); break;    // C_MAX
            }
        }
        acc[(size_t)(1 + c) * cap + s] = v;
    }
}
__global__ void tag_min_prep(const unsigned long long *__restrict__ tags, size_t cap, unsigned long long *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) {
        const unsigned long long t = tags[i];
        out[i] = t == 0ULL ? ~0ULL : t;
    }
}
// the round-2 first version: one grouped launch of an all-reduce per operator run (B200_AGG_MERGE=allreduce)
static void allreduce_agg_table_grouped(b200_comm *c, unsigned long long *acc, size_t cap, const std::vector<int> &cell_ops, const std::vector<char> &keep_local,
                         unsigned long long *tagmin, int *status_flag) {
    require_device();
    if (!c) fail("null communicator");
    cudaStream_t st = ctx().stream;
    B200_NCCL(nccl().GroupStart());
    B200_NCCL(nccl().AllReduce(acc, acc, cap, ncclUint64, ncclMax, c->comm, st));
    B200_NCCL(nccl().AllReduce(tagmin, tagmin, cap, ncclUint64, ncclMin, c->comm, st));
    B200_NCCL(nccl().AllReduce(status_flag, status_flag, 1, ncclInt32, ncclMax, c->comm, st));
    size_t i = 0;
    while (i < cell_ops.size()) {
        if (keep_local[i]) {
            i++;
            continue;
        }
        size_t j = i + 1;
        while (j < cell_ops.size() && !keep_local[j] && cell_ops[j] == cell_ops[i]) j++;
        unsigned long long *base = acc + (1 + i) * cap;
        const size_t count = (j - i) * cap;
        switch (cell_ops[i]) {
            case 0: B200_NCCL(nccl().AllReduce(base, base, count, ncclDouble, ncclSum, c->comm, st)); break;   // C_ADD_F64
            case 2: B200_NCCL(nccl().AllReduce(base, base, count, ncclUint64, ncclSum, c->comm, st)); break;   // C_ADD_I64 (wraps like the kernel)
            case 3: B200_NCCL(nccl().AllReduce(base, base, count, ncclUint64, ncclMin, c->comm, st)); break;   // C_MIN
            case 4: B200_NCCL(nccl().AllReduce(base, base, count, ncclUint64, ncclMax, c->comm, st)); break;   // C_MAX
            default: fail("agg all-reduce: unsupported cell operator");
        }
        i = j;
    }
    B200_NCCL(nccl().GroupEnd());
    count_launch();
}

void allreduce_agg_table(b200_comm *c, unsigned long long *acc, size_t cap, const std::vector<int> &cell_ops, const std::vector<char> &keep_local,
                         unsigned long long *tagmin, int *status_flag) {
    require_device();
    if (!c) fail("null communicator");
    cudaStream_t st = ctx().stream;
    KernelTimer kt("agg_table_allreduce");
    {
        const char *env = getenv("B200_AGG_MERGE");
        if (env && !strcmp(env, "allreduce")) {
            tag_min_prep<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(acc, cap, tagmin);
            count_launch();
            allreduce_agg_table_grouped(c, acc, cap, cell_ops, keep_local, tagmin, status_flag);
            return;
        }
    }
    MergeOps ops;
    if (cell_ops.size() > sizeof(ops.op)) fail("agg all-reduce: too many cells");
    ops.ncells = (int)cell_ops.size();
    for (size_t i = 0; i < cell_ops.size(); i++) {
        if (!keep_local[i] && cell_ops[i] != 0 && cell_ops[i] != 2 && cell_ops[i] != 3 && cell_ops[i] != 4) fail("agg all-reduce: unsupported cell operator");
        ops.op[i] = keep_local[i] ? (signed char)-1 : (signed char)cell_ops[i];
    }
    const int W = c->world;
    const size_t words = 2 + (1 + cell_ops.size()) * cap;  // [status, pad, tags[cap], cells[ncells][cap]]
    Buf mine = dev_alloc(words * 8), all = dev_alloc(words * 8 * (size_t)W);
    fill_u8(mine->ptr, 0, 16);
    B200_CUDA(cudaMemcpyAsync(mine->ptr, status_flag, 4, cudaMemcpyDeviceToDevice, st));
    B200_CUDA(cudaMemcpyAsync(as<unsigned long long>(mine) + 2, acc, (words - 2) * 8, cudaMemcpyDeviceToDevice, st));
    B200_NCCL(nccl().AllGather(mine->ptr, all->ptr, words, ncclUint64, c->comm, st));
    count_launch();
    merge_agg_tables<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(as<unsigned long long>(all), W, words, cap, ops, acc, tagmin, status_flag);
    count_launch();
}

Batch *all_gather_batches(b200_comm *c, const Batch &b) {
    require_device();
    if (!c) fail("null communicator");
    if (Batch *small = all_gather_small(c, b)) return small;
    const int W = c->world;
    std::vector<int64_t> soff(W, 0), scnt(W, b.num_rows);
    const Route rows = make_route(c, soff, scnt);
    return route_batch(c, b, rows, nullptr);
}

}  // namespace b200

"""Multi-GPU host logic: one process per GPU (torch.distributed for rendezvous/plumbing, NCCL for data).

Replaces the Flotilla shapes the hot path needs (src/daft-distributed/src/pipeline_node/aggregate.rs:281-340,
shuffles/repartition.rs:26-167): scan partitions = contiguous row ranges per rank; low-cardinality
aggregates exchange only their tiny partial results; high-cardinality group-bys and joins hash-repartition
their inputs with the NCCL all-to-all in libdaftb200 (b200_exchange_by_hash), whose placement is the
same hash_rows % world as the reference's partition_by_hash.
"""
from __future__ import annotations

import ctypes
import io
from typing import Callable, Sequence

import pyarrow as pa


def shard_rows(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous scan partition of rank `rank`: sizes differ by at most one row."""
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _to_ipc(tbl: pa.Table) -> bytes:
    sink = io.BytesIO()
    with pa.ipc.new_stream(sink, tbl.schema) as w:
        w.write_table(tbl)
    return sink.getvalue()


def _from_ipc(b: bytes) -> pa.Table:
    return pa.ipc.open_stream(io.BytesIO(b)).read_all()


_SMALL_CAP = 1 << 16  # bytes: partial aggregates are a few hundred bytes


def gather_tables(tbl: pa.Table, group=None) -> pa.Table:
    """Every rank receives the concatenation (rank order) of all ranks' small host tables
    (partial aggregates, top-N candidates).  Works on any torch.distributed backend (nccl or gloo).
    Small payloads travel in ONE fixed-size all-gather (length-prefixed Arrow IPC), large ones fall back
    to all_gather_object."""
    import numpy as np
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tbl
    world = dist.get_world_size(group)
    payload = _to_ipc(tbl)
    small = len(payload) + 8 <= _SMALL_CAP
    use_cuda = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    buf = np.zeros(_SMALL_CAP, dtype=np.uint8)
    if small:
        buf[:8] = np.frombuffer(np.int64(len(payload)).tobytes(), dtype=np.uint8)
        buf[8:8 + len(payload)] = np.frombuffer(payload, dtype=np.uint8)
    else:
        buf[:8] = np.frombuffer(np.int64(-1).tobytes(), dtype=np.uint8)
    mine = torch.from_numpy(buf).to(dev)
    out = torch.empty(world * _SMALL_CAP, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    host = out.cpu().numpy().reshape(world, _SMALL_CAP)
    lens = [int(np.frombuffer(host[r, :8].tobytes(), dtype=np.int64)[0]) for r in range(world)]
    if all(l >= 0 for l in lens):
        return pa.concat_tables([_from_ipc(host[r, 8:8 + lens[r]].tobytes()) for r in range(world)])
    parts = [None] * world  # some rank has a large table: every rank takes the slow path together
    dist.all_gather_object(parts, payload, group=group)
    return pa.concat_tables([_from_ipc(p) for p in parts])


def two_stage_aggregate(partial: Callable[[], pa.Table], final: Callable[[pa.Table], pa.Table], group=None) -> pa.Table:
    """partial aggregate per rank -> all-gather of the partials -> identical final merge on every rank
    (populate_aggregation_stages, src/daft-local-plan/src/agg.rs:13-25)."""
    return final(gather_tables(partial(), group))


def partitioned_join(left, right, left_on: Sequence[str], right_on: Sequence[str], world: int, rank: int, partition, exchange, join):
    """The shape of a distributed hash join (src/daft-distributed/src/pipeline_node/join/hash_join.rs, translated from
    src/daft-logical-plan's HashJoin with both sides repartitioned by hash of the join keys): rows whose keys are equal hash alike, so
    after both sides are exchanged by hash(keys) % world the local joins partition the global join.  `partition(table, keys, world)`
    -> world pieces, `exchange(pieces)` -> this rank's rows from every rank, `join(l, r)` the local join.  Written over callables so
    the world-2 gloo test drives it with the oracle (tests/test_distributed_gloo.py) and Communicator.hash_join with the device
    primitives.  A null key hashes like any value but matches nothing: where it lands does not matter for inner / semi joins; for
    left / right / outer / anti joins every row is still emitted exactly once, on the rank its key hashes to."""
    lp, rp = partition(left, left_on, world), partition(right, right_on, world)
    return join(exchange(lp), exchange(rp))


class Communicator:
    """NCCL communicator owned by libdaftb200 (b200_comm): rank 0 creates the unique id, torch.distributed
    broadcasts it (out-of-band rendezvous only), every rank calls ncclCommInitRank."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        from . import _lib
        _lib.ensure_init()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        ident = [None]
        if self.rank == 0:
            buf = (ctypes.c_uint8 * 128)()
            _lib.call("b200_comm_unique_id", buf)
            ident[0] = bytes(buf)
        dist.broadcast_object_list(ident, src=0, group=group)
        self._h = ctypes.c_void_p()
        arr = (ctypes.c_uint8 * 128).from_buffer_copy(ident[0])
        _lib.call("b200_comm_init", arr, ctypes.c_int32(self.rank), ctypes.c_int32(self.world), ctypes.byref(self._h))

    def __del__(self):
        try:
            from . import _lib
            if self._h and self._h.value:
                _lib.lib().b200_comm_free(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    def exchange_by_hash(self, batch, exprs: Sequence):
        """Hash-repartition `batch` across the ranks: this rank receives every row whose
        hash_rows(exprs) % world == rank, in (source rank, source row) order."""
        from . import _lib
        from .recordbatch import RecordBatch, _i32
        both, idx = batch._resolve_keys(exprs)
        out = ctypes.c_void_p()
        _lib.call("b200_exchange_by_hash", self._h, both._h, _i32(idx), ctypes.c_int32(len(idx)), ctypes.byref(out))
        return RecordBatch._wrap(out).select_columns(list(range(batch.num_columns())))

    def exchange_random(self, batch, seed: int = 0):
        """RepartitionSpec::Random over the ranks (src/daft-local-execution/src/sinks/repartition.rs:169-171)."""
        from . import _lib
        from .recordbatch import RecordBatch
        out = ctypes.c_void_p()
        _lib.call("b200_exchange_random", self._h, batch._h, ctypes.c_uint64(seed & (2**64 - 1)), ctypes.byref(out))
        return RecordBatch._wrap(out, like=batch)

    def exchange_by_range(self, batch, exprs: Sequence, boundaries, descending: Sequence[bool]):
        """RepartitionSpec::Range: rank r receives the rows whose search_sorted position among the world - 1 boundaries is r."""
        from . import _lib
        from .recordbatch import RecordBatch, _i32, _u8
        both, idx = batch._resolve_keys(exprs)
        out = ctypes.c_void_p()
        _lib.call("b200_exchange_by_range", self._h, both._h, _i32(idx), ctypes.c_int32(len(idx)), boundaries._h, _u8(list(descending)),
                  ctypes.byref(out))
        return RecordBatch._wrap(out).select_columns(list(range(batch.num_columns())))

    def sort(self, batch, sort_keys: Sequence, descending=None, nulls_first=None, sample_size: int = 20):
        """Distributed sort (src/daft-distributed/src/pipeline_node/sort.rs): every rank samples `sample_size` rows of the sort
        keys (sample_size_for_sort = 20), the samples are gathered and sorted, RecordBatch::quantiles picks world - 1
        boundaries, the rows are range-exchanged and each rank sorts its range: rank order == global order."""
        from .recordbatch import RecordBatch
        desc, nf = RecordBatch._sort_flags(sort_keys, descending, nulls_first)
        keys = batch.eval_expression_list([e.alias(f"__sk{i}") for i, e in enumerate(sort_keys)])
        n = len(keys)
        step = max(n // max(sample_size, 1), 1)
        sample = keys.take(pa.array(list(range(0, n, step))[:sample_size], type=pa.int64()))
        from .expressions import col
        kcols = [col(f"__sk{i}") for i in range(len(sort_keys))]
        gathered = self.all_gather(sample).sort(kcols, desc, nf)
        boundaries = gathered.quantiles(self.world)
        if len(boundaries) != self.world - 1:   # no sampled rows at all (every rank is empty or nearly so)
            return self._sort_fallback(batch, sort_keys, desc, nf)
        both = batch.union(keys)
        moved = self.exchange_by_range(both, kcols, boundaries, desc)
        return moved.sort(kcols, desc, nf).select_columns(list(range(batch.num_columns())))

    def _sort_fallback(self, batch, sort_keys, desc, nf):
        """Tiny inputs: gather everything and keep this rank's contiguous share of the sorted rows."""
        allb = self.all_gather(batch).sort(sort_keys, desc, nf)
        b, e = shard_rows(len(allb), self.rank, self.world)
        return allb.slice(b, e)

    def agg(self, batch, to_agg: Sequence, group_by: Sequence | None = None, predicate: Sequence | None = None):
        """HashAggregate over one scan partition per rank; every rank returns the global result.
        Low-cardinality results take ONE stage: the per-GPU group tables are all-reduced over NVLink
        (b200_agg_allreduce).  Anything else (decided collectively inside the library) runs the reference's two-stage
        plan: first-stage aggregate per rank -> all-gather of the partials -> second stage + final projection
        (populate_aggregation_stages; src/daft-distributed/src/pipeline_node/aggregate.rs:281-340)."""
        from . import _lib
        from .aggstages import populate_aggregation_stages
        from .expressions import _agg_inner, col, encode_agg
        from .recordbatch import RecordBatch
        group_by = list(group_by or [])
        staged_only = any(_agg_inner(a).op in ("stddev", "var", "skew", "count_distinct") for a in to_agg)
        if not staged_only:
            plan = batch._cached_plan("agg", [to_agg, group_by, predicate],
                                      lambda: encode_agg(to_agg, group_by, batch.column_names(), batch.dtypes(), predicate))
            out, merged = ctypes.c_void_p(), ctypes.c_int32(0)
            _lib.call("b200_agg_allreduce", self._h, batch._h, plan, ctypes.c_size_t(len(plan)), ctypes.byref(out), ctypes.byref(merged))
            if merged.value:
                onames = [g.name() for g in group_by] + [a.name() for a in to_agg]
                odt = (ctypes.c_int32 * len(onames))()
                _lib.call("b200_batch_dtypes", out, odt, ctypes.c_int32(len(onames)))
                return RecordBatch._wrap(out, names=onames, dtypes=list(odt))
        first, second, final = populate_aggregation_stages(to_agg, group_by)
        partial = batch.agg(first, group_by, predicate)
        keys = [col(g.name()) for g in group_by]
        return self.all_gather(partial).agg(second, keys).eval_expression_list(final)

    def hash_join(self, left, right, left_on: Sequence, right_on: Sequence, how: str = "inner", broadcast_rows: int = 1 << 16):
        """Distributed hash join over one scan partition of each side per rank; the union of the ranks' outputs is the global join.
        A right side of at most `broadcast_rows` rows per rank is broadcast (all-gather, src/daft-distributed/src/pipeline_node/join/
        broadcast_join.rs) when the join type allows it (inner / left / semi / anti keep every left row where it is); otherwise both
        sides are exchanged by the hash of their join keys (hash_join.rs) and joined locally.  Join keys are cast to their common
        supertype BEFORE hashing (ops/joins/mod.rs:22-46): an Int32 and an Int64 of the same value hash differently."""
        import torch
        import torch.distributed as dist
        from . import datatypes as dt
        from .expressions import bind
        lnames, ldt, rnames, rdt = left.column_names(), left.dtypes(), right.column_names(), right.dtypes()
        kd = [dt.supertype(bind(le, lnames, ldt).dtype, bind(re_, rnames, rdt).dtype) for le, re_ in zip(left_on, right_on)]
        small = torch.tensor([1 if len(right) <= broadcast_rows else 0], device="cuda")
        dist.all_reduce(small, op=dist.ReduceOp.MIN)     # every rank takes the same route
        if int(small.item()) and how in ("inner", "left", "semi", "anti"):
            return left.hash_join(self.all_gather(right), left_on, right_on, how=how)
        lk = [(e if bind(e, lnames, ldt).dtype == t else e.cast(t)) for e, t in zip(left_on, kd)]
        rk = [(e if bind(e, rnames, rdt).dtype == t else e.cast(t)) for e, t in zip(right_on, kd)]
        return self.exchange_by_hash(left, lk).hash_join(self.exchange_by_hash(right, rk), left_on, right_on, how=how)

    def all_gather(self, batch):
        from . import _lib
        from .recordbatch import RecordBatch
        out = ctypes.c_void_p()
        _lib.call("b200_all_gather", self._h, batch._h, ctypes.byref(out))
        return RecordBatch._wrap(out, like=batch)

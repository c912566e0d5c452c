"""Running the SAME query compositions (daft_b200/tpch.py, clickbench.py) over scan partitions spread across the ranks of a communicator
(BASELINE config 4: "all 22 queries, scan partitions sharded across 8 x B200 with NVLink all-to-all shuffle").

`Sharded` has the RecordBatch verb surface the queries use and decides, verb by verb, what the reference's distributed planner decides
(src/daft-distributed/src/pipeline_node/*): filters and projections stay local; an aggregate runs its first stage per rank and then either
gathers the partials (few groups: every rank finishes the same answer, `aggregate.rs:281-340`) or repartitions them by the group keys
(many groups); a join broadcasts a small side (`join/broadcast_join.rs`) or repartitions both sides by the hash of the join keys
(`join/hash_join.rs`); top-N / sort / dedup combine local results; a window repartitions by its
PARTITION BY keys.  A table is either *sharded* (every rank holds a disjoint part) or
*replicated* (every rank holds all of it, e.g. a gathered aggregate): the state decides which of those is needed.

The communication itself is a small backend object, so the same planner drives
  * the device: `DeviceBackend(Communicator)` -- `b200_exchange_by_hash` / `b200_all_gather` over NVLink (daft_b200/distributed.py), and
  * the CPU check: world-2 gloo processes with the oracle standing in for the kernels (tests/test_distributed_gloo.py), which is how all
    22 TPC-H queries were verified against their single-process answers; the device backend has not been run on N GPUs this round.
Every rank must call the same verbs in the same order (they contain collectives), which the query functions do by construction."""
from __future__ import annotations

from typing import Sequence

from .aggstages import populate_aggregation_stages
from .expressions import Expression, _agg_inner, col

COMPOSITE = ("count_distinct", "percentile", "stddev", "var", "skew", "list", "set", "concat")


class DeviceBackend:
    """The collectives of daft_b200.distributed.Communicator plus a scalar all-reduce (torch.distributed, rendezvous only)."""

    def __init__(self, comm):
        self.comm, self.rank, self.world = comm, comm.rank, comm.world

    def exchange_by_hash(self, batch, key_exprs):
        return self.comm.exchange_by_hash(batch, key_exprs)

    def all_gather(self, batch):
        return self.comm.all_gather(batch)

    def allreduce_sum(self, value: int) -> int:
        import torch
        import torch.distributed as dist
        t = torch.tensor([int(value)], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        return int(t.item())


class Sharded:
    def __init__(self, local, backend, replicated: bool = False, gather_rows: int = 1 << 20):
        self.local, self.backend, self.replicated, self.gather_rows = local, backend, replicated, gather_rows

    def _like(self, local, replicated=None) -> "Sharded":
        return Sharded(local, self.backend, self.replicated if replicated is None else replicated, self.gather_rows)

    # ---- schema / size ----
    def column_names(self): return self.local.column_names()
    def num_columns(self): return self.local.num_columns()
    def dtypes(self): return self.local.dtypes()

    def __len__(self) -> int:
        return len(self.local) if self.replicated else self.backend.allreduce_sum(len(self.local))

    # ---- local verbs ----
    def filter(self, exprs): return self._like(self.local.filter(exprs))
    def eval_expression_list(self, exprs): return self._like(self.local.eval_expression_list(exprs))

    # ---- state changes ----
    def gathered(self) -> "Sharded":
        return self if self.replicated else self._like(self.backend.all_gather(self.local), True)

    def resharded(self) -> "Sharded":
        """a replicated table cut into one contiguous slice per rank (so that a repartitioning verb sees every row exactly once)"""
        if not self.replicated:
            return self
        from .distributed import shard_rows
        b, e = shard_rows(len(self.local), self.backend.rank, self.backend.world)
        return self._like(self.local.slice(b, e), False)

    def to_arrow_table(self):
        return self.gathered().local.to_arrow_table()

    # ---- aggregation ----
    def agg(self, to_agg: Sequence[Expression], group_by: Sequence[Expression] | None = None, predicate=None) -> "Sharded":
        group_by = list(group_by or [])
        if self.replicated:
            return self._like(self.local.agg(to_agg, group_by, predicate))
        if any(_agg_inner(a).op in COMPOSITE for a in to_agg) or any(self._string_minmax(a) for a in to_agg):
            # aggregates that are not a fold of partials: bring every group onto one rank first (or, ungrouped, every row onto all)
            base = self.filter(predicate) if predicate else self
            if not group_by:
                return base.gathered().agg(to_agg, [])
            moved = base.backend.exchange_by_hash(base.local, group_by)
            return self._like(moved.agg(to_agg, group_by), False)
        first, second, final = populate_aggregation_stages(to_agg, group_by)
        partial = self.local.agg(first, group_by, predicate)
        keys = [col(g.name()) for g in group_by]
        if not group_by or self.backend.allreduce_sum(len(partial)) <= self.gather_rows:
            merged = self.backend.all_gather(partial).agg(second, keys).eval_expression_list(final)
            return self._like(merged, True)
        moved = self.backend.exchange_by_hash(partial, keys)          # many groups: every group's partials meet on one rank
        return self._like(moved.agg(second, keys).eval_expression_list(final), False)

    def _string_minmax(self, a: Expression) -> bool:
        from . import datatypes as dt
        from .expressions import bind
        inner = _agg_inner(a)
        return inner.op in ("min", "max") and dt.is_string(bind(inner.args[0], self.column_names(), self.dtypes()).dtype)

    def dedup(self, exprs) -> "Sharded":
        if self.replicated:
            return self._like(self.local.dedup(exprs))
        return self._like(self.backend.exchange_by_hash(self.local.dedup(exprs), exprs).dedup(exprs), False)

    # ---- joins ----
    def hash_join(self, right: "Sharded", left_on, right_on, how: str = "inner", null_equals_null=None, build_on_left=None) -> "Sharded":
        left = self
        kw = dict(how=how, null_equals_null=null_equals_null, build_on_left=build_on_left)
        if left.replicated and right.replicated:
            return left._like(left.local.hash_join(right.local, left_on, right_on, **kw), True)
        # a small sharded side is broadcast where the join type lets every row of the OTHER side stay where it is
        if not right.replicated and how in ("inner", "left", "semi", "anti") and not left.replicated and len(right) <= self.gather_rows:
            right = right.gathered()
        elif not left.replicated and how in ("inner", "right") and not right.replicated and len(left) <= self.gather_rows:
            left = left.gathered()
        if right.replicated and not left.replicated and how in ("inner", "left", "semi", "anti"):
            return left._like(left.local.hash_join(right.local, left_on, right_on, **kw), False)
        if left.replicated and not right.replicated and how in ("inner", "right"):
            return left._like(left.local.hash_join(right.local, left_on, right_on, **kw), False)
        # repartition both sides by the hash of their join keys (cast to the common key type first: equal values must hash alike)
        left, right = left.resharded(), right.resharded()
        lk, rk = _supertype_keys(left.local, right.local, left_on, right_on)
        lm = left.backend.exchange_by_hash(left.local, lk)
        rm = left.backend.exchange_by_hash(right.local, rk)
        return left._like(lm.hash_join(rm, left_on, right_on, **kw), False)

    # ---- windows ----
    def window(self, partition_by, order_by=(), descending=None, nulls_first=None, functions=()) -> "Sharded":
        """every partition of the window on one rank (repartition by the PARTITION BY keys, as the reference's distributed window node does:
        src/daft-distributed/src/pipeline_node/window.rs), then the local operator; without partition keys the rows are gathered"""
        from .window import window
        args = (list(partition_by), list(order_by), descending, nulls_first, list(functions))
        if self.replicated:
            return self._like(window(self.local, *args))
        if not args[0]:
            g = self.gathered()
            return g._like(window(g.local, *args), True)
        return self._like(window(self.backend.exchange_by_hash(self.local, args[0]), *args), False)

    # ---- order ----
    def sort(self, sort_keys, descending=None, nulls_first=None) -> "Sharded":
        g = self.gathered()        # query results: small by the time they are sorted
        return g._like(g.local.sort(sort_keys, descending, nulls_first), True)

    def top_n(self, sort_keys, limit: int, descending=None, nulls_first=None, offset: int = 0) -> "Sharded":
        if self.replicated:
            return self._like(self.local.top_n(sort_keys, limit, descending, nulls_first, offset))
        cand = self.backend.all_gather(self.local.top_n(sort_keys, limit + offset, descending, nulls_first))
        return self._like(cand.top_n(sort_keys, limit, descending, nulls_first, offset), True)

    def head(self, num: int) -> "Sharded":
        g = self.gathered()
        return g._like(g.local.head(num), True)


def _supertype_keys(left, right, left_on, right_on):
    from . import datatypes as dt
    from .expressions import bind
    ln, ld, rn, rd = left.column_names(), left.dtypes(), right.column_names(), right.dtypes()
    lk, rk = [], []
    for le, re_ in zip(left_on, right_on):
        a, b = bind(le, ln, ld).dtype, bind(re_, rn, rd).dtype
        t = dt.supertype(a, b)
        lk.append(le if a == t else le.cast(t))
        rk.append(re_ if b == t else re_.cast(t))
    return lk, rk

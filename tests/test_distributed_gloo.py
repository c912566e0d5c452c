"""N>1 host logic on CPU: world_size-2 gloo processes run the sharded two-stage Q1 (row-range scan
partitions -> per-rank partial aggregate -> all-gather of partials -> final merge) with the oracle standing in
for the GPU partial aggregate, and must reproduce the single-process answer."""
import os
import sys

import pyarrow as pa
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from daft_b200 import distributed, tpch
    from helpers import oracle_agg
    from oracle import expr_eval as E
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = tpch.num_rows("lineitem", 0.01)
        b, e = distributed.shard_rows(n, rank, world)
        shard = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS, b, e)
        # the plan shapes Communicator.agg falls back to: the generic rewrite of the single-stage Q1 aggregation list
        from daft_b200 import col
        from daft_b200.aggstages import populate_aggregation_stages
        pred, gb, aggs = tpch.q1_exprs()
        first, second, fproj = populate_aggregation_stages(aggs, gb)
        fgb = [col(g.name()) for g in gb]

        def final(parts):
            merged = oracle_agg(parts, second, fgb)
            return E.eval_expression_list(merged, fproj).sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")])

        out = distributed.two_stage_aggregate(lambda: oracle_agg(shard, first, gb, pred), final)
        q.put((rank, distributed._to_ipc(out), e - b))
    finally:
        dist.destroy_process_group()


def test_shard_rows_cover_everything():
    from daft_b200.distributed import shard_rows
    for n in (0, 1, 7, 100, 59986052):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1


@pytest.mark.timeout(180)
def test_two_stage_q1_world2_gloo():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from daft_b200 import tpch
    from helpers import assert_tables_equal, oracle_agg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    full = tpch.generate_host("lineitem", 0.01, tpch.Q1_COLUMNS)
    assert sum(r[2] for r in results) == full.num_rows
    pred, gb, aggs = tpch.q1_exprs()
    want = oracle_agg(full, aggs, gb, pred).sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")])
    for _, got, _ in results:
        from daft_b200.distributed import _from_ipc
        assert_tables_equal(_from_ipc(got), want, float_rtol=1e-12)


# ---- the distributed hash join's shape (distributed.partitioned_join): both sides exchanged by hash(keys) % world, joined locally ----
def _join_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from daft_b200 import distributed, tpch
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        for how in ("inner", "left", "anti", "outer"):
            lo, ro = ["O_ORDERKEY"], ["L_ORDERKEY"]
            left = tpch.generate_host("orders", 0.005, ["O_ORDERKEY", "O_CUSTKEY"], *distributed.shard_rows(tpch.num_rows("orders", 0.005), rank, world))
            right = tpch.generate_host("lineitem", 0.002, ["L_ORDERKEY", "L_QUANTITY"], *distributed.shard_rows(tpch.num_rows("lineitem", 0.002), rank, world))

            def exchange(pieces):
                got = [None] * world
                for dst in range(world):          # an all-to-all out of `world` gathers: piece `dst` of every rank goes to rank dst
                    parts = [None] * world
                    dist.all_gather_object(parts, distributed._to_ipc(pieces[dst]))
                    if dst == rank:
                        got = parts
                return pa.concat_tables([distributed._from_ipc(p) for p in got])
            res = distributed.partitioned_join(left, right, lo, ro, world, rank, O.partition_by_hash, exchange,
                                               lambda l, r: O.hash_join(l, r, lo, ro, how))
            out[how] = distributed._to_ipc(res)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_partitioned_join_world2_gloo():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from daft_b200 import tpch
    from daft_b200.distributed import _from_ipc
    from helpers import assert_tables_equal
    from oracle import oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 17) % 1000)
    procs = [ctx.Process(target=_join_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=150) for _ in procs)
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    left = tpch.generate_host("orders", 0.005, ["O_ORDERKEY", "O_CUSTKEY"])
    right = tpch.generate_host("lineitem", 0.002, ["L_ORDERKEY", "L_QUANTITY"])
    for how in ("inner", "left", "anti", "outer"):
        got = pa.concat_tables([_from_ipc(results[r][how]) for r in range(2)])
        want = O.hash_join(left, right, ["O_ORDERKEY"], ["L_ORDERKEY"], how)
        keys = [(c, "ascending") for c in want.column_names]
        assert_tables_equal(got.sort_by(keys), want.sort_by(keys))

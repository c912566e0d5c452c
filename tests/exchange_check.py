"""Run under torchrun (one process per GPU): NCCL hash-repartition exchange + all-gather vs the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # this file lives in tests/: it checks against the oracle
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import pyarrow as pa
import torch
import torch.distributed as dist

import daft_b200
from daft_b200 import col, distributed
from helpers import assert_tables_equal, random_table
from oracle import oracle as O


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    daft_b200.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = distributed.Communicator()
    rng = np.random.default_rng(100 + rank)
    n = 20000 + 1000 * rank
    tbl = random_table(rng, n, 0.1)
    rb = daft_b200.RecordBatch.from_arrow_table(tbl)
    for keys in (["i64"], ["s", "u8"]):
        got = comm.exchange_by_hash(rb, [col(k) for k in keys]).to_arrow_table()
        parts = [None] * world
        dist.all_gather_object(parts, distributed._to_ipc(tbl))
        want = pa.concat_tables([O.partition_by_hash(distributed._from_ipc(p), keys, world)[rank] for p in parts])
        assert_tables_equal(got, want)
    allg = comm.all_gather(rb).to_arrow_table()
    parts = [None] * world
    dist.all_gather_object(parts, distributed._to_ipc(tbl))
    assert_tables_equal(allg, pa.concat_tables([distributed._from_ipc(p) for p in parts]))
    # small batches take the packed single-collective path (exchange.cu all_gather_small): nulls, Booleans, strings,
    # a rank with zero rows, and a null-typed column
    for nsmall in (7 + 3 * rank, 0 if rank == world - 1 else 33, 300):
        small = random_table(np.random.default_rng(7 * rank + nsmall), nsmall, 0.3)
        small = small.append_column("allnull", pa.nulls(nsmall))
        got = comm.all_gather(daft_b200.RecordBatch.from_arrow_table(small)).to_arrow_table()
        sparts = [None] * world
        dist.all_gather_object(sparts, distributed._to_ipc(small))
        assert_tables_equal(got, pa.concat_tables([distributed._from_ipc(p) for p in sparts]))
    # high-cardinality group-by through the exchange == single-table oracle
    full = pa.concat_tables([distributed._from_ipc(p) for p in parts])
    mine = comm.exchange_by_hash(rb, [col("i32")]).agg([col("f64").sum().alias("s"), col("i32").count("all").alias("c")], [col("i32")]).to_arrow_table()
    gathered = distributed.gather_tables(mine).sort_by([("i32", "ascending")])
    from helpers import oracle_agg
    want = oracle_agg(full, [col("f64").sum().alias("s"), col("i32").count("all").alias("c")], [col("i32")]).sort_by([("i32", "ascending")])
    assert_tables_equal(gathered, want, float_rtol=1e-9)
    # distributed low-cardinality aggregate: the group tables all-reduced (b200_agg_allreduce) == single-table oracle;
    # plans the table merge cannot carry (any_value, float min / max folds, high cardinality) take the two-stage rewrite
    from daft_b200 import tpch
    aggs = [col("i64").sum().alias("sm"), col("f64").mean().alias("m"), col("i32").count("null").alias("cn"), col("u16").min().alias("lo"),
            col("i64").max().alias("hi"), col("i8").mean().alias("m8")]
    for gb, pred in (([col("k")], None), ([col("b"), col("u8")], [col("i32") > 0]), ([], None), ([col("s")], None), ([col("i32")], None)):
        got = comm.agg(rb, aggs, gb, pred).to_arrow_table()
        want = oracle_agg(full, aggs, gb, pred)
        keys = [g.name() for g in gb]
        assert_tables_equal(got, want, float_rtol=1e-9, sort_by=keys or None)
    fl = [col("f64").min().alias("mn"), col("f32").max().alias("mx"), col("s").any_value().alias("av")]
    got = comm.agg(rb, fl[:2], [col("k")]).to_arrow_table()
    assert_tables_equal(got, oracle_agg(full, fl[:2], [col("k")]), float_rtol=1e-9, sort_by=["k"])
    li = tpch.generate("lineitem", 0.1, tpch.Q1_COLUMNS, *distributed.shard_rows(tpch.num_rows("lineitem", 0.1), rank, world))
    q1 = tpch.q1_distributed(li, comm).to_arrow_table()
    pred, gb, q1aggs = tpch.q1_exprs()
    want = oracle_agg(tpch.generate_host("lineitem", 0.1, tpch.Q1_COLUMNS), q1aggs, gb, pred).sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")])
    assert_tables_equal(q1, want, float_rtol=1e-9)
    # RepartitionSpec::Random and ::Range, and the distributed sort built on the range exchange
    rnd = comm.exchange_random(rb, 7).to_arrow_table()
    cnt = torch.tensor([rnd.num_rows], device="cuda")
    dist.all_reduce(cnt)
    assert int(cnt.item()) == sum(20000 + 1000 * r for r in range(world))
    allr = [None] * world
    dist.all_gather_object(allr, sorted(rnd.column("i64").to_pylist(), key=lambda v: (v is None, v)))
    assert sorted((v for part_ in allr for v in part_), key=lambda v: (v is None, v)) == sorted(full.column("i64").to_pylist(), key=lambda v: (v is None, v))
    for keys, desc in ((["i64"], [False]), (["s", "i32"], [True, False])):
        srt = comm.sort(rb, [col(k) for k in keys], descending=desc).to_arrow_table()
        pieces = [None] * world
        dist.all_gather_object(pieces, distributed._to_ipc(srt))
        got_all = pa.concat_tables([distributed._from_ipc(p) for p in pieces])
        want_all = full.take(pa.array(O.argsort([full.column(k) for k in keys], desc, desc), type=pa.int64()))
        assert got_all.num_rows == want_all.num_rows
        for k in keys:   # rank order == global order on the sort keys (ties between equal keys are free)
            assert got_all.column(k).to_pylist() == want_all.column(k).to_pylist(), k
    # distributed hash join: exchanged by key hash (large right side) and broadcast (small right side) == single-table oracle
    lt = tbl.select(["i32", "i64", "s"]).rename_columns(["k", "lv", "ls"])
    rt = random_table(np.random.default_rng(500 + rank), 5000 + 100 * rank, 0.1).select(["i32", "f64"]).rename_columns(["k", "rv"])
    lparts, rparts = [None] * world, [None] * world
    dist.all_gather_object(lparts, distributed._to_ipc(lt))
    dist.all_gather_object(rparts, distributed._to_ipc(rt))
    lfull, rfull = (pa.concat_tables([distributed._from_ipc(p) for p in ps]) for ps in (lparts, rparts))
    for how in ("inner", "left", "anti"):
        for thr in (0, 1 << 20):
            j = comm.hash_join(daft_b200.RecordBatch.from_arrow_table(lt), daft_b200.RecordBatch.from_arrow_table(rt), [col("k")], [col("k")], how, thr).to_arrow_table()
            jparts = [None] * world
            dist.all_gather_object(jparts, distributed._to_ipc(j))
            got_j = pa.concat_tables([distributed._from_ipc(p) for p in jparts])
            want_j = O.hash_join(lfull, rfull, ["k"], ["k"], how)
            order = [(c, "ascending") for c in want_j.column_names]
            assert_tables_equal(got_j.sort_by(order), want_j.sort_by(order))
    # the distributed planner over the device backend: a few whole queries on sharded scan partitions == the single-GPU answer
    from daft_b200.distributed_plan import DeviceBackend, Sharded
    backend = DeviceBackend(comm)
    for qname, sf in (("q3", 0.05), ("q5", 0.05), ("q12", 0.05), ("q13", 0.05), ("q18", 0.05), ("q21", 0.05)):
        cols = tpch.Q5_COLUMNS if qname == "q5" else tpch.QUERY_COLUMNS[qname]
        for thr in (1 << 20, 300):
            shards = {t: Sharded(tpch.generate(t, sf, c, *distributed.shard_rows(tpch.num_rows(t, sf), rank, world)), backend, gather_rows=thr) for t, c in cols.items()}
            got_q = getattr(tpch, qname)(shards).to_arrow_table()
            want_q = getattr(tpch, qname)({t: tpch.generate(t, sf, c) for t, c in cols.items()}).to_arrow_table()
            assert_tables_equal(got_q, want_q, float_rtol=1e-9)
    dist.barrier()
    if rank == 0:
        print(f"exchange ok on {world} GPUs")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""All 22 TPC-H queries over N GPUs through the distributed planner (daft_b200/distributed_plan.py); run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 scripts/tpch22_distributed.py --sf 100

Each rank generates its scan partition (a contiguous row range) of every table straight into HBM; every query runs once to warm up and
twice timed (CUDA events on the library stream, max over ranks); rank 0 prints one JSON line.  NOT part of bench.py: the device backend of the
planner has only been verified with gloo + the oracle (tests/test_distributed_gloo.py), and an unverified collective sequence must not be
able to hang the driver's bench run."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10.0)
    ap.add_argument("--queries", default="")
    ap.add_argument("--gather-rows", type=int, default=1 << 20)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import daft_b200
    from daft_b200 import _lib, distributed, tpch
    from daft_b200.distributed_plan import DeviceBackend, Sharded
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    daft_b200.init(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    backend = DeviceBackend(distributed.Communicator())
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(), device=torch.device("cuda", local))
    wanted = [int(x) for x in args.queries.split(",") if x.strip()] or list(range(1, 23))
    per_query = {}
    for q in wanted:
        name = f"q{q}"
        cols = {"lineitem": tpch.Q1_COLUMNS} if q in (1, 6) else (tpch.Q5_COLUMNS if q == 5 else tpch.QUERY_COLUMNS[name])
        shards = {t: Sharded(tpch.generate(t, args.sf, c, *distributed.shard_rows(tpch.num_rows(t, args.sf), rank, world)), backend, gather_rows=args.gather_rows)
                  for t, c in cols.items()}
        arg = shards["lineitem"] if q in (1, 6) else shards
        fn = getattr(tpch, name)
        out = fn(arg)
        dist.barrier(); _lib.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(2):
            out = fn(arg)
        e1.record(stream)
        dist.barrier(); _lib.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 2], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_query[name] = {"ms": float(t.item()), "rows": len(out)}
        del shards, arg, out
    if rank == 0:
        print(json.dumps({"metric": "tpch_total_wall_clock_sec", "value": sum(v["ms"] for v in per_query.values()) / 1e3, "unit": "s", "n_gpus": world,
                          "scale_factor": args.sf, "per_query": per_query}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

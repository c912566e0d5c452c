"""Dev probe (torchrun): the distributed Q1 step with the group tables merged by one all-gather + merge kernel vs by grouped
all-reduces, same box, same data; and the exchange step with the staged / plain peer scatter."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
import daft_b200
from daft_b200 import _lib, tpch, col, distributed
daft_b200.init(lr)
comm = distributed.Communicator()
n = tpch.num_rows("lineitem", 10.0)
li = tpch.generate("lineitem", 10.0 * world, tpch.Q1_COLUMNS, rank * n, (rank + 1) * n)
ex = tpch.generate("lineitem", 10.0 * world, ["L_ORDERKEY", "L_QUANTITY"], rank * n, (rank + 1) * n)
def timed(fn, reps=40):
    for _ in range(5): fn()
    _lib.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    _lib.synchronize(); dist.barrier()
    t = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 4)
out = {"world": world}
for mode in ("allgather", "allreduce", "allgather", "allreduce"):
    os.environ["B200_AGG_MERGE"] = mode
    out.setdefault("q1_" + mode, []).append(timed(lambda: tpch.q1_distributed(li, comm)))
out["q1_local_only"] = timed(lambda: tpch.q1(li))
for mode in ("1", "0", "1", "0"):
    os.environ["B200_PEER_SCATTER_STAGED"] = mode
    out.setdefault("exchange_staged" + mode, []).append(timed(lambda: comm.exchange_by_hash(ex, [col("L_ORDERKEY")]), 10))
_lib.lib().b200_kernel_timing(1); _lib.kernel_timing_report()
os.environ["B200_PEER_SCATTER_STAGED"] = "1"
for _ in range(5): comm.exchange_by_hash(ex, [col("L_ORDERKEY")])
_lib.synchronize()
out["regions_ms"] = {k: round(v["ms"] / 5, 4) for k, v in _lib.kernel_timing_report().items()}
if rank == 0: print(json.dumps(out))
dist.barrier(); dist.destroy_process_group()

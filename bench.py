#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric ("hash-join/groupby rows/sec at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference algorithm (oracle/)

Top-level line = BASELINE config[1], "TPC-H SF10 Q1 hash-aggregate kernel on 1xB200": one step = one pass of
Filter -> Project -> HashAggregate (-> 4-row Sort) over one lineitem batch;
  value : inputs already resident in HBM when the timed region starts;
  e2e   : the same step through the public API with HOST (pinned) Arrow buffers: host->device copy of the
          7 input columns and device->host read of the result inside the timed region.
"q5" (N = 1) = BASELINE config[2], "TPC-H SF100 Q5 six-way hash-join pipeline on 1xB200": one step = the whole query
(2 filters, 5 hash joins, aggregate, sort) over the six resident tables; same value / e2e / roofline / cpu_baseline keys.
N > 1: weak scaling -- every rank owns the next contiguous SF10-sized row range of an SF(10*N) lineitem
(scan partitions = row ranges).  Q1 runs as a distributed low-cardinality aggregate (per-GPU aggregate, group tables
all-reduced over NVLink, identical final result on every rank; src/daft-distributed/src/pipeline_node/aggregate.rs:281-340).
"exchange" (N > 1) = a high-cardinality group-by (sum(L_QUANTITY) by L_ORDERKEY, the Q18 shape) through the NCCL
hash-repartition exchange: the data-path collective, reported with the NVLink bytes it moves and checked against the
same aggregate computed without the exchange ("exchange_parity").
Both arms run the same config; every result is checked against the CPU engine (oracle/cpu_engine.cpp) outside the
timed regions.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "tpch_q1_groupby_rows_per_sec"
UNIT = "rows/s"
Q1_BYTES_PER_ROW = 54  # 2 x (8 B offset + 1 B) + i64 + 3 x f64 + date32: the Arrow layout the reference holds
Q5_BYTES = {"lineitem": 32, "orders": 20, "customer": 16, "supplier": 16}  # compulsory input bytes per row


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=float, default=10.0, help="Q1 scale factor per GPU (BASELINE config[1] = 10)")
    ap.add_argument("--q5-sf", type=float, default=100.0, help="Q5 scale factor (BASELINE config[2] = 100)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-q5", action="store_true", help="skip the config[2] leg")
    ap.add_argument("--tpch22-sf", type=float, default=100.0, help="scale factor of the all-22-queries leg (N = 1; BASELINE metric: TPC-H SF100 total wall-clock); 0 skips it")
    ap.add_argument("--tpch22-queries", default="", help="comma-separated query numbers for that leg (default: all 22)")
    ap.add_argument("--tpch22-budget-s", type=float, default=150.0, help="wall-clock budget of that leg; queries not reached are listed as skipped")
    ap.add_argument("--parquet-sf", type=float, default=1.0, help="scale factor of the Parquet -> HBM scan leg (N = 1); 0 skips it")
    ap.add_argument("--no-exchange", action="store_true", help="N > 1: skip the exchange workload")
    ap.add_argument("--no-e2e", action="store_true", help="kernel sweeps only: skip the host-buffer leg (the line is then not a bench result)")
    return ap.parse_args()


def q1_config(sf: float) -> dict:
    """Identical in both arms (the driver compares them)."""
    return {"workload": f"TPC-H SF{sf:g} Q1: lineitem Filter(shipdate) -> Project -> HashAggregate(2 Utf8 keys, 8 aggs) -> Sort",
            "scale_factor_per_gpu": sf, "bytes_per_row": Q1_BYTES_PER_ROW,
            "layout": "Arrow as the reference holds it: flags LargeUtf8 (i64 offsets + 1 byte), qty Int64, money Float64, date Date32",
            "l2": "inputs (3.24 GB per GPU at SF10) are larger than the 126 MB L2; no flush needed",
            "parallelism": "scan partitions (row ranges) sharded over the GPUs; per-GPU aggregate, group tables all-reduced, identical result on every rank"}


def q5_config(sf: float) -> dict:
    return {"workload": f"TPC-H SF{sf:g} Q5: region x nation x supplier x lineitem x orders(1994) x customer -> HashAggregate(N_NAME) -> Sort",
            "scale_factor": sf, "bytes_per_row": Q5_BYTES,
            "l2": "inputs (22.46 GB at SF100) are larger than the 126 MB L2; no flush needed"}


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    @staticmethod
    def hold(step, ms_step: float, seconds: float = 0.45) -> int:
        """nvidia-smi answers every ~100 ms; a timed region of 20 sub-millisecond steps ends before its first answer.  Keep the SAME
        step running (un-timed) for `seconds`: a step COUNT derived from the agreed step time, so every rank of a collective step
        runs the same number."""
        extra = max(1, min(5000, int(seconds / max(ms_step * 1e-3, 1e-6)) + 1))
        for _ in range(extra):
            step()
        return extra

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": "the timed steps plus 0.45 s of the same step repeated un-timed right behind them "
                                              "(nvidia-smi answers every 100 ms)"}


def host_mem_available_gb() -> float:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def peak_hbm() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------
# CPU arm: the morsel-parallel C++ restatement of the reference algorithm (oracle/cpu_engine.cpp), all host cores.
# Nothing of daft_b200/ is imported on this path.
# --------------------------------------------------------------------------------------------------
def cpu_q1_run(sf: float, rows_begin: int, rows_end: int, steps: int, warmup: int, threads: int = 0):
    from oracle import cpu_engine as CE
    cols = CE.generate_columns("lineitem", sf, CE.Q1_COLUMNS, rows_begin, rows_end)
    n = rows_end - rows_begin
    ans, tasks = None, 0
    for _ in range(warmup):
        ans, tasks = CE.q1(cols, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        ans, tasks = CE.q1(cols, threads)
    dt = time.perf_counter() - t0
    return {"rows": n, "rows_per_s": n * steps / dt, "ms_per_step": dt / steps * 1e3, "tasks": tasks, "answer": ans, "steps": steps,
            "threads": threads or CE.usable_threads()}


def cpu_q5_run(sf: float, steps: int, warmup: int, threads: int = 0):
    from oracle import cpu_engine as CE
    cols = {t: CE.generate_columns(t, sf, c) for t, c in CE.Q5_COLUMNS.items()}
    n = len(cols["lineitem"]["L_ORDERKEY"][0])
    ans, tasks, probes = None, 0, 0
    for _ in range(warmup):
        ans, tasks, probes = CE.q5(cols, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        ans, tasks, probes = CE.q5(cols, threads)
    dt = time.perf_counter() - t0
    return {"rows": n, "rows_per_s": n * steps / dt, "probe_rows_per_s": probes * steps / dt, "ms_per_step": dt / steps * 1e3, "tasks": tasks,
            "answer": ans, "steps": steps, "threads": threads or CE.usable_threads(), "cols": cols}


def q5_cpu_sf(q5_sf: float) -> float:
    """The CPU engine needs the tables (22.5 GB at SF100) plus its probe tables in host RAM."""
    need_gb = 0.45 * q5_sf + 8
    return q5_sf if host_mem_available_gb() > need_gb else min(q5_sf, 10.0)


def host_cpu_info():
    """What the host gives this process: logical CPUs and the cgroup quota (a box can show 128 CPUs and allow a dozen)."""
    info = {"logical_cpus": os.cpu_count()}
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            info["cgroup_" + os.path.basename(path)] = open(path).read().strip()
            break
        except OSError:
            pass
    return info


CPU_NOTE = "CPU restatement of the reference algorithm (oracle/cpu_engine.cpp: 131072-row morsels, one worker per host thread, " \
           "per-bucket finalize), not Daft itself: no Rust toolchain in this image"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_engine as CE
    threads = CE.usable_threads()   # CPUs of the affinity mask, capped by the cgroup CPU quota
    n = CE.num_rows("lineitem", args.sf)
    steps, warmup = max(1, args.steps), max(1, args.warmup)
    r = cpu_q1_run(args.sf, 0, n, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["rows_per_s"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": q1_config(args.sf),
        "cpu_baseline": {"value": r["rows_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "tasks_per_step": r["tasks"],
                         "sample": f"the whole workload: lineitem SF{args.sf:g} ({n} rows) per step, {steps} steps", "note": CPU_NOTE},
        "e2e": {"value": r["rows_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": host_cpu_info(),
    }
    if not args.no_q5 and args.gpus == 1:
        sf5 = q5_cpu_sf(args.q5_sf)
        s5 = max(1, min(steps, 3))
        q = cpu_q5_run(sf5, s5, 1)
        line["q5"] = {"metric": "tpch_q5_join_pipeline_lineitem_rows_per_sec", "value": q["rows_per_s"], "unit": UNIT, "ms_per_step": q["ms_per_step"],
                      "probe_rows_per_s": q["probe_rows_per_s"], "steps": s5, "config": q5_config(args.q5_sf),
                      "cpu_baseline": {"value": q["rows_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "tasks_per_step": q["tasks"],
                                       "sample": f"SF{sf5:g} ({q['rows']} lineitem rows) per step, {s5} steps"
                                                 + ("" if sf5 == args.q5_sf else f" (host RAM too small for SF{args.q5_sf:g})"), "note": CPU_NOTE}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
def pinned_table(tbl):
    """Copy an Arrow table's buffers into pinned host memory (so the H2D copies run at full PCIe speed)."""
    import numpy as np
    import pyarrow as pa
    import torch

    keep, cols, total = [], [], 0
    for c in tbl.columns:
        a = c.chunk(0) if c.num_chunks == 1 else c.combine_chunks()
        bufs = []
        for b in a.buffers():
            if b is None:
                bufs.append(None)
                continue
            t = torch.empty(b.size, dtype=torch.uint8, pin_memory=True)
            t.numpy()[:] = np.frombuffer(b, dtype=np.uint8)
            keep.append(t)
            total += b.size
            bufs.append(pa.foreign_buffer(t.data_ptr(), b.size, base=t))
        cols.append(pa.Array.from_buffers(a.type, len(a), bufs, null_count=a.null_count))
    return pa.table(cols, names=tbl.column_names), total, keep


def check_answer(got, want, keys, what: str):
    """GPU answer vs the CPU engine's on the SAME rows: key and integer columns exact, floats within 1e-6 relative
    (north_star's tolerance; the summation orders differ)."""
    import math
    g, w = got.to_pydict(), want.to_pydict()
    assert list(g.keys()) == list(w.keys()), (what, list(g.keys()), list(w.keys()))
    assert got.num_rows == want.num_rows, (what, got.num_rows, want.num_rows)
    for name in g:
        for i, (a, b) in enumerate(zip(g[name], w[name])):
            if isinstance(b, float):
                assert math.isclose(a, b, rel_tol=1e-6, abs_tol=0.0), (what, name, i, a, b)
            else:
                assert a == b, (what, name, i, a, b)
    return "ok"


def time_steps(fn, steps, warmup, stream, barrier):
    import torch
    out = None
    for _ in range(warmup):
        out = fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        out = fn()
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1), out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist

    import daft_b200
    from daft_b200 import _lib, col, distributed, tpch

    torch.cuda.set_device(local_rank)
    daft_b200.init(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(), device=torch.device("cuda", local_rank))
    peak, peak_src = peak_hbm()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ================================================================================================
    # config[1]: Q1
    # ================================================================================================
    sf = args.sf
    n_rows = tpch.num_rows("lineitem", sf)
    total_sf = sf * world if world > 1 else sf
    begin = rank * n_rows  # rank r scans rows [r*n, (r+1)*n) of the SF(sf*world) table
    lineitem = tpch.generate("lineitem", total_sf, tpch.Q1_COLUMNS, begin, begin + n_rows)
    assert len(lineitem) == n_rows
    comm = distributed.Communicator() if world > 1 else None
    kernel_ms = []

    def step_q1(batch):
        out = tpch.q1(batch) if world == 1 else tpch.q1_distributed(batch, comm)
        kernel_ms.append(L.b200_last_agg_kernel_ms())
        return out

    L.b200_kernel_timing(1)
    for _ in range(args.warmup):
        res = step_q1(lineitem)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = daft_b200.launch_count()
    kernel_ms.clear()
    ms_total, res = time_steps(lambda: step_q1(lineitem), args.steps, 0, stream, barrier)
    launches = daft_b200.launch_count() - launches0
    kernel_ms_timed = list(kernel_ms)  # (the e2e leg below launches the same kernel again)
    ms_step = max_over_ranks(ms_total) / args.steps
    sampler.hold(lambda: step_q1(lineitem), ms_step)
    _lib.synchronize()   # the library's collectives are drained before torch.distributed's communicator is used again
    barrier()
    clocks = sampler.stop()
    L.b200_kernel_timing(0)
    value = world * n_rows / (ms_step * 1e-3)
    result = res.to_arrow_table()

    if args.no_e2e:
        if rank == 0:
            k_ms = [x for x in kernel_ms_timed if x > 0]
            print(json.dumps({"sweep_only": True, "ms_per_step": ms_step, "value": value, "kernel_ms": sum(k_ms) / max(len(k_ms), 1),
                              "gbps": Q1_BYTES_PER_ROW * n_rows / (sum(k_ms) / max(len(k_ms), 1) * 1e-3) / 1e9 if k_ms else None,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("B200_")}, "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- e2e: host buffers, H2D + D2H inside the timed region ----------------
    host = tpch.generate_host("lineitem", total_sf, tpch.Q1_COLUMNS, begin, begin + n_rows)
    host_pinned, h2d_bytes, _keep = pinned_table(host)
    del host

    def step_e2e():
        return step_q1(daft_b200.RecordBatch.from_arrow_table(host_pinned)).to_arrow_table()

    ms_e2e, out = time_steps(step_e2e, args.e2e_steps, 2, stream, barrier)
    ms_e2e = max_over_ranks(ms_e2e) / args.e2e_steps
    e2e_value = world * n_rows / (ms_e2e * 1e-3)
    d2h_bytes = int(out.nbytes)
    del host_pinned, _keep

    # ---------------- the answer, checked against the CPU engine on the SAME rows (outside every timed region) ----------------
    # rank 0 computes the CPU answer over the union of all ranks' row ranges = rows [0, world * n) of the SF(total) table
    cpu = None
    checks = {}
    if rank == 0 and not args.no_cpu_baseline:
        r = cpu_q1_run(total_sf, 0, world * n_rows, 2 if world == 1 else 1, 1)
        checks["q1_answer_vs_cpu_engine"] = check_answer(result, r["answer"], ["L_RETURNFLAG", "L_LINESTATUS"], "q1 value leg")
        checks["q1_e2e_answer_vs_cpu_engine"] = check_answer(out, r["answer"], ["L_RETURNFLAG", "L_LINESTATUS"], "q1 e2e leg")
        if world == 1:
            cpu = {"value": r["rows_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port", "tasks_per_step": r["tasks"],
                   "sample": f"the whole workload: lineitem SF{sf:g} ({r['rows']} rows) per step, {r['steps']} steps, {r['ms_per_step']:.0f} ms/step, "
                             f"all {r['threads']} host threads", "note": CPU_NOTE}

    # ---------------- roofline of the dominant kernel (fused filter+project+hash-aggregate) ----------------
    k_ms = [x for x in kernel_ms_timed if x > 0]
    k_avg = sum(k_ms) / len(k_ms) if k_ms else None
    alg_bytes = Q1_BYTES_PER_ROW * n_rows
    achieved = alg_bytes / (k_avg * 1e-3) / 1e9 if k_avg else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "q1_agg_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "b200_agg_k (JIT fused filter+project+hash-aggregate)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": k_avg, "algorithmic_bytes_per_launch": alg_bytes, "frac_of_8TBs_spec": (achieved / 8000.0) if achieved else None,
                "step_frac": alg_bytes / (ms_step * 1e-3) / 1e9 / peak}

    # ================================================================================================
    # N > 1: the data-path collective -- a high-cardinality group-by through the hash-repartition exchange
    # ================================================================================================
    exchange = None
    if world > 1 and not args.no_exchange:
        exchange = run_exchange(args, rank, world, comm, total_sf, begin, n_rows, stream, barrier, max_over_ranks)

    # ================================================================================================
    # config[2]: Q5 (one GPU)
    # ================================================================================================
    q5 = None
    if world == 1 and not args.no_q5:
        del lineitem
        q5 = run_q5(args, stream, barrier, peak, peak_src)

    # ================================================================================================
    # config[3] on one GPU: all 22 queries, tables resident in HBM
    # ================================================================================================
    tpch22 = None
    if world == 1 and args.tpch22_sf > 0:
        try:
            del lineitem
        except NameError:
            pass
        tpch22 = run_tpch22(args, stream, barrier)

    parquet = None
    if world == 1 and args.parquet_sf > 0:
        parquet = run_parquet(args, stream, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cfg = q1_config(sf)
    cfg.update({"rows_per_gpu": n_rows, "result_groups": result.num_rows})
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": ms_e2e, "steps": args.e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "checks": checks,
    }
    if q5 is not None:
        line["q5"] = q5
    if tpch22 is not None:
        line["tpch22"] = tpch22
    if parquet is not None:
        line["parquet_scan"] = parquet
    if exchange is not None:
        line["exchange"] = exchange
        line["exchange_parity"] = exchange["parity"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_exchange(args, rank, world, comm, total_sf, begin, n_rows, stream, barrier, max_over_ranks):
    """sum(L_QUANTITY) by L_ORDERKEY over this rank's scan partition, through b200_exchange_by_hash (NCCL all-to-all)."""
    import torch
    import torch.distributed as dist
    import daft_b200
    from daft_b200 import col, tpch
    part = tpch.generate("lineitem", total_sf, ["L_ORDERKEY", "L_QUANTITY"], begin, begin + n_rows)
    aggs, gb = [col("L_QUANTITY").sum().alias("sum_qty"), col("L_QUANTITY").count().alias("n")], [col("L_ORDERKEY")]

    def step():
        return comm.exchange_by_hash(part, gb).agg(aggs, gb)

    steps = max(3, min(args.steps, 10))
    ms, out = time_steps(step, steps, 2, stream, barrier)
    ms = max_over_ranks(ms) / steps
    # un-timed exchange step alone (the collective's own bytes / time)
    from daft_b200 import _lib
    L = _lib.lib()
    L.b200_kernel_timing(1)
    _lib.kernel_timing_report()
    ms_x, ex = time_steps(lambda: comm.exchange_by_hash(part, gb), steps, 1, stream, barrier)
    kt = _lib.kernel_timing_report()   # device time of the library's timed regions on this rank (nested: peer_store contains the scatter)
    L.b200_kernel_timing(0)
    ms_x = max_over_ranks(ms_x) / steps
    # parity (un-timed): placement = hash_rows % world as the reference's partition_by_hash; every key lands on exactly
    # one rank, so the per-rank groups partition the global groups: (groups, rows, sum) must add up to the values of the
    # SAME aggregate computed locally without the exchange, and every received key must hash to this rank.
    import numpy as np
    h = np.array(ex.hash_rows(gb).to_pydict()["hash"][:100000], dtype=np.uint64)
    placement_ok = bool((h % np.uint64(world) == np.uint64(rank)).all())
    got = out.agg([col("sum_qty").sum().alias("s"), col("n").sum().alias("c"), col("n").count("all").alias("g")], []).to_pydict()
    loc = part.agg([col("L_QUANTITY").sum().alias("s"), col("L_QUANTITY").count().alias("c")], []).to_pydict()
    t = torch.tensor([got["s"][0], got["c"][0], got["g"][0], loc["s"][0], loc["c"][0], 0 if placement_ok else 1], dtype=torch.int64, device="cuda")
    dist.all_reduce(t)
    tot = t.tolist()
    # distinct order keys among the first n_all lineitem rows: the generator emits 28 lines per 7 consecutive orders
    # (4, 1, 7, 2, 6, 3, 5 lines; daft_b200/csrc/tpch_gen_core.h line_of), and n_all <= 4 x orders, so no order index wraps
    n_all = world * n_rows
    starts = (0, 4, 5, 12, 14, 20, 23)
    expect_groups = (n_all // 28) * 7 + sum(1 for s0 in starts if s0 < n_all % 28)
    parity = "ok" if (tot[0] == tot[3] and tot[1] == tot[4] and tot[5] == 0 and tot[1] == n_all and tot[2] == expect_groups) else \
        f"MISMATCH {tot} expected_groups={expect_groups}"
    sent = n_rows * 16 * (world - 1) / world  # bytes each GPU sends (and receives): two Int64 columns
    return {"workload": f"TPC-H SF{total_sf:g} lineitem: sum(L_QUANTITY), count by L_ORDERKEY through exchange_by_hash (hash_rows % {world}) -> local HashAggregate",
            "value": n_all / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "exchange_ms": ms_x,
            "nvlink_bytes_sent_per_gpu": int(sent), "nvlink_GBps_per_gpu_per_direction": sent / (ms_x * 1e-3) / 1e9,
            "nvlink_peak_GBps_per_direction": 900.0, "groups": tot[2], "parity": parity,
            "regions_ms_per_step_rank0": {k: round(v["ms"] / (steps + 1), 4) for k, v in kt.items()}}


def run_parquet(args, stream, barrier):
    """SURVEY 8f-1 on the clock: TPC-H lineitem (the 7 Q1 columns) written by pyarrow as the reference's data generation writes it
    (Snappy, dictionary pages, 1 Mi-row row groups) to local disk, then `read_parquet` -> HBM (footer, page headers and file reads on
    the host, decompression and decoding on the device) and Q1 over the scanned batch.  Wall clock around synchronised calls (the scan
    includes host file reads), device stage times from the library's timing registry, the answer compared with Q1 over the generated
    table, pyarrow's reader on the host cores beside it.  Any failure is reported in place of the numbers."""
    import tempfile
    import daft_b200
    from daft_b200 import _lib, tpch
    try:
        import pyarrow.parquet as pq
        sf = args.parquet_sf
        host = tpch.generate_host("lineitem", sf, tpch.Q1_COLUMNS)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "lineitem.parquet")
            pq.write_table(host, path, compression="snappy", row_group_size=1 << 20)
            size = os.path.getsize(path)
            groups, rows = daft_b200.parquet_file_info(path)
            for _ in range(2):
                rb = daft_b200.read_parquet(path)
            L = _lib.lib()
            L.b200_kernel_timing(1)
            _lib.kernel_timing_report()
            steps = 3
            _lib.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                rb = daft_b200.read_parquet(path)
            _lib.synchronize()
            scan_s = (time.perf_counter() - t0) / steps
            kt = _lib.kernel_timing_report()
            L.b200_kernel_timing(0)
            t0 = time.perf_counter()
            for _ in range(steps):
                ans = tpch.q1(daft_b200.read_parquet(path)).to_arrow_table()
            e2e_s = (time.perf_counter() - t0) / steps
            want = tpch.q1(daft_b200.RecordBatch.from_arrow_table(host)).to_arrow_table()
            check = check_answer(ans, want, ["L_RETURNFLAG", "L_LINESTATUS"], "q1 over the scanned file")
            pq.read_table(path)
            t0 = time.perf_counter()
            pq.read_table(path)
            pa_s = time.perf_counter() - t0
        return {"workload": f"TPC-H SF{sf:g} lineitem, Q1 columns: Parquet (Snappy, dictionary, {groups} row groups) on local disk -> HBM -> Q1",
                "rows": rows, "file_bytes": size, "arrow_bytes": int(host.nbytes), "scan_ms": scan_s * 1e3, "scan_rows_per_s": rows / scan_s,
                "scan_arrow_GBps": host.nbytes / scan_s / 1e9, "scan_plus_q1_ms": e2e_s * 1e3, "device_stage_ms": {k: v["ms"] / steps for k, v in kt.items()},
                "answer_vs_generated_table": check, "pyarrow_read_table_ms": pa_s * 1e3, "pyarrow_cpus": os.cpu_count()}
    except Exception as e:  # noqa: BLE001 -- this leg never takes the bench line down
        return {"error": f"{type(e).__name__}: {str(e)[:300]}"}


def run_tpch22(args, stream, barrier):
    """BASELINE metric, first half ("TPC-H total wall-clock sec") at one GPU: each of the 22 queries of benchmarking/tpch/answers.py
    (daft_b200/tpch.py) over tables generated into HBM at --tpch22-sf, one warm-up (it also JIT-compiles the query's kernels) and
    two timed runs with CUDA events.  The answers of these compositions are held to SQLite and the oracle at SF 0.01-0.05 by
    tests/test_gpu_sqlite_truth.py and tests/test_gpu_zzz_tpch_all.py; here only the result's row count is reported.  A query that
    fails is reported with its error instead of a time, and queries beyond the wall-clock budget are listed as skipped."""
    import daft_b200
    from daft_b200 import _lib, tpch
    sf = args.tpch22_sf
    t_begin = time.perf_counter()
    per_query, skipped, failed = {}, [], {}
    launches = 0
    wanted = [int(x) for x in args.tpch22_queries.split(",") if x.strip()] or list(range(1, 23))
    for q in wanted:
        name = f"q{q}"
        if time.perf_counter() - t_begin > args.tpch22_budget_s:
            skipped.append(name)
            continue
        try:
            if q in (1, 6):
                tables = tpch.generate("lineitem", sf, tpch.Q1_COLUMNS)
            else:
                cols = tpch.Q5_COLUMNS if q == 5 else tpch.QUERY_COLUMNS[name]
                tables = {t: tpch.generate(t, sf, c) for t, c in cols.items()}
            fn = getattr(tpch, name)
            l0 = daft_b200.launch_count()
            ms, out = time_steps(lambda: fn(tables), 2, 1, stream, barrier)
            n_launch = (daft_b200.launch_count() - l0) // 3
            launches += n_launch * 2
            per_query[name] = {"ms": ms / 2, "rows": len(out), "launches": int(n_launch)}
            del tables, out
        except Exception as e:  # noqa: BLE001 -- the other queries' numbers stand
            failed[name] = f"{type(e).__name__}: {str(e)[:300]}"
            try:
                _lib.synchronize()
            except Exception:  # noqa: BLE001
                pass
    total_ms = sum(v["ms"] for v in per_query.values())
    lineitem_rows = tpch.num_rows("lineitem", sf)
    return {"metric": "tpch_total_wall_clock_sec", "value": total_ms / 1e3 if not failed and not skipped and len(per_query) == 22 else None, "unit": "s",
            "higher_is_better": False, "sum_of_completed_queries_s": total_ms / 1e3, "queries_completed": len(per_query),
            "queries_failed": failed, "queries_skipped": skipped, "per_query": per_query,
            "config": {"workload": f"TPC-H SF{sf:g}, all 22 queries of benchmarking/tpch/answers.py, 1 x B200, tables resident in HBM "
                                   "(generated on the device, only the columns a query reads)", "scale_factor": sf,
                       "lineitem_rows": lineitem_rows, "timing": "per query: 1 warm-up + 2 timed runs, CUDA events on the library stream"},
            "gpu_launches": int(launches), "cpu_baseline": None,
            "note": "no CPU arm for this leg: the CPU engine (oracle/cpu_engine.cpp) implements Q1 and Q5 only"}


def run_q5(args, stream, barrier, peak, peak_src):
    import numpy as np
    import pyarrow as pa
    import torch
    import daft_b200
    from daft_b200 import _lib, tpch
    L = _lib.lib()
    sf = args.q5_sf
    tables = {t: tpch.generate(t, sf, cols) for t, cols in tpch.Q5_COLUMNS.items()}
    _lib.synchronize()
    rows = {t: len(b) for t, b in tables.items()}
    alg_bytes = sum(rows[t] * w for t, w in Q5_BYTES.items())
    n_li = rows["lineitem"]
    steps = max(3, min(args.steps, 20))
    # ---- value: tables resident in HBM
    ms_w, res = time_steps(lambda: tpch.q5(tables), 1, max(args.warmup, 3), stream, barrier)
    L.b200_kernel_timing(1)
    _lib.kernel_timing_report()
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    launches0 = daft_b200.launch_count()
    ms, res = time_steps(lambda: tpch.q5(tables), steps, 0, stream, barrier)
    launches = daft_b200.launch_count() - launches0
    kt = _lib.kernel_timing_report()
    sampler.hold(lambda: tpch.q5(tables), ms / steps)
    _lib.synchronize()
    clocks = sampler.stop()
    _lib.kernel_timing_report()
    L.b200_kernel_timing(0)
    ms /= steps
    result = res.to_arrow_table()
    # ---- roofline of the dominant kernel: the probe of 600 M L_SUPPKEY values (count pass + write pass)
    roof = {"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src, "traffic": None,
            "pipeline": {"algorithmic_bytes_per_step": alg_bytes, "achieved": alg_bytes / (ms * 1e-3) / 1e9,
                         "frac": alg_bytes / (ms * 1e-3) / 1e9 / peak, "floor_ms": alg_bytes / peak / 1e6},
            "kernels_ms_per_step": {k: v["ms"] / steps for k, v in kt.items()},
            "kernel_launches_per_step": {k: v["launches"] / steps for k, v in kt.items()}}
    pc = kt.get("join_probe_count")
    if pc:
        # per step the count pass streams: L_SUPPKEY (600 M x 8 B), L_ORDERKEY read through the selection vector
        # (120 M x (8 B row id + 8 B key)), O_CUSTKEY / N_NATIONKEY hashes of the last join (18 M x 8 B)
        roof.update({"kernel": "probe_kernel<*, count> (all three big joins)", "kernel_ms": pc["ms"] / steps,
                     "algorithmic_bytes_per_launch": n_li * 8, "achieved": None, "frac": None})
    # ---- host tables: generated once by the CPU engine's generator (the same values as the device generator,
    # tests/test_cpu_engine.py), page-locked in place, used by the e2e leg AND by the CPU baseline / answer check
    from oracle import cpu_engine as CE
    sf_cpu = q5_cpu_sf(sf) if not args.no_cpu_baseline else sf
    cols = {t: CE.generate_columns(t, sf, c) for t, c in CE.Q5_COLUMNS.items()}
    rt = torch.cuda.cudart()
    pinned, h2d, registered = {}, 0, []
    for t, names in tpch.Q5_COLUMNS.items():
        if t in cols:
            arrays = []
            for c in names:
                a = cols[t][c][0]
                if rt.cudaHostRegister(a.ctypes.data, a.nbytes, 0) == 0:
                    registered.append(a)
                typ = pa.date32() if a.dtype == np.int32 else (pa.float64() if a.dtype == np.float64 else pa.int64())
                arrays.append(pa.Array.from_buffers(typ, len(a), [None, pa.py_buffer(a)]))
                h2d += a.nbytes
            pinned[t] = pa.table(arrays, names=names)
        else:
            pinned[t] = tpch.generate_host(t, sf, names)  # nation, region: a few rows
            h2d += pinned[t].nbytes

    def step_e2e():
        return tpch.q5({t: daft_b200.RecordBatch.from_arrow_table(p) for t, p in pinned.items()}).to_arrow_table()

    e2e_steps = max(2, min(args.e2e_steps, 5))
    ms_e2e, out = time_steps(step_e2e, e2e_steps, 1, stream, barrier)
    ms_e2e /= e2e_steps
    for a in registered:
        rt.cudaHostUnregister(a.ctypes.data)
    # ---- CPU engine on the same tables: the answer check and the cpu_baseline
    cpu, checks = None, {}
    checks["q5_e2e_equals_value_leg"] = check_answer(out, result, ["N_NAME"], "q5 e2e vs value")
    if not args.no_cpu_baseline:
        if sf_cpu != sf:   # host RAM too small to run the CPU engine next to the pinned tables: a bounded sample instead
            del cols, pinned
            cols = {t: CE.generate_columns(t, sf_cpu, c) for t, c in CE.Q5_COLUMNS.items()}
        ans, tasks, probes = CE.q5(cols, 0)   # warm-up
        t0 = time.perf_counter()
        ans, tasks, probes = CE.q5(cols, 0)
        dt = time.perf_counter() - t0
        n_cpu = len(cols["lineitem"]["L_ORDERKEY"][0])
        threads = CE.usable_threads()
        if sf_cpu == sf:
            checks["q5_answer_vs_cpu_engine"] = check_answer(result, ans, ["N_NAME"], "q5 value leg")
        else:
            small = tpch.q5({t: tpch.generate(t, sf_cpu, c) for t, c in tpch.Q5_COLUMNS.items()}).to_arrow_table()
            checks[f"q5_sf{sf_cpu:g}_answer_vs_cpu_engine"] = check_answer(small, ans, ["N_NAME"], "q5 reduced")
        cpu = {"value": n_cpu / dt, "unit": UNIT, "cores": threads, "kind": "port", "tasks_per_step": tasks, "probe_rows_per_s": probes / dt,
               "sample": f"SF{sf_cpu:g} ({n_cpu} lineitem rows), 1 step after 1 warm-up, {dt * 1e3:.0f} ms/step, all {threads} host threads",
               "note": CPU_NOTE}
    return {"metric": "tpch_q5_join_pipeline_lineitem_rows_per_sec", "value": n_li / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps,
            "config": dict(q5_config(sf), rows=rows, compulsory_input_bytes=alg_bytes),
            "e2e": {"value": n_li / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(out.nbytes),
                    "ms_per_step": ms_e2e, "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "checks": checks,
            "result": result.to_pydict()}


if __name__ == "__main__":
    main()

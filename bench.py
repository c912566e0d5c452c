#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: group-by rows/s on TPC-H Q1 (config[1]:
"TPC-H SF10 Q1 hash-aggregate kernel on 1xB200").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU port of the reference algorithm (oracle/)

One step = one pass of Filter -> Project -> HashAggregate (-> 4-row Sort) over one lineitem batch:
  value : inputs already resident in HBM when the timed region starts;
  e2e   : the same step through the public API with HOST (pinned) Arrow buffers: host->device copy of the
          7 input columns and device->host read of the result inside the timed region.
N > 1: weak scaling -- every rank owns the next contiguous SF10-sized row range of an SF(10*N) lineitem
(scan partitions = row ranges), runs the partial aggregate on its GPU, the 4-row partials are all-gathered
and every rank runs the final merge aggregate (the reference's two-stage plan shape,
src/daft-distributed/src/pipeline_node/aggregate.rs:281-340).  No data-path collective beyond that.
"""
from __future__ import annotations

import argparse
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


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=float, default=10.0, help="scale factor per GPU (BASELINE config[1] = 10)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel sweeps only: skip the host-buffer leg (the line is then not a bench result)")
    return ap.parse_args()


def workload_name(sf: float) -> str:
    return f"TPC-H SF{sf:g} Q1: lineitem Filter(shipdate) -> Project -> HashAggregate(2 Utf8 keys, 8 aggs) -> Sort"


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
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU port of the reference algorithm (oracle/): per-morsel partial aggregate on a thread pool, then the
# final merge -- the scheme of GroupedAggregateSink (grouped_aggregate.rs:355-450), 131 072-row morsels.
# --------------------------------------------------------------------------------------------------
def cpu_q1(tbl, threads: int):
    import pyarrow as pa
    from concurrent.futures import ThreadPoolExecutor
    from daft_b200 import tpch
    from helpers import oracle_agg

    pred, gb, aggs = tpch.q1_partial_exprs()
    fgb, faggs, fproj = tpch.q1_final_exprs()
    morsel = 131072 * 8  # the sink concatenates queued morsels before aggregating (blocking_sink.rs:123-137)
    pieces = [tbl.slice(o, morsel) for o in range(0, tbl.num_rows, morsel)]
    if threads <= 1:
        parts = [oracle_agg(p, aggs, gb, pred) for p in pieces]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(lambda p: oracle_agg(p, aggs, gb, pred), pieces))
    merged = oracle_agg(pa.concat_tables(parts), faggs, fgb)
    from oracle import expr_eval as E
    return E.eval_expression_list(merged, fproj).sort_by([("L_RETURNFLAG", "ascending"), ("L_LINESTATUS", "ascending")])


def time_cpu(sample_sf: float, threads: int, steps: int, warmup: int):
    from daft_b200 import tpch
    tbl = tpch.generate_host("lineitem", sample_sf, tpch.Q1_COLUMNS)
    for _ in range(warmup):
        cpu_q1(tbl, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_q1(tbl, threads)
    dt = time.perf_counter() - t0
    return tbl.num_rows * steps / dt, dt / steps * 1e3, tbl.num_rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_sf = 1.0
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    val, ms, rows = time_cpu(sample_sf, threads, steps, warmup)
    sample = f"lineitem SF{sample_sf:g} ({rows} rows) per step, {steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.sf), "scale_factor_per_gpu": args.sf},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU restatement of the reference algorithm (oracle/), not Daft itself: no Rust toolchain in this image"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


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
    import pyarrow as pa

    import daft_b200
    from daft_b200 import _lib, distributed, tpch

    torch.cuda.set_device(local_rank)
    daft_b200.init(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    L = _lib.lib()
    stream = torch.cuda.ExternalStream(_lib.stream_ptr(), device=torch.device("cuda", local_rank))

    sf = args.sf
    n_rows = tpch.num_rows("lineitem", sf)
    begin = rank * n_rows
    # rank r scans rows [r*n, (r+1)*n) of the SF(sf*world) table
    lineitem = tpch.generate("lineitem", sf * world if world > 1 else sf, tpch.Q1_COLUMNS, begin, begin + n_rows)
    assert len(lineitem) == n_rows

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        _lib.synchronize()

    # N > 1: partial aggregate per GPU -> device-side all-gather of the (4-row) partials over NCCL (one fixed-size
    # ncclAllGather inside libdaftb200, exchange.cu all_gather_small) -> identical final merge on every rank
    comm = distributed.Communicator() if world > 1 else None
    kernel_ms = []

    def step_device(batch):
        if world == 1:
            out = tpch.q1(batch)
            kernel_ms.append(L.b200_last_agg_kernel_ms())
            return out
        part = tpch.q1_partial(batch)
        kernel_ms.append(L.b200_last_agg_kernel_ms())
        return tpch.q1_final(comm.all_gather(part))

    # ---------------- value: inputs resident in HBM ----------------
    L.b200_kernel_timing(1)
    for _ in range(args.warmup):
        res = step_device(lineitem)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = daft_b200.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms.clear()
    e0.record(stream)
    for _ in range(args.steps):
        res = step_device(lineitem)
    e1.record(stream)
    barrier()
    kernel_ms_timed = list(kernel_ms)  # (the e2e leg below launches the same kernel again)
    clocks = sampler.stop()
    launches = daft_b200.launch_count() - launches0
    ms_total = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total.item()) / args.steps
    value = world * n_rows / (ms_step * 1e-3)
    result = res.to_arrow_table()

    if args.no_e2e:
        if rank == 0:
            k_ms = [x for x in kernel_ms_timed if x > 0]
            print(json.dumps({"sweep_only": True, "ms_per_step": ms_step, "value": value, "kernel_ms": sum(k_ms) / max(len(k_ms), 1),
                              "gbps": tpch.Q1_BYTES_PER_ROW * n_rows / (sum(k_ms) / max(len(k_ms), 1) * 1e-3) / 1e9 if k_ms else None,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("B200_")}, "clocks": clocks}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- e2e: host buffers, H2D + D2H inside the timed region ----------------
    host = tpch.generate_host("lineitem", sf * world if world > 1 else sf, tpch.Q1_COLUMNS, begin, begin + n_rows)
    host_pinned, h2d_bytes, _keep = pinned_table(host)
    del host

    def step_e2e():
        rb = daft_b200.RecordBatch.from_arrow_table(host_pinned)
        out = step_device(rb).to_arrow_table()
        return out

    out = step_e2e()
    out = step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    for _ in range(args.e2e_steps):
        out = step_e2e()
    e3.record(stream)
    barrier()
    ms_e2e = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * n_rows / (float(ms_e2e.item()) / args.e2e_steps * 1e-3)
    d2h_bytes = int(out.nbytes)
    assert out.equals(result) or out.num_rows == result.num_rows

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (fused filter+project+hash-aggregate) ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    k_ms = [x for x in kernel_ms_timed if x > 0]
    k_avg = sum(k_ms) / len(k_ms) if k_ms else None
    alg_bytes = tpch.Q1_BYTES_PER_ROW * n_rows
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
                "kernel_ms": k_avg, "algorithmic_bytes_per_launch": alg_bytes, "frac_of_8TBs_spec": (achieved / 8000.0) if achieved else None}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, ms, rows = time_cpu(1.0, 1, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first SF1 slice of the workload ({rows} rows/step, 3 steps, {ms:.0f} ms/step), single thread; "
                         "CPU restatement of the reference algorithm (oracle/), not Daft itself"}

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(sf), "scale_factor_per_gpu": sf, "rows_per_gpu": n_rows, "bytes_per_row": tpch.Q1_BYTES_PER_ROW,
                   "layout": "Arrow as the reference holds it: flags LargeUtf8 (i64 offsets + 1 byte), qty Int64, money Float64, date Date32",
                   "l2": "inputs (%.2f GB per GPU) are larger than the 126 MB L2; no flush needed" % (alg_bytes / 1e9),
                   "parallelism": f"scan partitions sharded over {world} GPU(s); partial agg per GPU, all-gather of 4-row partials, final merge",
                   "result_groups": result.num_rows},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": float(ms_e2e.item()) / args.e2e_steps, "steps": args.e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

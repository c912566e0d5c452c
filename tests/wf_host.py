"""ctypes loader of tests/wf_host.cpp (the ROWS-frame window state machine's host build; test infrastructure)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libwf_host.so")
_lib = None
AGG = {"sum": 0, "count_all": 1, "count_valid": 2, "count_null": 3, "mean": 4, "min": 5, "max": 6}
VTYPE = {pa.int64(): 0, pa.uint64(): 1, pa.float32(): 2, pa.float64(): 3}


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "wf_host.cpp"), os.path.join(_HERE, "..", "daft_b200", "csrc", "window_frame_core.h")]
        if not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
            tmp = f"{_SO}.{os.getpid()}.tmp"
            # -ffp-contract=off: a sliding float sum is a sequence of single adds and subtracts on the device too
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-Wall", *os.environ.get("B200_EMU_FLAGS", "").split(), "-o", tmp, srcs[0]])
            os.replace(tmp, _SO)
        _lib = ctypes.CDLL(_SO)
    return _lib


def _bits(arr: pa.Array):
    if arr is None or not arr.null_count:
        return None
    return np.packbits(np.asarray(arr.is_valid()), bitorder="little")


def run(values: pa.Array | None, partition_begin: np.ndarray, agg: int, start, end, min_periods: int, kind: int = 1, order: pa.Array | None = None,
        descending: bool = False) -> pa.Array:
    """what b200_window_frame returns for the same arguments (values already in the accumulate type, `order` as Int64 for RANGE frames);
    `kind`: 1 = start / end are row offsets, 2 = offsets on the order value; None = unbounded on that side"""
    n = len(partition_begin)
    pb = np.ascontiguousarray(partition_begin, dtype=np.int64)
    counts = agg in (1, 2, 3)
    vt, vbuf = 0, None
    out_type = pa.uint64()
    bits = _bits(values)
    if not counts:
        vt = VTYPE[values.type]
        vbuf = np.ascontiguousarray(values.fill_null(0).to_numpy(zero_copy_only=False))
        out_type = pa.float64() if agg == 4 else values.type
    obuf = np.ascontiguousarray(order.fill_null(0).to_numpy(zero_copy_only=False), dtype=np.int64) if order is not None else None
    obits = _bits(order)
    out = np.zeros(max(n, 1), dtype={pa.uint64(): np.uint64, pa.int64(): np.int64, pa.float32(): np.float32, pa.float64(): np.float64}[out_type])
    ok = np.zeros(max(n, 1), dtype=np.uint8)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    rc = lib().wf_host_run(ctypes.c_int(agg), ctypes.c_int(vt), ctypes.c_int(kind if start is not None else 0), ctypes.c_longlong(start or 0),
                           ctypes.c_int(kind if end is not None else 0), ctypes.c_longlong(end or 0), ctypes.c_int(bool(descending)), ctypes.c_longlong(min_periods),
                           ptr(vbuf), ptr(bits), ptr(obuf), ptr(obits), ptr(pb), ctypes.c_longlong(n), ptr(out), ptr(ok))
    if rc == 2:
        raise ValueError("Negative frame size is not allowed")
    assert rc == 0
    return pa.array(out[:n], type=out_type, mask=ok[:n] == 0)

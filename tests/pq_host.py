"""ctypes loader of tests/pq_host.cpp (the Parquet scan's host build; test infrastructure) -> pyarrow Tables."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import pyarrow as pa

from daft_b200 import datatypes as dt

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpq_host.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, "pq_host.cpp"), os.path.join(_HERE, "..", "daft_b200", "csrc", "parquet_core.h"),
                os.path.join(_HERE, "..", "daft_b200", "csrc", "parquet_read.h")]
        if not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fno-strict-aliasing", "-shared", "-fPIC", "-Wall", "-pthread", "-o", _SO, srcs[0]])
        L = ctypes.CDLL(_SO)
        L.pq_host_read.restype = ctypes.c_void_p
        L.pq_host_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, ctypes.c_longlong, ctypes.c_longlong]
        L.pq_host_error.restype = ctypes.c_char_p
        L.pq_host_error.argtypes = [ctypes.c_void_p]
        L.pq_host_rows.restype = ctypes.c_longlong
        L.pq_host_rows.argtypes = [ctypes.c_void_p]
        L.pq_host_num_columns.argtypes = [ctypes.c_void_p]
        L.pq_host_column.restype = ctypes.c_char_p
        L.pq_host_column.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                                     ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p)]
        L.pq_host_free.argtypes = [ctypes.c_void_p]
        L.pq_host_row_groups.restype = ctypes.c_longlong
        L.pq_host_row_groups.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong)]
        _lib = L
    return _lib


def _copy(ptr, nbytes) -> pa.Buffer:
    if not ptr or nbytes == 0:
        return pa.py_buffer(b"")
    return pa.py_buffer(ctypes.string_at(ptr, nbytes))


def read_table(path: str, columns=None, row_groups: tuple[int, int] = (0, -1)) -> pa.Table:
    L = lib()
    arr = (ctypes.c_char_p * len(columns))(*[c.encode() for c in columns]) if columns else None
    h = L.pq_host_read(path.encode(), arr, len(columns) if columns else 0, row_groups[0], row_groups[1])
    try:
        err = L.pq_host_error(h)
        if err:
            raise ValueError(err.decode())
        n = L.pq_host_rows(h)
        cols, names = [], []
        for i in range(L.pq_host_num_columns(h)):
            dtype, nulls, nbytes = ctypes.c_int(), ctypes.c_longlong(), ctypes.c_longlong()
            data, valid, offs = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
            name = L.pq_host_column(h, i, ctypes.byref(dtype), ctypes.byref(nulls), ctypes.byref(nbytes), ctypes.byref(data), ctypes.byref(valid), ctypes.byref(offs))
            typ = dt.to_arrow(dtype.value)
            vbuf = _copy(valid.value, (n + 7) // 8) if nulls.value else None
            if dt.is_string(dtype.value):
                bufs = [vbuf, _copy(offs.value, (n + 1) * 8), _copy(data.value, nbytes.value)]
            else:
                bufs = [vbuf, _copy(data.value, nbytes.value)]
            cols.append(pa.Array.from_buffers(typ, n, bufs, null_count=nulls.value))
            names.append(name.decode())
        return pa.table(cols, names=names)
    finally:
        L.pq_host_free(h)


def row_groups(path: str) -> tuple[int, int]:
    rows = ctypes.c_longlong()
    g = lib().pq_host_row_groups(path.encode(), ctypes.byref(rows))
    if g < 0:
        raise ValueError("not a Parquet file")
    return int(g), int(rows.value)

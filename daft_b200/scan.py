"""Parquet -> HBM scan: the scan-task source of the reference (src/daft-local-execution/src/sources/scan_task.rs:33,
src/daft-parquet/src/read.rs:553 stream_parquet) for local files.  The host reads bytes and page headers; decompression and
decoding run on the device (daft_b200/csrc/parquet.cu, parquet_core.h)."""
from __future__ import annotations

import ctypes
from typing import Sequence

from . import _lib
from .recordbatch import RecordBatch


def parquet_file_info(path: str) -> tuple[int, int]:
    """(row groups, rows) from the footer"""
    groups, rows = ctypes.c_int64(), ctypes.c_int64()
    _lib.call("b200_parquet_file_info", path.encode(), ctypes.byref(groups), ctypes.byref(rows))
    return int(groups.value), int(rows.value)


def read_parquet(path: str, columns: Sequence[str] | None = None, row_groups: tuple[int, int] | None = None) -> RecordBatch:
    """Row groups [begin, end) (default: all) of `columns` (default: all) decoded straight into HBM.  Scan partitions are row-group
    ranges: rank r of N reads row_groups = distributed.shard_rows(num_row_groups, r, N)."""
    _lib.ensure_init()
    arr = (ctypes.c_char_p * len(columns))(*[c.encode() for c in columns]) if columns else None
    begin, end = row_groups if row_groups is not None else (0, -1)
    out = ctypes.c_void_p()
    _lib.call("b200_parquet_read", path.encode(), arr, ctypes.c_int32(len(columns) if columns else 0), ctypes.c_int64(begin), ctypes.c_int64(end),
              ctypes.byref(out))
    return RecordBatch._wrap(out)

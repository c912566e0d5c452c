"""ctypes binding of libdaftb200.so (the C ABI in include/daft_b200.h).

The library is the product: if it is missing, fails to load, or no CUDA device is present, every
operator raises -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdaftb200.so")
CSRC = os.path.join(_HERE, "csrc")


class B200Error(RuntimeError):
    pass


class ArrowSchema(ctypes.Structure):
    pass


class ArrowArray(ctypes.Structure):
    pass


ArrowSchema._fields_ = [
    ("format", ctypes.c_char_p), ("name", ctypes.c_char_p), ("metadata", ctypes.c_char_p), ("flags", ctypes.c_int64),
    ("n_children", ctypes.c_int64), ("children", ctypes.POINTER(ctypes.POINTER(ArrowSchema))),
    ("dictionary", ctypes.POINTER(ArrowSchema)), ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p),
]
ArrowArray._fields_ = [
    ("length", ctypes.c_int64), ("null_count", ctypes.c_int64), ("offset", ctypes.c_int64), ("n_buffers", ctypes.c_int64),
    ("n_children", ctypes.c_int64), ("buffers", ctypes.POINTER(ctypes.c_void_p)),
    ("children", ctypes.POINTER(ctypes.POINTER(ArrowArray))), ("dictionary", ctypes.POINTER(ArrowArray)),
    ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p),
]


class ColumnDesc(ctypes.Structure):
    _fields_ = [
        ("dtype", ctypes.c_int32), ("_reserved", ctypes.c_int32), ("length", ctypes.c_int64), ("null_count", ctypes.c_int64),
        ("validity", ctypes.c_void_p), ("offsets", ctypes.c_void_p), ("data", ctypes.c_void_p), ("data_bytes", ctypes.c_int64),
        ("name", ctypes.c_char_p),
    ]


# every symbol include/daft_b200.h declares (tests/test_abi_cpu.py checks the .so exports all of them)
SYMBOLS = [
    "b200_abi_version", "b200_init", "b200_free_string", "b200_stream", "b200_synchronize", "b200_launch_count",
    "b200_set_jit_cache_dir", "b200_precompile", "b200_jit_describe", "b200_kernel_timing", "b200_last_agg_kernel_ms", "b200_kernel_timing_report",
    "b200_batch_import", "b200_batch_export", "b200_batch_from_device", "b200_batch_describe", "b200_batch_materialize", "b200_batch_num_rows",
    "b200_batch_num_columns", "b200_batch_size_bytes", "b200_batch_free", "b200_batch_select", "b200_batch_rename",
    "b200_batch_union", "b200_batch_slice",
    "b200_parquet_read", "b200_parquet_file_info", "b200_merge_join_sorted", "b200_asof_join_sorted", "b200_cross_join", "b200_window_ranks", "b200_window_offset", "b200_window_frame",
    "b200_hash_rows", "b200_partition_by_hash", "b200_partition_by_random", "b200_exchange_random", "b200_exchange_by_range", "b200_search_sorted", "b200_partition_by_range", "b200_batch_dtypes", "b200_filter", "b200_eval_expression_list", "b200_agg", "b200_dedup", "b200_groupby_create", "b200_groupby_push", "b200_groupby_finalize", "b200_groupby_free",
    "b200_take", "b200_concat", "b200_argsort", "b200_sort", "b200_top_n",
    "b200_probe_table_create", "b200_probe_table_add", "b200_probe_table_build", "b200_probe_table_num_rows",
    "b200_probe_table_free", "b200_probe_exists", "b200_probe_indices", "b200_join_state_create", "b200_join_state_free",
    "b200_hash_join_probe", "b200_hash_join_finalize",
    "b200_comm_unique_id", "b200_comm_init", "b200_comm_free", "b200_exchange_by_hash", "b200_all_gather", "b200_agg_allreduce",
    "b200_tpch_generate", "b200_tpch_num_rows", "b200_tpch_generate_host",
]


def build(force: bool = False) -> str:
    """Compile libdaftb200.so in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    if force:
        subprocess.check_call(["make", "-s", "-C", CSRC, "clean"])
    subprocess.check_call(["make", "-s", "-j8", "-C", CSRC])
    return LIB_PATH


_lib = None
_initialised = False


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(or make -C daft_b200/csrc).  daft_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        L.b200_batch_num_rows.restype = ctypes.c_int64
        L.b200_batch_size_bytes.restype = ctypes.c_int64
        L.b200_probe_table_num_rows.restype = ctypes.c_int64
        L.b200_launch_count.restype = ctypes.c_int64
        L.b200_tpch_num_rows.restype = ctypes.c_int64
        L.b200_tpch_num_rows.argtypes = [ctypes.c_int32, ctypes.c_double]
        L.b200_stream.restype = ctypes.c_void_p
        L.b200_last_agg_kernel_ms.restype = ctypes.c_double
        L.b200_free_string.argtypes = [ctypes.c_void_p]
        L.b200_batch_free.argtypes = [ctypes.c_void_p]
        L.b200_probe_table_free.argtypes = [ctypes.c_void_p]
        L.b200_join_state_free.argtypes = [ctypes.c_void_p]
        L.b200_comm_free.argtypes = [ctypes.c_void_p]
        L.b200_groupby_free.argtypes = [ctypes.c_void_p]
        L.b200_batch_num_rows.argtypes = [ctypes.c_void_p]
        L.b200_batch_num_columns.argtypes = [ctypes.c_void_p]
        L.b200_batch_size_bytes.argtypes = [ctypes.c_void_p]
        L.b200_probe_table_num_rows.argtypes = [ctypes.c_void_p]
        _lib = L
    return _lib


def check(rc: int, err: ctypes.c_void_p) -> None:
    if rc != 0:
        msg = "unknown error"
        if err.value:
            msg = ctypes.string_at(err.value).decode(errors="replace")
            lib().b200_free_string(err)
        raise B200Error(msg)


def call(fn_name: str, *args) -> None:
    """Call an `int f(..., char** errmsg)` entry point, raising B200Error with the library's message."""
    err = ctypes.c_void_p()
    rc = getattr(lib(), fn_name)(*args, ctypes.byref(err))
    check(rc, err)


def init(device: int | None = None) -> None:
    """Select the CUDA device (LOCAL_RANK by default).  Raises when no GPU is present."""
    global _initialised
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    call("b200_init", ctypes.c_int(device))
    _initialised = True


def ensure_init() -> None:
    if not _initialised:
        init()


def launch_count() -> int:
    return int(lib().b200_launch_count())


def stream_ptr() -> int:
    ensure_init()
    return int(lib().b200_stream())


def synchronize() -> None:
    call("b200_synchronize")


def kernel_timing_report() -> dict:
    """{"kernel": {"launches": n, "ms": total}} since the last report (enable with lib().b200_kernel_timing(1))."""
    import json
    out = ctypes.c_void_p()
    call("b200_kernel_timing_report", ctypes.byref(out))
    try:
        return json.loads(ctypes.string_at(out.value).decode())
    finally:
        lib().b200_free_string(out)

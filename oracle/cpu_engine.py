"""Python face of the morsel-parallel CPU engine (oracle/cpu_engine.cpp).

TEST INFRASTRUCTURE / CPU BASELINE ONLY: used by bench.py (`cpu_baseline`, `--impl reference`) and tests/.  It loads
oracle/libdaft_cpu_engine.so and nothing from daft_b200/: the reference arm of the benchmark runs without the product
library in the process.
"""
from __future__ import annotations

import ctypes
import datetime
import os
import subprocess
from typing import Sequence

import numpy as np
import pyarrow as pa

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdaft_cpu_engine.so")
TABLES = {"lineitem": 0, "orders": 1, "customer": 2, "supplier": 3, "nation": 4, "region": 5}
DEFAULT_SEED = 20260921
Q1_COLUMNS = ["L_QUANTITY", "L_EXTENDEDPRICE", "L_DISCOUNT", "L_TAX", "L_RETURNFLAG", "L_LINESTATUS", "L_SHIPDATE"]
Q5_COLUMNS = {
    "lineitem": ["L_ORDERKEY", "L_SUPPKEY", "L_EXTENDEDPRICE", "L_DISCOUNT"],
    "orders": ["O_ORDERKEY", "O_CUSTKEY", "O_ORDERDATE"],
    "customer": ["C_CUSTKEY", "C_NATIONKEY"],
    "supplier": ["S_SUPPKEY", "S_NATIONKEY"],
}
_EPOCH = datetime.date(1970, 1, 1)
_DT = {5: np.int64, 11: np.float64, 12: np.int32}


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, "cpu_engine.cpp"), os.path.join(_HERE, "daft_oracle.c"),
            os.path.join(_HERE, "..", "daft_b200", "csrc", "tpch_gen_core.h")]
    stale = not os.path.exists(_SO) or any(os.path.exists(s) and os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libdaft_cpu_engine.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        L.cpu_tpch_num_rows.restype = ctypes.c_int64
        L.cpu_tpch_num_rows.argtypes = [ctypes.c_int, ctypes.c_double]
        L.cpu_tpch_name.restype = ctypes.c_char_p
        L.cpu_tpch_name.argtypes = [ctypes.c_int, ctypes.c_int64]
        L.cpu_q1.restype = ctypes.c_int64
        L.cpu_q5.restype = ctypes.c_int64
        _lib = L
    return _lib


def num_rows(table: str, sf: float) -> int:
    return int(lib().cpu_tpch_num_rows(TABLES[table], sf))


def _ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def generate_columns(table: str, sf: float, columns: Sequence[str], row_begin: int = 0, row_end: int = -1, seed: int = DEFAULT_SEED,
                     threads: int = 0) -> dict[str, tuple]:
    """numpy buffers per column: fixed-width -> (values,), flag strings -> (bytes uint8[n], offsets int64[n + 1])."""
    total = num_rows(table, sf)
    if row_end < 0:
        row_end = total
    n = row_end - row_begin
    out = {}
    for c in columns:
        if c in ("N_NAME", "R_NAME"):
            names = [lib().cpu_tpch_name(1 if c == "R_NAME" else 0, r).decode() for r in range(row_begin, row_end)]
            out[c] = ("names", names)
            continue
        if c in ("L_RETURNFLAG", "L_LINESTATUS"):
            data, offs = np.empty(n + 8, dtype=np.uint8), np.empty(n + 1, dtype=np.int64)
            dt = lib().cpu_tpch_fill(c.encode(), ctypes.c_double(sf), ctypes.c_int64(row_begin), ctypes.c_int64(row_end), ctypes.c_uint64(seed),
                                     _ptr(data), _ptr(offs), threads)
            assert dt == 14, (c, dt)
            out[c] = (data[:n], offs)
            continue
        # dtype is known only after the call for unknown columns: probe with a one-row call is not needed, the name decides
        np_dt = np.int32 if c.endswith("DATE") else (np.float64 if c in ("L_EXTENDEDPRICE", "L_DISCOUNT", "L_TAX", "O_TOTALPRICE", "C_ACCTBAL", "S_ACCTBAL")
                                                    else np.int64)
        data = np.empty(max(n, 1), dtype=np_dt)
        dt = lib().cpu_tpch_fill(c.encode(), ctypes.c_double(sf), ctypes.c_int64(row_begin), ctypes.c_int64(row_end), ctypes.c_uint64(seed),
                                 _ptr(data), None, threads)
        if dt < 0:
            raise KeyError(f"tpch: unknown column {c}")
        assert _DT[dt] == np_dt, (c, dt)
        out[c] = (data[:n],)
    return out


def generate_table(table: str, sf: float, columns: Sequence[str], row_begin: int = 0, row_end: int = -1, seed: int = DEFAULT_SEED) -> pa.Table:
    """The same rows as daft_b200.tpch.generate / generate_host, as a pyarrow Table (zero-copy over the numpy buffers)."""
    cols = generate_columns(table, sf, columns, row_begin, row_end, seed)
    arrays = []
    for c in columns:
        v = cols[c]
        if isinstance(v[0], str):
            arrays.append(pa.array(v[1], type=pa.large_utf8()))
        elif len(v) == 2:
            n = len(v[1]) - 1
            arrays.append(pa.Array.from_buffers(pa.large_utf8(), n, [None, pa.py_buffer(v[1]), pa.py_buffer(v[0])]))
        elif v[0].dtype == np.int32:
            arrays.append(pa.Array.from_buffers(pa.date32(), len(v[0]), [None, pa.py_buffer(v[0])]))
        else:
            arrays.append(pa.array(v[0]))
    return pa.table(arrays, names=list(columns))


def _days(d: datetime.date) -> int:
    return (d - _EPOCH).days


def usable_threads() -> int:
    """Workers a query uses by default: CPUs of the affinity mask, capped by the cgroup CPU quota (cpu_engine.cpp usable_threads)."""
    return int(lib().cpu_usable_threads())


def q1(cols: dict[str, tuple], threads: int = 0) -> tuple[pa.Table, int]:
    """TPC-H Q1 over numpy column buffers (generate_columns('lineitem', sf, Q1_COLUMNS)); returns (answer, morsel tasks)."""
    qty, price, disc, tax = cols["L_QUANTITY"][0], cols["L_EXTENDEDPRICE"][0], cols["L_DISCOUNT"][0], cols["L_TAX"][0]
    rf, rf_off = cols["L_RETURNFLAG"]
    ls, ls_off = cols["L_LINESTATUS"]
    ship = cols["L_SHIPDATE"][0]
    n, cap = len(qty), 64
    o_rf, o_ls = np.zeros(cap, dtype=np.uint8), np.zeros(cap, dtype=np.uint8)
    s_qty = np.zeros(cap, dtype=np.int64)
    f = [np.zeros(cap, dtype=np.float64) for _ in range(6)]
    cnt = np.zeros(cap, dtype=np.uint64)
    tasks = ctypes.c_int64()
    ng = lib().cpu_q1(ctypes.c_int64(n), _ptr(qty), _ptr(price), _ptr(disc), _ptr(tax), _ptr(rf_off), _ptr(rf), _ptr(ls_off), _ptr(ls), _ptr(ship),
                      ctypes.c_int32(_days(datetime.date(1998, 9, 2))), ctypes.c_int(threads), ctypes.c_int64(cap), _ptr(o_rf), _ptr(o_ls), _ptr(s_qty),
                      _ptr(f[0]), _ptr(f[1]), _ptr(f[2]), _ptr(f[3]), _ptr(f[4]), _ptr(f[5]), _ptr(cnt), ctypes.byref(tasks))
    assert 0 <= ng <= cap
    tbl = pa.table({
        "L_RETURNFLAG": pa.array([chr(x) for x in o_rf[:ng]], type=pa.large_utf8()),
        "L_LINESTATUS": pa.array([chr(x) for x in o_ls[:ng]], type=pa.large_utf8()),
        "sum_qty": pa.array(s_qty[:ng]), "sum_base_price": pa.array(f[0][:ng]), "sum_disc_price": pa.array(f[1][:ng]),
        "sum_charge": pa.array(f[2][:ng]), "avg_qty": pa.array(f[3][:ng]), "avg_price": pa.array(f[4][:ng]), "avg_disc": pa.array(f[5][:ng]),
        "count_order": pa.array(cnt[:ng]),
    })
    return tbl, int(tasks.value)


def q5(t: dict[str, dict[str, tuple]], threads: int = 0) -> tuple[pa.Table, int, int]:
    """TPC-H Q5 over numpy column buffers {table: generate_columns(table, sf, Q5_COLUMNS[table])};
    returns (answer sorted by revenue descending, morsel tasks, probe rows over the three big joins)."""
    li, od, cu, su = t["lineitem"], t["orders"], t["customer"], t["supplier"]
    cap = 32
    names = np.zeros(cap * 32, dtype=np.uint8)
    rev = np.zeros(cap, dtype=np.float64)
    tasks, probes = ctypes.c_int64(), ctypes.c_int64()
    ng = lib().cpu_q5(
        ctypes.c_int64(len(li["L_ORDERKEY"][0])), _ptr(li["L_ORDERKEY"][0]), _ptr(li["L_SUPPKEY"][0]), _ptr(li["L_EXTENDEDPRICE"][0]),
        _ptr(li["L_DISCOUNT"][0]),
        ctypes.c_int64(len(od["O_ORDERKEY"][0])), _ptr(od["O_ORDERKEY"][0]), _ptr(od["O_CUSTKEY"][0]), _ptr(od["O_ORDERDATE"][0]),
        ctypes.c_int64(len(cu["C_CUSTKEY"][0])), _ptr(cu["C_CUSTKEY"][0]), _ptr(cu["C_NATIONKEY"][0]),
        ctypes.c_int64(len(su["S_SUPPKEY"][0])), _ptr(su["S_SUPPKEY"][0]), _ptr(su["S_NATIONKEY"][0]),
        ctypes.c_int32(_days(datetime.date(1994, 1, 1))), ctypes.c_int32(_days(datetime.date(1995, 1, 1))), ctypes.c_int(threads),
        ctypes.c_int64(cap), _ptr(names), _ptr(rev), ctypes.byref(tasks), ctypes.byref(probes))
    assert 0 <= ng <= cap
    nm = [bytes(names[g * 32:(g + 1) * 32]).split(b"\0")[0].decode() for g in range(ng)]
    return pa.table({"N_NAME": pa.array(nm, type=pa.large_utf8()), "revenue": pa.array(rev[:ng])}), int(tasks.value), int(probes.value)

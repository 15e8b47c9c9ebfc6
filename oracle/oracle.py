"""ctypes binding of oracle/libbk_oracle.so — the CPU restatement of the reference row engine.

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of bench.py.  Nothing under baikaldb_b200/ may
import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Sequence

import numpy as np

from baikaldb_b200.column import Column, unpack_validity
from baikaldb_b200.plan import PrimitiveType, storage_dtype

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbk_oracle.so")


class _BkoColumn(ctypes.Structure):
    _fields_ = [("tuple_id", ctypes.c_int32), ("slot_id", ctypes.c_int32), ("prim_type", ctypes.c_int32),
                ("elem_size", ctypes.c_int32), ("values", ctypes.c_void_p), ("validity", ctypes.c_void_p),
                ("length", ctypes.c_int64)]


class _BkoResult(ctypes.Structure):
    _fields_ = [("ncols", ctypes.c_int32), ("nrows", ctypes.c_int64), ("cols", ctypes.POINTER(_BkoColumn)),
                ("rows_scanned", ctypes.c_int64), ("rows_filtered", ctypes.c_int64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "bk_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libbk_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.bko_execute.restype = ctypes.c_int
        L.bko_execute.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(_BkoColumn), ctypes.c_int,
                                  ctypes.POINTER(ctypes.POINTER(_BkoResult)), ctypes.c_char_p, ctypes.c_size_t]
        L.bko_free_result.argtypes = [ctypes.POINTER(_BkoResult)]
        L.bko_ev_compare.restype = ctypes.c_int64
        L.bko_ev_compare.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
        L.bko_ev_cast.restype = ctypes.c_uint64
        L.bko_ev_cast.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int]
        L.bko_key_encode.restype = ctypes.c_int
        L.bko_key_encode.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p]
        _lib = L
    return _lib


class OracleResult:
    def __init__(self, columns: List[Column], rows_scanned: int, rows_filtered: int):
        self.columns = columns
        self.rows_scanned = rows_scanned
        self.rows_filtered = rows_filtered

    @property
    def nrows(self) -> int:
        return len(self.columns[0]) if self.columns else 0


def execute(plan_bytes: bytes, columns: Sequence[Column]) -> OracleResult:
    """Run the plan over ``columns`` with the reference row engine's semantics."""
    L = lib()
    n = len(columns)
    arr = (_BkoColumn * max(n, 1))()
    keep = []
    for i, c in enumerate(columns):
        vals = np.ascontiguousarray(c.values)
        bitmap = c.validity_bitmap()
        keep += [vals, bitmap]
        arr[i].tuple_id, arr[i].slot_id, arr[i].prim_type = c.tuple_id, c.slot_id, c.prim_type
        arr[i].elem_size = 16 if c.prim_type == PrimitiveType.STRING else vals.dtype.itemsize
        arr[i].values = vals.ctypes.data
        arr[i].validity = bitmap.ctypes.data if bitmap is not None else None
        arr[i].length = len(c)
    out = ctypes.POINTER(_BkoResult)()
    err = ctypes.create_string_buffer(512)
    rc = L.bko_execute(plan_bytes, len(plan_bytes), arr, n, ctypes.byref(out), err, 512)
    if rc != 0:
        raise RuntimeError(f"oracle: rc={rc}: {err.value.decode(errors='replace')}")
    res = out.contents
    cols: List[Column] = []
    nrows = res.nrows
    for i in range(res.ncols):
        oc = res.cols[i]
        if oc.prim_type == PrimitiveType.STRING:
            raw = np.ctypeslib.as_array(ctypes.cast(oc.values, ctypes.POINTER(ctypes.c_uint8)), shape=(max(nrows, 1) * 16,))
            vals = raw[: nrows * 16].reshape(nrows, 16).copy()
        else:
            dt = np.dtype(storage_dtype(oc.prim_type))
            raw = np.ctypeslib.as_array(ctypes.cast(oc.values, ctypes.POINTER(ctypes.c_uint8)),
                                        shape=(max(nrows, 1) * dt.itemsize,))
            vals = raw[: nrows * dt.itemsize].copy().view(dt)
        bm = np.ctypeslib.as_array(ctypes.cast(oc.validity, ctypes.POINTER(ctypes.c_uint8)), shape=((nrows + 7) // 8 + 1,))
        valid = unpack_validity(bm.copy(), nrows)
        cols.append(Column(oc.tuple_id, oc.slot_id, oc.prim_type, vals, None if valid.all() else valid))
    r = OracleResult(cols, res.rows_scanned, res.rows_filtered)
    L.bko_free_result(out)
    return r

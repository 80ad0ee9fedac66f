"""ctypes wrapper around oracle/_build/liboracle.so — the CPU restatement of the reference hot path.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  Nothing under galaxysql_b200/ imports this module.

Columns are ``(numpy array, nulls)`` pairs, ``nulls`` = None or a uint8/bool array (1 = NULL), the same
convention as the reference's ``Block`` (values array + ``boolean[] isNull``, EX/chunk/AbstractBlock.java:27-47).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

T_INT32, T_INT64, T_FP64, T_DEC128 = 0, 1, 2, 3
JOIN_INNER, JOIN_LEFT, JOIN_RIGHT, JOIN_SEMI, JOIN_ANTI = 0, 1, 2, 3, 4
AGG_COUNT_STAR, AGG_COUNT, AGG_SUM, AGG_AVG, AGG_MIN, AGG_MAX, AGG_SUM0 = range(7)
OK, ERR_MORE_THAN_ONE_ROW, ERR_UNSUPPORTED = 0, 1, 2
MAX_KEYS = 8

_NP_OF = {T_INT32: np.int32, T_INT64: np.int64, T_FP64: np.float64}
_T_OF = {np.dtype(np.int32): T_INT32, np.dtype(np.int64): T_INT64, np.dtype(np.float64): T_FP64}

Col = Tuple[np.ndarray, Optional[np.ndarray]]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (a few hundred ms).  Returns the .so path."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(
        os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "oracle.h"))
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _SO


class _OrcCol(C.Structure):
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("data", C.c_void_p), ("nulls", C.c_void_p)]


class _JoinSpec(C.Structure):
    _fields_ = [
        ("join_type", C.c_int32), ("max_one_row", C.c_int32), ("build_outer", C.c_int32), ("nkeys", C.c_int32),
        ("outer_key", C.c_int32 * MAX_KEYS), ("inner_key", C.c_int32 * MAX_KEYS), ("key_type", C.c_int32 * MAX_KEYS),
        ("n_anti_operands", C.c_int32), ("anti_operands", C.c_int32 * MAX_KEYS),
        ("n_cond", C.c_int32), ("cond_col", C.c_int32 * 4), ("cond_ne_value", C.c_int64 * 4),
    ]


class _AggCall(C.Structure):
    _fields_ = [("kind", C.c_int32), ("ncols", C.c_int32), ("cols", C.c_int32 * 4), ("filter_arg", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_mix.restype = C.c_int32
        _lib.orc_mix.argtypes = [C.c_int32]
        _lib.orc_murmur_hash3.restype = C.c_int32
        _lib.orc_murmur_hash3.argtypes = [C.c_int32]
        _lib.orc_array_size.restype = C.c_int32
        _lib.orc_array_size.argtypes = [C.c_int32, C.c_float]
        _lib.orc_max_fill.restype = C.c_int32
        _lib.orc_max_fill.argtypes = [C.c_int32, C.c_float]
        _lib.orc_partition.restype = C.c_int32
        _lib.orc_partition.argtypes = [C.c_int32, C.c_int32]
        _lib.orc_result_rows.restype = C.c_int64
        _lib.orc_result_rows.argtypes = [C.c_void_p]
        _lib.orc_result_ncols.restype = C.c_int32
        _lib.orc_result_ncols.argtypes = [C.c_void_p]
        _lib.orc_result_col.restype = C.c_int32
        _lib.orc_result_col.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _lib.orc_result_free.argtypes = [C.c_void_p]
    return _lib


class _Cols:
    """Keeps numpy buffers alive while C reads them."""

    def __init__(self, cols: Sequence[Col]):
        self.keep = []
        n = len(cols)
        self.arr = (_OrcCol * max(n, 1))()
        self.rows = 0
        for i, (data, nulls) in enumerate(cols):
            data = np.ascontiguousarray(data)
            if data.dtype not in _T_OF:
                raise TypeError(f"unsupported dtype {data.dtype}")
            self.keep.append(data)
            self.arr[i].type = _T_OF[data.dtype]
            self.arr[i].data = data.ctypes.data if data.size else 0
            if nulls is not None:
                nb = np.ascontiguousarray(np.asarray(nulls).astype(np.uint8))
                assert nb.shape == data.shape
                self.keep.append(nb)
                self.arr[i].nulls = nb.ctypes.data if nb.size else 0
            else:
                self.arr[i].nulls = 0
            self.rows = data.shape[0]
        self.n = n


def _take_result(res_ptr) -> List[Col]:
    L = lib()
    rows = L.orc_result_rows(res_ptr)
    out: List[Col] = []
    for i in range(L.orc_result_ncols(res_ptr)):
        d, nl = C.c_void_p(), C.c_void_p()
        t = L.orc_result_col(res_ptr, i, C.byref(d), C.byref(nl))
        if t == T_DEC128:
            raw = np.ctypeslib.as_array(C.cast(d, C.POINTER(C.c_uint64)), shape=(max(rows, 1) * 2,))[: rows * 2].copy()
            lo = raw[0::2].astype(object)
            hi = raw[1::2].astype(np.int64).astype(object)
            vals = np.array([int(h) * (1 << 64) + int(l) for h, l in zip(hi, lo)], dtype=object)
        else:
            ct = {T_INT32: C.c_int32, T_INT64: C.c_int64, T_FP64: C.c_double}[t]
            vals = np.ctypeslib.as_array(C.cast(d, C.POINTER(ct)), shape=(max(rows, 1),))[:rows].copy()
        nulls = np.ctypeslib.as_array(C.cast(nl, C.POINTER(C.c_uint8)), shape=(max(rows, 1),))[:rows].copy()
        out.append((vals, nulls.astype(bool)))
    L.orc_result_free(res_ptr)
    return out


# ------------------------------------------------------------------------------------------------- scalar / vector
def mix(x: int) -> int:
    return lib().orc_mix(C.c_int32(np.int32(x)))


def murmur_hash3(x: int) -> int:
    return lib().orc_murmur_hash3(C.c_int32(np.int32(x)))


def array_size(expected: int, f: float) -> int:
    return lib().orc_array_size(expected, f)


def max_fill(n: int, f: float) -> int:
    return lib().orc_max_fill(n, f)


def partition(h: int, nparts: int) -> int:
    return lib().orc_partition(C.c_int32(np.int32(h)), nparts)


def hash_rows(keycols: Sequence[Col], unified_types: Optional[Sequence[int]] = None) -> np.ndarray:
    cc = _Cols(keycols)
    types = list(unified_types) if unified_types is not None else [cc.arr[i].type for i in range(cc.n)]
    out = np.empty(cc.rows, dtype=np.int32)
    lib().orc_hash_rows(cc.arr, C.c_int32(cc.n), (C.c_int32 * max(cc.n, 1))(*types), C.c_int64(cc.rows),
                        C.c_void_p(out.ctypes.data if out.size else 0))
    return out


def partition_ids(hashes: np.ndarray, nparts: int) -> np.ndarray:
    hashes = np.ascontiguousarray(hashes, dtype=np.int32)
    out = np.empty_like(hashes)
    lib().orc_partition_ids(C.c_void_p(hashes.ctypes.data if hashes.size else 0), C.c_int64(hashes.size),
                            C.c_int32(nparts), C.c_void_p(out.ctypes.data if out.size else 0))
    return out


# ------------------------------------------------------------------------------------------------- operators
@dataclass
class JoinSpec:
    join_type: int = JOIN_INNER
    outer_keys: Sequence[int] = (0,)
    inner_keys: Sequence[int] = (0,)
    key_types: Sequence[int] = (T_INT32,)
    max_one_row: bool = False
    build_outer: bool = False
    anti_operands: Optional[Sequence[int]] = None
    cond_ne: Sequence[Tuple[int, int]] = field(default_factory=tuple)  # (join-row column, value): col != value

    def to_c(self) -> _JoinSpec:
        s = _JoinSpec()
        s.join_type = self.join_type
        s.max_one_row = int(self.max_one_row)
        s.build_outer = int(self.build_outer)
        s.nkeys = len(self.outer_keys)
        for i, (o, n, t) in enumerate(zip(self.outer_keys, self.inner_keys, self.key_types)):
            s.outer_key[i], s.inner_key[i], s.key_type[i] = o, n, t
        ops = list(self.anti_operands or [])
        s.n_anti_operands = len(ops)
        for i, o in enumerate(ops):
            s.anti_operands[i] = o
        s.n_cond = len(self.cond_ne)
        for i, (c, v) in enumerate(self.cond_ne):
            s.cond_col[i], s.cond_ne_value[i] = c, v
        return s


class MoreThanOneRow(RuntimeError):
    """ErrorCode.ERR_SCALAR_SUBQUERY_RETURN_MORE_THAN_ONE_ROW (AbstractBufferedJoinExec.java:217-219)."""


def hash_join(spec: JoinSpec, outer: Sequence[Col], inner: Sequence[Col]) -> List[Col]:
    co, ci = _Cols(outer), _Cols(inner)
    res = C.c_void_p()
    cs = spec.to_c()
    rc = lib().orc_hash_join(C.byref(cs), co.arr, C.c_int32(co.n), C.c_int64(co.rows), ci.arr, C.c_int32(ci.n),
                             C.c_int64(ci.rows), C.byref(res))
    if rc == ERR_MORE_THAN_ONE_ROW:
        lib().orc_result_free(res)
        raise MoreThanOneRow()
    if rc != OK:
        if res:
            lib().orc_result_free(res)
        raise NotImplementedError(f"oracle join rc={rc}")
    return _take_result(res)


@dataclass
class AggCall:
    kind: int
    cols: Sequence[int] = ()
    filter_arg: int = -1

    def to_c(self) -> _AggCall:
        a = _AggCall()
        a.kind = self.kind
        a.ncols = len(self.cols)
        for i, c in enumerate(self.cols):
            a.cols[i] = c
        a.filter_arg = self.filter_arg
        return a


def hash_agg(cols: Sequence[Col], groups: Sequence[int], aggs: Sequence[AggCall], expected_groups: int = 1024,
             chunk_size: int = 1000) -> List[Col]:
    cc = _Cols(cols)
    ca = (_AggCall * max(len(aggs), 1))(*[a.to_c() for a in aggs])
    res = C.c_void_p()
    rc = lib().orc_hash_agg(cc.arr, C.c_int32(cc.n), C.c_int64(cc.rows), (C.c_int32 * max(len(groups), 1))(*groups),
                            C.c_int32(len(groups)), ca, C.c_int32(len(aggs)), C.c_int32(expected_groups),
                            C.c_int32(chunk_size), C.byref(res))
    if rc != OK:
        raise NotImplementedError(f"oracle agg rc={rc}")
    return _take_result(res)


def partition_exchange(cols: Sequence[Col], channels: Sequence[int], nparts: int,
                       key_types: Optional[Sequence[int]] = None) -> Tuple[List[Col], np.ndarray]:
    cc = _Cols(cols)
    kt = list(key_types) if key_types is not None else [cc.arr[c].type for c in channels]
    counts = np.zeros(nparts, dtype=np.int64)
    res = C.c_void_p()
    rc = lib().orc_partition_exchange(cc.arr, C.c_int32(cc.n), C.c_int64(cc.rows),
                                      (C.c_int32 * max(len(channels), 1))(*channels), C.c_int32(len(channels)),
                                      (C.c_int32 * max(len(kt), 1))(*kt), C.c_int32(nparts), C.byref(res),
                                      C.c_void_p(counts.ctypes.data))
    if rc != OK:
        raise NotImplementedError(f"oracle exchange rc={rc}")
    return _take_result(res), counts


def chunk_row_open_hash_map(build: Sequence[Col], probe: Sequence[Col]) -> Tuple[np.ndarray, np.ndarray]:
    cb, cp = _Cols(build), _Cols(probe)
    put = np.empty(cb.rows, dtype=np.int32)
    get = np.empty(cp.rows, dtype=np.int32)
    lib().orc_chunk_row_open_hash_map(cb.arr, C.c_int32(cb.n), C.c_int32(cb.rows), cp.arr, C.c_int32(cp.rows),
                                      C.c_void_p(put.ctypes.data), C.c_void_p(get.ctypes.data))
    return put, get


# ------------------------------------------------------------------------------------------------- CPU baseline
def mt_join(spec: JoinSpec, outer: Sequence[Col], inner: Sequence[Col], nthreads: int, chunk_rows: int = 1000):
    """Reference-shaped parallel INNER join.  Returns dict(build_s, probe_s, out_rows, checksum)."""
    co, ci = _Cols(outer), _Cols(inner)
    cs = spec.to_c()
    b, p = C.c_double(), C.c_double()
    n, ck = C.c_int64(), C.c_uint64()
    rc = lib().orc_mt_join(C.byref(cs), co.arr, C.c_int32(co.n), C.c_int64(co.rows), ci.arr, C.c_int32(ci.n),
                           C.c_int64(ci.rows), C.c_int32(nthreads), C.c_int32(chunk_rows), C.byref(b), C.byref(p),
                           C.byref(n), C.byref(ck))
    if rc != OK:
        raise NotImplementedError(f"oracle mt_join rc={rc}")
    return {"build_s": b.value, "probe_s": p.value, "out_rows": n.value, "checksum": ck.value}


def mt_hash_agg(cols: Sequence[Col], groups: Sequence[int], aggs: Sequence[AggCall], expected_groups: int,
                nthreads: int, chunk_rows: int = 1000):
    cc = _Cols(cols)
    ca = (_AggCall * max(len(aggs), 1))(*[a.to_c() for a in aggs])
    s, g = C.c_double(), C.c_int64()
    rc = lib().orc_mt_hash_agg(cc.arr, C.c_int32(cc.n), C.c_int64(cc.rows),
                               (C.c_int32 * max(len(groups), 1))(*groups), C.c_int32(len(groups)), ca,
                               C.c_int32(len(aggs)), C.c_int32(expected_groups), C.c_int32(nthreads),
                               C.c_int32(chunk_rows), C.byref(s), C.byref(g))
    if rc != OK:
        raise NotImplementedError(f"oracle mt_hash_agg rc={rc}")
    return {"seconds": s.value, "groups": g.value}

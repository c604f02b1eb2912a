"""ctypes wrapper over oracle/kvc_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.

The C port executes whole layers (norm -> full sort -> take k -> sort -> gather) with OpenMP over
(batch, head) rows.  It is driven by the descriptors of oracle/kvc_oracle.py (``select=None``), is
cross-checked against the numpy oracle in tests/test_oracle_c.py, and is what bench.py times as
the CPU baseline ("port") and as the ``--impl reference`` arm.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
import time
from typing import List, Tuple

import numpy as np

from . import kvc_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "kvc_oracle.c")
_LIB = os.path.join(_HERE, "libkvc_oracle.so")
_DT = {"f32": 0, "f16": 1, "bf16": 2}
_MODE = {"none": 0, "low": 1, "high": 2, "snapkv": 3}
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-o", _LIB, _SRC, "-lm"],
                       check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.kvc_oracle_layer.restype = ctypes.c_int
        _lib.kvc_oracle_layer.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 7 + \
            [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.kvc_oracle_norms.restype = ctypes.c_int
        _lib.kvc_oracle_norms.argtypes = [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                               ctypes.c_void_p, ctypes.c_int]
        _lib.kvc_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    """Host threads this process may use.  Launchers such as torchrun export OMP_NUM_THREADS=1, which
    would silently make the CPU baseline single-threaded: go by the CPU affinity mask instead and pass
    the count explicitly to the C port (omp_set_num_threads)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # not Linux
        return max(1, os.cpu_count() or int(lib().kvc_oracle_max_threads()))


def norms(K: np.ndarray, dtype: str, lo: int, hi: int, nthreads: int = 0) -> np.ndarray:
    K = np.ascontiguousarray(K)
    B, H, S, D = K.shape
    out = np.empty((B, H, hi - lo), dtype=np.float32)
    rc = lib().kvc_oracle_norms(_DT[dtype], B, H, S, D, K.ctypes.data, lo, hi, out.ctypes.data, nthreads)
    assert rc == 0
    return out


def run_layer(K: np.ndarray, V: np.ndarray, dtype: str, res: O.LayerResult, nthreads: int = 0,
              out: Tuple[np.ndarray, np.ndarray] = None):
    """Execute one layer's descriptor in C. Returns (K_out, V_out, rows[B,H,C] int32)."""
    assert not res.untouched and res.mode != "random"
    K = np.ascontiguousarray(K)
    V = np.ascontiguousarray(V)
    B, H, S, D = K.shape
    sink, tail = len(res.head), len(res.tail)
    assert np.array_equal(res.head, np.arange(sink)) and np.array_equal(res.tail, np.arange(S - tail, S))
    C = sink + res.k_sel + tail
    if out is None:
        out = (np.empty((B, H, C, D), dtype=K.dtype), np.empty((B, H, C, D), dtype=V.dtype))
    rows = np.empty((B, H, C), dtype=np.int32)
    lo, hi = res.region
    rc = lib().kvc_oracle_layer(_DT[dtype], B, H, S, D, K.ctypes.data, V.ctypes.data, sink, lo, hi, res.k_sel, tail,
                                _MODE[res.mode], res.pool_kernel, out[0].ctypes.data, out[1].ctypes.data,
                                rows.ctypes.data, nthreads)
    assert rc == 0, "kvc_oracle_layer failed"
    return out[0], out[1], rows


def run_method(method: str, layers, dtype: str, nthreads: int = 0, **kwargs):
    """The whole compress call on the C port. Returns (list of (K, V), descriptors, seconds in C)."""
    results = O.METHODS[method](layers, dtype, select=None, **kwargs)
    outs: List[Tuple[np.ndarray, np.ndarray]] = []
    t0 = time.perf_counter()
    for (K, V), res in zip(layers, results):
        if res.untouched:
            outs.append((K, V))
        elif res.is_view:
            n = len(res.tail)
            outs.append((K[:, :, K.shape[2] - n:], V[:, :, V.shape[2] - n:]))
        else:
            ko, vo, rows = run_layer(K, V, dtype, res, nthreads)
            res.sel = rows[..., len(res.head):len(res.head) + res.k_sel].astype(np.int64)
            outs.append((ko, vo))
    return outs, results, time.perf_counter() - t0

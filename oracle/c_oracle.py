"""ctypes binding of oracle/cpu_scan.c (libsema_oracle.so).

TEST INFRASTRUCTURE ONLY — see the header of cpu_scan.c.  Used by tests/ as the
checker at sizes NumPy is too slow for, and by bench.py as the timed CPU baseline.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsema_oracle.so")
_lib = None

METRIC_DOT = 0
METRIC_L2 = 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cpu_scan.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libsema_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        f32p, u64p, u8p, u32p = (C.POINTER(C.c_float), C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint8), C.POINTER(C.c_uint32))
        L.sema_oracle_normalize.argtypes = [f32p, C.c_uint64, C.c_uint32]
        L.sema_oracle_normalize.restype = None
        L.sema_oracle_mean_pool.argtypes = [f32p, f32p, C.c_uint64, C.c_uint32, C.c_uint32, f32p]
        L.sema_oracle_mean_pool.restype = None
        L.sema_oracle_synth.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]
        L.sema_oracle_synth.restype = None
        L.sema_oracle_scan.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint64, u8p, f32p,
                                       C.c_uint32, C.c_int, C.c_uint64, u64p, f32p]
        L.sema_oracle_scan.restype = C.c_uint32
        L.sema_oracle_scan_batch.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_uint64, u8p, f32p,
                                             C.c_uint32, C.c_uint32, C.c_int, C.c_uint64, u64p,
                                             f32p, u32p]
        L.sema_oracle_scan_batch.restype = None
        L.sema_oracle_merge.argtypes = [f32p, u64p, u32p, C.c_uint32, C.c_uint32, C.c_int, u64p, f32p]
        L.sema_oracle_merge.restype = C.c_uint32
        L.sema_oracle_threads.restype = C.c_int
        L.sema_oracle_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def threads() -> int:
    return int(lib().sema_oracle_threads())


def set_threads(n: int) -> None:
    lib().sema_oracle_set_threads(int(n))


def use_all_cores() -> int:
    """One OpenMP thread per core this process may run on (sched_getaffinity), whatever
    OMP_NUM_THREADS says — torchrun exports OMP_NUM_THREADS=1 to every rank, which would silently
    turn the timed CPU baseline into a single-threaded one.  Returns the thread count now in use."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    set_threads(max(n, 1))
    return threads()


def normalize(rows: np.ndarray) -> np.ndarray:
    x = np.array(rows, dtype=np.float32, copy=True, order="C")
    n, d = x.shape
    lib().sema_oracle_normalize(_p(x, C.c_float), n, d)
    return x


def normalize_inplace(x: np.ndarray) -> np.ndarray:
    """normalize() without the copy (the 15 GB full-size corpora of bench.py / the full-size tests)."""
    assert x.dtype == np.float32 and x.flags.c_contiguous and x.ndim == 2
    lib().sema_oracle_normalize(_p(x, C.c_float), x.shape[0], x.shape[1])
    return x


def mean_pool(tokens: np.ndarray, mask: np.ndarray) -> np.ndarray:
    tokens = np.ascontiguousarray(tokens, dtype=np.float32)
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    n, seq, hidden = tokens.shape
    out = np.empty((n, hidden), dtype=np.float32)
    lib().sema_oracle_mean_pool(_p(tokens, C.c_float), _p(mask, C.c_float), n, seq, hidden, _p(out, C.c_float))
    return out


def synth(seed: int, row0: int, n: int, d: int, out: np.ndarray | None = None) -> np.ndarray:
    if out is None:
        out = np.empty((n, d), dtype=np.float32)
    assert out.dtype == np.float32 and out.flags.c_contiguous and out.shape == (n, d)
    lib().sema_oracle_synth(_p(out, C.c_float), seed, row0, n, d)
    return out


def scan(X: np.ndarray, q: np.ndarray, k: int, metric: int = METRIC_DOT, valid=None,
         id_base: int = 0):
    X = np.ascontiguousarray(X, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = X.shape
    ids = np.zeros(max(k, 1), dtype=np.uint64)
    sc = np.zeros(max(k, 1), dtype=np.float32)
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    nf = lib().sema_oracle_scan(_p(X, C.c_float), n, d, d, _p(v, C.c_uint8), _p(q, C.c_float), k,
                                metric, id_base, _p(ids, C.c_uint64), _p(sc, C.c_float))
    return ids[:nf].copy(), sc[:nf].copy()


def scan_batch(X: np.ndarray, Q: np.ndarray, k: int, metric: int = METRIC_DOT, valid=None,
               id_base: int = 0):
    X = np.ascontiguousarray(X, dtype=np.float32)
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    n, d = X.shape
    nq = Q.shape[0]
    ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
    sc = np.zeros((nq, max(k, 1)), dtype=np.float32)
    nf = np.zeros(nq, dtype=np.uint32)
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    lib().sema_oracle_scan_batch(_p(X, C.c_float), n, d, d, _p(v, C.c_uint8), _p(Q, C.c_float), nq,
                                 k, metric, id_base, _p(ids, C.c_uint64), _p(sc, C.c_float),
                                 _p(nf, C.c_uint32))
    return ids, sc, nf


def merge(scores: np.ndarray, ids: np.ndarray, lens, k: int, metric: int = METRIC_DOT):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    G = scores.shape[0]
    oi = np.zeros(max(k, 1), dtype=np.uint64)
    os_ = np.zeros(max(k, 1), dtype=np.float32)
    nf = lib().sema_oracle_merge(_p(scores, C.c_float), _p(ids, C.c_uint64), _p(lens, C.c_uint32),
                                 G, k, metric, _p(oi, C.c_uint64), _p(os_, C.c_float))
    return oi[:nf].copy(), os_[:nf].copy()

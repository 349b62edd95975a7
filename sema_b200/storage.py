"""Python handle over include/sema_store.h — the host-side mirror of the reference's
storage boundary (StorageManager / LanceIndexer / Engine post-processing).

Class and method names follow the reference (``src/storage/mod.rs``,
``src/storage/lance_indexer.rs``, ``src/tui/engine.rs``, ``src/types/mod.rs``) so that
tests read like tests of the reference would.  All logic lives in the C++ layer
(``csrc/host/storage.cpp``) and the CUDA kernels; this file only marshals arguments.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import SemaError

SEARCH_RESULTS_LIMIT = 50  # src/tui/engine.rs:11


@dataclass
class Chunk:  # src/types/mod.rs:40-47
    id: str
    file_path: str
    start_line: int
    end_line: int
    content: str


@dataclass
class SearchResult:  # src/types/mod.rs:55-60
    chunk: Chunk
    score: float
    total_matches_in_file: int


class _Hit(C.Structure):
    _fields_ = [("row", C.c_uint64), ("score", C.c_float)]


class _Grouped(C.Structure):
    _fields_ = [("row", C.c_uint64), ("score", C.c_float), ("total", C.c_uint64)]


EMBED_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_char_p, C.POINTER(C.c_float), C.c_uint32)

_vp = C.c_void_p
_cpp = C.POINTER(C.c_char_p)
_u64p = C.POINTER(C.c_uint64)
STORE_SIGNATURES = {
    "sema_store_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "sema_store_create_multi": (C.c_int, [C.POINTER(C.c_int), C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "sema_store_destroy": (C.c_int, [_vp]),
    "sema_store_set_embedder": (C.c_int, [_vp, EMBED_FN, _vp]),
    "sema_store_index_chunks": (C.c_int, [_vp, C.c_uint64, _cpp, _cpp, _u64p, _u64p, _cpp, _vp, _vp]),
    "sema_store_index_chunks_embed": (C.c_int, [_vp, C.c_uint64, _cpp, _cpp, _u64p, _u64p, _cpp]),
    "sema_store_search_vector": (C.c_int, [_vp, _vp, C.c_uint32, C.POINTER(_Hit), C.POINTER(C.c_uint32)]),
    "sema_store_search": (C.c_int, [_vp, C.c_char_p, C.c_uint32, C.POINTER(_Hit), C.POINTER(C.c_uint32)]),
    "sema_store_execute_search": (C.c_int, [_vp, C.c_char_p, C.POINTER(_Grouped), C.c_uint32, C.POINTER(C.c_uint32)]),
    "sema_store_group_results_by_file": (C.c_int, [_vp, C.POINTER(_Hit), C.c_uint32, C.POINTER(_Grouped), C.c_uint32,
                                                   C.POINTER(C.c_uint32)]),
    "sema_store_remove_file_chunks": (C.c_int, [_vp, C.c_char_p, _u64p]),
    "sema_store_compact": (C.c_int, [_vp, _u64p]),
    "sema_store_chunk": (C.c_int, [_vp, C.c_uint64, _cpp, _cpp, _u64p, _u64p, _cpp]),
    "sema_store_len": (C.c_uint64, [_vp]),
    "sema_store_last_error": (C.c_char_p, []),
    "sema_group_results_by_file": (C.c_int, [C.c_uint32, _cpp, _u64p, C.POINTER(C.c_float), C.POINTER(C.c_uint32),
                                             _u64p, C.POINTER(C.c_uint32)]),
    "sema_like_contains": (C.c_int, [C.c_char_p, C.c_char_p]),
    "sema_store_index": (_vp, [_vp]),
}

_bound = False


def _L():
    global _bound
    L = _lib.lib()
    if not _bound:
        for name, (res, args) in STORE_SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _bound = True
    return L


def _check(rc: int) -> None:
    if rc != 0:
        raise SemaError(rc, _L().sema_store_last_error().decode("utf-8", "replace"))


def _strs(items):
    arr = (C.c_char_p * len(items))(*[s.encode("utf-8") for s in items])
    return arr


def group_results_by_file(file_paths, start_lines, scores):
    """Pure-host ``Engine::group_results_by_file`` (src/tui/engine.rs:156-182) over ranked hits
    -> list of (index of the representative hit, total_matches_in_file), best first."""
    n = len(file_paths)
    sl = np.ascontiguousarray(start_lines, dtype=np.uint64)
    sc = np.ascontiguousarray(scores, dtype=np.float32)
    rep = np.zeros(max(n, 1), dtype=np.uint32)
    tot = np.zeros(max(n, 1), dtype=np.uint64)
    ng = C.c_uint32()
    _check(_L().sema_group_results_by_file(n, _strs(file_paths), sl.ctypes.data_as(_u64p),
                                           sc.ctypes.data_as(C.POINTER(C.c_float)),
                                           rep.ctypes.data_as(C.POINTER(C.c_uint32)), tot.ctypes.data_as(_u64p),
                                           C.byref(ng)))
    return [(int(rep[i]), int(tot[i])) for i in range(ng.value)]


def like_contains(content: str, needle: str) -> bool:
    return bool(_L().sema_like_contains(content.encode("utf-8"), needle.encode("utf-8")))


class StorageManager:
    """``StorageManager`` (src/storage/mod.rs:13-132), vector route, on one GPU."""

    def __init__(self, dim: int = 384, capacity_rows: int = 1 << 20, device: int = 0, normalize: bool = True,
                 embedder=None, devices=None):
        """devices: a list of GPU ordinals spreads the table over them (one process, one handle: the rows are dealt
        out in contiguous ranges of capacity_rows / len(devices), every search is one fused multi-GPU scan)."""
        self._lib = _L()
        self._h = _vp()
        if devices is not None and len(devices) > 1:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            _check(self._lib.sema_store_create_multi(arr, len(devices), dim, capacity_rows, int(normalize), C.byref(self._h)))
        else:
            if devices:
                device = int(devices[0])
            _check(self._lib.sema_store_create(device, dim, capacity_rows, int(normalize), C.byref(self._h)))
        self.dim = dim
        self._cb = None
        if embedder is not None:
            self.set_embedder(embedder)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.sema_store_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_embedder(self, embedder) -> None:
        """embedder(text) -> sequence of dim floats, or None when the embedding fails."""
        dim = self.dim

        def _cb(_user, text, out, d):
            try:
                v = embedder(text.decode("utf-8"))
                if v is None:
                    return 1
                a = np.ascontiguousarray(v, dtype=np.float32)
                if a.shape != (dim,):
                    return 1
                C.memmove(out, a.ctypes.data, dim * 4)
                return 0
            except Exception:
                return 1

        self._cb = EMBED_FN(_cb)
        _check(self._lib.sema_store_set_embedder(self._h, self._cb, None))

    # -- LanceIndexer::index_chunks -------------------------------------------------
    def index_chunks(self, chunks, vectors=None, valid=None) -> None:
        n = len(chunks)
        if n == 0:
            return
        cols = (_strs([c.id for c in chunks]), _strs([c.file_path for c in chunks]),
                np.array([c.start_line for c in chunks], dtype=np.uint64),
                np.array([c.end_line for c in chunks], dtype=np.uint64), _strs([c.content for c in chunks]))
        if vectors is None:
            _check(self._lib.sema_store_index_chunks_embed(self._h, n, cols[0], cols[1], cols[2].ctypes.data_as(_u64p),
                                                           cols[3].ctypes.data_as(_u64p), cols[4]))
            return
        v = np.ascontiguousarray(vectors, dtype=np.float32)
        if v.shape != (n, self.dim):
            raise ValueError(f"vectors must be [{n}, {self.dim}]")
        ok = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
        _check(self._lib.sema_store_index_chunks(self._h, n, cols[0], cols[1], cols[2].ctypes.data_as(_u64p),
                                                 cols[3].ctypes.data_as(_u64p), cols[4], _vp(v.ctypes.data),
                                                 None if ok is None else _vp(ok.ctypes.data)))

    # -- searches ------------------------------------------------------------------------
    def chunk(self, row: int) -> Chunk:
        i, f, c = C.c_char_p(), C.c_char_p(), C.c_char_p()
        s, e = C.c_uint64(), C.c_uint64()
        _check(self._lib.sema_store_chunk(self._h, row, C.byref(i), C.byref(f), C.byref(s), C.byref(e), C.byref(c)))
        return Chunk(i.value.decode(), f.value.decode(), s.value, e.value, c.value.decode())

    def _hits(self, hits, n):
        return [(self.chunk(hits[i].row), float(hits[i].score)) for i in range(n)]

    def search_vector(self, query_embedding, limit: int):
        """nearest_to(query_embedding).limit(limit) -> [(Chunk, score)] best first."""
        q = np.ascontiguousarray(query_embedding, dtype=np.float32)
        hits = (_Hit * max(limit, 1))()
        nf = C.c_uint32()
        _check(self._lib.sema_store_search_vector(self._h, _vp(q.ctypes.data), limit, hits, C.byref(nf)))
        return self._hits(hits, nf.value)

    def search(self, query: str, limit: int):
        """``StorageManager::search(query, limit) -> Vec<(Chunk, f32)>`` (src/storage/mod.rs:112-125)."""
        hits = (_Hit * max(limit, 1))()
        nf = C.c_uint32()
        _check(self._lib.sema_store_search(self._h, query.encode("utf-8"), limit, hits, C.byref(nf)))
        return self._hits(hits, nf.value)

    def execute_search(self, query: str):
        """``Engine::execute_search`` (src/tui/engine.rs:102-154) -> [SearchResult] grouped by file."""
        out = (_Grouped * SEARCH_RESULTS_LIMIT)()
        n = C.c_uint32()
        _check(self._lib.sema_store_execute_search(self._h, query.encode("utf-8"), out, SEARCH_RESULTS_LIMIT, C.byref(n)))
        return [SearchResult(self.chunk(out[i].row), float(out[i].score), int(out[i].total)) for i in range(n.value)]

    def remove_file_chunks(self, file_path: str) -> int:
        removed = C.c_uint64()
        _check(self._lib.sema_store_remove_file_chunks(self._h, file_path.encode("utf-8"), C.byref(removed)))
        return removed.value

    def compact(self) -> int:
        """Drop removed chunks and renumber rows; returns the live chunk count."""
        n = C.c_uint64()
        _check(self._lib.sema_store_compact(self._h, C.byref(n)))
        return n.value

    def __len__(self) -> int:
        return int(self._lib.sema_store_len(self._h))

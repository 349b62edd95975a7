"""GpuIndex — thin Python handle over the C ABI (include/sema_b200.h).

Used by the tests, the benchmark and the host-side mirror of the reference's
storage layer.  All arithmetic happens in the CUDA kernels behind the ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import METRIC_COSINE, METRIC_L2, SemaError, check  # noqa: F401


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class GpuIndex:
    """The chunk-embedding matrix resident in one GPU's HBM.

    Reference counterpart: the ``chunks`` Lance table's ``vector`` column opened by
    ``LanceIndexer`` (``src/storage/lance_indexer.rs:19-28, 92-101``).
    """

    def __init__(self, dim: int, capacity_rows: int, device: int = 0, metric: int = METRIC_COSINE,
                 growable: bool = False):
        """growable=True: ``capacity_rows`` only reserves address space; HBM is mapped as rows arrive
        (``sema_index_create_growable``), like the reference's table, which grows without a declared size."""
        self._L = _lib.lib()
        self._h = C.c_void_p()
        create = self._L.sema_index_create_growable if growable else self._L.sema_index_create
        check(create(device, dim, capacity_rows, metric, C.byref(self._h)))
        self.dim = dim
        self.metric = metric
        self.device = device

    # -- lifecycle ------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.sema_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self) -> C.c_void_p:
        return self._h

    # -- ingest (K1) ------------------------------------------------------------
    def append(self, rows: np.ndarray, valid: np.ndarray | None = None, normalize: bool = True,
               asynchronous: bool = False) -> int:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        if rows.ndim != 2 or rows.shape[1] != self.dim:
            raise ValueError(f"rows must be [n, {self.dim}], got {rows.shape}")
        v = None
        if valid is not None:
            v = np.ascontiguousarray(valid, dtype=np.uint8)
            if v.shape != (rows.shape[0],):
                raise ValueError("valid must hold one byte per row")
        first = C.c_uint64()
        fn = self._L.sema_index_append_async if asynchronous else self._L.sema_index_append
        check(fn(self._h, _ptr(rows), rows.shape[0], _ptr(v), int(normalize), C.byref(first)))
        if asynchronous:
            self._keepalive = getattr(self, "_keepalive", []) + [rows, v]
        return first.value

    def flush(self) -> None:
        check(self._L.sema_index_flush(self._h))
        self._keepalive = []

    def append_device(self, rows_ptr: int, n: int, valid_ptr: int | None = None,
                      normalize: bool = True) -> int:
        first = C.c_uint64()
        check(self._L.sema_index_append_device(self._h, C.c_void_p(rows_ptr), n,
                                               C.c_void_p(valid_ptr) if valid_ptr else None,
                                               int(normalize), C.byref(first)))
        return first.value

    def append_synthetic(self, seed: int, row0: int, n: int, normalize: bool = True) -> int:
        first = C.c_uint64()
        check(self._L.sema_index_append_synthetic(self._h, seed, row0, n, int(normalize), C.byref(first)))
        return first.value

    # -- the step before the path: mean pooling (K0) ------------------------------------
    def mean_pool(self, tokens: np.ndarray, mask: np.ndarray, skip_masked: bool = False) -> np.ndarray:
        """``mean_pool`` of ``src/semantic/embeddings.rs:61-91`` for n texts: tokens [n, seq, dim] fp32,
        mask [n, seq] fp32 -> [n, dim] pooled + L2-normalised, bit-identical to the reference's sums."""
        tokens = np.ascontiguousarray(tokens, dtype=np.float32)
        mask = np.ascontiguousarray(mask, dtype=np.float32)
        if tokens.ndim != 3 or tokens.shape[2] != self.dim or mask.shape != tokens.shape[:2]:
            raise ValueError(f"tokens must be [n, seq, {self.dim}] and mask [n, seq]; got {tokens.shape}, {mask.shape}")
        out = np.empty((tokens.shape[0], self.dim), dtype=np.float32)
        check(self._L.sema_mean_pool(self._h, _ptr(tokens), _ptr(mask), tokens.shape[0], tokens.shape[1],
                                     int(skip_masked), _ptr(out)))
        return out

    def mean_pool_device(self, tokens_ptr: int, mask_ptr: int, n: int, seq_len: int, out_ptr: int,
                         skip_masked: bool = False) -> None:
        check(self._L.sema_mean_pool_device(self._h, C.c_void_p(tokens_ptr), C.c_void_p(mask_ptr), n, seq_len,
                                            int(skip_masked), C.c_void_p(out_ptr)))

    def append_pooled_device(self, tokens_ptr: int, mask_ptr: int, n: int, seq_len: int,
                             valid_ptr: int | None = None, skip_masked: bool = False) -> int:
        """mean_pool fused with the append (K0 writes straight into the matrix rows)."""
        first = C.c_uint64()
        check(self._L.sema_index_append_pooled_device(self._h, C.c_void_p(tokens_ptr), C.c_void_p(mask_ptr), n, seq_len,
                                                      C.c_void_p(valid_ptr) if valid_ptr else None,
                                                      int(skip_masked), C.byref(first)))
        return first.value

    def tombstone(self, rows) -> None:
        r = np.ascontiguousarray(rows, dtype=np.uint64)
        check(self._L.sema_index_tombstone(self._h, _ptr(r), r.shape[0]))

    def compact(self, out: np.ndarray | None = None) -> np.ndarray:
        """Drop null / tombstoned rows; returns new_row_of_old (uint64, 2**64-1 = dropped).  `out`: a caller-owned
        uint64 array of at least len(self) entries to receive the map (a buffer that has been written before takes no
        page faults during the copy)."""
        n = len(self)
        if out is not None and (out.dtype != np.uint64 or out.ndim != 1 or out.shape[0] < n or not out.flags.c_contiguous):
            raise ValueError("out must be a contiguous uint64 array with at least len(index) entries")
        m = out if out is not None else np.zeros(max(n, 1), dtype=np.uint64)
        live = C.c_uint64()
        check(self._L.sema_index_compact(self._h, _ptr(m), C.byref(live)))
        return m[:n]

    def compact_keep(self, keep, want_map: bool = False):
        """Keep exactly the rows whose flag is non-zero (null-vector rows included when flagged): sema_index_compact_keep.
        Returns the number of rows left, or (rows left, new_row_of_old) with want_map."""
        n = len(self)
        kf = np.ascontiguousarray(keep, dtype=np.uint8)
        if kf.shape[0] != n:
            raise ValueError(f"keep has {kf.shape[0]} flags for {n} rows")
        m = np.zeros(max(n, 1), dtype=np.uint64) if want_map else None
        live = C.c_uint64()
        check(self._L.sema_index_compact_keep(self._h, _ptr(kf) if n else None, _ptr(m) if want_map else None, C.byref(live)))
        return (int(live.value), m[:n]) if want_map else int(live.value)

    def save(self, path: str) -> None:
        check(self._L.sema_index_save(self._h, path.encode("utf-8")))

    @classmethod
    def load(cls, path: str, device: int = 0, capacity_rows: int = 0) -> "GpuIndex":
        L = _lib.lib()
        h = C.c_void_p()
        check(L.sema_index_load(path.encode("utf-8"), device, capacity_rows, C.byref(h)))
        self = cls.__new__(cls)
        self._L, self._h = L, h
        self.dim = int(L.sema_index_dim(h))
        self.device = device
        self.metric = None
        return self

    # -- search (K2 / K3) -----------------------------------------------------------
    def search(self, q: np.ndarray, k: int):
        """-> (row_ids uint64[n_found], scores float32[n_found]) best first."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.shape != (self.dim,):
            raise ValueError(f"query must be [{self.dim}], got {q.shape}")
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        nf = C.c_uint32()
        check(self._L.sema_index_search(self._h, _ptr(q), k, _ptr(ids), _ptr(sc), C.byref(nf)))
        return ids[:nf.value].copy(), sc[:nf.value].copy()

    def search_into(self, q: np.ndarray, k: int, ids: np.ndarray, sc: np.ndarray) -> int:
        """Allocation-free variant for timing loops; returns n_found."""
        nf = C.c_uint32()
        check(self._L.sema_index_search(self._h, _ptr(q), k, _ptr(ids), _ptr(sc), C.byref(nf)))
        return nf.value

    def search_ptr(self, q_ptr: C.c_void_p, k: int, ids_ptr: C.c_void_p, sc_ptr: C.c_void_p) -> int:
        """sema_index_search on pre-built ctypes pointers (host buffers): the binding adds nothing but
        the foreign call itself; returns n_found."""
        nf = C.c_uint32()
        check(self._L.sema_index_search(self._h, q_ptr, k, ids_ptr, sc_ptr, C.byref(nf)))
        return nf.value

    def submit_ptr(self, q_ptr: C.c_void_p, k: int) -> int:
        """Asynchronous search (``sema_index_search_submit``): returns a ticket once the scan is enqueued."""
        t = C.c_uint64()
        check(self._L.sema_index_search_submit(self._h, q_ptr, k, C.byref(t)))
        return t.value

    def collect_ptr(self, ticket: int, ids_ptr: C.c_void_p, sc_ptr: C.c_void_p) -> int:
        nf = C.c_uint32()
        check(self._L.sema_index_search_collect(self._h, ticket, ids_ptr, sc_ptr, C.byref(nf)))
        return nf.value

    def submit(self, q: np.ndarray, k: int) -> int:
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.shape != (self.dim,):
            raise ValueError(f"query must be [{self.dim}], got {q.shape}")
        return self.submit_ptr(_ptr(q), k)          # q is consumed before submit returns

    def collect(self, ticket: int, k: int):
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        nf = self.collect_ptr(ticket, _ptr(ids), _ptr(sc))
        return ids[:nf].copy(), sc[:nf].copy()

    def search_batch(self, Q: np.ndarray, k: int):
        """-> (row_ids uint64[nq,k], scores float32[nq,k], n_found uint32[nq])."""
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        if Q.ndim != 2 or Q.shape[1] != self.dim:
            raise ValueError(f"queries must be [nq, {self.dim}], got {Q.shape}")
        nq = Q.shape[0]
        ids = np.zeros((nq, max(k, 1)), dtype=np.uint64)
        sc = np.zeros((nq, max(k, 1)), dtype=np.float32)
        nf = np.zeros(max(nq, 1), dtype=np.uint32)
        check(self._L.sema_index_search_batch(self._h, _ptr(Q), nq, k, _ptr(ids), _ptr(sc), _ptr(nf)))
        return ids, sc, nf[:nq]

    def search_batch_device(self, q_ptr: int, nq: int, k: int, ids_ptr: int, scores_ptr: int,
                            nfound_ptr: int) -> None:
        check(self._L.sema_index_search_batch_device(self._h, C.c_void_p(q_ptr), nq, k, C.c_void_p(ids_ptr),
                                                     C.c_void_p(scores_ptr), C.c_void_p(nfound_ptr)))

    def search_stream_device(self, q_ptr: int, nq: int, k: int, ids_ptr: int, scores_ptr: int,
                             nfound_ptr: int) -> None:
        """A stream of nq independent single-query scans (K2: one persistent launch for the whole stream when a
        scan is long enough, else one launch per query chained with programmatic dependent launch; identical
        results); device pointers, results nq x k, nothing synchronises."""
        check(self._L.sema_index_search_stream_device(self._h, C.c_void_p(q_ptr), nq, k, C.c_void_p(ids_ptr),
                                                      C.c_void_p(scores_ptr), C.c_void_p(nfound_ptr)))

    def set_normalize_queries(self, on: bool) -> int:
        """Apply the reference's normalise tail (K1) to host queries before scanning."""
        return self._L.sema_index_set_normalize_queries(self._h, int(on))

    def set_batch_mode(self, mode: int) -> int:
        """0 = automatic (K3 precision cascade), 1 = K2 once per query, 2 = K3 bf16x3, 3 = K3 single pass."""
        return self._L.sema_index_set_batch_mode(self._h, mode)

    def set_batch_precision(self, prec: int) -> int:
        """0 = automatic (fp16 halves when every stored element is <= 1024 in magnitude, else bf16), 1 = bf16 always."""
        return self._L.sema_index_set_batch_precision(self._h, prec)

    @property
    def batch_precision_active(self) -> int:
        """Format the K3 planes are in: 0 bf16, 1 fp16, -1 none built yet."""
        return int(self._L.sema_index_batch_precision_active(self._h))

    def batch_stats(self) -> tuple[int, int]:
        """-> (queries served by K3, of which re-run through K2 for lack of an exactness proof)."""
        a, b = C.c_uint64(), C.c_uint64()
        check(self._L.sema_index_batch_stats(self._h, C.byref(a), C.byref(b), None))
        return a.value, b.value

    @property
    def batch_cascaded(self) -> int:
        """Queries that went from the single-pass K3 stage to the bf16x3 stage (automatic mode)."""
        c = C.c_uint64()
        check(self._L.sema_index_batch_stats(self._h, None, None, C.byref(c)))
        return c.value

    def search_keys_device(self, q_ptr: int, k: int, keys_ptr: int) -> None:
        check(self._L.sema_index_search_keys_device(self._h, C.c_void_p(q_ptr), k, C.c_void_p(keys_ptr)))

    def search_device(self, q_ptr: int, k: int, ids_ptr: int, scores_ptr: int, nfound_ptr: int) -> None:
        check(self._L.sema_index_search_device(self._h, C.c_void_p(q_ptr), k, C.c_void_p(ids_ptr),
                                               C.c_void_p(scores_ptr), C.c_void_p(nfound_ptr)))

    def merge_device(self, keys_ptr: int, n_lists: int, k: int, ids_ptr: int, scores_ptr: int,
                     nfound_ptr: int) -> None:
        check(self._L.sema_topk_merge_device(self._h, C.c_void_p(keys_ptr), n_lists, k,
                                             C.c_void_p(ids_ptr), C.c_void_p(scores_ptr),
                                             C.c_void_p(nfound_ptr)))

    def search_batch_keys_device(self, q_ptr: int, nq: int, k: int, keys_ptr: int) -> None:
        check(self._L.sema_index_search_batch_keys_device(self._h, C.c_void_p(q_ptr), nq, k, C.c_void_p(keys_ptr)))

    def merge_batch_device(self, keys_ptr: int, n_lists: int, nq: int, k: int, ids_ptr: int, scores_ptr: int,
                           nfound_ptr: int) -> None:
        check(self._L.sema_topk_merge_batch_device(self._h, C.c_void_p(keys_ptr), n_lists, nq, k,
                                                   C.c_void_p(ids_ptr), C.c_void_p(scores_ptr),
                                                   C.c_void_p(nfound_ptr)))

    # -- properties -------------------------------------------------------------------
    def set_row_base(self, base: int) -> None:
        check(self._L.sema_index_set_row_base(self._h, base))

    def set_stream(self, cuda_stream: int | None) -> None:
        """Run searches on the caller's stream (an int cudaStream_t handle; 0 = the CUDA
        default stream) or, with None, on the handle's own query stream."""
        if cuda_stream is None:
            check(self._L.sema_index_set_stream(self._h, None, 0))
        else:
            check(self._L.sema_index_set_stream(self._h, C.c_void_p(cuda_stream), 1))

    def set_scan_variant(self, variant: int) -> int:
        return self._L.sema_index_set_scan_variant(self._h, variant)

    def __len__(self) -> int:
        return int(self._L.sema_index_size(self._h))

    @property
    def visible(self) -> int:
        return int(self._L.sema_index_visible(self._h))

    @property
    def capacity(self) -> int:
        return int(self._L.sema_index_capacity(self._h))

    @property
    def last_snapshot(self) -> int:
        return int(self._L.sema_index_last_snapshot(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.sema_index_launch_count(self._h))

    def read_rows(self, first: int, n: int) -> np.ndarray:
        out = np.empty((n, self.dim), dtype=np.float32)
        check(self._L.sema_index_read_rows(self._h, first, n, _ptr(out)))
        return out


class ShardGroup:
    """Fused scan + peer exchange + merge over a row-sharded corpus (include/sema_b200.h:
    sema_shard_group_*).  One per process / GPU; `exchange_handles` is any callable that
    all-gathers a bytes object across the ranks and returns the list ordered by rank."""

    HANDLE_BYTES = 64

    def __init__(self, index: GpuIndex, world: int, rank: int, exchange_handles=None):
        self._L = _lib.lib()
        self.index, self.world, self.rank = index, world, rank
        self._g = C.c_void_p()
        check(self._L.sema_shard_group_create(index.handle, world, rank, C.byref(self._g)))
        if world > 1:
            if exchange_handles is None:
                raise ValueError("world > 1 needs an exchange_handles callable")
            mine = (C.c_ubyte * self.HANDLE_BYTES)()
            check(self._L.sema_shard_group_local_handle(self._g, mine))
            handles = exchange_handles(bytes(mine))
            blob = b"".join(handles)
            if len(blob) != world * self.HANDLE_BYTES:
                raise ValueError("handle exchange returned the wrong number of bytes")
            buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
            check(self._L.sema_shard_group_connect(self._g, buf))

    @classmethod
    def local(cls, shards) -> "ShardGroup":
        """Single-process group over per-GPU shards (``sema_shard_group_create_local``): one handle, one
        host call per search, the N fused kernels launched together by the library's worker threads."""
        self = cls.__new__(cls)
        self._L = _lib.lib()
        self.index, self.world, self.rank = shards[0], len(shards), 0
        self.shards = list(shards)                       # keep the shard handles alive
        arr = (C.c_void_p * len(shards))(*[s.handle for s in shards])
        self._g = C.c_void_p()
        check(self._L.sema_shard_group_create_local(arr, len(shards), C.byref(self._g)))
        return self

    def close(self) -> None:
        if getattr(self, "_g", None) is not None and self._g.value:
            self._L.sema_shard_group_destroy(self._g)
            self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def search(self, q: np.ndarray, k: int):
        """Every rank calls this with the same query; every rank gets the global top-k."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        ids = np.zeros(max(k, 1), dtype=np.uint64)
        sc = np.zeros(max(k, 1), dtype=np.float32)
        nf = C.c_uint32()
        check(self._L.sema_shard_group_search(self._g, _ptr(q), k, _ptr(ids), _ptr(sc), C.byref(nf)))
        return ids[:nf.value].copy(), sc[:nf.value].copy()

    def search_into(self, q: np.ndarray, k: int, ids: np.ndarray, sc: np.ndarray) -> int:
        nf = C.c_uint32()
        check(self._L.sema_shard_group_search(self._g, _ptr(q), k, _ptr(ids), _ptr(sc), C.byref(nf)))
        return nf.value

    def search_ptr(self, q_ptr: C.c_void_p, k: int, ids_ptr: C.c_void_p, sc_ptr: C.c_void_p) -> int:
        nf = C.c_uint32()
        check(self._L.sema_shard_group_search(self._g, q_ptr, k, ids_ptr, sc_ptr, C.byref(nf)))
        return nf.value

    def submit_ptr(self, q_ptr: C.c_void_p, k: int) -> int:
        t = C.c_uint64()
        check(self._L.sema_shard_group_search_submit(self._g, q_ptr, k, C.byref(t)))
        return t.value

    def collect_ptr(self, ticket: int, ids_ptr: C.c_void_p, sc_ptr: C.c_void_p) -> int:
        nf = C.c_uint32()
        check(self._L.sema_shard_group_search_collect(self._g, ticket, ids_ptr, sc_ptr, C.byref(nf)))
        return nf.value

    def search_device(self, q_ptr: int, k: int, ids_ptr: int, scores_ptr: int, nfound_ptr: int) -> None:
        check(self._L.sema_shard_group_search_device(self._g, C.c_void_p(q_ptr), k, C.c_void_p(ids_ptr),
                                                     C.c_void_p(scores_ptr), C.c_void_p(nfound_ptr)))

    def search_stream_device(self, q_ptr: int, nq: int, k: int, ids_ptr: int, scores_ptr: int,
                             nfound_ptr: int) -> None:
        """nq group searches issued back to back (every rank: the same queries in the same order)."""
        check(self._L.sema_shard_group_search_stream_device(self._g, C.c_void_p(q_ptr), nq, k, C.c_void_p(ids_ptr),
                                                            C.c_void_p(scores_ptr), C.c_void_p(nfound_ptr)))

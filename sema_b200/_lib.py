"""ctypes loader for libsema_b200.so — the C ABI declared in include/sema_b200.h.

The library is built in-tree (``sema_b200/csrc/Makefile`` -> ``sema_b200/libsema_b200.so``)
for sm_100a.  There is no fallback: if the shared object is missing this module
raises, and every compute entry point fails when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SEMA_B200_LIB: another build of the same library (scripts/k3_probe.py points it at the probe build)
SO_PATH = os.environ.get("SEMA_B200_LIB") or os.path.join(_HERE, "libsema_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sema_b200.h")

SEMA_OK = 0
SEMA_ERR_INVALID = -1
SEMA_ERR_CUDA = -2
SEMA_ERR_CAPACITY = -3
SEMA_ERR_NOMEM = -4
SEMA_ERR_UNSUPPORTED = -5
METRIC_COSINE = 0
METRIC_L2 = 1
MAX_K = 1024
MAX_DIM = 8192

_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p

# name -> (restype, argtypes); mirrors include/sema_b200.h one to one
SIGNATURES = {
    "sema_index_create": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "sema_index_create_growable": (C.c_int, [C.c_int, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(_vp)]),
    "sema_index_destroy": (C.c_int, [_vp]),
    "sema_index_append": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_int, _u64p]),
    "sema_index_append_async": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_int, _u64p]),
    "sema_index_flush": (C.c_int, [_vp]),
    "sema_index_append_device": (C.c_int, [_vp, _vp, C.c_uint64, _vp, C.c_int, _u64p]),
    "sema_index_append_synthetic": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, _u64p]),
    "sema_mean_pool": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_int, _vp]),
    "sema_mean_pool_device": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, C.c_int, _vp]),
    "sema_index_append_pooled_device": (C.c_int, [_vp, _vp, _vp, C.c_uint64, C.c_uint32, _vp, C.c_int, _u64p]),
    "sema_index_tombstone": (C.c_int, [_vp, _vp, C.c_uint64]),
    "sema_index_compact": (C.c_int, [_vp, _vp, _u64p]),
    "sema_index_compact_keep": (C.c_int, [_vp, _vp, _vp, _u64p]),
    "sema_index_save": (C.c_int, [_vp, C.c_char_p]),
    "sema_index_load": (C.c_int, [C.c_char_p, C.c_int, C.c_uint64, C.POINTER(_vp)]),
    "sema_index_search": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp, _u32p]),
    "sema_index_search_submit": (C.c_int, [_vp, _vp, C.c_uint32, _u64p]),
    "sema_index_search_collect": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _u32p]),
    "sema_index_search_batch": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "sema_index_search_batch_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "sema_index_search_stream_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "sema_index_set_normalize_queries": (C.c_int, [_vp, C.c_int]),
    "sema_index_set_batch_mode": (C.c_int, [_vp, C.c_int]),
    "sema_index_set_batch_precision": (C.c_int, [_vp, C.c_int]),
    "sema_index_batch_precision_active": (C.c_int, [_vp]),
    "sema_index_batch_stats": (C.c_int, [_vp, _u64p, _u64p, _u64p]),
    "sema_index_search_keys_device": (C.c_int, [_vp, _vp, C.c_uint32, _vp]),
    "sema_index_search_device": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp, _vp]),
    "sema_topk_merge_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "sema_index_search_batch_keys_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp]),
    "sema_topk_merge_batch_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "sema_shard_group_create": (C.c_int, [_vp, C.c_uint32, C.c_uint32, C.POINTER(_vp)]),
    "sema_shard_group_create_local": (C.c_int, [C.POINTER(_vp), C.c_uint32, C.POINTER(_vp)]),
    "sema_shard_group_local_handle": (C.c_int, [_vp, _vp]),
    "sema_shard_group_connect": (C.c_int, [_vp, _vp]),
    "sema_shard_group_search": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp, _u32p]),
    "sema_shard_group_search_submit": (C.c_int, [_vp, _vp, C.c_uint32, _u64p]),
    "sema_shard_group_search_collect": (C.c_int, [_vp, C.c_uint64, _vp, _vp, _u32p]),
    "sema_shard_group_search_device": (C.c_int, [_vp, _vp, C.c_uint32, _vp, _vp, _vp]),
    "sema_shard_group_search_stream_device": (C.c_int, [_vp, _vp, C.c_uint32, C.c_uint32, _vp, _vp, _vp]),
    "sema_shard_group_destroy": (C.c_int, [_vp]),
    "sema_index_set_row_base": (C.c_int, [_vp, C.c_uint64]),
    "sema_index_set_stream": (C.c_int, [_vp, _vp, C.c_int]),
    "sema_index_size": (C.c_uint64, [_vp]),
    "sema_index_visible": (C.c_uint64, [_vp]),
    "sema_index_capacity": (C.c_uint64, [_vp]),
    "sema_index_dim": (C.c_uint32, [_vp]),
    "sema_index_device": (C.c_int, [_vp]),
    "sema_index_last_snapshot": (C.c_uint64, [_vp]),
    "sema_index_read_rows": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp]),
    "sema_index_set_scan_variant": (C.c_int, [_vp, C.c_int]),
    "sema_index_launch_count": (C.c_uint64, [_vp]),
    "sema_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_size_t]),
    "sema_host_free": (C.c_int, [_vp]),
    "sema_device_count": (C.c_int, []),
    "sema_last_error": (C.c_char_p, []),
    "sema_version": (C.c_char_p, []),
}

_lib = None


class SemaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"sema_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `make -C sema_b200/csrc` "
                "(or __graft_entry__.build()).  sema_b200 has no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI and the binding drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != SEMA_OK:
        raise SemaError(rc, lib().sema_last_error().decode("utf-8", "replace"))

"""CPU tests of the drop-in boundary: the shared library loads, exports exactly the
symbols include/sema_b200.h declares, and refuses to compute without a GPU."""
import ctypes as C
import os
import re

import pytest

from sema_b200 import _lib

pytestmark = pytest.mark.skipif(not os.path.exists(_lib.SO_PATH),
                                reason="libsema_b200.so not built (run __graft_entry__.build())")


def _declared():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(sema_[a-z0-9_]+)\s*\(", src))


def test_header_and_binding_agree():
    assert _declared() == set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(_lib.SO_PATH)
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in include/sema_b200.h but not exported"


def test_library_is_blackwell_native():
    # the fatbin must hold sm_100a code and nothing else
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_version_and_error_string():
    L = _lib.lib()
    assert b"sm_100a" in L.sema_version()
    assert isinstance(L.sema_last_error(), bytes)


def test_no_cpu_fallback():
    L = _lib.lib()
    if L.sema_device_count() > 0:
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = L.sema_index_create(0, 384, 1000, 0, C.byref(h))
    assert rc == _lib.SEMA_ERR_CUDA and not h.value
    assert b"no CPU fallback" in L.sema_last_error()


def test_argument_validation_without_gpu():
    L = _lib.lib()
    h = C.c_void_p()
    assert L.sema_index_create(0, 0, 10, 0, C.byref(h)) == _lib.SEMA_ERR_INVALID
    assert L.sema_index_create(0, 384, 10, 7, C.byref(h)) == _lib.SEMA_ERR_INVALID
    assert L.sema_index_create(0, 384, 1 << 33, 0, C.byref(h)) == _lib.SEMA_ERR_INVALID
    assert L.sema_index_destroy(None) == _lib.SEMA_OK
    assert L.sema_index_size(None) == 0


def test_new_entry_points_validate_arguments_without_gpu():
    # every entry point added for query streams / async searches / mean-pool / growable indexes rejects
    # null handles before touching CUDA, and the growable create has no CPU fallback either
    L = _lib.lib()
    t = C.c_uint64()
    nf = C.c_uint32()
    inv = _lib.SEMA_ERR_INVALID
    assert L.sema_index_search_submit(None, None, 10, C.byref(t)) == inv
    assert L.sema_index_search_collect(None, 1, None, None, C.byref(nf)) == inv
    assert L.sema_index_search_stream_device(None, None, 4, 10, None, None, None) == inv
    assert L.sema_index_search_batch_keys_device(None, None, 4, 10, None) == inv
    assert L.sema_topk_merge_batch_device(None, None, 2, 4, 10, None, None, None) == inv
    assert L.sema_mean_pool(None, None, None, 1, 16, 0, None) == inv
    assert L.sema_mean_pool_device(None, None, None, 1, 16, 0, None) == inv
    assert L.sema_shard_group_search_submit(None, None, 10, C.byref(t)) == inv
    assert L.sema_shard_group_search_stream_device(None, None, 4, 10, None, None, None) == inv
    h = C.c_void_p()
    assert L.sema_index_create_growable(0, 384, 0, 0, C.byref(h)) == inv
    if L.sema_device_count() == 0:
        assert L.sema_index_create_growable(0, 384, 1000, 0, C.byref(h)) == _lib.SEMA_ERR_CUDA and not h.value

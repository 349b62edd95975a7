"""Host-side view of the packed 64-bit ranking keys the kernels exchange.

    key = ordered_u32(rank value) << 32 | (0xFFFFFFFF - global_row_id)

(csrc/common.cuh: make_key).  rank value = the cosine score, or minus the squared L2
distance; ordered_u32 maps fp32 to an unsigned that sorts the same way.  A plain
unsigned compare therefore ranks by score, then by lower row id; 0 marks an empty slot.
The sharded path all-gathers these keys; this module lets host code (and the CPU tests
of the multi-rank logic) build, merge and decode them without a GPU.
"""
from __future__ import annotations

import numpy as np

METRIC_COSINE = 0
METRIC_L2 = 1


def pack_keys(scores, ids, metric: int = METRIC_COSINE) -> np.ndarray:
    s = np.asarray(scores, dtype=np.float32)
    rank = (-s if metric == METRIC_L2 else s) + np.float32(0.0)     # -0 -> +0
    u = rank.view(np.uint32)
    ordered = np.where(u & np.uint32(0x80000000), ~u, u | np.uint32(0x80000000)).astype(np.uint64)
    gid = np.asarray(ids, dtype=np.uint64)
    if gid.size and gid.max() >= 0xFFFFFFFF:
        raise ValueError("global row ids must stay below 2^32-1")
    return (ordered << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - gid)


def unpack_keys(keys, metric: int = METRIC_COSINE):
    """-> (ids uint64, scores float32) of the non-empty keys, in the given order."""
    k = np.asarray(keys).astype(np.uint64)
    k = k[k != 0]
    o = (k >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o).astype(np.uint32)
    rank = u.view(np.float32)
    ids = np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))
    scores = (-rank + np.float32(0.0)) if metric == METRIC_L2 else rank
    return ids, scores.astype(np.float32)


def merge_keys(key_lists, k: int) -> np.ndarray:
    """Global top-k of several key lists: the k largest distinct non-empty keys, descending
    (what kernel K4 computes on the device)."""
    allk = np.concatenate([np.asarray(x).astype(np.uint64).ravel() for x in key_lists]) if len(key_lists) else np.zeros(0, np.uint64)
    allk = allk[allk != 0]
    return np.sort(allk)[::-1][:k].copy()

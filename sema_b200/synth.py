"""Host mirror of the device corpus generator (k1_ingest.cuh: synth_value).

value(seed,row,col) = (sum of the four 16-bit fields of a splitmix64 hash) - 131070,
an exact integer in fp32 — so any row of the on-device synthetic corpus (SURVEY.md
§8(d)) can be regenerated on the host from its row id alone.  Used by bench.py to make
query vectors and by callers that want planted neighbours.
"""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def synth_rows(seed: int, row0: int, n: int, d: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        rows = (np.arange(n, dtype=np.uint64) + np.uint64(row0))[:, None]
        cols = np.arange(d, dtype=np.uint64)[None, :]
        x = np.uint64(seed) * _M1 + rows * _M2 + cols * _M3 + np.uint64(1)
        x ^= x >> np.uint64(30)
        x *= _M2
        x ^= x >> np.uint64(27)
        x *= _M3
        x ^= x >> np.uint64(31)
        m = np.uint64(0xFFFF)
        s = (x & m) + ((x >> np.uint64(16)) & m) + ((x >> np.uint64(32)) & m) + (x >> np.uint64(48))
    return (s.astype(np.int64) - 131070).astype(np.float32)

"""Generates tests/golden/*.npz — known-answer fixtures for the search hot path.

PARITY UNPINNED BY THE REFERENCE: akshitsinha/sema ships no golden vectors and
cannot be executed here, so these fixtures are produced by this
repo's own oracle (oracle/oracle.py).  Each case stores the inputs, the fp32 oracle
answer and the independent float64-shadow answer; the script refuses to write a case
where the two disagree on ids outside the 1e-5 tie tolerance.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, n, d, nq, k, seed
    ("g384", 1000, 384, 8, 10, 11),
    ("g768", 400, 768, 4, 100, 12),
    ("g130", 777, 130, 4, 50, 13),   # dim not a multiple of 4 -> padded rows
    ("g50", 300, 50, 4, 10, 14),
]


def main():
    for name, n, d, nq, k, seed in CASES:
        raw = O.synth(seed, 0, n, d)
        valid = np.ones(n, dtype=np.uint8)
        valid[5::97] = 0                      # null vectors (failed embeddings)
        raw[3] = 0.0                          # a zero row stays zero
        raw[n // 2] = raw[n // 3]             # an exact duplicate -> an exact tie
        X = O.normalize(raw)
        Q = O.normalize(O.synth(seed + 1000, 0, nq, d))
        Q[0] = X[n // 3]                      # a query that hits the duplicated rows exactly
        out = {"valid": valid, "X": X, "Q": Q, "k": np.int64(k), "seed": np.int64(seed)}
        for metric, tag in ((O.METRIC_DOT, "dot"), (O.METRIC_L2, "l2")):
            ids = np.zeros((nq, k), dtype=np.uint64)
            sc = np.zeros((nq, k), dtype=np.float32)
            sc64 = np.zeros((nq, k), dtype=np.float64)
            for i in range(nq):
                a, b = O.scan(X, Q[i], k, metric, valid)
                a64, b64 = O.scan(X, Q[i], k, metric, valid, f64=True)
                assert len(a) == k and len(a64) == k
                O.check_parity(a, b, a64, b64)   # f32 oracle vs f64 shadow
                ids[i], sc[i], sc64[i] = a, b, b64
            out[f"ids_{tag}"] = ids
            out[f"scores_{tag}"] = sc
            out[f"scores64_{tag}"] = sc64
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print("wrote", name, n, d, nq, k)


if __name__ == "__main__":
    main()

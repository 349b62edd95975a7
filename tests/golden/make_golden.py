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


def pool_case():
    """mean_pool (src/semantic/embeddings.rs:61-91): token embeddings with the reference's shapes
    (seq_len 32 here to keep the fixture small; hidden 384), 0/1 attention masks of varying length,
    one all-padding text (mask_sum = 0) and one all-zero text (norm = 0).  The fp32 answer is
    cross-checked against a float64 evaluation of the same formula before it is written."""
    n, seq, hidden = 6, 32, 384
    tok = (O.synth(77, 0, n * seq, hidden) / np.float32(65536.0)).reshape(n, seq, hidden)
    lens = [32, 17, 1, 0, 9, 25]
    mask = np.zeros((n, seq), dtype=np.float32)
    for t, ln in enumerate(lens):
        mask[t, :ln] = 1.0
    tok[4] = 0.0
    out = O.mean_pool(tok, mask)
    t64, m64 = tok.astype(np.float64), mask.astype(np.float64)
    p64 = (t64 * m64[:, :, None]).sum(axis=1)
    ms = m64.sum(axis=1)
    p64[ms > 0] /= ms[ms > 0, None]
    nr = np.sqrt((p64 * p64).sum(axis=1))
    p64[nr > 0] /= nr[nr > 0, None]
    assert np.allclose(out, p64, rtol=2e-5, atol=1e-7)
    assert np.array_equal(out[3], np.zeros(hidden, np.float32)) and np.array_equal(out[4], np.zeros(hidden, np.float32))
    np.savez_compressed(os.path.join(HERE, "pool384.npz"), tokens=tok, mask=mask, pooled=out, pooled64=p64)
    print("wrote pool384", n, seq, hidden)


def main():
    pool_case()
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

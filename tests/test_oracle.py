"""CPU tests of the oracle itself (oracle/oracle.py, oracle/cpu_scan.c).

The reference has no tests for this path; these are the known-answer
tests derived from its behaviour, each naming the reference lines it restates.
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "g*.npz")))


def test_normalize_kat():
    # src/semantic/embeddings.rs:83-88: norm = sqrt(sum x^2); x /= norm iff norm > 0
    x = np.zeros((3, 8), dtype=np.float32)
    x[0, :2] = [3.0, 4.0]
    x[2, :] = 2.0
    y = O.normalize(x)
    assert np.array_equal(y[0, :2], np.array([3.0, 4.0], np.float32) / np.float32(5.0))
    assert np.array_equal(y[1], np.zeros(8, np.float32))          # zero row left as zeros
    assert np.allclose(y[2], 1.0 / np.sqrt(8.0), rtol=1e-7)


def test_normalize_sequential_sum_matches_c(oracle_c):
    raw = O.synth(7, 0, 257, 384)
    assert np.array_equal(O.normalize(raw), oracle_c.normalize(raw))


def test_synth_matches_c_and_is_row_addressable(oracle_c):
    a = O.synth(3, 100, 50, 130)
    assert np.array_equal(a, oracle_c.synth(3, 100, 50, 130))
    assert np.array_equal(a[10:20], O.synth(3, 110, 10, 130))      # shard-count invariant corpus
    assert np.abs(a).max() <= 131070 and a.dtype == np.float32
    assert np.array_equal(a, np.round(a))                          # exact integers


def test_scan_one_hot_kat():
    # rows are one-hot basis vectors scaled; the query picks coordinates
    n, d = 40, 16
    X = np.zeros((n, d), dtype=np.float32)
    for r in range(n):
        X[r, r % d] = 1.0 + r
    X = O.normalize(X)                      # all unit one-hot rows
    q = np.zeros(d, dtype=np.float32)
    q[3] = 1.0
    ids, sc = O.scan(X, q, 5)
    assert ids.tolist() == [3, 19, 35, 0, 1]      # three exact hits (ties -> lower id), then zeros
    assert sc.tolist() == [1.0, 1.0, 1.0, 0.0, 0.0]
    ids2, d2 = O.scan(X, q, 5, O.METRIC_L2)
    assert ids2.tolist() == ids.tolist()
    assert d2.tolist() == [0.0, 0.0, 0.0, 2.0, 2.0]   # L2^2 = 2 - 2 cos


def test_metric_equivalence_on_unit_rows():
    # the reference relies on LanceDB's default L2 over normalised vectors
    # (src/storage/lance_indexer.rs:121-126 + src/semantic/embeddings.rs:83-88)
    X = O.normalize(O.synth(1, 0, 2000, 384))
    q = O.normalize(O.synth(2, 0, 1, 384))[0]
    i_dot, s_dot = O.scan(X, q, 50, O.METRIC_DOT, f64=True)
    i_l2, s_l2 = O.scan(X, q, 50, O.METRIC_L2, f64=True)
    assert np.array_equal(i_dot, i_l2)
    assert np.allclose(s_l2, O.l2sq_from_cosine(s_dot), atol=2e-6)


def test_limit_semantics():
    X = O.normalize(O.synth(1, 0, 7, 32))
    q = X[2].copy()
    ids, sc = O.scan(X, q, 50)                # limit > N returns every row
    assert len(ids) == 7 and ids[0] == 2
    assert np.all(np.diff(sc) <= 0)
    ids, _ = O.scan(X[:0], q, 10)             # missing/empty table => empty (lance_indexer.rs:108-111)
    assert len(ids) == 0
    ids, _ = O.scan(X, q, 0)
    assert len(ids) == 0


def test_null_rows_skipped():
    # null vectors come from failed embeddings (lance_indexer.rs:66-70) and are never returned
    X = O.normalize(O.synth(1, 0, 100, 64))
    valid = np.ones(100, np.uint8)
    q = X[10].copy()
    valid[10] = 0
    ids, _ = O.scan(X, q, 100, valid=valid)
    assert 10 not in ids.tolist() and len(ids) == 99


def test_ties_lower_id_first():
    X = O.normalize(O.synth(1, 0, 64, 48))
    X[50] = X[7]
    X[20] = X[7]
    ids, sc = O.scan(X, X[7].copy(), 3)
    assert ids.tolist() == [7, 20, 50] and sc[0] == sc[1] == sc[2]


@pytest.mark.parametrize("metric", [O.METRIC_DOT, O.METRIC_L2])
@pytest.mark.parametrize("n,d,k", [(1, 384, 10), (1000, 384, 10), (513, 768, 100), (300, 50, 7)])
def test_c_oracle_equals_numpy_oracle(oracle_c, n, d, k, metric):
    X = O.normalize(O.synth(5, 0, n, d))
    valid = np.ones(n, np.uint8)
    valid[::13] = 0
    q = O.normalize(O.synth(6, 0, 1, d))[0]
    a_ids, a_sc = O.scan(X, q, k, metric, valid, id_base=1000)
    b_ids, b_sc = oracle_c.scan(X, q, k, metric, valid, id_base=1000)
    assert np.array_equal(a_ids, b_ids)
    assert np.array_equal(a_sc, b_sc)          # same operation order => bit exact


def test_c_oracle_thread_count_invariant(oracle_c):
    X = O.normalize(O.synth(5, 0, 5000, 384))
    q = O.normalize(O.synth(6, 0, 1, 384))[0]
    nt = oracle_c.threads()
    oracle_c.set_threads(1)
    a = oracle_c.scan(X, q, 50)
    oracle_c.set_threads(max(nt, 2))
    b = oracle_c.scan(X, q, 50)
    oracle_c.set_threads(nt)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_merge_equals_single_scan(oracle_c):
    # top-k of the union of per-shard top-k == global top-k
    n, d, k, G = 4000, 384, 10, 4
    X = O.normalize(O.synth(1, 0, n, d))
    q = O.normalize(O.synth(2, 0, 1, d))[0]
    full_ids, full_sc = O.scan(X, q, k)
    per = n // G
    sc = np.zeros((G, k), np.float32)
    ids = np.zeros((G, k), np.uint64)
    lens = []
    for g in range(G):
        i, s = O.scan(X[g * per:(g + 1) * per], q, k, id_base=g * per)
        ids[g, :len(i)], sc[g, :len(i)] = i, s
        lens.append(len(i))
    for impl in (O, oracle_c):
        m_ids, m_sc = impl.merge(sc, ids, lens, k)
        assert np.array_equal(m_ids, full_ids) and np.array_equal(m_sc, full_sc)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracles_reproduce_golden(oracle_c, path):
    g = np.load(path)
    X, Q, valid, k = g["X"], g["Q"], g["valid"], int(g["k"])
    for metric, tag in ((O.METRIC_DOT, "dot"), (O.METRIC_L2, "l2")):
        for i in range(Q.shape[0]):
            for impl in (O, oracle_c):
                ids, sc = impl.scan(X, Q[i], k, metric, valid)
                assert np.array_equal(ids, g[f"ids_{tag}"][i])
                assert np.array_equal(sc, g[f"scores_{tag}"][i])
            O.check_parity(g[f"ids_{tag}"][i], g[f"scores_{tag}"][i], g[f"ids_{tag}"][i],
                           g[f"scores64_{tag}"][i])


def test_check_parity_rejects_wrong_order():
    ids = np.array([1, 2, 3], np.uint64)
    sc = np.array([0.9, 0.5, 0.1], np.float32)
    O.check_parity(ids, sc, ids, sc)
    with pytest.raises(AssertionError):
        O.check_parity(np.array([2, 1, 3], np.uint64), sc, ids, sc)
    with pytest.raises(AssertionError):
        O.check_parity(ids, sc * np.float32(1.001), ids, sc)
    # a swap inside the tie tolerance is accepted
    sc_t = np.array([0.9, 0.500001, 0.5], np.float32)
    O.check_parity(np.array([1, 3, 2], np.uint64), sc_t, ids, sc_t)


# ---------------------------------------------------------------- mean_pool (the step before the path)
def _mean_pool_literal(tok, mask):
    """Scalar transliteration of src/semantic/embeddings.rs:61-91, one text, np.float32 scalars."""
    seq, hidden = tok.shape
    pooled = [np.float32(0.0)] * hidden
    mask_sum = np.float32(0.0)
    for i in range(seq):
        mv = np.float32(mask[i])
        mask_sum = np.float32(mask_sum + mv)
        for j in range(hidden):
            pooled[j] = np.float32(pooled[j] + np.float32(tok[i, j] * mv))
    if mask_sum > 0:
        pooled = [np.float32(v / mask_sum) for v in pooled]
    ss = np.float32(0.0)
    for v in pooled:
        ss = np.float32(ss + np.float32(v * v))
    norm = np.float32(np.sqrt(ss))
    if norm > 0:
        pooled = [np.float32(v / norm) for v in pooled]
    return np.array(pooled, dtype=np.float32)


def test_mean_pool_kat():
    # two tokens attended, one padded: pooled = mean of the attended rows, then unit norm
    tok = np.array([[[3.0, 0.0], [0.0, 4.0], [100.0, 100.0]]], dtype=np.float32)
    mask = np.array([[1.0, 1.0, 0.0]], dtype=np.float32)
    out = O.mean_pool(tok, mask)
    assert np.array_equal(out[0], np.array([1.5, 2.0], np.float32) / np.float32(2.5))
    # an all-padding text: mask_sum = 0 -> no division, pooled = 0 -> norm = 0 -> stays zero
    assert np.array_equal(O.mean_pool(tok, np.zeros((1, 3), np.float32))[0], np.zeros(2, np.float32))


def test_mean_pool_vectorised_equals_literal_loop_and_c(oracle_c):
    tok = (O.synth(9, 0, 3 * 11, 20) / np.float32(4096.0)).reshape(3, 11, 20)
    mask = np.zeros((3, 11), dtype=np.float32)
    mask[0, :11] = 1.0
    mask[1, :4] = 1.0
    mask[2, :7] = 0.5                       # the code multiplies by the mask value, whatever it is
    got = O.mean_pool(tok, mask)
    for t in range(3):
        assert np.array_equal(got[t], _mean_pool_literal(tok[t], mask[t]))
    assert np.array_equal(got, oracle_c.mean_pool(tok, mask))


def test_mean_pool_golden_fixture(oracle_c):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pool384.npz"))
    assert np.array_equal(O.mean_pool(g["tokens"], g["mask"]), g["pooled"])
    assert np.array_equal(oracle_c.mean_pool(g["tokens"], g["mask"]), g["pooled"])
    np.testing.assert_allclose(g["pooled"], g["pooled64"], rtol=2e-5, atol=1e-7)


# ------------------------------------------------------------------ a third, independent implementation
@pytest.mark.parametrize("n,d,k,metric", [(200_000, 384, 10, 0), (200_000, 384, 50, 1), (60_000, 768, 100, 0), (30_000, 130, 10, 1)])
def test_oracles_agree_with_an_independent_float64_torch_implementation(oracle_c, n, d, k, metric):
    """The oracle is the builder's restatement, and the golden fixtures are its own output; so it is also held
    against code that shares nothing with it: torch's CPU GEMV and torch.topk in float64 (different summation
    order, different selection algorithm, twice the precision).  BASELINE.json's rule must hold between them:
    same ids and order except ties within 1e-5, scores within 1e-5 relative."""
    import torch
    X = oracle_c.normalize(oracle_c.synth(11, 0, n, d))
    Q = oracle_c.normalize(oracle_c.synth(12, 0, 6, d))
    valid = np.ones(n, np.uint8)
    valid[::29] = 0
    Xt = torch.from_numpy(X).double()
    dead = torch.from_numpy(valid == 0)
    for q in Q:
        qt = torch.from_numpy(q).double()
        if metric == 0:
            s = Xt @ qt
            s[dead] = -float("inf")
            top = torch.topk(s, k)
            t_ids, t_sc = top.indices.numpy().astype(np.uint64), top.values.numpy()
        else:
            s = ((Xt - qt) ** 2).sum(dim=1)
            s[dead] = float("inf")
            top = torch.topk(s, k, largest=False)
            t_ids, t_sc = top.indices.numpy().astype(np.uint64), top.values.numpy()
        c_ids, c_sc = oracle_c.scan(X, q, k, metric=metric, valid=valid)
        O.check_parity(c_ids, c_sc, t_ids, t_sc)                    # C oracle vs the independent implementation
        if n <= 60_000:
            p_ids, p_sc = O.scan(X, q, k, metric=metric, valid=valid)
            O.check_parity(p_ids, p_sc, t_ids, t_sc)                # NumPy oracle vs the independent implementation


def test_normalize_agrees_with_float64_torch():
    import torch
    raw = O.synth(13, 0, 2000, 384)
    want = torch.nn.functional.normalize(torch.from_numpy(raw).double(), dim=1).numpy()
    np.testing.assert_allclose(O.normalize(raw), want, rtol=3e-6, atol=1e-9)   # fp32 sequential sum vs fp64: a few ulp

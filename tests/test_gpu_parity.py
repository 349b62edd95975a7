"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle.

Acceptance rule (BASELINE.json): identical top-k id set and order except for ties
within 1e-5; scores within 1e-5 relative (oracle.check_parity).  Exact-tie cases and
row ids are checked bit-exactly.
"""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "g*.npz")))


@pytest.fixture(scope="module")
def sema():
    import sema_b200
    from sema_b200 import _lib
    if _lib.lib().sema_device_count() == 0:
        pytest.fail("gpu-marked test run without a CUDA device")
    return sema_b200


def _unit(seed, n, d):
    return O.normalize(O.synth(seed, 0, n, d))


# ---------------------------------------------------------------- K1 ingest
@pytest.mark.parametrize("d", [384, 768, 130, 50, 4, 1])
def test_k1_normalize_matches_reference_rule(sema, d):
    # src/semantic/embeddings.rs:83-88
    n = 257
    raw = O.synth(21, 0, n, d)
    raw[3] = 0.0
    if d >= 2:
        raw[4] = 0.0
        raw[4, :2] = [3.0, 4.0]
    with sema.GpuIndex(d, n) as idx:
        assert idx.append(raw, normalize=True) == 0
        got = idx.read_rows(0, n)
    want = O.normalize(raw)
    assert np.array_equal(got[3], np.zeros(d, np.float32))     # zero row stays zero
    if d >= 2:
        assert np.array_equal(got[4, :2], np.array([3.0, 4.0], np.float32) / np.float32(5.0))
    # K1 folds the squares in the reference's sequential order (multiply, then add; IEEE sqrt and divide):
    # the stored rows are the reference's bits, not merely close to them
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,d", [(1, 384), (31, 384), (33, 384), (4099, 384), (1000, 768), (129, 1024), (77, 132), (65, 7)])
def test_k1_is_bit_identical_to_the_sequential_reference_sum(sema, oracle_c, n, d):
    # src/semantic/embeddings.rs:83-88 — ragged row counts (warps own groups of 32 rows), chunked columns,
    # large dynamic range inside a row (the fold order matters most there), host / device / query entry points
    rng = np.random.default_rng(n * 1000 + d)
    raw = (rng.standard_normal((n, d)) * np.exp(rng.uniform(-12, 12, (n, d)))).astype(np.float32)
    want = oracle_c.normalize(raw)
    assert np.array_equal(want, O.normalize(raw))              # C and NumPy restatements agree
    with sema.GpuIndex(d, 2 * n + 8) as idx:
        idx.append(raw, normalize=True)
        assert np.array_equal(idx.read_rows(0, n), want)
        import torch
        dev_rows = torch.from_numpy(raw).cuda()
        idx.append_device(dev_rows.data_ptr(), n, None, normalize=True)
        assert np.array_equal(idx.read_rows(n, n), want)


def test_k1_no_normalize_is_bit_exact_copy(sema):
    X = _unit(5, 300, 384)
    with sema.GpuIndex(384, 300) as idx:
        idx.append(X, normalize=False)
        assert np.array_equal(idx.read_rows(0, 300), X)


def test_k1_synthetic_generator_is_bit_identical_to_oracle(sema):
    with sema.GpuIndex(384, 1000) as idx:
        idx.append_synthetic(seed=1, row0=12345, n=1000, normalize=False)
        assert np.array_equal(idx.read_rows(0, 1000), O.synth(1, 12345, 1000, 384))


# ---------------------------------------------------------------- K2 scan
CASES = [
    # n, d, k
    (1, 384, 10), (3, 384, 1), (31, 384, 10), (33, 384, 50), (1000, 384, 10), (4097, 384, 50),
    (50001, 384, 10), (50001, 384, 100), (20000, 384, 128), (20000, 768, 100), (5000, 768, 10),
    (3000, 130, 50), (3000, 50, 10), (2000, 1024, 10), (500, 2048, 33), (2000, 4, 10),
    (6000, 384, 200), (6000, 384, 1024),
]


@pytest.mark.parametrize("metric", [0, 1], ids=["cosine", "l2"])
@pytest.mark.parametrize("n,d,k", CASES)
def test_k2_search_matches_oracle(sema, oracle_c, n, d, k, metric):
    X = _unit(1, n, d)
    Q = _unit(2, 3, d)
    with sema.GpuIndex(d, n + 5, metric=metric) as idx:
        idx.append(X, normalize=False)         # bit-identical matrix on both sides
        for q in Q:
            ids, sc = idx.search(q, k)
            r_ids, r_sc = oracle_c.scan(X, q, k, metric)
            assert len(ids) == min(k, n)
            O.check_parity(ids, sc, r_ids, r_sc)
            # ranking direction
            assert np.all(np.diff(sc) >= 0) if metric else np.all(np.diff(sc) <= 0)


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_k2_variants_agree(sema, oracle_c, variant):
    n, d = 30011, 384
    X = _unit(3, n, d)
    q = _unit(4, 1, d)[0]
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        idx.set_scan_variant(variant)
        ids, sc = idx.search(q, 10)
    r_ids, r_sc = oracle_c.scan(X, q, 10)
    O.check_parity(ids, sc, r_ids, r_sc)


def test_empty_index_and_k_zero(sema):
    # missing table => Ok(empty) (src/storage/lance_indexer.rs:108-111)
    with sema.GpuIndex(384, 10) as idx:
        ids, sc = idx.search(np.zeros(384, np.float32), 10)
        assert len(ids) == 0 and len(sc) == 0
        idx.append(_unit(1, 4, 384), normalize=False)
        ids, _ = idx.search(np.zeros(384, np.float32), 0)
        assert len(ids) == 0


def test_limit_larger_than_table(sema, oracle_c):
    X = _unit(1, 7, 384)
    with sema.GpuIndex(384, 7) as idx:
        idx.append(X, normalize=False)
        ids, sc = idx.search(X[2].copy(), 50)      # SEARCH_RESULTS_LIMIT = 50, src/tui/engine.rs:11
    r_ids, r_sc = oracle_c.scan(X, X[2].copy(), 50)
    assert len(ids) == 7 and ids[0] == 2
    O.check_parity(ids, sc, r_ids, r_sc)


def test_null_and_nonfinite_rows_never_returned(sema, oracle_c):
    # nullable vector column: src/storage/lance_indexer.rs:41-45, 66-70
    n, d = 2000, 384
    raw = O.synth(9, 0, n, d)
    valid = np.ones(n, np.uint8)
    valid[::7] = 0
    raw[11, 5] = np.nan
    raw[13, 0] = np.inf
    ok = valid.copy()
    ok[11] = ok[13] = 0
    q = O.normalize(raw[14:15].copy())[0]          # 14 is a null row (14 % 7 == 0): must not match itself
    with sema.GpuIndex(d, n) as idx:
        idx.append(raw, valid=valid, normalize=True)
        ids, sc = idx.search(q, 100)
        all_ids, _ = idx.search(q, 1024)
        X = idx.read_rows(0, n)
    assert not (set(ids.tolist()) & set(np.nonzero(ok == 0)[0].tolist()))
    assert np.isnan(X[0]).all() and np.isnan(X[11]).all()
    Xo = O.normalize(np.where(ok[:, None] == 1, raw, 0).astype(np.float32))
    r_ids, r_sc = oracle_c.scan(Xo, q, 100, valid=ok)
    O.check_parity(ids, sc, r_ids, r_sc)
    assert len(all_ids) == min(1024, int(ok.sum()))


def test_all_rows_null(sema):
    raw = O.synth(9, 0, 64, 384)
    with sema.GpuIndex(384, 64) as idx:
        idx.append(raw, valid=np.zeros(64, np.uint8))
        ids, _ = idx.search(O.normalize(raw[:1])[0], 10)
        assert len(ids) == 0


def test_exact_ties_rank_lower_row_id_first(sema):
    n, d = 5000, 384
    X = _unit(1, n, d)
    for r in (4000, 17, 2500, 4999):
        X[r] = X[100]
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        ids, sc = idx.search(X[100].copy(), 6)
    assert ids[:5].tolist() == [17, 100, 2500, 4000, 4999]
    assert len(set(sc[:5].tolist())) == 1


def test_metric_equivalence_l2_vs_cosine(sema):
    # LanceDB default L2 over unit rows orders like cosine; _distance = 2 - 2 cos
    n, d, k = 20000, 384, 50
    X = _unit(1, n, d)
    q = _unit(2, 1, d)[0]
    with sema.GpuIndex(d, n, metric=0) as a, sema.GpuIndex(d, n, metric=1) as b:
        a.append(X, normalize=False)
        b.append(X, normalize=False)
        ia, sa = a.search(q, k)
        ib, sb = b.search(q, k)
    assert np.array_equal(ia, ib)
    np.testing.assert_allclose(sb, 2.0 - 2.0 * sa.astype(np.float64), atol=2e-6)


def test_l2_metric_on_unnormalised_rows_and_zero_row(sema, oracle_c):
    # under LanceDB's L2 a zero row has distance |q|^2 = 1 (SURVEY.md §7 edge semantics)
    n, d = 3000, 384
    X = (O.synth(1, 0, n, d) / np.float32(65536.0 * 20)).astype(np.float32)
    X[7] = 0.0
    q = _unit(2, 1, d)[0]
    with sema.GpuIndex(d, n, metric=1) as idx:
        idx.append(X, normalize=False)
        ids, dist = idx.search(q, 20)
    r_ids, r_dist = oracle_c.scan(X, q, 20, 1)
    O.check_parity(ids, dist, r_ids, r_dist)


def test_row_base_offsets_ids(sema):
    X = _unit(1, 100, 384)
    with sema.GpuIndex(384, 100) as idx:
        idx.set_row_base(1_000_000)
        idx.append(X, normalize=False)
        ids, _ = idx.search(X[42].copy(), 3)
    assert ids[0] == 1_000_042


def test_tombstone_removes_rows(sema, oracle_c):
    # table.delete(predicate): src/storage/lance_indexer.rs:234-250
    n, d = 4000, 384
    X = _unit(1, n, d)
    q = X[123].copy()
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        ids, _ = idx.search(q, 10)
        assert ids[0] == 123
        dead = ids[:3].copy()
        idx.tombstone(dead)
        ids2, sc2 = idx.search(q, 10)
    valid = np.ones(n, np.uint8)
    valid[dead.astype(np.int64)] = 0
    r_ids, r_sc = oracle_c.scan(X, q, 10, valid=valid)
    O.check_parity(ids2, sc2, r_ids, r_sc)


def test_append_then_search_sees_new_rows(sema, oracle_c):
    # table.add then nearest_to: src/storage/lance_indexer.rs:92-95, 121-126
    d = 384
    X = _unit(1, 3000, d)
    q = X[2900].copy()
    with sema.GpuIndex(d, 3000) as idx:
        assert idx.append(X[:1000], normalize=False) == 0
        ids, _ = idx.search(q, 5)
        assert 2900 not in ids.tolist() and idx.last_snapshot == 1000
        assert idx.append(X[1000:], normalize=False) == 1000
        ids, sc = idx.search(q, 5)
        assert ids[0] == 2900 and idx.last_snapshot == 3000 and len(idx) == 3000
        r_ids, r_sc = oracle_c.scan(X, q, 5)
        O.check_parity(ids, sc, r_ids, r_sc)
        with pytest.raises(sema.SemaError):
            idx.append(X[:1], normalize=False)       # capacity exceeded


def test_async_ingest_snapshot_semantics(sema, oracle_c):
    # config 5: ingest interleaved with queries; each query equals the oracle on the
    # snapshot (visible row count) it observed
    d, batch, nb = 384, 2048, 8
    X = _unit(1, batch * nb, d)
    q = _unit(2, 1, d)[0]
    with sema.GpuIndex(d, batch * nb) as idx:
        for b in range(nb):
            idx.append(X[b * batch:(b + 1) * batch], normalize=False, asynchronous=True)
            ids, sc = idx.search(q, 10)
            snap = idx.last_snapshot
            assert snap % batch == 0 and snap <= (b + 1) * batch
            r_ids, r_sc = oracle_c.scan(X[:snap], q, 10)
            O.check_parity(ids, sc, r_ids, r_sc)
        idx.flush()
        assert idx.visible == batch * nb


def test_batch_search_matches_single_queries(sema, oracle_c):
    n, d, k, nq = 20000, 384, 10, 17
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        ids, sc, nf = idx.search_batch(Q, k)
    r_ids, r_sc, r_nf = oracle_c.scan_batch(X, Q, k)
    assert np.array_equal(nf, r_nf)
    for i in range(nq):
        O.check_parity(ids[i], sc[i], r_ids[i], r_sc[i])


# ---------------------------------------------------------------- K4 merge
@pytest.mark.parametrize("G,k", [(2, 10), (8, 10), (8, 100), (4, 300), (3, 1)])
def test_k4_virtual_shards_equal_single_index(sema, oracle_c, G, k):
    # G row ranges on one device, allgather replaced by a concatenation
    import torch
    n, d = 24000, 384
    X = _unit(1, n, d)
    q = _unit(2, 1, d)[0]
    per = n // G
    dev = torch.device("cuda:0")
    q_dev = torch.from_numpy(q).to(dev)
    keys = torch.zeros(G * k, dtype=torch.int64, device=dev)
    ids_d = torch.zeros(k, dtype=torch.int64, device=dev)
    sc_d = torch.zeros(k, dtype=torch.float32, device=dev)
    nf_d = torch.zeros(1, dtype=torch.int32, device=dev)
    shards = []
    try:
        for g in range(G):
            lo, hi = g * per, (n if g == G - 1 else (g + 1) * per)
            s = sema.GpuIndex(d, hi - lo)
            s.set_row_base(lo)
            s.append(X[lo:hi], normalize=False)
            s.set_stream(torch.cuda.current_stream().cuda_stream)
            shards.append(s)
        for g, s in enumerate(shards):
            s.search_keys_device(q_dev.data_ptr(), k, keys.data_ptr() + g * k * 8)
        shards[0].merge_device(keys.data_ptr(), G, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
        torch.cuda.synchronize()
    finally:
        for s in shards:
            s.close()
    nf = int(nf_d.item())
    r_ids, r_sc = oracle_c.scan(X, q, k)
    assert nf == len(r_ids)
    O.check_parity(ids_d.cpu().numpy().astype(np.uint64)[:nf], sc_d.cpu().numpy()[:nf], r_ids, r_sc)


# ---------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_golden_fixtures(sema, path):
    g = np.load(path)
    X, Q, valid, k = g["X"], g["Q"], g["valid"], int(g["k"])
    d = X.shape[1]
    for metric, tag in ((0, "dot"), (1, "l2")):
        with sema.GpuIndex(d, X.shape[0], metric=metric) as idx:
            idx.append(X, valid=valid, normalize=False)
            for i in range(Q.shape[0]):
                ids, sc = idx.search(Q[i], k)
                O.check_parity(ids, sc, g[f"ids_{tag}"][i], g[f"scores_{tag}"][i])
            ids, sc, nf = idx.search_batch(Q, k)
            for i in range(Q.shape[0]):
                O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], g[f"ids_{tag}"][i], g[f"scores_{tag}"][i])


# ---------------------------------------------------------------- BASELINE sizes
def test_config2_1m_x_384_matches_oracle(sema, oracle_c):
    # BASELINE.json configs[1]: 1M x 384 synthetic unit-norm, single-query top-10.
    # The corpus is generated on the device and, independently, on the host by the
    # oracle's generator; K1 and the oracle normalise their own copies.
    n, d, k = 1_000_000, 384, 10
    Xo = oracle_c.normalize(oracle_c.synth(1, 0, n, d))
    Q = oracle_c.normalize(oracle_c.synth(2, 0, 8, d))
    with sema.GpuIndex(d, n) as idx:
        idx.append_synthetic(seed=1, row0=0, n=n, normalize=True)
        for q in Q:
            ids, sc = idx.search(q, k)
            r_ids, r_sc = oracle_c.scan(Xo, q, k)
            O.check_parity(ids, sc, r_ids, r_sc)
        ids, sc = idx.search(Q[0], 100)
        r_ids, r_sc = oracle_c.scan(Xo, Q[0], 100)
        O.check_parity(ids, sc, r_ids, r_sc)


def test_full_size_10m_x_384_properties(sema, oracle_c):
    # BASELINE.json headline size.  Size-independent properties:
    #  (1) a stored row used as the query returns itself first with score ~ 1;
    #  (2) the result over N rows equals the K4 merge of results over row sub-ranges
    #      (checked through the oracle on the rows the GPU returned);
    #  (3) every returned score equals the oracle's fp32 dot on that row (regenerated
    #      on the host from the row id alone) and the list is sorted.
    n, d, k = 10_000_000, 384, 10
    Q = oracle_c.normalize(oracle_c.synth(2, 0, 4, d))
    Xo = oracle_c.normalize(oracle_c.synth(1, 0, 1_000_000, d))
    with sema.GpuIndex(d, n) as idx:
        idx.append_synthetic(seed=1, row0=0, n=n, normalize=True)
        for probe in (0, 1, 4_999_999, n - 1):
            ids, sc = idx.search(idx.read_rows(probe, 1)[0], 3)
            assert ids[0] == probe and abs(sc[0] - 1.0) < 1e-5
        for q in Q:
            ids, sc = idx.search(q, k)
            assert len(ids) == k and np.all(np.diff(sc) <= 0) and len(set(ids.tolist())) == k
            for i, s in zip(ids, sc):
                row = oracle_c.normalize(oracle_c.synth(1, int(i), 1, d))
                want = O.row_keys(row, q)[0]
                assert abs(s - want) <= 1e-5 * abs(want)
            # the first 1M-row prefix: every oracle hit that beats the GPU's k-th score
            # must be in the GPU list (no better row was missed in that prefix)
            p_ids, p_sc = oracle_c.scan(Xo, q, k)
            for i, s in zip(p_ids, p_sc):
                if s > sc[-1] + 1e-5:
                    assert i in ids


# ---------------------------------------------------------------- K3 batched tensor-core path
K3_CASES = [
    # n, nq, k
    (64, 4, 10), (1000, 17, 10), (50001, 128, 10), (50001, 300, 50), (200000, 64, 100), (20000, 5, 1),
    (130, 129, 16), (70000, 1024, 10),
]


@pytest.mark.parametrize("prec", [0, 1], ids=["fp16", "bf16"])
@pytest.mark.parametrize("mode", [2, 3], ids=["x3", "x1"])
@pytest.mark.parametrize("n,nq,k", K3_CASES)
def test_k3_batch_matches_oracle_and_k2(sema, oracle_c, n, nq, k, mode, prec):
    d = 384
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        assert idx.set_batch_mode(mode) == mode           # tensor-core path (3 passes / 1 pass)
        assert idx.set_batch_precision(prec) == prec      # automatic -> fp16 halves on unit rows; 1 = bf16 halves
        ids3, sc3, nf3 = idx.search_batch(Q, k)
        served, fallbacks = idx.batch_stats()
        assert served == nq                                # K3 really ran
        assert idx.batch_precision_active == (1 if prec == 0 else 0)
        idx.set_batch_mode(1)                              # K2 once per query
        ids2, sc2, nf2 = idx.search_batch(Q, k)
    assert np.array_equal(nf3, nf2) and (nf3 == min(k, n)).all()
    assert np.array_equal(ids3, ids2)
    assert np.array_equal(sc3, sc2)                        # same fp32 arithmetic after re-scoring
    r_ids, r_sc, _ = oracle_c.scan_batch(X, Q, k)
    for i in range(nq):
        O.check_parity(ids3[i, :nf3[i]], sc3[i, :nf3[i]], r_ids[i, :nf3[i]], r_sc[i, :nf3[i]])
    assert fallbacks <= max(1, nq // 50)                   # exactness is proven for ~all random queries


def test_k3_ties_and_duplicates_fall_back_to_exact_path(sema, oracle_c):
    # 300 copies of one row: more exact ties than any candidate list holds -> the exactness
    # proof fails for queries near that row and those queries are re-run through K2
    n, d, k = 20000, 384, 10
    X = _unit(1, n, d)
    X[5000:5300] = X[123]
    Q = _unit(2, 8, d)
    Q[0] = X[123]
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        idx.set_batch_mode(2)
        ids, sc, nf = idx.search_batch(Q, k)
        served, fallbacks = idx.batch_stats()
        idx.set_batch_mode(3)                             # the single-pass filter must stay exact too
        ids1, sc1, nf1 = idx.search_batch(Q, k)
    assert np.array_equal(ids, ids1) and np.array_equal(sc, sc1)
    assert served == 8 and fallbacks >= 1
    assert ids[0].tolist() == [123] + list(range(5000, 5009))
    r_ids, r_sc, _ = oracle_c.scan_batch(X, Q, k)
    for i in range(8):
        O.check_parity(ids[i], sc[i], r_ids[i], r_sc[i])


PAIR_CASES = [
    # n, d, nq, k: the single-pass stage as CTA pairs (nq > 128: at least two query tiles).  Tile counts odd and
    # even, a last tile of one row, dims from one k-block to the TMEM limit, every candidate-list size
    (65, 384, 129, 10), (129, 384, 200, 10), (4160, 384, 256, 10), (4161, 64, 130, 16), (50001, 384, 300, 50),
    (20000, 512, 257, 100), (9999, 128, 512, 10), (70000, 384, 1024, 10), (30000, 256, 384, 32),
]


@pytest.mark.parametrize("variant", [701, 700], ids=["pairs", "single-cta"])
@pytest.mark.parametrize("n,d,nq,k", PAIR_CASES)
def test_k3_pair_kernel_matches_oracle_and_k2(sema, oracle_c, n, d, nq, k, variant):
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        idx.set_batch_mode(3)
        assert idx.set_scan_variant(variant) == variant
        ids3, sc3, nf3 = idx.search_batch(Q, k)
        served, fallbacks = idx.batch_stats()
        assert served == nq
        idx.set_batch_mode(1)
        ids2, sc2, nf2 = idx.search_batch(Q, k)
    assert np.array_equal(nf3, nf2) and (nf3 == min(k, n)).all()
    assert np.array_equal(ids3, ids2) and np.array_equal(sc3, sc2)
    r_ids, r_sc, _ = oracle_c.scan_batch(X, Q, k)
    for i in range(0, nq, 7):
        O.check_parity(ids3[i, :nf3[i]], sc3[i, :nf3[i]], r_ids[i, :nf3[i]], r_sc[i, :nf3[i]])
    assert fallbacks <= max(1, nq // 50)                   # the pair kernel's lists prove ~all random queries


@pytest.mark.parametrize("qmag,xmag,want_fmt", [(1e6, 1.0, 1), (1e-6, 1.0, 1), (3.0, 500.0, 1), (1.0, 5000.0, 0), (1e30, 1e-3, 1)])
def test_k3_query_scaling_and_format_choice(sema, oracle_c, qmag, xmag, want_fmt):
    """Queries of any magnitude are scaled by a power of two inside the kernel (fp16's range is narrow); rows
    decide the format: fp16 halves while every element is <= 1024, bf16 beyond."""
    n, d, nq, k = 20000, 384, 160, 10
    X = (_unit(1, n, d) * np.float32(xmag)).astype(np.float32)
    Q = (_unit(2, nq, d) * np.float32(qmag)).astype(np.float32)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        for mode in (3, 2, 0):
            idx.set_batch_mode(mode)
            q0, f0 = idx.batch_stats()
            ids3, sc3, nf3 = idx.search_batch(Q, k)
            q1, f1 = idx.batch_stats()
            assert q1 - q0 == nq and f1 - f0 <= 3
            assert idx.batch_precision_active == want_fmt
            idx.set_batch_mode(1)
            ids2, sc2, nf2 = idx.search_batch(Q, k)
            assert np.array_equal(ids3, ids2) and np.array_equal(sc3, sc2) and (nf3 == k).all()
    r_ids, r_sc, _ = oracle_c.scan_batch(X, Q[:16], k)
    for i in range(16):
        O.check_parity(ids3[i], sc3[i], r_ids[i], r_sc[i])


def test_k3_probe_settings_are_refused_by_the_shipped_library(sema):
    with sema.GpuIndex(384, 64) as idx:
        for bad in (301, 302, 303, 316, 332, 364, 702, 1000, 1102, 1300):
            assert idx.set_scan_variant(bad) == -1
        assert idx.set_scan_variant(308) == 308 and idx.set_scan_variant(300) == 300


def test_k3_null_rows_appends_and_tombstones(sema, oracle_c):
    n, d, k = 30000, 384, 10
    raw = O.synth(1, 0, n, d)
    valid = np.ones(n, np.uint8)
    valid[::11] = 0
    Q = _unit(2, 32, d)
    with sema.GpuIndex(d, n) as idx:
        idx.set_batch_mode(2)
        idx.append(raw[:10000], valid=valid[:10000], normalize=True)
        X = idx.read_rows(0, 10000)
        ids, sc, nf = idx.search_batch(Q, k)
        r = oracle_c.scan_batch(np.nan_to_num(X), Q, k, valid=valid[:10000])
        for i in range(32):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
        # planes follow later appends (the partial last tile is re-tiled) ...
        idx.append(raw[10000:], valid=valid[10000:], normalize=True)
        X = idx.read_rows(0, n)
        ids, sc, nf = idx.search_batch(Q, k)
        r = oracle_c.scan_batch(np.nan_to_num(X), Q, k, valid=valid)
        for i in range(32):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
        # ... and tombstones
        dead = np.unique(ids[:, 0])
        l0 = idx.launch_count
        ids, sc, nf = idx.search_batch(Q, k)
        per_batch = idx.launch_count - l0                   # kernels of one batch on up-to-date planes
        idx.tombstone(dead)
        v2 = valid.copy()
        v2[dead.astype(np.int64)] = 0
        l0 = idx.launch_count
        ids, sc, nf = idx.search_batch(Q, k)
        # the dead rows were NaN-poisoned inside the planes: no re-tiling (split_planes launch) follows a tombstone
        assert idx.launch_count - l0 == per_batch
        r = oracle_c.scan_batch(np.nan_to_num(X), Q, k, valid=v2)
        for i in range(32):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
        assert not np.isin(ids, dead).any()
        assert idx.batch_stats()[0] == 128


@pytest.mark.parametrize("growable", [False, True], ids=["fixed", "growable"])
@pytest.mark.parametrize("variant", [700, 701], ids=["single-cta", "pairs"])
def test_k3_single_pass_kernels_follow_appends_and_tombstones(sema, oracle_c, variant, growable):
    """Both single-pass kernels (the pair form reads the planes through a tensor map encoded once per planes
    allocation, also for the growable index whose planes are an address-space reservation) must see later appends,
    the re-tiled partial last tile, and rows poisoned in place by tombstones."""
    n, d, k, nq = 40000, 384, 10, 200
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    cap = 2_000_000 if growable else n
    with sema.GpuIndex(d, cap, growable=growable) as idx:
        idx.set_batch_mode(3)
        assert idx.set_scan_variant(variant) == variant
        idx.append(X[:15001], normalize=False)                      # 235 tiles: odd, the last one holds 25 rows
        ids, sc, nf = idx.search_batch(Q, k)
        r = oracle_c.scan_batch(X[:15001], Q, k)
        for i in range(0, nq, 5):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
        idx.append(X[15001:], normalize=False)
        ids, sc, nf = idx.search_batch(Q, k)
        r = oracle_c.scan_batch(X, Q, k)
        for i in range(0, nq, 5):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
        dead = np.unique(ids[:, :2])                                # every query loses its two best rows
        idx.tombstone(dead)
        valid = np.ones(n, np.uint8)
        valid[dead] = 0
        ids, sc, nf = idx.search_batch(Q, k)
        r = oracle_c.scan_batch(X, Q, k, valid=valid)
        for i in range(0, nq, 5):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
        assert not np.isin(ids, dead).any()
        served, fallbacks = idx.batch_stats()
        assert served == 3 * nq and fallbacks <= 6


def test_k3_prefetch_knob_does_not_change_results(sema, oracle_c):
    n, d, k, nq = 30000, 384, 10, 150
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        out = []
        for mode in (3, 2):
            idx.set_batch_mode(mode)
            for pf in (800, 807, 840):
                assert idx.set_scan_variant(pf) == pf
                out.append(idx.search_batch(Q, k))
        idx.set_scan_variant(800)
    for ids, sc, nf in out[1:]:
        assert np.array_equal(ids, out[0][0]) and np.array_equal(sc, out[0][1])
    r = oracle_c.scan_batch(X, Q[:8], k)
    for i in range(8):
        O.check_parity(out[0][0][i], out[0][1][i], r[0][i], r[1][i])


def test_k3_format_follows_later_appends(sema, oracle_c):
    """fp16 halves while every element is <= 1024; an append that brings larger elements re-splits every plane as
    bf16 on the next batch (and results stay exact on both sides of the switch)."""
    n, d, k, nq = 20000, 384, 10, 160
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    big = (_unit(3, 2000, d) * np.float32(5000.0)).astype(np.float32)
    with sema.GpuIndex(d, n + 2000) as idx:
        idx.append(X, normalize=False)
        ids, sc, nf = idx.search_batch(Q, k)
        assert idx.batch_precision_active == 1
        r = oracle_c.scan_batch(X, Q, k)
        for i in range(0, nq, 9):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
        idx.append(big, normalize=False)
        X2 = np.concatenate([X, big])
        for mode in (0, 2, 3):
            idx.set_batch_mode(mode)
            ids, sc, nf = idx.search_batch(Q, k)
            assert idx.batch_precision_active == 0
            idx.set_batch_mode(1)
            ids2, sc2, nf2 = idx.search_batch(Q, k)
            assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2)
        r = oracle_c.scan_batch(X2, Q[:8], k)
        for i in range(8):
            O.check_parity(ids[i], sc[i], r[0][i], r[1][i])


def test_k3_repeated_batches_are_stable(sema, oracle_c):
    """A stage is two concurrent launches on two streams joined by events, its candidate buffers are reused by the
    next batch, and the cascade re-enters the stage for sub-batches: 60 batches back to back, modes alternating, device
    buffers, every result identical to the first of its mode (any race between the streams or between consecutive
    batches shows up as a difference)."""
    import torch
    n, d, k, nq = 200_000, 384, 10, 300
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    dev = torch.device("cuda:0")
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        Qd = torch.from_numpy(Q).to(dev)
        ids_d = torch.zeros(nq * k, dtype=torch.int64, device=dev)
        sc_d = torch.zeros(nq * k, dtype=torch.float32, device=dev)
        nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)
        idx.set_stream(torch.cuda.current_stream().cuda_stream)
        first = {}
        for it in range(60):
            mode = (0, 3, 2)[it % 3]
            idx.set_batch_mode(mode)
            ids_d.zero_(); sc_d.zero_()
            idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
            torch.cuda.synchronize()
            got = (ids_d.cpu().numpy().copy(), sc_d.cpu().numpy().copy())
            if mode not in first:
                first[mode] = got
            else:
                assert np.array_equal(got[0], first[mode][0]) and np.array_equal(got[1], first[mode][1]), f"batch {it} (mode {mode}) differs"
        idx.set_stream(None)
    assert np.array_equal(first[0][0], first[2][0]) and np.array_equal(first[0][0], first[3][0])
    r = oracle_c.scan_batch(X, Q[:8], k)
    ids = first[0][0].astype(np.uint64).reshape(nq, k); sc = first[0][1].reshape(nq, k)
    for i in range(8):
        O.check_parity(ids[i], sc[i], r[0][i], r[1][i])


def test_k3_k50_with_all_neighbours_in_one_partition(sema, oracle_c):
    """k = 50 (the reference's SEARCH_RESULTS_LIMIT) with the single-pass stage's short lists (32 entries per row
    partition): when all 50 neighbours of a query sit next to each other (the chunks of one document) one partition's
    list overflows, the proof fails for that query and the cascade's next stage (long lists) or K2 answers it — exactly."""
    n, d, k, nq = 60000, 384, 50, 140
    rng = np.random.default_rng(11)
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    for qi in (0, 1, 2, 3, 4, 5):
        u = Q[qi].astype(np.float64)
        for i in range(80):                                       # 80 planted rows, contiguous, cosines 0.95 .. 0.79
            v = rng.standard_normal(d)
            v -= v.dot(u) * u
            v /= np.linalg.norm(v)
            a = 0.95 - i * 2e-3
            X[7000 * qi + 100 + i] = (a * u + np.sqrt(1 - a * a) * v).astype(np.float32)
    X = O.normalize(X)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        ids, sc, nf = idx.search_batch(Q, k)                      # automatic mode: cascade
        served, fallbacks = idx.batch_stats()
        cascaded = idx.batch_cascaded
        idx.set_batch_mode(1)
        ids2, sc2, nf2 = idx.search_batch(Q, k)
    assert served == nq and (nf == k).all()
    assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2)
    assert cascaded + fallbacks >= 6                              # the six clustered queries needed more than the short lists
    for qi in range(6):
        assert ids[qi, :50].tolist() == [7000 * qi + 100 + i for i in range(50)]
    r = oracle_c.scan_batch(X, Q[:10], k)
    for i in range(10):
        O.check_parity(ids[i], sc[i], r[0][i], r[1][i])


@pytest.mark.parametrize("nq", [20000, 40000])
def test_k3_very_large_batches(sema, oracle_c, nq):
    """More query tiles than resident clusters: the stage splits the query axis over several launches (40 000 queries
    = 157 CTAs of two tiles > 74 clusters of 2); every query still equals the single-query kernel and the oracle."""
    n, d, k = 6000, 384, 10
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        for mode in (0, 2):
            idx.set_batch_mode(mode)
            q0, _ = idx.batch_stats()
            ids, sc, nf = idx.search_batch(Q, k)
            assert idx.batch_stats()[0] - q0 == nq and (nf == k).all()
            sel = np.arange(0, nq, 397)
            for i in sel[:40]:
                r_ids, r_sc = idx.search(Q[i], k)
                assert np.array_equal(ids[i], r_ids) and np.array_equal(sc[i], r_sc)
    r = oracle_c.scan_batch(X, Q[nq - 8:], k)
    for j in range(8):
        O.check_parity(ids[nq - 8 + j], sc[nq - 8 + j], r[0][j], r[1][j])


def test_k3_auto_mode_and_unsupported_shapes_use_k2(sema, oracle_c):
    # dim 1024 is served by K2 (one pass per query); the L2 metric over unit rows by K3: same results
    X = _unit(1, 5000, 1024)
    Q = _unit(2, 9, 1024)
    with sema.GpuIndex(1024, 5000) as idx:
        idx.append(X, normalize=False)
        ids, sc, nf = idx.search_batch(Q, 10)
        assert idx.batch_stats()[0] == 0
    r = oracle_c.scan_batch(X, Q, 10)
    for i in range(9):
        O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
    X = _unit(1, 5000, 384)
    with sema.GpuIndex(384, 5000, metric=1) as idx:
        idx.append(X, normalize=False)
        Q = _unit(2, 9, 384)
        ids, sc, nf = idx.search_batch(Q, 10)
        assert idx.batch_stats()[0] == 9                  # unit-norm rows: K3 selects by dot product, re-scores as distances
    r = oracle_c.scan_batch(X, Q, 10, metric=1)
    for i in range(9):
        O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
    # dim 768 (config 5's width): K3 with the single bf16 pass, k = 100
    X = _unit(1, 40000, 768)
    Q = _unit(2, 130, 768)
    with sema.GpuIndex(768, 40000) as idx:
        idx.append(X, normalize=False)
        ids, sc, nf = idx.search_batch(Q, 100)
        assert idx.batch_stats() == (130, 0)
        idx.set_batch_mode(1)
        ids2, sc2, nf2 = idx.search_batch(Q, 100)
    assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2)
    r = oracle_c.scan_batch(X, Q, 100)
    for i in range(130):
        O.check_parity(ids[i], sc[i], r[0][i], r[1][i])
    X = _unit(1, 5000, 384)
    Q = _unit(2, 9, 384)
    with sema.GpuIndex(384, 5000) as idx:                 # auto mode, nq >= 4, cosine -> K3
        idx.append(X, normalize=False)
        ids, sc, nf = idx.search_batch(Q, 10)
        assert idx.batch_stats()[0] == 9
    r = oracle_c.scan_batch(X, Q, 10)
    for i in range(9):
        O.check_parity(ids[i], sc[i], r[0][i], r[1][i])


# ---------------------------------------------------------------- fused shard exchange
def test_shard_group_world1_equals_plain_search(sema, oracle_c):
    # the fused publish / wait / merge path with a single rank (real multi-rank runs: bench.py --gpus N)
    n, d, k = 40000, 384, 10
    X = _unit(1, n, d)
    Q = _unit(2, 5, d)
    with sema.GpuIndex(d, n) as idx:
        idx.set_row_base(7_000_000)
        idx.append(X, normalize=False)
        g = sema.ShardGroup(idx, 1, 0)
        try:
            for q in Q:
                ids, sc = g.search(q, k)
                r_ids, r_sc = oracle_c.scan(X, q, k, id_base=7_000_000)
                O.check_parity(ids, sc, r_ids, r_sc)
                ids100, sc100 = g.search(q, 100)
                r_ids, r_sc = oracle_c.scan(X, q, 100, id_base=7_000_000)
                O.check_parity(ids100, sc100, r_ids, r_sc)
            with pytest.raises(sema.SemaError):
                g.search(Q[0], 200)                      # the fused exchange covers k <= 128
        finally:
            g.close()


@pytest.mark.parametrize("d,k", [(384, 10), (384, 50), (768, 100), (130, 10)])
def test_single_process_shard_group_over_all_visible_gpus(sema, oracle_c, d, k):
    """sema_shard_group_create_local: ONE process, ONE handle, one shard per visible GPU (the reference is a
    single process owning a single StorageManager, src/storage/mod.rs:13-16).  Runs with however many GPUs
    the box shows (1 on the default test box; `gpurun --gpus N` exercises the peer exchange for real)."""
    from sema_b200 import _lib
    from sema_b200.sharded import shard_range
    G = min(_lib.lib().sema_device_count(), 8)
    n = 60011
    X = _unit(1, n, d)
    valid = np.ones(n, np.uint8)
    valid[::13] = 0
    Q = _unit(2, 6, d)
    shards = []
    try:
        for g in range(G):
            lo, hi = shard_range(n, G, g)
            idx = sema.GpuIndex(d, max(hi - lo, 1), device=g)
            idx.set_row_base(lo)
            if hi > lo:
                idx.append(X[lo:hi], valid=valid[lo:hi], normalize=False)
            shards.append(idx)
        grp = sema.ShardGroup.local(shards)
        try:
            for q in Q:
                ids, sc = grp.search(q, k)                           # one host call, G fused kernels
                r_ids, r_sc = oracle_c.scan(X, q, k, valid=valid)
                O.check_parity(ids, sc, r_ids, r_sc)
            if d in (384, 768):                                      # submit / collect: several searches in flight
                import ctypes as C
                ids_h, sc_h = np.zeros(k, np.uint64), np.zeros(k, np.float32)
                tickets = [grp.submit_ptr(C.c_void_p(Q[i].ctypes.data), k) for i in range(4)]
                for i, t in enumerate(tickets):
                    nf = grp.collect_ptr(t, C.c_void_p(ids_h.ctypes.data), C.c_void_p(sc_h.ctypes.data))
                    r_ids, r_sc = oracle_c.scan(X, Q[i], k, valid=valid)
                    O.check_parity(ids_h[:nf], sc_h[:nf], r_ids, r_sc)
            # tombstones on one shard are seen by the next group search
            dead = oracle_c.scan(X, Q[0], k, valid=valid)[0][:3]
            for row in dead:
                g = next(i for i in range(G) if shard_range(n, G, i)[0] <= row < shard_range(n, G, i)[1])
                shards[g].tombstone(np.array([row - shard_range(n, G, g)[0]], dtype=np.uint64))
            v2 = valid.copy()
            v2[dead.astype(np.int64)] = 0
            ids, sc = grp.search(Q[0], k)
            r_ids, r_sc = oracle_c.scan(X, Q[0], k, valid=v2)
            O.check_parity(ids, sc, r_ids, r_sc)
            with pytest.raises(sema.SemaError):
                grp.search(Q[0], 200)                                # the fused exchange covers k <= 128
            import torch
            qd = torch.from_numpy(Q[0]).cuda()
            out = torch.zeros(k, dtype=torch.int64, device="cuda"), torch.zeros(k, device="cuda"), torch.zeros(1, dtype=torch.int32, device="cuda")
            with pytest.raises(sema.SemaError):                      # device-resident queries live on one GPU
                grp.search_device(qd.data_ptr(), k, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
        finally:
            grp.close()
    finally:
        for idx in shards:
            idx.close()


def test_single_process_shard_group_rejects_bad_arguments(sema):
    a = sema.GpuIndex(384, 16)
    b = sema.GpuIndex(384, 16)
    c = sema.GpuIndex(768, 16)
    try:
        with pytest.raises(sema.SemaError):
            sema.ShardGroup.local([a, b])                            # two shards on one device
        if __import__("sema_b200")._lib.lib().sema_device_count() >= 2:
            c2 = sema.GpuIndex(768, 16, device=1)
            try:
                with pytest.raises(sema.SemaError):
                    sema.ShardGroup.local([a, c2])                   # dims differ
            finally:
                c2.close()
    finally:
        a.close(); b.close(); c.close()


# ---------------------------------------------------------------- next rows: compaction, disk cache, Arrow
def test_compaction_drops_dead_rows_and_keeps_order(sema, oracle_c):
    n, d, k = 20000, 384, 10
    X = _unit(1, n, d)
    valid = np.ones(n, np.uint8)
    valid[::9] = 0
    Q = _unit(2, 6, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, valid=valid, normalize=False)
        dead = np.arange(5, n, 13, dtype=np.uint64)
        idx.tombstone(dead)
        valid[dead.astype(np.int64)] = 0
        mapping = idx.compact()
        live_rows = np.nonzero(valid)[0]
        assert len(idx) == len(live_rows)
        assert np.array_equal(mapping[live_rows], np.arange(len(live_rows), dtype=np.uint64))
        assert (mapping[valid == 0] == np.uint64(2**64 - 1)).all()
        assert np.array_equal(idx.read_rows(0, len(live_rows)), X[live_rows])
        Xc = X[live_rows]
        for q in Q:
            ids, sc = idx.search(q, k)
            r_ids, r_sc = oracle_c.scan(Xc, q, k)
            O.check_parity(ids, sc, r_ids, r_sc)
        idx.set_batch_mode(2)                               # K3 planes are re-tiled after the move
        bids, bsc, bnf = idx.search_batch(Q, k)
        for i, q in enumerate(Q):
            r_ids, r_sc = oracle_c.scan(Xc, q, k)
            O.check_parity(bids[i], bsc[i], r_ids, r_sc)
        first = idx.append(X[:100], normalize=False)        # the freed capacity is usable again
        assert first == len(live_rows)


@pytest.mark.parametrize("pattern", ["file_then_scatter", "scatter_only", "tail_only", "head_block"])
def test_compaction_over_many_chunks_bounce_and_direct(sema, oracle_c, pattern):
    """Compaction is an ordered gather in chunks of 65 536 rows: a chunk goes through the bounce buffer while fewer
    than a chunk's worth of rows has been dropped before it, and straight into place afterwards.  300 000 rows cover
    an untouched prefix, bounce chunks, the switch to direct chunks and a ragged last chunk."""
    n, d, k = 300000, 64, 10
    with sema.GpuIndex(d, n) as idx:
        idx.append_synthetic(seed=1, row0=0, n=n, normalize=True)
        X = idx.read_rows(0, n)
        rng = np.random.default_rng(11)
        if pattern == "file_then_scatter":      # prefix kept, 70 001 rows in one block (direct from there on), then 3 % scattered
            dead = np.unique(np.concatenate([np.arange(40000, 110001), rng.choice(n, n // 33, replace=False)]))
        elif pattern == "scatter_only":         # 30 % scattered: bounce chunks first, direct once 65 536 rows are gone
            dead = np.sort(rng.choice(n, 3 * n // 10, replace=False))
        elif pattern == "tail_only":            # nothing moves
            dead = np.arange(n - 1234, n)
        else:                                   # the first 200 000 rows go: every chunk is direct, sources far ahead
            dead = np.arange(0, 200000)
        idx.tombstone(dead.astype(np.uint64))
        keep = np.ones(n, bool)
        keep[dead] = False
        mapping = idx.compact()
        live = int(keep.sum())
        assert len(idx) == live
        assert np.array_equal(mapping[keep], np.arange(live, dtype=np.uint64))
        assert (mapping[~keep] == np.uint64(2**64 - 1)).all()
        assert np.array_equal(idx.read_rows(0, live), X[keep])           # every surviving row, in order, bit for bit
        q = _unit(2, 1, d)[0]
        ids, sc = idx.search(q, k)
        r_ids, r_sc = oracle_c.scan(X[keep], q, k)
        O.check_parity(ids, sc, r_ids, r_sc)


def test_compact_keep_follows_the_callers_flags(sema, oracle_c):
    # the store keeps chunks whose embedding failed (null vector: the LIKE fallback still finds them) and drops live ones
    n, d, k = 70000, 384, 10
    X = _unit(1, n, d)
    valid = np.ones(n, np.uint8)
    valid[3::10] = 0                                   # null vectors
    keep = np.ones(n, np.uint8)
    keep[1::4] = 0                                     # dropped whatever their validity
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, valid=valid, normalize=False)
        live, mapping = idx.compact_keep(keep, want_map=True)
        kept = np.nonzero(keep)[0]
        assert live == len(kept) == len(idx)
        assert np.array_equal(mapping[kept], np.arange(live, dtype=np.uint64)) and (mapping[keep == 0] == np.uint64(2**64 - 1)).all()
        q = _unit(2, 1, d)[0]
        ids, sc = idx.search(q, k)
        r_ids, r_sc = oracle_c.scan(X[kept], q, k, valid=valid[kept])      # kept null rows still never rank
        O.check_parity(ids, sc, r_ids, r_sc)
        assert idx.compact_keep(np.zeros(live, np.uint8)) == 0 and len(idx) == 0
        assert idx.search(q, k)[0].size == 0
        first = idx.append(X[:500], normalize=False)
        assert first == 0
        ids, sc = idx.search(q, k)
        r_ids, r_sc = oracle_c.scan(X[:500], q, k)
        O.check_parity(ids, sc, r_ids, r_sc)


def test_save_and_load_round_trip(sema, oracle_c, tmp_path):
    n, d, k = 5000, 130, 10                                 # padded rows (dim % 4 != 0)
    X = _unit(1, n, d)
    valid = np.ones(n, np.uint8)
    valid[::17] = 0
    q = _unit(2, 1, d)[0]
    path = str(tmp_path / "chunks.semaidx")
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, valid=valid, normalize=False)
        want = idx.search(q, k)
        idx.save(path)
    assert os.path.getsize(path) == 64 + n + n * d * 4
    idx2 = sema.GpuIndex.load(path, capacity_rows=n + 10)
    try:
        got = idx2.search(q, k)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
        assert len(idx2) == n and idx2.capacity == n + 10
        r_ids, r_sc = oracle_c.scan(X, q, k, valid=valid)
        O.check_parity(got[0], got[1], r_ids, r_sc)
    finally:
        idx2.close()
    with pytest.raises(sema.SemaError):
        sema.GpuIndex.load(str(tmp_path / "missing.semaidx"))
    # a damaged file is an error code, never an allocation failure thrown across the C boundary:
    blob = open(path, "rb").read()
    cases = {
        "truncated": blob[:len(blob) - 1000],
        "trailing": blob + b"\0" * 8,
        "huge_rows": blob[:16] + (2 ** 40).to_bytes(8, "little") + blob[24:],     # n_rows far beyond the file
        "rows_over_32bit_ids": blob[:16] + (2 ** 33).to_bytes(8, "little") + blob[24:],
        "zero_dim": blob[:8] + (0).to_bytes(4, "little") + blob[12:],
        "bad_metric": blob[:12] + (7).to_bytes(4, "little") + blob[16:],
        "future_version": blob[:24] + (99).to_bytes(4, "little") + blob[28:],
        "header_only_half": blob[:30],
    }
    for name, data in cases.items():
        bad = str(tmp_path / f"{name}.semaidx")
        open(bad, "wb").write(data)
        with pytest.raises(sema.SemaError) as ei:
            sema.GpuIndex.load(bad)
        assert ei.value.code == -1, name                                    # SEMA_ERR_INVALID
    # files written before the version / byte-order fields existed (both zero) still load
    old = str(tmp_path / "v0.semaidx")
    open(old, "wb").write(blob[:24] + b"\0" * 8 + blob[32:])
    idx3 = sema.GpuIndex.load(old)
    try:
        got = idx3.search(q, k)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    finally:
        idx3.close()


def test_arrow_fixed_size_list_import(sema, oracle_c):
    # the reference's vector column: FixedSizeList<Float32,384>, nullable (src/storage/lance_indexer.rs:41-45, 75-76)
    import pyarrow as pa
    from sema_b200.arrow_import import append_arrow
    n, d = 3000, 384
    X = _unit(1, n, d)
    valid = np.ones(n, np.uint8)
    valid[::5] = 0
    vectors = [X[i].tolist() if valid[i] else None for i in range(n)]
    arr = pa.array(vectors, type=pa.list_(pa.float32(), d))
    assert arr.null_count == int((valid == 0).sum())
    q = _unit(2, 1, d)[0]
    with sema.GpuIndex(d, 2 * n) as idx:
        assert append_arrow(idx, arr) == 0
        assert append_arrow(idx, arr.slice(100, 50)) == n       # a sliced column (non-zero offset)
        ids, sc = idx.search(q, 10)
    Xall = np.concatenate([X, X[100:150]])
    vall = np.concatenate([valid, valid[100:150]])
    r_ids, r_sc = oracle_c.scan(Xall, q, 10, valid=vall)
    O.check_parity(ids, sc, r_ids, r_sc)


# ---------------------------------------------------------------- randomized sweep
def test_randomized_shapes_against_oracle(sema, oracle_c):
    """Seeded random sweep over ragged sizes, padded dims, both metrics, null rows, k up to 300,
    single and batched searches: every result must satisfy the BASELINE acceptance rule."""
    rng = np.random.default_rng(20261018)
    dims = [4, 7, 50, 130, 384, 384, 384, 768, 1000]
    for case in range(36):
        d = int(rng.choice(dims))
        n = int(rng.integers(1, 6000))
        k = int(rng.choice([1, 3, 10, 32, 33, 50, 64, 100, 128, 129, 300]))
        metric = int(rng.integers(0, 2))
        X = _unit(100 + case, n, d)
        valid = (rng.random(n) > 0.1).astype(np.uint8)
        if rng.random() < 0.3:                               # duplicates -> exact ties
            src = rng.integers(0, n, size=max(1, n // 20))
            dst = rng.integers(0, n, size=len(src))
            X[dst] = X[src]
        nq = int(rng.choice([1, 2, 5, 9]))
        Q = _unit(200 + case, nq, d)
        if rng.random() < 0.5:
            Q[0] = X[int(rng.integers(0, n))]
        with sema.GpuIndex(d, n + 3, metric=metric) as idx:
            half = n // 2
            idx.append(X[:half], valid=valid[:half], normalize=False)
            idx.append(X[half:], valid=valid[half:], normalize=False)
            for i in range(nq):
                ids, sc = idx.search(Q[i], k)
                r_ids, r_sc = oracle_c.scan(X, Q[i], k, metric, valid)
                assert len(ids) == len(r_ids), (case, d, n, k, metric)
                O.check_parity(ids, sc, r_ids, r_sc)
            bids, bsc, bnf = idx.search_batch(Q, k)
            for i in range(nq):
                r_ids, r_sc = oracle_c.scan(X, Q[i], k, metric, valid)
                assert bnf[i] == len(r_ids)
                O.check_parity(bids[i, :bnf[i]], bsc[i, :bnf[i]], r_ids, r_sc)


def test_shard_group_with_an_empty_shard(sema):
    # a rank whose row range is empty still takes part in the exchange (n = 0: nothing scanned)
    with sema.GpuIndex(384, 8) as idx:
        g = sema.ShardGroup(idx, 1, 0)
        try:
            ids, sc = g.search(np.ones(384, np.float32) / np.sqrt(384.0), 10)
            assert len(ids) == 0 and len(sc) == 0
            idx.append(_unit(1, 3, 384), normalize=False)
            ids, sc = g.search(_unit(1, 3, 384)[1], 10)
            assert ids.tolist()[0] == 1 and len(ids) == 3
        finally:
            g.close()


@pytest.mark.parametrize("prec", [0, 1], ids=["fp16", "bf16"])
def test_k3_precision_cascade_in_automatic_mode(sema, oracle_c, prec):
    """Automatic mode: single bf16 pass first; queries whose proof fails under its loose bound move to
    the bf16x3 stage; only what neither proves goes through K2.  Dense neighbourhood: 200 rows whose
    cosine to the query steps down by 1e-4 from 0.9 — the 10th and the 32nd candidate are 2.2e-3 apart
    (inside the single-pass bound 8.5e-3, outside the bf16x3 bound 2.5e-4)."""
    n, d, k = 30000, 384, 10
    rng = np.random.default_rng(7)
    X = _unit(1, n, d)
    Q = _unit(2, 8, d)
    u = Q[0].astype(np.float64)
    for i in range(200):
        v = rng.standard_normal(d)
        v -= v.dot(u) * u
        v /= np.linalg.norm(v)
        a = 0.9 - i * 1e-4
        X[1000 + i] = (a * u + np.sqrt(1 - a * a) * v).astype(np.float32)   # contiguous: more than KC of them per row partition
    X = O.normalize(X)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        idx.set_batch_precision(prec)
        ids, sc, nf = idx.search_batch(Q, k)                  # automatic mode
        served, fallbacks = idx.batch_stats()
        cascaded = idx.batch_cascaded
        idx.set_batch_mode(3)
        ids1, sc1, _ = idx.search_batch(Q, k)                 # single pass only: the dense query falls back to K2
        fb1 = idx.batch_stats()[1] - fallbacks
    assert served == 8
    r_ids, r_sc, _ = oracle_c.scan_batch(X, Q, k)
    for i in range(8):
        O.check_parity(ids[i], sc[i], r_ids[i], r_sc[i])
    assert ids[0].tolist() == [1000 + i for i in range(10)]
    assert np.array_equal(ids, ids1) and np.array_equal(sc, sc1)
    assert fb1 >= 1                                           # the loose bound alone cannot prove query 0
    # with fewer than 4 open queries the cascade goes straight to K2; with >= 4 it uses the bf16x3 stage
    assert cascaded in (0, fallbacks) or fallbacks == 0


@pytest.mark.parametrize("prec", [0, 1], ids=["fp16", "bf16"])
def test_k3_cascade_uses_bf16x3_stage_for_many_dense_queries(sema, oracle_c, prec):
    n, d, k = 30000, 384, 10
    rng = np.random.default_rng(8)
    X = _unit(1, n, d)
    Q = _unit(2, 16, d)
    for qi in range(6):                                       # six queries with a dense neighbourhood each
        u = Q[qi].astype(np.float64)
        for i in range(120):
            v = rng.standard_normal(d)
            v -= v.dot(u) * u
            v /= np.linalg.norm(v)
            a = 0.9 - i * 1e-4
            X[2000 * qi + i + 3] = (a * u + np.sqrt(1 - a * a) * v).astype(np.float32)   # contiguous: one row partition
    X = O.normalize(X)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        idx.set_batch_precision(prec)
        ids, sc, nf = idx.search_batch(Q, k)
        served, fallbacks = idx.batch_stats()
        cascaded = idx.batch_cascaded
    assert served == 16 and cascaded >= 6 and fallbacks == 0  # proven by the bf16x3 stage, no K2 pass needed
    r_ids, r_sc, _ = oracle_c.scan_batch(X, Q, k)
    for i in range(16):
        O.check_parity(ids[i], sc[i], r_ids[i], r_sc[i])


# ---------------------------------------------------------------- query streams (chained K2 launches) and the host-query path
def _stream_search(sema, idx, Q, k, group=None):
    import torch
    dev = torch.device("cuda", idx.device)
    nq = len(Q)
    Qd = torch.from_numpy(np.ascontiguousarray(Q)).to(dev)
    ids = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    sc = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    nf = torch.zeros(nq, dtype=torch.int32, device=dev)
    idx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    try:
        (group or idx).search_stream_device(Qd.data_ptr(), nq, k, ids.data_ptr(), sc.data_ptr(), nf.data_ptr())
        torch.cuda.synchronize()
    finally:
        idx.set_stream(None)
    return ids.cpu().numpy().astype(np.uint64), sc.cpu().numpy(), nf.cpu().numpy()


STREAM_CASES = [
    # n, d, k, nq, metric
    (1, 384, 10, 3, 0), (31, 384, 10, 5, 0), (33, 384, 50, 7, 1), (4737, 384, 10, 40, 0), (4737, 768, 100, 9, 0),
    (150001, 384, 10, 64, 0), (150001, 384, 128, 6, 1), (60000, 768, 10, 33, 0),
    (3000, 130, 10, 6, 0),      # generic kernel: the stream is not chained
    (6000, 384, 300, 3, 0),     # k > 128: multi-pass launches are never chained
]


@pytest.mark.parametrize("n,d,k,nq,metric", STREAM_CASES)
def test_query_stream_matches_oracle(sema, oracle_c, n, d, k, nq, metric):
    """sema_index_search_stream_device: nq scans chained with programmatic dependent launch; every
    query's result must be the oracle's (and therefore the one-query-per-call result)."""
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n, metric=metric) as idx:
        idx.append(X, normalize=False)
        ids, sc, nf = _stream_search(sema, idx, Q, k)
        for i in range(nq):
            r_ids, r_sc = oracle_c.scan(X, Q[i], k, metric)
            assert nf[i] == len(r_ids) == min(k, n)
            O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)
            one_ids, one_sc = idx.search(Q[i], k)
            assert np.array_equal(one_ids, ids[i, :nf[i]]) and np.array_equal(one_sc, sc[i, :nf[i]])


def test_query_stream_chained_equals_unchained_bitwise(sema):
    # the dynamic tile scheduler changes which block scans which tile, never the result
    n, d, k, nq = 300007, 384, 10, 200
    with sema.GpuIndex(d, n) as idx:
        idx.append_synthetic(seed=1, row0=0, n=n, normalize=True)
        Q = _unit(2, nq, d)
        a = _stream_search(sema, idx, Q, k)
        idx.set_scan_variant(600)              # unchained
        b = _stream_search(sema, idx, Q, k)
        idx.set_scan_variant(601)
        idx.set_scan_variant(2)                # register-fed kernel, static schedule
        c = _stream_search(sema, idx, Q, k)
    for x, y, z in zip(a, b, c):
        assert np.array_equal(x, y) and np.array_equal(x, z)


def test_query_stream_with_nulls_and_tombstones(sema, oracle_c):
    n, d, k, nq = 20011, 384, 10, 12
    X = _unit(1, n, d)
    valid = np.ones(n, np.uint8)
    valid[::7] = 0
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, valid=valid, normalize=False)
        dead = np.arange(1, n, 11, dtype=np.uint64)
        idx.tombstone(dead)
        valid[dead] = 0
        ids, sc, nf = _stream_search(sema, idx, Q, k)
        for i in range(nq):
            r_ids, r_sc = oracle_c.scan(X, Q[i], k, 0, valid)
            O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)


def test_shard_group_stream_world1(sema, oracle_c):
    n, d, k, nq = 40000, 384, 10, 25
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.set_row_base(1_000_000)
        idx.append(X, normalize=False)
        g = sema.ShardGroup(idx, 1, 0)
        try:
            ids, sc, nf = _stream_search(sema, idx, Q, k, group=g)
            for i in range(nq):
                r_ids, r_sc = oracle_c.scan(X, Q[i], k, id_base=1_000_000)
                O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)
            one_ids, _ = g.search(Q[3], k)          # host call after a stream: sequence numbers stay in step
            assert np.array_equal(one_ids, ids[3, :nf[3]])
        finally:
            g.close()


# ---- the whole stream as ONE persistent launch (k2_stream.cuh): producer never drains, finisher warp merges
PERSISTENT_CASES = [
    # n, d, k, nq, metric — tiny corpora (fewer tiles than blocks, blocks that see no tile of a query), both dims,
    # every list width (k <= 16 by selection rounds, M = 1 / 2 / 4 by bitonic merges), odd and even query counts
    (1, 384, 10, 3, 0), (31, 384, 10, 5, 0), (33, 384, 50, 7, 1), (4737, 384, 10, 40, 0), (4737, 768, 100, 9, 0),
    (150001, 384, 10, 64, 0), (150001, 384, 128, 6, 1), (60000, 768, 10, 33, 0), (20000, 384, 16, 2, 0),
    (20000, 384, 17, 11, 0), (9000, 768, 64, 5, 1),
]


@pytest.mark.parametrize("n,d,k,nq,metric", PERSISTENT_CASES)
def test_persistent_stream_matches_oracle_and_single_calls(sema, oracle_c, n, d, k, nq, metric):
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n, metric=metric) as idx:
        idx.append(X, normalize=False)
        assert idx.set_scan_variant(901) == 901          # one persistent launch whatever the corpus size
        before = idx.launch_count
        ids, sc, nf = _stream_search(sema, idx, Q, k)
        assert idx.launch_count - before == 1
        for i in range(nq):
            r_ids, r_sc = oracle_c.scan(X, Q[i], k, metric)
            assert nf[i] == len(r_ids) == min(k, n)
            O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)
            one_ids, one_sc = idx.search(Q[i], k)
            assert np.array_equal(one_ids, ids[i, :nf[i]]) and np.array_equal(one_sc, sc[i, :nf[i]])


def test_persistent_stream_equals_launch_per_query_bitwise(sema):
    # 300 007 rows x 384: by default k = 10 streams take the persistent launch (a scan is long against the finisher
    # warp's work), k = 50 streams need a longer scan and stay one launch per query
    n, d, nq = 300007, 384, 200
    res = {}
    with sema.GpuIndex(d, n) as idx:
        idx.append_synthetic(seed=1, row0=0, n=n, normalize=True)
        Q = _unit(2, nq, d)
        for k in (10, 50):
            before = idx.launch_count
            a = _stream_search(sema, idx, Q, k)
            assert idx.launch_count - before == (1 if k == 10 else nq)
            idx.set_scan_variant(902)              # one launch per query, chained
            before = idx.launch_count
            b = _stream_search(sema, idx, Q, k)
            assert idx.launch_count - before == nq
            idx.set_scan_variant(901)              # persistent whatever the size
            before = idx.launch_count
            c = _stream_search(sema, idx, Q, k)
            assert idx.launch_count - before == 1
            idx.set_scan_variant(900)
            for x, y, z in zip(a, b, c):
                assert np.array_equal(x, y) and np.array_equal(x, z)
            res[k] = a
        # persistent launches back to back on one handle: the control words are re-zeroed on the stream every time
        idx.set_scan_variant(901)
        for _ in range(3):
            c = _stream_search(sema, idx, Q[:7], 10)
            assert np.array_equal(c[0], res[10][0][:7]) and np.array_equal(c[1], res[10][1][:7])
    assert np.array_equal(res[10][0], res[50][0][:, :10]) and np.array_equal(res[10][1], res[50][1][:, :10])


def test_persistent_stream_with_nulls_tombstones_and_shard_exchange(sema, oracle_c):
    n, d, k, nq = 20011, 384, 10, 13
    X = _unit(1, n, d)
    valid = np.ones(n, np.uint8)
    valid[::7] = 0
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.set_row_base(1_000_000)
        idx.append(X, valid=valid, normalize=False)
        dead = np.arange(1, n, 11, dtype=np.uint64)
        idx.tombstone(dead)
        valid[dead] = 0
        idx.set_scan_variant(901)
        ids, sc, nf = _stream_search(sema, idx, Q, k)
        g = sema.ShardGroup(idx, 1, 0)             # world = 1: the fused exchange runs against the rank's own buffer
        try:
            gids, gsc, gnf = _stream_search(sema, idx, Q, k, group=g)
            gids2, gsc2, gnf2 = _stream_search(sema, idx, Q[:6], k, group=g)   # sequence numbers carry on
            one_ids, _ = g.search(Q[3], k)         # host call after a persistent stream: still in step
        finally:
            g.close()
        for i in range(nq):
            r_ids, r_sc = oracle_c.scan(X, Q[i], k, 0, valid, id_base=1_000_000)
            O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)
        assert np.array_equal(ids, gids) and np.array_equal(sc, gsc) and np.array_equal(nf, gnf)
        assert np.array_equal(ids[:6], gids2) and np.array_equal(sc[:6], gsc2)
        assert np.array_equal(one_ids, ids[3, :nf[3]])


def test_two_persistent_streams_on_one_device_at_once(sema):
    """Two handles on one GPU, each running a persistent (cooperative, one CTA per SM) stream launch on its own CUDA
    stream with nothing synchronising in between: the two grids cannot be co-resident, so the launches must be
    serialised by the device rather than interleaved into a deadlock; both must give the single-call results."""
    import torch
    dev = torch.device("cuda:0")
    n, d, k, nq = 120001, 384, 10, 60
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as a, sema.GpuIndex(d, n) as b:
        a.append_synthetic(seed=1, row0=0, n=n, normalize=True)
        b.append_synthetic(seed=7, row0=0, n=n, normalize=True)
        ref = [[h.search(q, k) for q in Q] for h in (a, b)]
        s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        Qd = torch.from_numpy(np.ascontiguousarray(Q)).to(dev)
        out = [(torch.zeros((nq, k), dtype=torch.int64, device=dev), torch.zeros((nq, k), dtype=torch.float32, device=dev),
                torch.zeros(nq, dtype=torch.int32, device=dev)) for _ in range(2)]
        torch.cuda.synchronize()
        a.set_scan_variant(901)
        b.set_scan_variant(901)
        a.set_stream(s1.cuda_stream)
        b.set_stream(s2.cuda_stream)
        try:
            for _ in range(3):          # A, B, A, B, ... queued without a synchronise
                for h, (ids, sc, nf) in zip((a, b), out):
                    h.search_stream_device(Qd.data_ptr(), nq, k, ids.data_ptr(), sc.data_ptr(), nf.data_ptr())
            torch.cuda.synchronize()
        finally:
            a.set_stream(None)
            b.set_stream(None)
        for r, (ids, sc, nf) in zip(ref, out):
            ids, sc, nf = ids.cpu().numpy().astype(np.uint64), sc.cpu().numpy(), nf.cpu().numpy()
            for i in range(nq):
                assert nf[i] == k and np.array_equal(ids[i], r[i][0]) and np.array_equal(sc[i], r[i][1])


def test_persistent_stream_while_rows_are_being_appended(sema, oracle_c):
    """config 5's interleave with the persistent kernel: the stream scans the snapshot it started with while K1 keeps
    appending on the ingest stream (its blocks compete for SMs with the cooperative grid)."""
    d, k, nq, n0, batch, nb = 384, 10, 40, 110000, 20000, 6
    X = _unit(1, n0 + batch * nb, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n0 + batch * nb) as idx:
        idx.append(X[:n0], normalize=False)
        idx.set_scan_variant(901)
        for b in range(nb):
            lo = n0 + b * batch
            idx.append(X[lo:lo + batch], normalize=False, asynchronous=True)
            ids, sc, nf = _stream_search(sema, idx, Q, k)
            snap = idx.last_snapshot
            assert n0 <= snap <= lo + batch and (snap - n0) % batch == 0
            for i in (0, nq // 2, nq - 1):
                r_ids, r_sc = oracle_c.scan(X[:snap], Q[i], k)
                O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)
        idx.flush()
        ids, sc, nf = _stream_search(sema, idx, Q, k)
        r_ids, r_sc = oracle_c.scan(X, Q[5], k)
        O.check_parity(ids[5, :nf[5]], sc[5, :nf[5]], r_ids, r_sc)


@pytest.mark.parametrize("d,k", [(384, 10), (384, 50), (384, 128), (768, 100)])
def test_host_query_path_equals_staged_path(sema, oracle_c, d, k):
    """sema_index_search: query by kernel parameter + results to mapped host memory (default) against
    the staged H2D / D2H path, many calls in a row (completion-flag sequencing)."""
    n = 30011
    X = _unit(1, n, d)
    Q = _unit(2, 40, d)
    for metric in (0, 1):
        with sema.GpuIndex(d, n, metric=metric) as idx:
            idx.append(X, normalize=False)
            fast = [idx.search(q, k) for q in Q]
            idx.set_scan_variant(500)
            slow = [idx.search(q, k) for q in Q]
            idx.set_scan_variant(501)
            for (a, b), (c, e), q in zip(fast, slow, Q):
                assert np.array_equal(a, c) and np.array_equal(b, e)
            r_ids, r_sc = oracle_c.scan(X, Q[7], k, metric)
            O.check_parity(fast[7][0], fast[7][1], r_ids, r_sc)


# ---------------------------------------------------------------- K0: mean pooling (the step before the path)
def _tokens(seed, n, seq, d):
    return (O.synth(seed, 0, n * seq, d) / np.float32(65536.0)).reshape(n, seq, d)


def _masks(n, seq, rng):
    mask = np.zeros((n, seq), dtype=np.float32)
    for t in range(n):
        mask[t, :int(rng.integers(0, seq + 1))] = 1.0
    return mask


@pytest.mark.parametrize("n,seq,d", [(37, 256, 384), (5, 100, 768), (3, 7, 130), (300, 16, 384), (2, 256, 2048)])
@pytest.mark.parametrize("skip", [False, True], ids=["readall", "skipmasked"])
def test_k0_mean_pool_is_bit_identical_to_the_oracle(sema, oracle_c, n, seq, d, skip):
    # src/semantic/embeddings.rs:61-91: every sum in the reference's order -> identical bits
    rng = np.random.default_rng(5)
    tok = _tokens(31, n, seq, d)
    mask = _masks(n, seq, rng)
    mask[0] = 1.0
    if n > 2:
        mask[1] = 0.0                       # all padding: mask_sum = 0
        tok[2] = 0.0                        # all-zero text: norm = 0
    want = oracle_c.mean_pool(tok, mask)
    with sema.GpuIndex(d, 4) as idx:
        got = idx.mean_pool(tok, mask, skip_masked=skip)
    assert np.array_equal(got, want)
    if n > 2:
        assert not got[1].any() and not got[2].any()


def test_k0_fractional_masks_and_golden_fixture(sema, oracle_c):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pool384.npz"))
    with sema.GpuIndex(384, 4) as idx:
        assert np.array_equal(idx.mean_pool(g["tokens"], g["mask"]), g["pooled"])
        assert np.array_equal(idx.mean_pool(g["tokens"], g["mask"], skip_masked=True), g["pooled"])
        tok = _tokens(8, 4, 19, 384)
        mask = np.random.default_rng(3).random((4, 19)).astype(np.float32)      # the code multiplies by any value
        assert np.array_equal(idx.mean_pool(tok, mask), oracle_c.mean_pool(tok, mask))


def test_k0_pooled_append_and_pooled_query_end_to_end(sema, oracle_c):
    """embed -> mean_pool -> store, and embed -> mean_pool -> search, all on the device: the stored rows
    are the oracle's pooled vectors bit for bit and the search over them matches the oracle's scan."""
    import torch
    n, seq, d, k = 3000, 24, 384, 10
    rng = np.random.default_rng(11)
    tok = _tokens(41, n, seq, d)
    mask = _masks(n, seq, rng)
    mask[:, 0] = 1.0                                    # [CLS] is always attended
    valid = np.ones(n, np.uint8)
    valid[::13] = 0                                     # failed embeddings (lance_indexer.rs:66-70)
    X = oracle_c.mean_pool(tok, mask)
    qtok, qmask = _tokens(42, 4, seq, d), _masks(4, seq, rng)
    qmask[:, 0] = 1.0
    Q = oracle_c.mean_pool(qtok, qmask)
    dev = torch.device("cuda:0")
    tok_d, mask_d, valid_d = torch.from_numpy(tok).to(dev), torch.from_numpy(mask).to(dev), torch.from_numpy(valid).to(dev)
    qtok_d, qmask_d = torch.from_numpy(qtok).to(dev), torch.from_numpy(qmask).to(dev)
    Q_d = torch.zeros((4, d), dtype=torch.float32, device=dev)
    ids = torch.zeros((4, k), dtype=torch.int64, device=dev)
    sc = torch.zeros((4, k), dtype=torch.float32, device=dev)
    nf = torch.zeros(4, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    with sema.GpuIndex(d, n + 10) as idx:
        assert idx.append_pooled_device(tok_d.data_ptr(), mask_d.data_ptr(), n, seq, valid_d.data_ptr(), skip_masked=True) == 0
        assert len(idx) == n and idx.visible == n
        got = idx.read_rows(0, n)
        live = valid != 0
        assert np.array_equal(got[live], X[live]) and np.isnan(got[~live]).all()
        idx.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        idx.mean_pool_device(qtok_d.data_ptr(), qmask_d.data_ptr(), 4, seq, Q_d.data_ptr())
        idx.search_stream_device(Q_d.data_ptr(), 4, k, ids.data_ptr(), sc.data_ptr(), nf.data_ptr())
        torch.cuda.synchronize()
        idx.set_stream(None)
    assert np.array_equal(Q_d.cpu().numpy(), Q)
    for i in range(4):
        r_ids, r_sc = oracle_c.scan(X, Q[i], k, 0, valid)
        O.check_parity(ids[i].cpu().numpy().astype(np.uint64), sc[i].cpu().numpy(), r_ids, r_sc)


def test_query_stream_edge_cases(sema, oracle_c):
    # empty index -> n_found = 0 for every query (lance_indexer.rs:108-111); a single query; k = 1; k > rows
    d = 384
    Q = _unit(2, 5, d)
    with sema.GpuIndex(d, 100) as idx:
        ids, sc, nf = _stream_search(sema, idx, Q, 10)
        assert not nf.any()
        g = sema.ShardGroup(idx, 1, 0)
        try:
            _, _, nf = _stream_search(sema, idx, Q, 10, group=g)      # an empty shard still runs the exchange
            assert not nf.any()
            X = _unit(1, 7, d)
            idx.append(X, normalize=False)
            ids, sc, nf = _stream_search(sema, idx, Q[:1], 1)
            r_ids, r_sc = oracle_c.scan(X, Q[0], 1)
            assert nf[0] == 1 and ids[0, 0] == r_ids[0]
            ids, sc, nf = _stream_search(sema, idx, Q, 50, group=g)   # limit > rows: every row, ranked
            for i in range(5):
                r_ids, r_sc = oracle_c.scan(X, Q[i], 50)
                assert nf[i] == 7
                O.check_parity(ids[i, :7], sc[i, :7], r_ids, r_sc)
        finally:
            g.close()


def test_query_stream_interleaved_with_appends_and_host_searches(sema, oracle_c):
    # streams, host-query calls and appends alternate on one handle: counters, flags and snapshots stay consistent
    d, k = 384, 10
    X = _unit(1, 9000, d)
    Q = _unit(2, 6, d)
    with sema.GpuIndex(d, 9000) as idx:
        for step, hi in enumerate((3000, 6000, 9000)):
            idx.append(X[hi - 3000:hi], normalize=False)
            ids, sc, nf = _stream_search(sema, idx, Q, k)
            for i in range(6):
                r_ids, r_sc = oracle_c.scan(X[:hi], Q[i], k)
                O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)
                h_ids, h_sc = idx.search(Q[i], k)
                assert np.array_equal(h_ids, ids[i, :nf[i]]) and np.array_equal(h_sc, sc[i, :nf[i]])
            assert idx.last_snapshot == hi


# ---------------------------------------------------------------- growable index (virtual-memory backed)
def test_growable_index_grows_in_place_and_matches_oracle(sema, oracle_c):
    """sema_index_create_growable: max_rows reserves address space only; appends map HBM in 256 MB steps
    (174 762 rows of dim 384 each), the matrix stays contiguous and every search sees exactly the oracle's rows."""
    import torch
    d, k = 384, 10
    free0, _ = torch.cuda.mem_get_info(0)
    with sema.GpuIndex(d, 60_000_000, growable=True) as idx:          # 92 GB of address space, nothing committed
        assert idx.capacity == 60_000_000 and len(idx) == 0
        free1, _ = torch.cuda.mem_get_info(0)
        assert free0 - free1 < (1 << 30)                                # creating it costs (almost) no HBM
        ids, sc = idx.search(_unit(2, 1, d)[0], k)
        assert len(ids) == 0
        total = 0
        Q = _unit(2, 3, d)
        for n_add in (1000, 173_000, 2_000, 190_000):                   # the 2nd and 4th appends cross a 256 MB step
            idx.append_synthetic(seed=1, row0=total, n=n_add, normalize=True)
            total += n_add
            X = oracle_c.normalize(oracle_c.synth(1, 0, total, d))
            for q in Q:
                ids, sc = idx.search(q, k)
                r_ids, r_sc = oracle_c.scan(X, q, k)
                O.check_parity(ids, sc, r_ids, r_sc)
        free2, _ = torch.cuda.mem_get_info(0)
        assert free1 - free2 < 3 * (256 << 20) + (64 << 20)             # 366 k rows = 562 MB -> three steps
        # the batched tensor-core path builds its planes in a growable buffer too
        Qb = _unit(3, 9, d)
        idx.set_batch_mode(2)
        b_ids, b_sc, b_nf = idx.search_batch(Qb, k)
        assert idx.batch_stats()[0] == 9
        for i in range(9):
            r_ids, r_sc = oracle_c.scan(X, Qb[i], k)
            O.check_parity(b_ids[i, :b_nf[i]], b_sc[i, :b_nf[i]], r_ids, r_sc)
        # deletions and compaction work on the mapped prefix
        idx.tombstone(np.arange(0, 1000, dtype=np.uint64))
        idx.compact()
        assert len(idx) == total - 1000
        ids, sc = idx.search(Q[0], k)
        valid = np.ones(total, np.uint8)
        valid[:1000] = 0
        r_ids, r_sc = oracle_c.scan(X, Q[0], k, 0, valid)
        O.check_parity(ids + 1000, sc, r_ids, r_sc)                     # rows moved down by the 1000 dropped


def test_growable_index_respects_max_rows(sema):
    with sema.GpuIndex(384, 100, growable=True) as idx:
        idx.append(_unit(1, 100, 384), normalize=False)
        with pytest.raises(sema.SemaError) as e:
            idx.append(_unit(1, 1, 384), normalize=False)
        assert e.value.code == -3                                        # SEMA_ERR_CAPACITY
        ids, _ = idx.search(_unit(1, 100, 384)[5], 3)
        assert ids[0] == 5


# ---------------------------------------------------------------- sharded batched search (batched keys + batched K4)
@pytest.mark.parametrize("G,k,nq,mode,metric", [(2, 10, 9, 0, 0), (4, 100, 5, 2, 0), (8, 10, 130, 3, 0), (3, 50, 6, 1, 1), (2, 128, 3, 1, 0)])
def test_k4_batched_virtual_shards_equal_single_index(sema, oracle_c, G, k, nq, mode, metric):
    # SURVEY.md §8(e), batched: every shard contributes nq x k packed keys (K3 or the K2 loop), the
    # all-gather is a concatenation here, the batched K4 merges per query
    import torch
    n, d = 24000, 384
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    valid = np.ones(n, np.uint8)
    valid[::9] = 0
    per = n // G
    dev = torch.device("cuda:0")
    Q_dev = torch.from_numpy(Q).to(dev)
    keys = torch.zeros(G * nq * k, dtype=torch.int64, device=dev)
    ids_d = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    sc_d = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)
    shards = []
    try:
        for g in range(G):
            lo, hi = g * per, (n if g == G - 1 else (g + 1) * per)
            s = sema.GpuIndex(d, hi - lo, metric=metric)
            s.set_row_base(lo)
            s.append(X[lo:hi], valid=valid[lo:hi], normalize=False)
            s.set_batch_mode(mode)
            s.set_stream(torch.cuda.current_stream().cuda_stream)
            shards.append(s)
        for g, s in enumerate(shards):
            s.search_batch_keys_device(Q_dev.data_ptr(), nq, k, keys.data_ptr() + g * nq * k * 8)
        shards[0].merge_batch_device(keys.data_ptr(), G, nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
        torch.cuda.synchronize()
    finally:
        for s in shards:
            s.close()
    ids, sc, nf = ids_d.cpu().numpy().astype(np.uint64), sc_d.cpu().numpy(), nf_d.cpu().numpy()
    for i in range(nq):
        r_ids, r_sc = oracle_c.scan(X, Q[i], k, metric, valid)
        assert nf[i] == len(r_ids)
        O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)


def test_sharded_searcher_batch_world1(sema, oracle_c):
    # sema_b200.sharded.ShardedSearcher.search_batch with a single rank (no process group): the same
    # code path bench.py --workload batch --gpus N drives under torchrun
    from sema_b200.sharded import ShardedSearcher
    n, d, k, nq = 30000, 384, 10, 20
    X = _unit(1, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n) as idx:
        idx.set_row_base(500)
        idx.append(X, normalize=False)
        sh = ShardedSearcher(idx, None, k)
        ids, sc, nf = sh.search_batch(Q)
        idx.set_stream(None)
    for i in range(nq):
        r_ids, r_sc = oracle_c.scan(X, Q[i], k, id_base=500)
        O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)


# ---------------------------------------------------------------- asynchronous host searches (submit / collect)
def test_submit_collect_matches_synchronous_search(sema, oracle_c):
    n, d, k = 50011, 384, 10
    X = _unit(1, n, d)
    Q = _unit(2, 40, d)
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        want = [idx.search(q, k) for q in Q]
        # depth-2 pipeline: submit i+1 before collecting i
        got = []
        t_prev = idx.submit(Q[0], k)
        for i in range(1, len(Q)):
            t = idx.submit(Q[i], k)
            got.append(idx.collect(t_prev, k))
            t_prev = t
        got.append(idx.collect(t_prev, k))
        for (a, b), (c, e) in zip(got, want):
            assert np.array_equal(a, c) and np.array_equal(b, e)
        r_ids, r_sc = oracle_c.scan(X, Q[5], k)
        O.check_parity(got[5][0], got[5][1], r_ids, r_sc)
        # eight in flight, collected out of order; the ninth submit is refused until a slot is free
        tickets = [idx.submit(Q[i], k) for i in range(8)]
        with pytest.raises(sema.SemaError):
            idx.submit(Q[8], k)
        with pytest.raises(sema.SemaError):
            idx.search(Q[8], k)                          # the synchronous call needs a slot too
        for i in (3, 0, 7, 1, 2, 6, 5, 4):
            a, b = idx.collect(tickets[i], k)
            assert np.array_equal(a, want[i][0]) and np.array_equal(b, want[i][1])
        with pytest.raises(sema.SemaError):
            idx.collect(tickets[0], k)                   # not outstanding any more
        a, b = idx.search(Q[8], k)
        assert np.array_equal(a, want[8][0])


def test_submit_collect_outside_the_fast_path_and_on_an_empty_index(sema, oracle_c):
    # dim 130 (generic kernel), k > 128 (multi-pass) and an empty index answer synchronously inside submit
    X = _unit(1, 3000, 130)
    Q = _unit(2, 3, 130)
    with sema.GpuIndex(130, 3000) as idx:
        t0 = idx.submit(Q[0], 10)
        assert idx.collect(t0, 10)[0].size == 0          # nothing appended yet
        idx.append(X, normalize=False)
        ts = [idx.submit(q, 10) for q in Q]
        for q, t in zip(Q, ts):
            ids, sc = idx.collect(t, 10)
            r_ids, r_sc = oracle_c.scan(X, q, 10)
            O.check_parity(ids, sc, r_ids, r_sc)
    X = _unit(1, 5000, 384)
    with sema.GpuIndex(384, 5000) as idx:
        idx.append(X, normalize=False)
        q = _unit(2, 1, 384)[0]
        t = idx.submit(q, 300)
        ids, sc = idx.collect(t, 300)
        r_ids, r_sc = oracle_c.scan(X, q, 300)
        O.check_parity(ids, sc, r_ids, r_sc)


def test_shard_group_submit_collect_world1(sema, oracle_c):
    import ctypes as C
    n, d, k = 30000, 384, 10
    X = _unit(1, n, d)
    Q = _unit(2, 12, d)
    with sema.GpuIndex(d, n) as idx:
        idx.set_row_base(100)
        idx.append(X, normalize=False)
        g = sema.ShardGroup(idx, 1, 0)
        try:
            ids_h, sc_h = np.zeros(k, np.uint64), np.zeros(k, np.float32)
            tickets = [g.submit_ptr(C.c_void_p(Q[i].ctypes.data), k) for i in range(4)]
            for i in range(4, 12):
                tickets.append(g.submit_ptr(C.c_void_p(Q[i].ctypes.data), k))
                nf = g.collect_ptr(tickets[i - 4], C.c_void_p(ids_h.ctypes.data), C.c_void_p(sc_h.ctypes.data))
                r_ids, r_sc = oracle_c.scan(X, Q[i - 4], k, id_base=100)
                O.check_parity(ids_h[:nf].copy(), sc_h[:nf].copy(), r_ids, r_sc)
            for i in range(8, 12):
                nf = g.collect_ptr(tickets[i], C.c_void_p(ids_h.ctypes.data), C.c_void_p(sc_h.ctypes.data))
                r_ids, r_sc = oracle_c.scan(X, Q[i], k, id_base=100)
                O.check_parity(ids_h[:nf].copy(), sc_h[:nf].copy(), r_ids, r_sc)
        finally:
            g.close()


# ---------------------------------------------------------------- K3 under the reference's literal metric (squared L2)
@pytest.mark.parametrize("mode", [0, 2, 3], ids=["cascade", "bf16x3", "bf16x1"])
def test_k3_l2_metric_on_unit_rows_matches_k2_and_oracle(sema, oracle_c, mode):
    """LanceDB's default `_distance` (squared L2, ascending; src/storage/lance_indexer.rs:121-126) over the
    unit-norm rows the reference stores: the tensor-core pass selects by dot product, the candidates are
    re-scored as exact distances, and the proof accounts for min |x|^2 — same bits as K2's L2 scan."""
    n, d, k, nq = 40000, 384, 10, 70
    raw = O.synth(1, 0, n, d)
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n, metric=1) as idx:
        idx.append(raw, normalize=True)                   # K1 normalises: every row is unit-norm
        X = idx.read_rows(0, n)
        idx.set_batch_mode(mode)
        ids3, sc3, nf3 = idx.search_batch(Q, k)
        served, fallbacks = idx.batch_stats()
        assert served == nq                               # K3 really ran under the L2 metric
        idx.set_batch_mode(1)
        ids2, sc2, nf2 = idx.search_batch(Q, k)
    assert np.array_equal(ids3, ids2) and np.array_equal(sc3, sc2) and np.array_equal(nf3, nf2)
    assert np.all(np.diff(sc3, axis=1) >= 0)              # ascending distance
    for i in range(0, nq, 7):
        r_ids, r_sc = oracle_c.scan(X, Q[i], k, 1)
        O.check_parity(ids3[i, :nf3[i]], sc3[i, :nf3[i]], r_ids, r_sc)
    assert fallbacks <= max(1, nq // 20)


def test_k3_l2_metric_with_varying_norms_stays_on_k2(sema, oracle_c):
    # rows of different length: the dot product does not rank like the distance, so batches run K2
    n, d, k, nq = 20000, 384, 10, 8
    X = _unit(1, n, d) * np.linspace(0.5, 2.0, n, dtype=np.float32)[:, None]
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n, metric=1) as idx:
        idx.append(X, normalize=False)
        idx.set_batch_mode(2)
        ids, sc, nf = idx.search_batch(Q, k)
        assert idx.batch_stats()[0] == 0                  # not served by K3
    for i in range(nq):
        r_ids, r_sc = oracle_c.scan(X, Q[i], k, 1)
        O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)


def test_k3_l2_metric_with_a_zero_row_stays_exact(sema, oracle_c):
    # a zero row (an all-padding text pools to zero) has norm 0: min |x|^2 = 0 disables the dot-product selection
    n, d, k, nq = 20000, 384, 10, 8
    raw = O.synth(1, 0, n, d)
    raw[77] = 0.0
    Q = _unit(2, nq, d)
    with sema.GpuIndex(d, n, metric=1) as idx:
        idx.append(raw, normalize=True)
        X = idx.read_rows(0, n)
        idx.set_batch_mode(0)
        ids, sc, nf = idx.search_batch(Q, k)
        assert idx.batch_stats()[0] == 0
    for i in range(nq):
        r_ids, r_sc = oracle_c.scan(X, Q[i], k, 1)
        O.check_parity(ids[i, :nf[i]], sc[i, :nf[i]], r_ids, r_sc)


def test_normalised_host_queries_take_the_fast_path_with_k1_arithmetic(sema, oracle_c):
    """sema_index_set_normalize_queries(1): the host-query path normalises the query in registers with
    K1's arithmetic, so it returns the same bits as the staged path (H2D + K1 + K2) and as a batch."""
    n, d, k = 30011, 384, 10
    X = _unit(1, n, d)
    Qraw = O.synth(2, 0, 12, d)                          # un-normalised queries
    Qraw[3] = 0.0                                        # a zero query stays zero: every score is 0
    with sema.GpuIndex(d, n) as idx:
        idx.append(X, normalize=False)
        idx.set_normalize_queries(True)
        fast = [idx.search(q, k) for q in Qraw]
        idx.set_scan_variant(500)                        # staged path: K1 normalises the query on the device
        slow = [idx.search(q, k) for q in Qraw]
        idx.set_scan_variant(501)
        bids, bsc, bnf = idx.search_batch(Qraw, k)
    for i, ((a, b), (c, e)) in enumerate(zip(fast, slow)):
        assert np.array_equal(a, c) and np.array_equal(b, e), i
        assert np.array_equal(a, bids[i, :bnf[i]]) and np.array_equal(b, bsc[i, :bnf[i]]), i
    Qn = O.normalize(Qraw)
    for i in (0, 5, 11):
        r_ids, r_sc = oracle_c.scan(X, Qn[i], k)
        O.check_parity(fast[i][0], fast[i][1], r_ids, r_sc)
    assert not fast[3][1].any()

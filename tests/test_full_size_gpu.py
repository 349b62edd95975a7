"""Parity at BASELINE.json's FULL sizes, against the oracle itself (not properties, not self-comparison).

The oracle (oracle/cpu_scan.c) generates its own copy of the 10M x 384 corpus on the host (15.36 GB) and scans
it; the CUDA path generates the corpus on the device and is called through the C ABI.  Since K1 folds the
squared norm in the reference's order, the two matrices are bit-identical, which is checked on sampled
blocks before any search is compared.

  * config "10M x 384 single query" (the headline): 8 queries, k = 10 and k = 50, host call and chained stream
  * config 3 (10M x 384, 1024-query batches on the tensor cores): 64 of the 1024 queries, all three batch modes
  * config 5 (d = 768, k = 100, streaming ingest interleaved with queries): every checked query against the
    oracle on exactly the snapshot it scanned, 1M rows
Acceptance rule everywhere: oracle.check_parity (BASELINE.json: same ids and order except ties within 1e-5,
scores within 1e-5 relative).
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROWS, DIM = 10_000_000, 384


@pytest.fixture(scope="module")
def sema():
    import sema_b200
    from sema_b200 import _lib
    if _lib.lib().sema_device_count() == 0:
        pytest.fail("gpu-marked test run without a CUDA device")
    return sema_b200


@pytest.fixture(scope="module")
def full(sema, oracle_c):
    """(GPU index, oracle's host corpus, 1024 queries) for the 10M x 384 configs."""
    oracle_c.use_all_cores()
    X = np.empty((ROWS, DIM), dtype=np.float32)
    for r0 in range(0, ROWS, 1 << 20):
        m = min(1 << 20, ROWS - r0)
        oracle_c.synth(1, r0, m, DIM, out=X[r0:r0 + m])
    oracle_c.normalize_inplace(X)
    Q = oracle_c.normalize(oracle_c.synth(2, 0, 1024, DIM))
    idx = sema.GpuIndex(DIM, ROWS)
    idx.append_synthetic(seed=1, row0=0, n=ROWS, normalize=True)
    cache = {}

    def want(qi, k):
        if (qi, k) not in cache:
            cache[(qi, k)] = oracle_c.scan(X, Q[qi], k)
        return cache[(qi, k)]

    yield idx, X, Q, want
    idx.close()


def test_device_corpus_is_bit_identical_to_the_oracles(full):
    idx, X, Q, _ = full
    for first in (0, 4_999_937, ROWS - 4096):
        assert np.array_equal(idx.read_rows(first, 4096), X[first:first + 4096])
    with __import__("sema_b200").GpuIndex(DIM, 1024) as qi:                     # queries through K1 as well
        from sema_b200.synth import synth_rows
        qi.append(synth_rows(2, 0, 1024, DIM), normalize=True)
        assert np.array_equal(qi.read_rows(0, 1024), Q)


@pytest.mark.parametrize("k", [10, 50])
def test_full_size_10m_x_384_single_query_matches_oracle(full, k):
    """lance_indexer.rs:121-126 at BASELINE's headline size: host call, and the chained device stream bench.py times."""
    import torch
    idx, X, Q, want = full
    nq = 8
    for i in range(nq):
        ids, sc = idx.search(Q[i], k)                                           # sema_index_search, host buffers
        O.check_parity(ids, sc, *want(i, k))
    dev = torch.device("cuda:0")
    Qd = torch.from_numpy(Q[:nq]).to(dev)
    ids_s = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    sc_s = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    nf_s = torch.zeros(nq, dtype=torch.int32, device=dev)
    idx.set_stream(torch.cuda.current_stream().cuda_stream)
    idx.search_stream_device(Qd.data_ptr(), nq, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
    torch.cuda.synchronize()
    idx.set_stream(None)
    for i in range(nq):
        assert int(nf_s[i]) == k
        O.check_parity(ids_s[i].cpu().numpy().astype(np.uint64), sc_s[i].cpu().numpy(), *want(i, k))


@pytest.mark.parametrize("mode", [2, 3, 0], ids=["bf16x3", "bf16x1", "cascade"])
def test_full_size_config3_batched_1024_queries_match_oracle(full, mode):
    """BASELINE configs[2]: 10M x 384, one batch of 1024 queries on the tensor cores, top-10; 64 of the 1024 queries
    (spread over all eight 128-query tiles) are re-derived by the oracle over the whole corpus."""
    idx, X, Q, want = full
    k = 10
    idx.set_batch_mode(mode)
    q0, f0 = idx.batch_stats()
    ids, sc, nf = idx.search_batch(Q, k)
    q1, f1 = idx.batch_stats()
    idx.set_batch_mode(0)
    assert q1 - q0 == 1024                                                      # served by K3, not by the K2 loop
    assert f1 - f0 <= 8                                                         # (almost) no K2 fallbacks on this corpus
    assert (nf == k).all()
    for i in range(0, 1024, 16):                                                # 64 queries, 8 per query tile
        O.check_parity(ids[i], sc[i], *want(i, k))


def test_config5_streaming_ingest_768_k100_every_query_matches_oracle_on_its_snapshot(sema, oracle_c):
    """BASELINE configs[4]: d = 768, k = 100, appends of 64k-row batches from the host (H2D + K1 on the ingest
    stream) interleaved with queries on the query stream.  Each query scanned exactly `last_snapshot` rows; the
    oracle re-derives its result on exactly that prefix."""
    d, k, B, nb = 768, 100, 65536, 16
    n = B * nb                                                                  # 1 048 576 rows
    oracle_c.use_all_cores()
    import torch
    raw_pinned = torch.from_numpy(oracle_c.synth(7, 0, n, d)).pin_memory()      # pinned: the H2D copies are truly asynchronous
    raw = raw_pinned.numpy()
    X = oracle_c.normalize(raw)
    Q = oracle_c.normalize(oracle_c.synth(8, 0, 32, d))
    results = []
    with sema.GpuIndex(d, n) as idx:
        # first batch synchronously + one search: the handle's one-time allocations (cudaMalloc / cudaHostAlloc
        # synchronise the device) must not sit between the asynchronous appends and the first timed query
        idx.append(raw[:B], normalize=True)
        ids, sc = idx.search(Q[0], k)
        results.append((0, idx.last_snapshot, ids, sc))
        for b in range(1, nb):
            idx.append(raw[b * B:(b + 1) * B], normalize=True, asynchronous=True)
        i = 1
        while idx.visible < n:
            if idx.visible == 0:
                continue
            ids, sc = idx.search(Q[i % 32], k)
            results.append((i % 32, idx.last_snapshot, ids, sc))
            i += 1
        idx.flush()
        for j in range(4):                                                      # and on the complete index
            ids, sc = idx.search(Q[j], k)
            results.append((j, idx.last_snapshot, ids, sc))
        assert np.array_equal(idx.read_rows(n - 1000, 1000), X[n - 1000:])      # K1 on the ingest stream: the oracle's bits
    snaps = [s for _, s, _, _ in results]
    assert snaps == sorted(snaps) and all(s % B == 0 and 0 < s <= n for s in snaps)
    assert len(set(snaps)) >= 3, "the queries should have interleaved with the ingest"
    step = max(len(results) // 40, 1)
    checked = results[::step] + results[-4:]
    for qi, s, ids, sc in checked:
        assert len(ids) == k
        O.check_parity(ids, sc, *oracle_c.scan(X[:s], Q[qi], k))

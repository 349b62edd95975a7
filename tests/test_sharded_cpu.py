"""CPU tests of the multi-rank host logic (world_size 2, gloo): shard ranges, the packed-key
exchange format and the all-gather + merge protocol.  The per-shard scan is done by the
oracle here (no GPU); on a B200 the same protocol runs with K2 / NCCL / K4
(tests/test_gpu_parity.py::test_k4_*, bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import oracle as O
from sema_b200 import keys as K
from sema_b200.sharded import shard_range


def test_shard_range_partitions_rows():
    for n in (0, 1, 7, 1000, 10_000_000, 100_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d


@pytest.mark.parametrize("metric", [0, 1])
def test_key_roundtrip_and_order(metric):
    rng = np.random.default_rng(0)
    s = rng.standard_normal(1000).astype(np.float32)
    s[:4] = [0.0, -0.0, 1.0, 1.0]
    if metric:
        s = np.abs(s)
    ids = rng.permutation(1000).astype(np.uint64)
    keys = K.pack_keys(s, ids, metric)
    assert (keys != 0).all()
    i2, s2 = K.unpack_keys(keys, metric)
    assert np.array_equal(i2, ids) and np.array_equal(s2, s + np.float32(0.0))
    # unsigned order of keys == oracle ranking (score, then lower id)
    order = np.argsort(keys)[::-1]
    rank = -s if metric else s
    want = np.lexsort((ids, -(rank + np.float32(0.0))))
    assert np.array_equal(order, want)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, k, metric, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = O.normalize(O.synth(1, 0, n, d))          # every rank can regenerate any row
        q = O.normalize(O.synth(2, 0, 1, d))[0]
        lo, hi = shard_range(n, world, rank)
        ids, sc = O.scan(X[lo:hi], q, k, metric, id_base=lo)   # stand-in for K2 on this shard
        mine = np.zeros(k, dtype=np.uint64)
        mine[:len(ids)] = K.pack_keys(sc, ids, metric)
        local = torch.from_numpy(mine.view(np.int64))
        gathered = torch.zeros(world * k, dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, local)
        merged = K.merge_keys([gathered.numpy().view(np.uint64)], k)   # stand-in for K4
        g_ids, g_sc = K.unpack_keys(merged, metric)
        r_ids, r_sc = O.scan(X, q, k, metric)
        assert np.array_equal(g_ids, r_ids), (rank, g_ids, r_ids)
        assert np.array_equal(g_sc, r_sc)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k,metric", [(5000, 10, 0), (5000, 100, 1), (3, 10, 0)])
def test_two_rank_allgather_merge_equals_global_scan(n, k, metric):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 64, k, metric, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _worker_batch(rank, world, port, n, d, k, nq, out):
    """Batched sharded search, host logic: every rank contributes [nq][k] keys, one all-gather yields
    [world][nq][k], the merge runs per query (kernel K4's batched form does this on the device)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        X = O.normalize(O.synth(1, 0, n, d))
        Q = O.normalize(O.synth(2, 0, nq, d))
        lo, hi = shard_range(n, world, rank)
        mine = np.zeros((nq, k), dtype=np.uint64)
        for i in range(nq):
            ids, sc = O.scan(X[lo:hi], Q[i], k, 0, id_base=lo)
            mine[i, :len(ids)] = K.pack_keys(sc, ids, 0)
        gathered = torch.zeros(world * nq * k, dtype=torch.int64)
        dist.all_gather_into_tensor(gathered, torch.from_numpy(mine.view(np.int64).ravel()))
        allk = gathered.numpy().view(np.uint64).reshape(world, nq, k)
        for i in range(nq):
            g_ids, g_sc = K.unpack_keys(K.merge_keys([allk[:, i, :]], k), 0)
            r_ids, r_sc = O.scan(X, Q[i], k, 0)
            assert np.array_equal(g_ids, r_ids) and np.array_equal(g_sc, r_sc), (rank, i)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_batched_allgather_merge_equals_global_scan():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_batch, args=(r, 2, port, 3001, 64, 10, 5, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

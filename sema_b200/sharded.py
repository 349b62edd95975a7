"""Corpus-sharded search: one process per GPU, each holding a contiguous row range.

No reference analogue (the reference is a single process); this is SURVEY.md §8(e):
every rank scans its shard (K2), the per-shard top-k key lists are exchanged with one
all-gather over NVLink (torch.distributed, NCCL on GPUs / gloo in the CPU tests of the
host logic), and kernel K4 merges the G x k candidates into the global top-k on every
rank.  Global row id = shard row_base + local row.
"""
from __future__ import annotations

import numpy as np


def make_shard_group(index, dist):
    """ShardGroup over torch.distributed: the 64-byte IPC handles travel by all_gather_object."""
    from .index import ShardGroup
    world, rank = dist.get_world_size(), dist.get_rank()

    def exchange(mine: bytes):
        out = [None] * world
        dist.all_gather_object(out, mine)
        return out

    g = ShardGroup(index, world, rank, exchange)
    dist.barrier()
    return g


def shard_range(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous row range [lo, hi) of shard `rank` (ceil split; trailing shards may be short or empty)."""
    per = (n_rows + world - 1) // world
    return min(rank * per, n_rows), min((rank + 1) * per, n_rows)


class ShardedSearcher:
    """Host-facing sharded search over an already-loaded per-rank GpuIndex."""

    def __init__(self, index, dist, k: int):
        import torch
        self.torch = torch
        self.idx, self.dist, self.k = index, dist, k
        self.world = dist.get_world_size() if dist is not None else 1
        self.dev = torch.device("cuda", index.device)
        self.stream = torch.cuda.current_stream(self.dev)
        index.set_stream(self.stream.cuda_stream)
        d = index.dim
        self.q_pin = torch.zeros(d, dtype=torch.float32).pin_memory()
        self.q_dev = torch.zeros(d, dtype=torch.float32, device=self.dev)
        self.keys_local = torch.zeros(k, dtype=torch.int64, device=self.dev)
        self.keys_all = torch.zeros(self.world * k, dtype=torch.int64, device=self.dev)
        self.ids_d = torch.zeros(k, dtype=torch.int64, device=self.dev)
        self.sc_d = torch.zeros(k, dtype=torch.float32, device=self.dev)
        self.nf_d = torch.zeros(1, dtype=torch.int32, device=self.dev)
        # one pinned block for the result: [n_found | ids | scores]
        self.ids_h = torch.zeros(k, dtype=torch.int64).pin_memory()
        self.sc_h = torch.zeros(k, dtype=torch.float32).pin_memory()
        self.nf_h = torch.zeros(1, dtype=torch.int32).pin_memory()

    def search(self, q: np.ndarray):
        """Every rank calls this with the same query; every rank gets the global top-k."""
        t = self.torch
        self.q_pin.copy_(t.from_numpy(np.ascontiguousarray(q, dtype=np.float32)))
        self.q_dev.copy_(self.q_pin, non_blocking=True)
        self.idx.search_keys_device(self.q_dev.data_ptr(), self.k, self.keys_local.data_ptr())
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.keys_all, self.keys_local)
            src = self.keys_all
        else:
            src = self.keys_local
        self.idx.merge_device(src.data_ptr(), self.world, self.k, self.ids_d.data_ptr(),
                              self.sc_d.data_ptr(), self.nf_d.data_ptr())
        self.ids_h.copy_(self.ids_d, non_blocking=True)
        self.sc_h.copy_(self.sc_d, non_blocking=True)
        self.nf_h.copy_(self.nf_d, non_blocking=True)
        self.stream.synchronize()
        nf = int(self.nf_h[0])
        return self.ids_h.numpy()[:nf].astype(np.uint64), self.sc_h.numpy()[:nf].copy()

    def search_batch(self, Q: np.ndarray):
        """Batched sharded search: every rank calls this with the same nq queries; each runs the batch
        on its shard (K3 on the tensor cores when the shape allows), one all-gather moves the
        nq x k packed keys of every shard, and the batched K4 merges per query.
        -> (ids uint64 [nq, k], scores float32 [nq, k], n_found int32 [nq]) on every rank."""
        t = self.torch
        Q = np.ascontiguousarray(Q, dtype=np.float32)
        nq, k = Q.shape[0], self.k
        Qd = t.from_numpy(Q).to(self.dev, non_blocking=False)
        keys_local = t.zeros(nq * k, dtype=t.int64, device=self.dev)
        ids_d = t.zeros((nq, k), dtype=t.int64, device=self.dev)
        sc_d = t.zeros((nq, k), dtype=t.float32, device=self.dev)
        nf_d = t.zeros(nq, dtype=t.int32, device=self.dev)
        self.idx.search_batch_keys_device(Qd.data_ptr(), nq, k, keys_local.data_ptr())
        if self.world > 1:
            keys_all = t.zeros(self.world * nq * k, dtype=t.int64, device=self.dev)
            self.dist.all_gather_into_tensor(keys_all, keys_local)       # [world][nq][k]
        else:
            keys_all = keys_local
        self.idx.merge_batch_device(keys_all.data_ptr(), self.world, nq, k, ids_d.data_ptr(), sc_d.data_ptr(),
                                    nf_d.data_ptr())
        self.stream.synchronize()
        return ids_d.cpu().numpy().astype(np.uint64), sc_d.cpu().numpy(), nf_d.cpu().numpy()

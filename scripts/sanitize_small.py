"""Small-shape exercise of K1/K2/K3/K4 for compute-sanitizer (memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sema_b200
from sema_b200.synth import synth_rows

rng = np.random.default_rng(0)
for d, n, k in [(384, 3001, 10), (768, 1000, 100), (130, 777, 50)]:
    X = synth_rows(1, 0, n, d)
    valid = np.ones(n, np.uint8); valid[::7] = 0
    with sema_b200.GpuIndex(d, n + 8) as idx:
        idx.append(X, valid=valid, normalize=True)
        q = idx.read_rows(1, 1)[0]
        ids, sc = idx.search(q, k)
        assert ids[0] == 1, ids[:3]
        ids, sc = idx.search(q, 300)
        idx.tombstone(np.array([1], dtype=np.uint64))
        ids, sc = idx.search(q, k)
        assert 1 not in ids.tolist()
        Q = idx.read_rows(2, 9)
        Q[np.isnan(Q)] = 0
        idx.set_batch_mode(2 if d == 384 else 0)
        bi, bs, bn = idx.search_batch(Q, min(k, 100))
        idx.compact()
        g = sema_b200.ShardGroup(idx, 1, 0)
        g.search(q, 10)
        g.close()
print("sanitize_small ok")

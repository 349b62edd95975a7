"""Alternating A/B timing of two tuning-knob settings of the batched path (K3) in ONE process.

The B200s of this pool run K3 at the 1 kW power cap and drift by several percent within a run and
between boxes, so two separate bench runs cannot resolve a few-percent change; alternating the two
settings round by round on the same corpus can.  Settings are sema_index_set_scan_variant values
(include/sema_b200.h): e.g. `--a 300 --b 308` compares the epilogue with / without its group
early-out, `--a 102 --b 104` cluster sizes 2 / 4, `--a 200 --b 201` two / one query tiles per CTA.

    python scripts/k3_ab.py --mode 3 --a 300 --b 308 --rounds 6 --batches 6
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import sema_b200  # noqa: E402
from sema_b200.synth import synth_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=10_000_000)
ap.add_argument("--dim", type=int, default=384)
ap.add_argument("--nq", type=int, default=1024)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--mode", type=int, default=3, help="batch mode: 0 cascade, 2 bf16x3, 3 single pass")
ap.add_argument("--a", type=int, required=True)
ap.add_argument("--b", type=int, required=True)
ap.add_argument("--rounds", type=int, default=6)
ap.add_argument("--batches", type=int, default=6)
a = ap.parse_args()

dev = torch.device("cuda:0")
idx = sema_b200.GpuIndex(a.dim, a.rows)
idx.append_synthetic(1, 0, a.rows, True)
with sema_b200.GpuIndex(a.dim, a.nq) as qi:
    qi.append(synth_rows(2, 0, a.nq, a.dim), normalize=True)
    Q = qi.read_rows(0, a.nq)
stream = torch.cuda.current_stream()
idx.set_stream(stream.cuda_stream)
idx.set_batch_mode(a.mode)
Qd = torch.from_numpy(Q).to(dev)
ids = torch.zeros(a.nq * a.k, dtype=torch.int64, device=dev)
sc = torch.zeros(a.nq * a.k, dtype=torch.float32, device=dev)
nf = torch.zeros(a.nq, dtype=torch.int32, device=dev)


def run(setting):
    idx.set_scan_variant(setting)
    for _ in range(2):
        idx.search_batch_device(Qd.data_ptr(), a.nq, a.k, ids.data_ptr(), sc.data_ptr(), nf.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.batches):
        idx.search_batch_device(Qd.data_ptr(), a.nq, a.k, ids.data_ptr(), sc.data_ptr(), nf.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.batches


ta, tb = [], []
for r in range(a.rounds):
    ta.append(run(a.a))
    tb.append(run(a.b))
    print(f"round {r}: A({a.a}) {ta[-1]:.3f} ms   B({a.b}) {tb[-1]:.3f} ms", flush=True)
skip = 1 if a.rounds > 2 else 0          # the first round also warms the box up
ma, mb = float(np.mean(ta[skip:])), float(np.mean(tb[skip:]))
wins = sum(x < y for x, y in zip(ta, tb))
print(f"mean after round {skip}: A {ma:.3f} ms, B {mb:.3f} ms, A/B = {ma / mb:.4f}; A faster in {wins}/{a.rounds} rounds")

"""K3 soak at full size: 10M x 384, 1024 queries, 150 batches back to back with the modes alternating (cascade, single
pass, three passes); every batch's results must equal the first batch of its mode, and the modes must agree."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sema_b200
from sema_b200.synth import synth_rows
rows, nq, k = int(os.environ.get("ROWS", 10_000_000)), 1024, 10
dev = torch.device("cuda:0")
idx = sema_b200.GpuIndex(384, rows); idx.append_synthetic(1, 0, rows, True)
with sema_b200.GpuIndex(384, nq) as qi:
    qi.append(synth_rows(2, 0, nq, 384), normalize=True); Q = qi.read_rows(0, nq)
stream = torch.cuda.current_stream(); idx.set_stream(stream.cuda_stream)
Qd = torch.from_numpy(Q).to(dev)
ids_d = torch.zeros(nq * k, dtype=torch.int64, device=dev); sc_d = torch.zeros(nq * k, dtype=torch.float32, device=dev); nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)
first, bad = {}, 0
for it in range(150):
    mode = (0, 3, 2)[it % 3]
    idx.set_batch_mode(mode)
    ids_d.zero_(); sc_d.zero_()
    idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    torch.cuda.synchronize()
    got = (ids_d.cpu().numpy().copy(), sc_d.cpu().numpy().copy())
    if mode not in first: first[mode] = got
    elif not (np.array_equal(got[0], first[mode][0]) and np.array_equal(got[1], first[mode][1])): bad += 1
agree = all(np.array_equal(first[0][0], first[m][0]) and np.array_equal(first[0][1], first[m][1]) for m in (2, 3))
q, f = idx.batch_stats()
print(f"k3 soak: 150 batches, {bad} differing, modes agree: {agree}, K3 queries {q}, K2 fallbacks {f}, format {idx.batch_precision_active}", flush=True)
sys.exit(0 if (bad == 0 and agree) else 1)

"""Time the K3 stages of whatever library SEMA_B200_LIB points at (default: the shipped one): 10M x 384 x 1024 queries,
3 batches per measurement after a 0.4 s idle gap, 4 rounds.  For A/B runs across builds (one process per build)."""
import os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sema_b200
from sema_b200.synth import synth_rows
rows, nq, k = int(os.environ.get("ROWS", 10_000_000)), 1024, 10
dev = torch.device("cuda:0")
idx = sema_b200.GpuIndex(384, rows)
idx.append_synthetic(1, 0, rows, True)
with sema_b200.GpuIndex(384, nq) as qi:
    qi.append(synth_rows(2, 0, nq, 384), normalize=True); Q = qi.read_rows(0, nq)
stream = torch.cuda.current_stream(); idx.set_stream(stream.cuda_stream)
Qd = torch.from_numpy(Q).to(dev)
ids_d = torch.zeros(nq * k, dtype=torch.int64, device=dev); sc_d = torch.zeros(nq * k, dtype=torch.float32, device=dev); nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)
def timed(mode, reps=3):
    idx.set_batch_mode(mode)
    idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    torch.cuda.synchronize(); time.sleep(0.4)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {3: [], 2: []}
for rnd in range(4):
    for m in (3, 2): res[m].append(timed(m))
tag = os.path.basename(os.environ.get("SEMA_B200_LIB", "libsema_b200.so"))
print(f"{tag:32s} 1pass median {statistics.median(res[3]):.3f} {[round(x, 2) for x in res[3]]}   3pass median {statistics.median(res[2]):.3f} {[round(x, 2) for x in res[2]]}", flush=True)

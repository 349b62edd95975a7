"""K3 epilogue timing probes (debug flags give wrong results; timing only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sema_b200
from sema_b200.synth import synth_rows
rows, nq, k = 10_000_000, 1024, 10
dev = torch.device("cuda:0")
idx = sema_b200.GpuIndex(384, rows)
idx.append_synthetic(1, 0, rows, True)
with sema_b200.GpuIndex(384, nq) as qi:
    qi.append(synth_rows(2, 0, nq, 384), normalize=True); Q = qi.read_rows(0, nq)
stream = torch.cuda.current_stream(); idx.set_stream(stream.cuda_stream)
Qd = torch.from_numpy(Q).to(dev)
ids_d = torch.zeros(nq * k, dtype=torch.int64, device=dev); sc_d = torch.zeros(nq * k, dtype=torch.float32, device=dev); nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)
def run(mode, dbg):
    idx.set_batch_mode(mode); idx.set_scan_variant(300 + dbg)
    for _ in range(2): idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3): idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    print(f"mode={mode} debug={dbg}: {e0.elapsed_time(e1)/3:.2f} ms   (0 full, 1 no scan, 2 mask pass only, 3 inserts without rescan)", flush=True)
for kc16 in (0, 1):
    idx.set_scan_variant(400 + kc16)
    print('kc16 =', kc16)
    for dbg in (0, 1, 2):
        run(3, dbg)
    run(0, 0)

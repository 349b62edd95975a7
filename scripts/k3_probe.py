"""K3 timing probes.  Needs the probe build (`make -C sema_b200/csrc PROBE=1`, loaded through SEMA_B200_LIB):
the debug flags skip work and give wrong results, which the shipped library refuses.
  pair kernel (variant 701 / 702): 1 = no scan (ld + release only), 2 = mask pass only, 16 = no tcgen05.ld either,
  32 = no MMAs (TMA + barriers only), 8 = no group early-out (correct)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("SEMA_B200_LIB", os.path.join(ROOT, "sema_b200", "libsema_b200_probe.so"))
sys.path.insert(0, ROOT)
import numpy as np, torch
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
def run(mode, variant, dbg, reps=4):
    idx.set_batch_mode(mode); idx.set_scan_variant(variant); assert idx.set_scan_variant(1000000 + dbg) == 1000000 + dbg
    for _ in range(2): idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    print(f"mode={mode} variant={variant} debug={dbg:2d}: {e0.elapsed_time(e1)/reps:.3f} ms", flush=True)
import time, statistics
def timed(mode, settings, reps=3):
    for v in settings: assert idx.set_scan_variant(v) == v
    idx.set_batch_mode(mode)
    idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    torch.cuda.synchronize(); time.sleep(0.4)                  # same short idle gap before every measurement
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps): idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cases = [(f"1pass mixed w={0.70 + w / 100:.2f}", 3, [700, 100, 1101, 1200 + w]) for w in (5, 15, 22, 30, 35)] + \
        [(f"3pass mixed w={0.70 + w / 100:.2f}", 2, [700, 100, 1101, 1200 + w]) for w in (5, 15, 22, 30, 35)]
res = {n: [] for n, _, _ in cases}
for rnd in range(4):
    for n, mode, st in (cases if rnd % 2 == 0 else cases[::-1]):
        res[n].append(timed(mode, st))
for n, v in res.items():
    print(f"{n:32s} min {min(v):7.3f}  median {statistics.median(v):7.3f}  all {[round(x, 2) for x in v]}", flush=True)

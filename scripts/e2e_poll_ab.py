"""Synchronous host-call latency of sema_index_search at 10M x 384 against the device stream rate, for different
stream-query intervals of the completion poll (SEMA_POLL_QUERY_SHIFT, one process per setting)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sema_b200
from sema_b200.synth import synth_rows
rows, k = 10_000_000, 10
idx = sema_b200.GpuIndex(384, rows); idx.append_synthetic(1, 0, rows, True)
with sema_b200.GpuIndex(384, 64) as qi:
    qi.append(synth_rows(2, 0, 64, 384), normalize=True); Q = qi.read_rows(0, 64)
ids_h, sc_h = np.zeros(k, dtype=np.uint64), np.zeros(k, dtype=np.float32)
qh = [ctypes.c_void_p(Q[i].ctypes.data) for i in range(64)]
ids_p, sc_p = ctypes.c_void_p(ids_h.ctypes.data), ctypes.c_void_p(sc_h.ctypes.data)
for i in range(10): idx.search_ptr(qh[i], k, ids_p, sc_p)
out = []
for rep in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(40): idx.search_ptr(qh[i % 64], k, ids_p, sc_p)
    out.append((time.perf_counter() - t0) * 1e3 / 40)
print("SEMA_POLL_QUERY_SHIFT", os.environ.get("SEMA_POLL_QUERY_SHIFT", "14 (default)"), "ms per synchronous call:", [round(x, 4) for x in out], flush=True)
# the same loop with bench.py's NVML clock sampler running beside it
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for period in (None, 0.02, 0.2, 0.02, None):
    out = []
    def loop():
        for rep in range(5):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for i in range(40): idx.search_ptr(qh[i % 64], k, ids_p, sc_p)
            out.append((time.perf_counter() - t0) * 1e3 / 40)
    if period is None:
        loop()
    else:
        with bench.ClockSampler(0, period) as clk:
            loop()
    print("clock sampler period", period, "ms per synchronous call:", [round(x, 4) for x in out], flush=True)

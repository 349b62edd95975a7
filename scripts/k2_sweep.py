"""K2 sweep over corpus sizes (device-resident queries, CUDA events): per-call latency (one search
call per query, launches serialised) and query-stream throughput (launches chained with PDL)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sema_b200
from sema_b200.synth import synth_rows

dev = torch.device("cuda:0")
k = int(os.environ.get("K", "10"))
variants = [int(v) for v in os.environ.get("VARIANTS", "0").split(",")]
with sema_b200.GpuIndex(384, 64) as qi:
    qi.append(synth_rows(2, 0, 64, 384), normalize=True)
    Q = qi.read_rows(0, 64)
Qd = torch.from_numpy(Q).to(dev)
ids_d = torch.zeros(k, dtype=torch.int64, device=dev)
sc_d = torch.zeros(k, dtype=torch.float32, device=dev)
nf_d = torch.zeros(1, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream()
out = []
for rows in [int(r) for r in os.environ.get("ROWS", "1000,10000,100000,250000,500000,1000000,2000000,5000000,10000000").split(",")]:
    idx = sema_b200.GpuIndex(384, rows)
    idx.append_synthetic(1, 0, rows, True)
    idx.set_stream(stream.cuda_stream)
    for v in variants:
        idx.set_scan_variant(v)
        steps = 400 if rows <= 2_000_000 else 100
        for i in range(20):
            idx.search_device(Qd[i % 64].data_ptr(), k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            idx.search_device(Qd[i % 64].data_ptr(), k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        Qs = Qd[torch.arange(steps, device=dev) % 64].contiguous()
        ids_s = torch.zeros((steps, k), dtype=torch.int64, device=dev)
        sc_s = torch.zeros((steps, k), dtype=torch.float32, device=dev)
        nf_s = torch.zeros(steps, dtype=torch.int32, device=dev)
        idx.search_stream_device(Qs.data_ptr(), 20, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
        torch.cuda.synchronize()
        e0.record(stream)
        idx.search_stream_device(Qs.data_ptr(), steps, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms_s = e0.elapsed_time(e1) / steps
        out.append((rows, v, ms, ms_s))
        print(f"rows={rows:>9} variant={v} per call {ms*1e3:9.1f} us {rows*1536/ms/1e6:8.1f} GB/s | "
              f"query stream {ms_s*1e3:9.1f} us {rows*1536/ms_s/1e6:8.1f} GB/s", flush=True)
    idx.close()

#!/usr/bin/env python
"""bench.py — headline benchmark of the Sema retrieval hot path on B200.

Metric (BASELINE.json): QPS of the exact top-10 cosine scan over 10M x 384 fp32
unit-norm embeddings, with the achieved fraction of the HBM roofline.  A "step" is one
single-query search over the whole corpus.

  python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path
  python bench.py --impl reference ...                       # CPU search path (oracle port)

N > 1 is launched by torchrun (one process per GPU, NCCL): the fixed corpus is row-
sharded over the N GPUs (strong scaling), every rank scans its shard, the per-shard
top-k lists are all-gathered over NVLink and merged by kernel K4 on every rank.

The reference (Rust + un-vendored LanceDB) cannot be built here, so the reference arm and
the cpu_baseline leg time oracle/cpu_scan.c — the CPU restatement of the reference's
search path — on the box's host cores ("kind": "port").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "qps_exact_top10_cosine_scan_10Mx384_fp32"
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="total corpus rows")
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=64, help="distinct query vectors cycled through")
    ap.add_argument("--variant", type=int, default=-1, help="K2 kernel variant (tuning)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: fused = P2P stores + merge inside K2's last block; nccl = all-gather + K4")
    ap.add_argument("--workload", default="single", choices=["single", "batch", "ingest", "config1", "pool"],
                    help="single = headline single-query scan (K2); batch = BASELINE config 3, nq-query batches (K3); "
                         "ingest = BASELINE config 4/5 style streaming ingest (K1) interleaved with queries; "
                         "config1 = ~10k-chunk synthetic markdown corpus through the StorageManager boundary")
    ap.add_argument("--ingest-batch", type=int, default=65536, help="rows per appended batch (--workload ingest)")
    ap.add_argument("--queries-per-batch", type=int, default=2, help="searches issued after each appended batch")
    ap.add_argument("--nq", type=int, default=1024, help="queries per batch (--workload batch)")
    ap.add_argument("--batch-mode", type=int, default=2, help="0 auto, 1 K2 per query, 2 K3 tensor cores")
    ap.add_argument("--k3-cluster", type=int, default=0, help="K3 cluster size (0 auto, 1, 2, 4) — tuning")
    ap.add_argument("--cpu-rows", type=int, default=2_000_000, help="rows of the CPU-baseline sample")
    ap.add_argument("--cpu-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--pool-texts", type=int, default=4096, help="texts per step (--workload pool)")
    ap.add_argument("--pool-full-mask", action="store_true", help="--workload pool: every token attended")
    ap.add_argument("--growable", action="store_true", help="corpus in a growable (virtual-memory backed) index")
    ap.add_argument("--no-stream", action="store_true", help="timed region: one search call per query instead of one query stream")
    ap.add_argument("--no-chain", action="store_true", help="query stream without programmatic dependent launch (comparison)")
    ap.add_argument("--staged-host-path", action="store_true", help="e2e through the staged H2D / D2H path (comparison)")
    return ap.parse_args()


def workload_name(a):
    return f"{a.rows}x{a.dim} fp32 unit-norm synthetic embeddings, single-query exact top-{a.k} cosine scan"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, device_index: int, period_s: float = 0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period = period_s
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if vis:
                try:
                    phys = int(vis.split(",")[device_index])
                except ValueError:
                    phys = device_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report that instead of inventing clocks
            self.nv = None
            self.err = str(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "error": getattr(self, "err", "nvml unavailable")}
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- CPU legs
def cpu_scan_qps(a, steps, warmup):
    """Times the oracle port (oracle/cpu_scan.c, all host threads) on a bounded sample:
    the first `cpu_rows` rows of the same synthetic corpus.  Returns QPS scaled to the
    full corpus (x sample_rows / rows: the scan is linear in rows) and a description."""
    from oracle import c_oracle  # the ONLY use of oracle/ in bench.py: the CPU baseline
    c_oracle.build()
    n = min(a.cpu_rows, a.rows)
    X = c_oracle.normalize(c_oracle.synth(1, 0, n, a.dim))
    Q = c_oracle.normalize(c_oracle.synth(2, 0, a.queries, a.dim))
    for i in range(warmup):
        c_oracle.scan(X, Q[i % len(Q)], a.k)
    t0 = time.perf_counter()
    for i in range(steps):
        c_oracle.scan(X, Q[i % len(Q)], a.k)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    qps_sample = 1.0 / dt
    scale = n / a.rows
    return {
        "value": qps_sample * scale, "unit": "queries/s", "cores": c_oracle.threads(), "kind": "port",
        "sample": (f"oracle/cpu_scan.c (C port of the reference's CPU search path; the Rust reference "
                   f"cannot be built here), {c_oracle.threads()} OpenMP threads, {steps} scans of the first "
                   f"{n} rows x {a.dim} of the same corpus at {dt * 1e3:.2f} ms/scan "
                   f"({n * a.dim * 4 / dt / 1e9:.1f} GB/s); QPS scaled by {n}/{a.rows} to the full corpus"),
        "ms_per_scan_sample": dt * 1e3,
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(a.steps, 1), max(a.warmup, 0)
    # bound the CPU work: the whole run must end within a few minutes
    steps, warm = min(steps, 50), min(warm, 5)
    base = cpu_scan_qps(a, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": "queries/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": 1e3 / base["value"], "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "rows": a.rows, "dim": a.dim, "k": a.k,
                   "note": "each step is one CPU scan of a bounded row sample, scaled linearly to the full corpus"},
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def make_queries(a, sema_b200, device):
    """Unit-norm query vectors: synthetic rows (seed 2) normalised by kernel K1."""
    from sema_b200.synth import synth_rows
    raw = synth_rows(2, 0, a.queries, a.dim)
    with sema_b200.GpuIndex(a.dim, a.queries, device=device) as qi:
        qi.append(raw, normalize=True)
        return qi.read_rows(0, a.queries)


def run_batch(a):
    """BASELINE.json configs[2]: 10M x 384, batched nq-query top-k on one B200 (kernel K3)."""
    import torch

    import sema_b200
    from sema_b200 import _lib
    from sema_b200.synth import synth_rows

    if _lib.lib().sema_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: sema_b200 has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    k, nq = a.k, a.nq
    idx = sema_b200.GpuIndex(a.dim, a.rows, device=0)
    idx.append_synthetic(seed=1, row0=0, n=a.rows, normalize=True)
    idx.set_batch_mode(a.batch_mode)
    if a.k3_cluster:
        idx.set_scan_variant(100 + a.k3_cluster)
    with sema_b200.GpuIndex(a.dim, nq, device=0) as qi:
        qi.append(synth_rows(2, 0, nq, a.dim), normalize=True)
        Q = qi.read_rows(0, nq)
    stream = torch.cuda.current_stream()
    idx.set_stream(stream.cuda_stream)
    Qd = torch.from_numpy(Q).to(dev)
    ids_d = torch.zeros(nq * k, dtype=torch.int64, device=dev)
    sc_d = torch.zeros(nq * k, dtype=torch.float32, device=dev)
    nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)

    def step():
        idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())

    steps, warm = a.steps, max(a.warmup, 3)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    l0 = idx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(0) as clk:
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        dev_ms = e0.elapsed_time(e1) / steps
        launches = idx.launch_count - l0
        idx.set_stream(None)
        for _ in range(2):
            idx.search_batch(Q, k)
        t0 = time.perf_counter()
        for _ in range(steps):
            ids_h, sc_h, nf_h = idx.search_batch(Q, k)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    served, fallbacks = idx.batch_stats()
    # the same batch in automatic mode (precision cascade: single pass -> bf16x3 -> K2), for the record
    auto = None
    if a.batch_mode != 0:
        idx.set_batch_mode(0)
        idx.set_stream(stream.cuda_stream)
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        auto_ms = e0.elapsed_time(e1) / steps
        auto = {"ms_per_step": auto_ms, "value": nq / (auto_ms * 1e-3), "unit": "queries/s",
                "algorithmic_tflops": 2.0 * nq * a.rows * a.dim / (auto_ms * 1e-3) / 1e12,
                "cascaded_queries": idx.batch_cascaded}
        idx.set_stream(None)
        idx.set_batch_mode(a.batch_mode)
    # spot-check against the single-query kernel
    ok = True
    for i in (0, nq // 2, nq - 1):
        r_ids, r_sc = idx.search(Q[i], k)
        ok &= bool(np.array_equal(ids_h[i, :nf_h[i]], r_ids) and np.array_equal(sc_h[i, :nf_h[i]], r_sc))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops", 1590.0))
    flop = 2.0 * nq * a.rows * a.dim
    achieved = flop / (dev_ms * 1e-3) / 1e12
    line = {
        "metric": f"qps_batched_{nq}q_exact_top{k}_cosine_{a.rows}x{a.dim}_fp32", "value": nq / (dev_ms * 1e-3),
        "unit": "queries/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": dev_ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16x3 split -> f32", "data": "synthetic",
        "config": {"workload": f"{a.rows}x{a.dim} fp32 corpus, batches of {nq} queries, exact top-{k} (BASELINE configs[2])",
                   "batch_mode": a.batch_mode, "k3_cluster": a.k3_cluster or "auto", "k3_queries": served, "k3_fallback_queries": fallbacks,
                   "l2_flush": "none needed: each batch streams the 15.36 GB bf16 hi/lo planes"},
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "issued_frac": 3 * achieved / peak, "traffic": None,
                     "note": "achieved = algorithmic 2*Q*N*d FLOP / device time per batch; the bf16x3 split issues 3x that"},
        "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * a.dim * 4,
                "d2h_bytes_per_step": nq * (k * 12 + 4), "ms_per_step": e2e_ms,
                "path": "sema_index_search_batch (C ABI) with host buffers"},
        "gpu_launches": int(launches), "clocks": clk.summary(), "verified_against_k2": ok,
        "auto_mode_cascade": auto,
    }
    print(json.dumps(line), flush=True)


def run_batch_sharded(a):
    """Config 3 over a row-sharded corpus (batched): every rank runs the batch on its
    shard (K3), one NCCL all-gather moves the nq x k packed keys of every shard, the batched K4 merges
    per query on every rank.  Launched with torchrun, one process per GPU."""
    import torch
    import torch.distributed as dist

    import sema_b200
    from sema_b200.sharded import ShardedSearcher
    from sema_b200.synth import synth_rows

    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    k, nq = a.k, a.nq
    per = (a.rows + world - 1) // world
    lo, hi = min(rank * per, a.rows), min((rank + 1) * per, a.rows)
    idx = sema_b200.GpuIndex(a.dim, max(hi - lo, 1), device=local)
    idx.set_row_base(lo)
    idx.append_synthetic(seed=1, row0=lo, n=hi - lo, normalize=True)
    idx.set_batch_mode(a.batch_mode)
    with sema_b200.GpuIndex(a.dim, nq, device=local) as qi:
        qi.append(synth_rows(2, 0, nq, a.dim), normalize=True)
        Q = qi.read_rows(0, nq)
    stream = torch.cuda.current_stream()
    idx.set_stream(stream.cuda_stream)
    Qd = torch.from_numpy(Q).to(dev)
    keys_local = torch.zeros(nq * k, dtype=torch.int64, device=dev)
    keys_all = torch.zeros(world * nq * k, dtype=torch.int64, device=dev)
    ids_d = torch.zeros((nq, k), dtype=torch.int64, device=dev)
    sc_d = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)

    def step():
        idx.search_batch_keys_device(Qd.data_ptr(), nq, k, keys_local.data_ptr())
        dist.all_gather_into_tensor(keys_all, keys_local)
        idx.merge_batch_device(keys_all.data_ptr(), world, nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    steps, warm = a.steps, max(a.warmup, 3)
    for _ in range(warm):
        step()
    barrier()
    l0 = idx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1) / steps
        launches = idx.launch_count - l0
        sh = ShardedSearcher(idx, dist, k)
        sh.search_batch(Q)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ids_h, sc_h, nf_h = sh.search_batch(Q)          # host queries in, host results out
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    ok = True
    for i in (0, nq // 2, nq - 1):                          # against the sharded single-query path (K2 + all-gather + K4)
        r_ids, r_sc = sh.search(Q[i])
        ok &= bool(np.array_equal(ids_h[i, :nf_h[i]], r_ids) and np.array_equal(sc_h[i, :nf_h[i]], r_sc))
    v = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.MIN)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops", 1590.0))
        achieved = 2.0 * nq * (hi - lo) * a.dim / (dev_ms * 1e-3) / 1e12        # per GPU
        line = {
            "metric": f"qps_batched_{nq}q_exact_top{k}_cosine_{a.rows}x{a.dim}_fp32", "value": nq / (dev_ms * 1e-3),
            "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": dev_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16 split -> f32", "data": "synthetic",
            "config": {"workload": f"{a.rows}x{a.dim} fp32 corpus row-sharded over {world} GPUs, batches of {nq} queries, exact top-{k}",
                       "batch_mode": a.batch_mode, "rows_per_gpu": hi - lo, "exchange": "nccl all-gather of nq x k packed keys + batched K4",
                       "timing": "CUDA events on the launching stream, barrier + synchronize both sides, max over ranks"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "note": "per GPU: algorithmic 2*Q*rows_per_gpu*d FLOP / device time per batch (exchange and merge included)"},
            "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * a.dim * 4,
                    "d2h_bytes_per_step": nq * (k * 12 + 4), "ms_per_step": e2e_ms, "path": "sharded.ShardedSearcher.search_batch"},
            "gpu_launches": int(launches), "clocks": clk.summary(), "verified": bool(v.item()),
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def pool_traffic(a, n, seq, d):
    """DRAM bytes per K0 launch from the committed ncu capture, when it is this exact workload."""
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        return tr.get(f"k0_pool_{n}x{seq}x{d}_full_masks") if a.pool_full_mask else None
    except Exception:
        return None


def run_pool(a):
    """Kernel K0 (the step before the path): mean_pool of src/semantic/embeddings.rs:61-91 fused with the
    append, for batches of texts whose token embeddings are already on the device.  One step pools and
    appends `--pool-texts` texts of seq_len 256 (the reference's MAX_LENGTH) x dim."""
    import torch

    import sema_b200
    from oracle import c_oracle

    torch.cuda.set_device(0)
    dev = torch.device("cuda:0")
    n, seq, d = a.pool_texts, 256, a.dim
    g = torch.Generator(device=dev).manual_seed(1)
    tok = torch.randn((n, seq, d), generator=g, dtype=torch.float32, device=dev)
    lens = torch.randint(8, seq + 1, (n,), generator=g, device=dev)
    mask = (torch.arange(seq, device=dev)[None, :] < lens[:, None]).to(torch.float32).contiguous()
    if a.pool_full_mask:
        mask.fill_(1.0)
    idx = sema_b200.GpuIndex(d, n * (a.steps + a.warmup + 1), device=0)
    for _ in range(a.warmup):
        idx.append_pooled_device(tok.data_ptr(), mask.data_ptr(), n, seq, None, skip_masked=True)
    torch.cuda.synchronize()
    l0 = idx.launch_count
    with ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for _ in range(a.steps):
            idx.append_pooled_device(tok.data_ptr(), mask.data_ptr(), n, seq, None, skip_masked=True)   # returns when the rows are visible
        dt = time.perf_counter() - t0
    launches = idx.launch_count - l0
    # parity spot check on the first 64 texts of the last batch
    first = len(idx) - n
    got = idx.read_rows(first, 64)
    want = c_oracle.mean_pool(tok[:64].cpu().numpy(), mask[:64].cpu().numpy())
    read_bytes = float(mask.sum().item()) * d * 4          # rows of attended tokens: what K0 has to read
    ms = dt * 1e3 / a.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    line = {
        "metric": f"mean_pool_append_texts_per_s_seq{seq}x{d}_fp32", "value": n / (ms * 1e-3), "unit": "texts/s",
        "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{n} texts x {seq} tokens x {d} fp32 token embeddings on the device, attention lengths uniform in [8, {seq}]"
                               + (" (full masks)" if a.pool_full_mask else "") + ", mean_pool + normalise + append per step (K0 + K1 bookkeeping)",
                   "timing": "host wall clock around sema_index_append_pooled_device (synchronous: rows visible on return)",
                   "l2_flush": f"none needed: each step reads {read_bytes / 1e9:.2f} GB"},
        "roofline": {"bound": "hbm", "achieved": read_bytes / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": read_bytes / (ms * 1e-3) / 1e9 / peak, "traffic": pool_traffic(a, n, seq, d), "algorithmic_bytes": read_bytes,
                     "kernel": "pool_kernel (K0)", "note": "algorithmic bytes = attended tokens x dim x 4 (padding rows are skipped)"},
        "gpu_launches": int(launches), "clocks": clk.summary(),
        "verified": bool(np.array_equal(got, want)),
    }
    print(json.dumps(line), flush=True)


def run_ingest(a):
    """BASELINE.json configs[4]: the index grows to rows x dim by appending batches from pinned host
    memory (H2D + K1 on the ingest stream) while top-k queries run on the query stream.  Every query
    scans exactly the rows whose ingest had completed when it started (snapshot semantics; checked
    against the oracle in tests/test_gpu_parity.py::test_async_ingest_snapshot_semantics)."""
    import torch

    import sema_b200
    from sema_b200 import _lib

    if _lib.lib().sema_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: sema_b200 has no CPU fallback")
    torch.cuda.set_device(0)
    k, B = a.k, a.ingest_batch
    nb = (a.rows + B - 1) // B
    idx = sema_b200.GpuIndex(a.dim, nb * B, device=0)
    g = torch.Generator().manual_seed(1)
    pool = [torch.randn(B, a.dim, generator=g, dtype=torch.float32).pin_memory() for _ in range(3)]
    host = [t.numpy() for t in pool]
    Q = torch.randn(64, a.dim, generator=g, dtype=torch.float32)
    Q = (Q / Q.norm(dim=1, keepdim=True)).numpy()
    ids_h = np.zeros(k, dtype=np.uint64)
    sc_h = np.zeros(k, dtype=np.float32)
    # warm-up: one batch + a few queries, then start over with a fresh index
    idx.append(host[0], normalize=True)
    for i in range(5):
        idx.search_into(Q[i], k, ids_h, sc_h)
    idx.close()
    idx = sema_b200.GpuIndex(a.dim, nb * B, device=0)
    snaps, nq = [], 0
    torch.cuda.synchronize()
    total_rows = nb * B
    with ClockSampler(0) as clk:
        t0 = time.perf_counter()
        for b in range(nb):                                    # H2D + K1 per batch, all on the ingest stream
            idx.append(host[b % len(host)], normalize=True, asynchronous=True)
        t_enq = time.perf_counter() - t0
        while True:                                            # K2 on the query stream, while the ingest runs
            v = idx.visible
            if v >= total_rows:
                break
            if v == 0:
                time.sleep(0.0002)
                continue
            idx.search_into(Q[nq % 64], k, ids_h, sc_h)
            snaps.append(idx.last_snapshot)
            nq += 1
        idx.flush()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    assert idx.visible == total_rows and all(s % B == 0 and s > 0 for s in snaps) and snaps == sorted(snaps)
    scanned_bytes = float(sum(snaps)) * a.dim * 4
    # after the stream: the index answers like any other (spot check: a stored row finds itself)
    probe = idx.read_rows(total_rows - 1, 1)[0]
    r_ids, r_sc = idx.search(probe, 1)
    line = {
        "metric": f"streaming_ingest_rows_per_s_with_interleaved_top{k}_queries_{total_rows}x{a.dim}_fp32",
        "value": total_rows / dt, "unit": "rows/s", "n_gpus": 1, "steps": nb, "warmup": 1,
        "ms_per_step": dt * 1e3 / nb, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"append {nb} batches of {B} x {a.dim} fp32 rows from pinned host memory (normalise on GPU) "
                               f"while top-{k} queries run back to back on the query stream (BASELINE configs[4])",
                   "rows": total_rows, "dim": a.dim, "k": k},
        "ingest": {"rows_per_s": total_rows / dt, "h2d_GBps": total_rows * a.dim * 4 / dt / 1e9,
                   "queries": nq, "qps_during_ingest": nq / dt, "enqueue_s": t_enq,
                   "mean_snapshot_rows": float(np.mean(snaps)) if snaps else 0.0,
                   "query_scan_GBps": scanned_bytes / dt / 1e9,
                   "snapshots_monotonic_and_batch_aligned": True,
                   "self_probe_ok": bool(len(r_ids) == 1 and abs(float(r_sc[0]) - 1.0) < 1e-5)},
        "e2e": {"value": total_rows / dt, "unit": "rows/s", "h2d_bytes_per_step": B * a.dim * 4, "d2h_bytes_per_step": 0,
                "note": "plus one query H2D and one result D2H per search"},
        "gpu_launches": int(idx.launch_count), "clocks": clk.summary(),
    }
    print(json.dumps(line), flush=True)


def run_config1(a):
    """BASELINE.json configs[0]: index + search over a small synthetic markdown corpus (~10k chunks,
    384-d, top-10) through the StorageManager boundary, next to the CPU oracle on the same vectors.
    The embedder is a deterministic STAND-IN (oracle/corpus.py): the reference's MiniLM model and
    ONNX Runtime are not available offline, so embedding time is excluded from both arms."""
    from oracle import corpus, c_oracle          # corpus builder + CPU arm (test infrastructure)
    from sema_b200 import _lib
    from sema_b200.storage import Chunk, StorageManager

    if _lib.lib().sema_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: sema_b200 has no CPU fallback")
    files = corpus.make_markdown_tree(1050, seed=3)
    chunks = [Chunk(c["id"], c["file_path"], c["start_line"], c["end_line"], c["content"]) for c in corpus.chunk_tree(files)]
    emb = np.stack([corpus.embed(c.content) for c in chunks])
    rng = np.random.default_rng(5)
    queries = [" ".join(corpus._WORDS[int(i)] for i in rng.integers(0, len(corpus._WORDS), 6)) for _ in range(100)]
    qvec = {q: corpus.embed(q) for q in queries}
    k = a.k
    with StorageManager(dim=corpus.DIM, capacity_rows=len(chunks) + 64, normalize=True,
                        embedder=lambda t: qvec.get(t)) as mgr:
        t0 = time.perf_counter()
        mgr.index_chunks(chunks, vectors=emb)
        t_index = time.perf_counter() - t0
        for q in queries[:10]:
            mgr.search(q, k)
        lat = []
        for q in queries:
            t0 = time.perf_counter()
            hits = mgr.search(q, k)
            lat.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        for q in queries:
            mgr.execute_search(q)
        t_exec = (time.perf_counter() - t0) / len(queries)
    X = c_oracle.normalize(emb)
    Qn = c_oracle.normalize(np.stack([qvec[q] for q in queries]))
    for q in Qn[:5]:
        c_oracle.scan(X, q, k)
    cl = []
    for q in Qn:
        t0 = time.perf_counter()
        c_oracle.scan(X, q, k)
        cl.append(time.perf_counter() - t0)
    ids_ok = [c.id for c, _ in hits] == [chunks[int(i)].id for i in c_oracle.scan(X, Qn[-1], k)[0]]
    line = {
        "metric": f"config1_search_latency_top{k}_{len(chunks)}_chunks_384d", "value": 1.0 / float(np.median(lat)),
        "unit": "queries/s", "n_gpus": 1, "steps": len(queries), "warmup": 10, "ms_per_step": float(np.median(lat)) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{len(files)} synthetic markdown files -> {len(chunks)} chunks (reference chunker constants), "
                               f"stand-in embedder (NOT MiniLM), 100 seeded queries, top-{k}, through StorageManager.search",
                   "index_chunks_s": t_index},
        "latency_ms": {"gpu_search_median": float(np.median(lat)) * 1e3, "gpu_search_p99": float(np.percentile(lat, 99)) * 1e3,
                       "gpu_execute_search_mean": t_exec * 1e3,
                       "cpu_oracle_median": float(np.median(cl)) * 1e3, "cpu_oracle_p99": float(np.percentile(cl, 99)) * 1e3,
                       "cpu_threads": c_oracle.threads()},
        "cpu_baseline": {"value": 1.0 / float(np.median(cl)), "unit": "queries/s", "cores": c_oracle.threads(), "kind": "port",
                         "sample": "oracle/cpu_scan.c on the same normalised vectors and queries (whole workload, not a sample)"},
        "e2e": {"value": 1.0 / float(np.median(lat)), "unit": "queries/s", "h2d_bytes_per_step": 1536, "d2h_bytes_per_step": 8 + 12 * k},
        "last_query_matches_oracle": bool(ids_ok),
    }
    print(json.dumps(line), flush=True)


def pipelined_e2e(obj, a, qh, k, ids_p, sc_p, sync):
    """Host queries in, host results out, through sema_*_search_submit / _collect with two searches in
    flight (what a search service does; the synchronous call is the reference's own pattern)."""
    for i in range(3):
        obj.collect_ptr(obj.submit_ptr(qh[i % a.queries], k), ids_p, sc_p)
    sync()
    t0 = time.perf_counter()
    prev = obj.submit_ptr(qh[0], k)
    for i in range(1, a.steps):
        t = obj.submit_ptr(qh[i % a.queries], k)
        obj.collect_ptr(prev, ids_p, sc_p)
        prev = t
    obj.collect_ptr(prev, ids_p, sc_p)
    sync()
    return (time.perf_counter() - t0) * 1e3


def run_ours(a):
    import torch

    import sema_b200
    from sema_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus != world:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun (one process per GPU)")
    if _lib.lib().sema_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: sema_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- corpus: rows [lo, hi) of the fixed synthetic corpus live on this GPU
    per = (a.rows + world - 1) // world
    lo, hi = min(rank * per, a.rows), min((rank + 1) * per, a.rows)
    idx = sema_b200.GpuIndex(a.dim, max(hi - lo, 1), device=local, growable=a.growable)
    idx.set_row_base(lo)
    idx.append_synthetic(seed=1, row0=lo, n=hi - lo, normalize=True)
    if a.variant >= 0:
        idx.set_scan_variant(a.variant)
    Q = make_queries(a, sema_b200, local)
    k = a.k

    stream = torch.cuda.current_stream()
    idx.set_stream(stream.cuda_stream)
    Qd = torch.from_numpy(Q).to(dev)
    qptr = [Qd[i].data_ptr() for i in range(a.queries)]
    ids_d = torch.zeros(k, dtype=torch.int64, device=dev)
    sc_d = torch.zeros(k, dtype=torch.float32, device=dev)
    nf_d = torch.zeros(1, dtype=torch.int32, device=dev)
    keys_local = torch.zeros(k, dtype=torch.int64, device=dev)
    keys_all = torch.zeros(world * k, dtype=torch.int64, device=dev)

    group = None
    exchange = "none"
    if world > 1:
        exchange = a.exchange
        if a.exchange == "fused":
            try:
                from sema_b200.sharded import make_shard_group
                group = make_shard_group(idx, dist)
            except Exception as e:      # e.g. CUDA IPC not permitted in this container
                ok = torch.tensor([0], device=dev)
                exchange = f"nccl (fused unavailable: {e})"
            else:
                ok = torch.tensor([1], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)          # all ranks or none
            if int(ok.item()) == 0 and group is not None:
                group.close()
                group = None
                exchange = "nccl (fused unavailable on a peer)"

    # the timed region issues its K queries as ONE query stream (one K2 launch per query, consecutive
    # launches chained with programmatic dependent launch); --no-stream issues K separate calls
    use_stream = (not a.no_stream) and (world == 1 or group is not None)
    n_s = max(a.steps, a.warmup, 1)
    Qs = Qd[torch.arange(n_s, device=dev) % a.queries].contiguous()      # query i of the stream = pool[i % pool]
    ids_s = torch.zeros((n_s, k), dtype=torch.int64, device=dev)
    sc_s = torch.zeros((n_s, k), dtype=torch.float32, device=dev)
    nf_s = torch.zeros(n_s, dtype=torch.int32, device=dev)
    if a.no_chain:
        idx.set_scan_variant(600)

    def run_device(nq):
        if nq <= 0:
            return
        if use_stream:
            (idx if world == 1 else group).search_stream_device(Qs.data_ptr(), nq, k, ids_s.data_ptr(), sc_s.data_ptr(),
                                                               nf_s.data_ptr())
        else:
            for i in range(nq):
                step_device(i)

    def step_device(i):
        q = qptr[i % a.queries]
        if world == 1:
            idx.search_device(q, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
        elif group is not None:
            group.search_device(q, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
        else:
            idx.search_keys_device(q, k, keys_local.data_ptr())
            dist.all_gather_into_tensor(keys_all, keys_local)
            idx.merge_device(keys_all.data_ptr(), world, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region: inputs resident in HBM, CUDA events on the launching stream
    run_device(a.warmup)
    barrier()
    l0 = idx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record(stream)
        run_device(a.steps)
        e1.record(stream)
        barrier()
        dev_ms = e0.elapsed_time(e1)
        launches = idx.launch_count - l0
        if use_stream:
            ids_d.copy_(ids_s[a.steps - 1])

        # ---- end-to-end region: host query in, host results out, through the public call
        ids_h = np.zeros(k, dtype=np.uint64)
        sc_h = np.zeros(k, dtype=np.float32)
        import ctypes
        qh = [ctypes.c_void_p(Q[i].ctypes.data) for i in range(a.queries)]      # host pointers, built once
        ids_p, sc_p = ctypes.c_void_p(ids_h.ctypes.data), ctypes.c_void_p(sc_h.ctypes.data)
        e2e_ms = None
        e2e_lat = None
        e2e_pipe_ms = None
        if a.staged_host_path:
            idx.set_scan_variant(500)
        if world == 1:
            idx.set_stream(None)
            for i in range(min(a.warmup, 5)):
                idx.search_ptr(qh[i % a.queries], k, ids_p, sc_p)
            torch.cuda.synchronize()
            lat = np.empty(a.steps)
            t0 = time.perf_counter()
            tp = t0
            for i in range(a.steps):
                idx.search_ptr(qh[i % a.queries], k, ids_p, sc_p)      # synchronous: results are in ids_h / sc_h on return
                tn = time.perf_counter()
                lat[i] = tn - tp
                tp = tn
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3
            e2e_lat = {"median_ms": float(np.median(lat)) * 1e3, "p99_ms": float(np.percentile(lat, 99)) * 1e3,
                       "max_ms": float(lat.max()) * 1e3}
            e2e_pipe_ms = pipelined_e2e(idx, a, qh, k, ids_p, sc_p, torch.cuda.synchronize)
        elif group is not None:
            for i in range(min(a.warmup, 5)):
                group.search_ptr(qh[i % a.queries], k, ids_p, sc_p)
            barrier()
            t0 = time.perf_counter()
            for i in range(a.steps):
                group.search_ptr(qh[i % a.queries], k, ids_p, sc_p)
            barrier()
            e2e_ms = (time.perf_counter() - t0) * 1e3
            e2e_pipe_ms = pipelined_e2e(group, a, qh, k, ids_p, sc_p, barrier)
        else:
            from sema_b200.sharded import ShardedSearcher
            sh = ShardedSearcher(idx, dist, k)
            for i in range(min(a.warmup, 5)):
                sh.search(Q[i % a.queries])
            barrier()
            t0 = time.perf_counter()
            for i in range(a.steps):
                sh.search(Q[i % a.queries])
            barrier()
            e2e_ms = (time.perf_counter() - t0) * 1e3
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms, e2e_pipe_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        e2e_pipe_ms = float(t[2]) or None

    ms_step = dev_ms / a.steps
    qps = 1e3 / ms_step
    e2e_qps = a.steps / (e2e_ms / 1e3)

    # ---- every result of the timed stream must equal the result of the same pool query earlier in
    # the stream (query i = pool[i % pool]): any race between chained launches would break this
    stream_consistent = None
    if use_stream and a.steps > a.queries:
        first = torch.arange(a.steps, device=dev) % a.queries
        same = (ids_s[:a.steps] == ids_s[first]).all() & (sc_s[:a.steps] == sc_s[first]).all() & (nf_s[:a.steps] == nf_s[first]).all()
        stream_consistent = bool(same.item())
        if dist is not None:
            c = torch.tensor([int(stream_consistent)], device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.MIN)
            stream_consistent = bool(c.item())

    # ---- verification of the last device-side result against a fresh host-API search
    verified = None
    if not a.no_verify and world == 1:
        i = (a.steps - 1) % a.queries
        r_ids, r_sc = idx.search(Q[i], k)
        verified = bool(np.array_equal(ids_d.cpu().numpy().astype(np.uint64), r_ids))
        if use_stream:                            # every result of the timed (chained) stream against a fresh host search
            ids_stream = ids_s.cpu().numpy().astype(np.uint64)
            for j in range(min(a.steps, a.queries, 8)):
                verified &= bool(np.array_equal(ids_stream[j], idx.search(Q[j], k)[0]))
    elif not a.no_verify:
        # multi-rank: the fused result must equal the NCCL all-gather + K4 result, and every global
        # hit that lives on this rank must be this rank's own local hit with the same score
        from sema_b200.sharded import ShardedSearcher
        sh = ShardedSearcher(idx, dist, k)
        verified = True
        ids_stream = ids_s.cpu().numpy().astype(np.uint64) if use_stream else None
        sc_stream = sc_s.cpu().numpy() if use_stream else None
        for i in range(4):
            n_ids, n_sc = sh.search(Q[i])
            if group is not None:
                f_ids, f_sc = group.search(Q[i], k)
                verified &= bool(np.array_equal(f_ids, n_ids) and np.array_equal(f_sc, n_sc))
            if use_stream and i < a.steps:        # query i of the timed (chained) stream was pool query i
                verified &= bool(np.array_equal(ids_stream[i], n_ids) and np.array_equal(sc_stream[i], n_sc))
            l_ids, l_sc = idx.search(Q[i], k)
            mine = (n_ids >= lo) & (n_ids < hi)
            verified &= bool(set(n_ids[mine].tolist()) <= set(l_ids.tolist()))
        v = torch.tensor([int(verified)], device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MIN)
        verified = bool(v.item())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    pk_src = "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        pk_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    shard_rows = hi - lo
    traffic = None
    try:   # DRAM bytes per launch from the committed `ncu --set full` capture of this exact workload
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if world == 1 and a.k == 10:
            traffic = tr.get(f"k2_single_{a.rows}x{a.dim}")
    except Exception:
        pass
    bytes_per_launch = shard_rows * a.dim * 4          # algorithmic bytes: the shard read once
    achieved = bytes_per_launch / (ms_step * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {
            "workload": workload_name(a), "rows": a.rows, "dim": a.dim, "k": k,
            "rows_per_gpu": shard_rows, "query_pool": a.queries,
            "l2_flush": f"none needed: each step streams {bytes_per_launch / 1e9:.2f} GB per GPU, "
                        f"{bytes_per_launch / L2_BYTES:.0f}x the 126 MB L2",
            "parallelism": "single GPU" if world == 1 else f"corpus row-sharded over {world} GPUs (one process each)",
            "exchange": exchange if world > 1 else None,
            "issue": ("one query stream of K queries (sema_index_search_stream_device / sema_shard_group_search_stream_device): "
                      "one K2 launch per query, " + ("unchained" if a.no_chain else "consecutive launches chained with programmatic dependent launch"))
                     if use_stream else "one search call per query",
            "timing": "CUDA events on the launching stream, barrier + synchronize both sides, max over ranks",
        },
        "roofline": {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": "profiles/r01_k2_scan_full_raw.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum)" if traffic else None,
            "algorithmic_bytes": bytes_per_launch, "peak_source": pk_src,
            "spec_peak": 8000.0, "frac_of_spec": achieved / 8000.0,     # HBM3e sheet number; a read-only stream can beat the measured read+write copy peak
            "kernel": "scan_topk_tma_kernel (K2, TMA ring)" if (a.variant <= 0 and a.dim in (384, 768)) else "scan_topk_kernel (K2, register-fed)",
            "note": "achieved = rows_per_gpu*dim*4 bytes / device time per step (one K2 launch per step"
                    + ("" if world == 1 else ", which includes the top-k exchange and the global merge") + ")",
        },
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": ((a.dim + 3) // 4) * 16,
                "d2h_bytes_per_step": 8 + 12 * k, "ms_per_step": e2e_ms / a.steps, "latency": e2e_lat,
                "pipelined": None if not e2e_pipe_ms else {
                    "value": a.steps / (e2e_pipe_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_pipe_ms / a.steps,
                    "path": "same host buffers through the submit / collect form of the call, two searches in flight"},
                "path": ("sema_index_search" if world == 1 else "sema_shard_group_search" if group is not None else "sharded.ShardedSearcher.search")
                        + " with host buffers: the query travels in the kernel parameters (these bytes), K2 (+ exchange + merge) stores"
                          " the result block into mapped host memory (these bytes), the call polls its completion flag"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
        "verified": verified if stream_consistent is None else bool(verified and stream_consistent),
        "stream_self_consistent": stream_consistent,
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_scan_qps(a, a.cpu_steps, 2)
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "batch":
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            run_batch_sharded(a)
        else:
            run_batch(a)
    elif a.workload == "ingest":
        run_ingest(a)
    elif a.workload == "config1":
        run_config1(a)
    elif a.workload == "pool":
        run_pool(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

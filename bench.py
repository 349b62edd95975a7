#!/usr/bin/env python
"""bench.py — headline benchmark of the Sema retrieval hot path on B200.

Metric (BASELINE.json): QPS of the exact top-10 cosine scan over 10M x 384 fp32
unit-norm embeddings, with the achieved fraction of the HBM roofline.  A "step" is one
single-query search over the whole corpus.

  python bench.py [--gpus N] [--steps K] [--warmup W]        # this repo's CUDA path
  python bench.py --impl reference ...                       # CPU search path (oracle port), whole corpus

The default run also measures the other BASELINE.json configs and reports them as compact
sub-records under "configs" in the same JSON line (each spot-checked against the oracle):
1M x 384 single query (configs[1]), 1024-query batches on the tensor cores (configs[2]),
10M x 768 / k = 100 with streaming ingest interleaved with queries (configs[4]) and, for
N > 1, the 100M x 384 corpus sharded over the N GPUs (configs[3]).  --no-extra skips them.

N > 1 is launched by torchrun (one process per GPU, NCCL): the fixed corpus is row-
sharded over the N GPUs (strong scaling), every rank scans its shard and the per-shard
top-k lists are exchanged and merged inside the scan kernel (fused peer exchange).

The reference (Rust + un-vendored LanceDB) cannot be built here, so the reference arm and
the cpu_baseline leg time oracle/cpu_scan.c — the CPU restatement of the reference's
search path — on the box's host cores ("kind": "port"), over the WHOLE corpus named in
`config`, with one thread per core of the process's affinity mask.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEADLINE = (10_000_000, 384, 10)
L2_BYTES = 126e6
N_ORACLE_QUERIES = 8          # queries of every measured workload that are re-derived by the oracle


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=HEADLINE[0], help="total corpus rows")
    ap.add_argument("--dim", type=int, default=HEADLINE[1])
    ap.add_argument("--k", type=int, default=HEADLINE[2])
    ap.add_argument("--queries", type=int, default=64, help="distinct query vectors cycled through")
    ap.add_argument("--repeats", type=int, default=5, help="the timed region of --steps steps is run this many times; the median is reported")
    ap.add_argument("--variant", type=int, default=-1, help="K2 kernel variant (tuning)")
    ap.add_argument("--exchange", default="fused", choices=["fused", "nccl"],
                    help="N > 1: fused = P2P stores + merge inside K2's last block; nccl = all-gather + K4")
    ap.add_argument("--workload", default="single", choices=["single", "batch", "ingest", "config1", "pool", "maintenance"],
                    help="single = headline single-query scan (K2) + the sub-configs; batch = BASELINE configs[2], nq-query "
                         "batches (K3); ingest = configs[4], streaming ingest (K1) interleaved with queries; "
                         "config1 = ~10k-chunk synthetic markdown corpus through the StorageManager boundary; pool = K0; "
                         "maintenance = SURVEY 8(f) rows 1-2: tombstones, compaction, disk cache save / load")
    ap.add_argument("--maint-rows", type=int, default=4_000_000, help="rows of the --workload maintenance index")
    ap.add_argument("--maint-save-rows", type=int, default=1_000_000, help="rows of the index saved to / loaded from disk")
    ap.add_argument("--ingest-batch", type=int, default=65536, help="rows per appended batch (--workload ingest)")
    ap.add_argument("--nq", type=int, default=1024, help="queries per batch (--workload batch)")
    ap.add_argument("--batch-mode", type=int, default=2, help="0 auto (cascade), 1 K2 per query, 2 K3 three-pass split, 3 K3 single pass")
    ap.add_argument("--precision", type=int, default=0, help="K3 split format: 0 automatic (fp16 halves on unit-norm rows), 1 bf16 halves")
    ap.add_argument("--k3-pair", type=int, default=-1, help="K3 single-pass stage: 0 single-CTA kernel (default), 1 CTA pairs (tuning)")
    ap.add_argument("--k3-cluster", type=int, default=0, help="K3 cluster size (0 auto, 1, 2, 4) — tuning")
    ap.add_argument("--cpu-steps", type=int, default=10, help="timed whole-corpus CPU scans of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle re-derivation of results (and the host corpus it needs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the sub-configs (configs[1], [2], [3], [4])")
    ap.add_argument("--rows-sharded", type=int, default=100_000_000, help="rows of the configs[3] sub-record (N > 1)")
    ap.add_argument("--pool-texts", type=int, default=4096, help="texts per step (--workload pool)")
    ap.add_argument("--pool-full-mask", action="store_true", help="--workload pool: every token attended")
    ap.add_argument("--growable", action="store_true", help="corpus in a growable (virtual-memory backed) index")
    ap.add_argument("--no-stream", action="store_true", help="timed region: one search call per query instead of one query stream")
    ap.add_argument("--no-chain", action="store_true", help="query stream without programmatic dependent launch (comparison)")
    ap.add_argument("--launch-per-query", action="store_true",
                    help="query stream as one K2 launch per query instead of one persistent launch (comparison)")
    ap.add_argument("--staged-host-path", action="store_true", help="e2e through the staged H2D / D2H path (comparison)")
    return ap.parse_args()


def metric_name(rows, dim, k):
    if (rows, dim, k) == HEADLINE:
        return "qps_exact_top10_cosine_scan_10Mx384_fp32"        # BASELINE.json's headline metric
    return f"qps_exact_top{k}_cosine_scan_{rows}x{dim}_fp32"


def shared_config(a):
    """The workload, stated identically by both arms (`--impl ours` and `--impl reference`)."""
    total = a.rows * a.dim * 4
    return {
        "workload": f"{a.rows}x{a.dim} fp32 unit-norm synthetic embeddings, single-query exact top-{a.k} cosine scan",
        "rows": a.rows, "dim": a.dim, "k": a.k, "query_pool": a.queries, "n_gpus": a.gpus,
        "corpus": "value(seed=1,row,col) of SURVEY.md §8(d), L2-normalised; queries: seed=2",
        "l2_flush": f"none needed: every step streams the whole corpus, {total / 1e9:.2f} GB "
                    f"({total / max(a.gpus, 1) / L2_BYTES:.0f}x the 126 MB L2 per GPU)",
    }


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {}, "fallback (B200_PROFILING.md: 6650 GB/s, 1590 TFLOP/s)"


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
class HostCorpus:
    """Host copy of the synthetic corpus, produced by the ORACLE (oracle/cpu_scan.c: generator +
    the reference's normalise rule) — the checker's own data, not a read-back of the GPU matrix.
    Used by the cpu_baseline / --impl reference legs (timed whole-corpus CPU scans) and to
    re-derive GPU results (oracle top-k over the same rows).  Oracle results are cached per
    (query, k, row prefix)."""

    def __init__(self, rows, dim, nq, seed=1, qseed=2):
        from oracle import c_oracle            # bench.py uses oracle/ only here: CPU baseline + checker
        self.c = c_oracle
        c_oracle.build()
        self.cores = c_oracle.use_all_cores()
        t0 = time.perf_counter()
        self.X = np.empty((rows, dim), dtype=np.float32)
        step = 1 << 20
        for r0 in range(0, rows, step):        # chunked: keeps the generator's temporaries small
            m = min(step, rows - r0)
            c_oracle.synth(seed, r0, m, dim, out=self.X[r0:r0 + m])
        c_oracle.normalize_inplace(self.X)
        self.Q = c_oracle.normalize(c_oracle.synth(qseed, 0, nq, dim))
        self.build_s = time.perf_counter() - t0
        self._cache = {}

    def topk(self, qi, k, prefix=None):
        key = (qi, k, prefix)
        if key not in self._cache:
            X = self.X if prefix is None else self.X[:prefix]
            self._cache[key] = self.c.scan(X, self.Q[qi], k)
        return self._cache[key]

    def check(self, qi, k, ids, scores, prefix=None):
        """BASELINE.json's acceptance rule (oracle.check_parity) for query qi; returns None or the failure text."""
        from oracle import oracle as O
        r_ids, r_sc = self.topk(qi, k, prefix)
        try:
            O.check_parity(np.asarray(ids, dtype=np.uint64), np.asarray(scores, dtype=np.float32), r_ids, r_sc)
            return None
        except AssertionError as e:
            return f"query {qi}: {e}"

    def time_scans(self, k, steps, warmup):
        """Whole-corpus CPU scans, all cores; returns (seconds elapsed over `steps` scans, per-scan list)."""
        nq = len(self.Q)
        for i in range(warmup):
            self.c.scan(self.X, self.Q[i % nq], k)
        per = []
        t0 = time.perf_counter()
        for i in range(steps):
            t1 = time.perf_counter()
            self.c.scan(self.X, self.Q[i % nq], k)
            per.append(time.perf_counter() - t1)
        return time.perf_counter() - t0, per


def cpu_baseline_record(hc, a, steps, warmup):
    el, per = hc.time_scans(a.k, steps, warmup)
    dt = el / max(steps, 1)
    n, d = hc.X.shape
    return {
        "value": 1.0 / dt, "unit": "queries/s", "cores": hc.cores, "kind": "port",
        "sample": (f"oracle/cpu_scan.c (C port of the reference's CPU search path; the Rust reference cannot be built "
                   f"here), {hc.cores} OpenMP threads = the process's affinity mask, {steps} scans of the WHOLE corpus "
                   f"({n} x {d}, {n * d * 4 / 1e9:.2f} GB on the host) after {warmup} warm-up scans: {dt * 1e3:.1f} ms/scan "
                   f"({n * d * 4 / dt / 1e9:.1f} GB/s); nothing extrapolated"),
        "ms_per_scan": dt * 1e3, "ms_per_scan_min": min(per) * 1e3, "ms_per_scan_max": max(per) * 1e3,
        "extrapolated": False, "host_corpus_build_s": hc.build_s,
    }


def host_ram_bytes():
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES")
    except (ValueError, OSError):
        return 0


def run_reference(a):
    """--impl reference: the reference's CPU search path (oracle port) over the whole corpus of `config`, on all the
    host cores this process may use.  Under torchrun rank 0 alone runs it."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    need = a.rows * a.dim * 4
    ram = host_ram_bytes()
    if ram and need > 0.7 * ram:
        print(json.dumps({"impl": "reference", "unavailable": f"the {need / 1e9:.1f} GB corpus does not fit this host's {ram / 1e9:.0f} GB of RAM"}), flush=True)
        return
    steps, warm = max(min(a.steps, 50), 1), max(min(a.warmup, 5), 0)   # bounded: 50 whole-corpus scans at most
    hc = HostCorpus(a.rows, a.dim, a.queries)
    base = cpu_baseline_record(hc, a, steps, warm)
    line = {
        "impl": "reference", "metric": metric_name(a.rows, a.dim, a.k), "value": base["value"], "unit": "queries/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": base["ms_per_scan"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": shared_config(a),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm: helpers
def need_gpu():
    from sema_b200 import _lib
    if _lib.lib().sema_device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: sema_b200 has no CPU fallback")


def make_queries(nq, dim, device):
    """Unit-norm query vectors: synthetic rows (seed 2) normalised by kernel K1."""
    import sema_b200
    from sema_b200.synth import synth_rows
    with sema_b200.GpuIndex(dim, nq, device=device) as qi:
        qi.append(synth_rows(2, 0, nq, dim), normalize=True)
        return qi.read_rows(0, nq)


def median_spread(xs):
    xs = [float(x) for x in xs]
    return float(np.median(xs)), {"n": len(xs), "min": min(xs), "max": max(xs), "all": [round(x, 6) for x in xs]}


def k2_stream_bench(idx, Q, k, steps, warmup, repeats, dev, stream):
    """Device-timed single-query stream on one GPU (inputs resident): -> (median ms/step, spread, ids, scores).
    k2_stream_bench.launches = kernel launches of the last timed region (1 = the persistent stream kernel)."""
    import torch
    nq = len(Q)
    Qd = torch.from_numpy(Q).to(dev)
    n_s = max(steps, warmup, 1)
    Qs = Qd[torch.arange(n_s, device=dev) % nq].contiguous()
    ids_s = torch.zeros((n_s, k), dtype=torch.int64, device=dev)
    sc_s = torch.zeros((n_s, k), dtype=torch.float32, device=dev)
    nf_s = torch.zeros(n_s, dtype=torch.int32, device=dev)
    idx.set_stream(stream.cuda_stream)
    idx.search_stream_device(Qs.data_ptr(), warmup, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(repeats):
        torch.cuda.synchronize()
        l0 = idx.launch_count
        e0.record(stream)
        idx.search_stream_device(Qs.data_ptr(), steps, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / steps)
        k2_stream_bench.launches = idx.launch_count - l0
    idx.set_stream(None)
    med, spread = median_spread(ms)
    return med, spread, ids_s.cpu().numpy().astype(np.uint64), sc_s.cpu().numpy()


def sync_e2e_bench(search_ptr, Q, k, steps, repeats, sync):
    """Host query in, host results out, one synchronous call per query: -> (median ms/step, spread, latency stats)."""
    ids_h, sc_h = np.zeros(k, dtype=np.uint64), np.zeros(k, dtype=np.float32)
    qh = [ctypes.c_void_p(Q[i].ctypes.data) for i in range(len(Q))]
    ids_p, sc_p = ctypes.c_void_p(ids_h.ctypes.data), ctypes.c_void_p(sc_h.ctypes.data)
    for i in range(min(5, steps)):
        search_ptr(qh[i % len(Q)], k, ids_p, sc_p)
    ms, lat = [], None
    for _ in range(repeats):
        sync()
        lat = np.empty(steps)
        t0 = tp = time.perf_counter()
        for i in range(steps):
            search_ptr(qh[i % len(Q)], k, ids_p, sc_p)      # synchronous: results are in ids_h / sc_h on return
            tn = time.perf_counter()
            lat[i] = tn - tp
            tp = tn
        sync()
        ms.append((time.perf_counter() - t0) * 1e3 / steps)
    med, spread = median_spread(ms)
    stats = {"median_ms": float(np.median(lat)) * 1e3, "p99_ms": float(np.percentile(lat, 99)) * 1e3, "max_ms": float(lat.max()) * 1e3}
    return med, spread, stats


def pipelined_e2e(obj, Q, steps, k, sync):
    """Host queries in, host results out, through sema_*_search_submit / _collect with two searches in
    flight (what a search service does; the synchronous call is the reference's own pattern)."""
    ids_h, sc_h = np.zeros(k, dtype=np.uint64), np.zeros(k, dtype=np.float32)
    qh = [ctypes.c_void_p(Q[i].ctypes.data) for i in range(len(Q))]
    ids_p, sc_p = ctypes.c_void_p(ids_h.ctypes.data), ctypes.c_void_p(sc_h.ctypes.data)
    for i in range(3):
        obj.collect_ptr(obj.submit_ptr(qh[i % len(Q)], k), ids_p, sc_p)
    sync()
    t0 = time.perf_counter()
    prev = obj.submit_ptr(qh[0], k)
    for i in range(1, steps):
        t = obj.submit_ptr(qh[i % len(Q)], k)
        obj.collect_ptr(prev, ids_p, sc_p)
        prev = t
    obj.collect_ptr(prev, ids_p, sc_p)
    sync()
    return (time.perf_counter() - t0) * 1e3 / steps


def hbm_roofline(rows, dim, ms_step, kernel, traffic=None):
    peaks, src = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    b = rows * dim * 4
    ach = b / (ms_step * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
            "algorithmic_bytes": b, "peak_source": src + " hbm_gbs (a read+write copy: a read-only stream can exceed it)",
            "spec_peak": 8000.0, "frac_of_spec": ach / 8000.0, "kernel": kernel}


# ----------------------------------------------------------------------------- sub-configs (default run)
def sub_single_1m(a, hc, dev, stream):
    """BASELINE.json configs[1]: 1M x 384, single-query top-10 on one B200 (K2)."""
    import sema_b200
    rows = min(1_000_000, a.rows)
    steps = 200
    with sema_b200.GpuIndex(a.dim, rows, device=dev.index or 0) as idx:
        idx.append_synthetic(seed=1, row0=0, n=rows, normalize=True)
        Q = hc.Q if hc is not None else make_queries(a.queries, a.dim, dev.index or 0)
        ms, spread, ids, sc = k2_stream_bench(idx, Q, a.k, steps, 20, 3, dev, stream)
        e2e_ms, _, lat = sync_e2e_bench(idx.search_ptr, Q, a.k, steps, 3, __import__("torch").cuda.synchronize)
        fails = []
        if hc is not None:
            fails = [f for f in (hc.check(i, a.k, ids[i], sc[i], prefix=rows) for i in range(min(N_ORACLE_QUERIES, len(Q)))) if f]
    rec = {"metric": metric_name(rows, a.dim, a.k), "value": 1e3 / ms, "unit": "queries/s", "ms_per_step": ms, "steps": steps,
           "repeats": spread, "gpu_launches": int(k2_stream_bench.launches),
           "roofline": hbm_roofline(rows, a.dim, ms, "scan_stream_kernel (K2, one persistent launch for the stream of %d queries)" % steps
                                    if k2_stream_bench.launches < steps else "scan_topk_tma_kernel (K2, one chained launch per query)"),
           "e2e": {"value": 1e3 / e2e_ms, "unit": "queries/s", "ms_per_step": e2e_ms, "latency": lat,
                   "h2d_bytes_per_step": a.dim * 4, "d2h_bytes_per_step": 8 + 12 * a.k},
           "oracle_checked_queries": 0 if hc is None else min(N_ORACLE_QUERIES, len(Q)), "oracle_failures": fails,
           "verified": (not fails) if hc is not None else None}
    return rec


BATCH_IDLE_GAP_S = 0.4


def batch_measure(idx, Qd, nq, k, mode, steps, stream):
    """One batched-search configuration on the tensor cores: -> (ms per batch, launches per batch, ids, scores, nf)."""
    import torch
    dev = Qd.device
    ids_d = torch.zeros(nq * k, dtype=torch.int64, device=dev)
    sc_d = torch.zeros(nq * k, dtype=torch.float32, device=dev)
    nf_d = torch.zeros(nq, dtype=torch.int32, device=dev)
    idx.set_batch_mode(mode)
    idx.set_stream(stream.cuda_stream)
    for _ in range(3):
        idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    torch.cuda.synchronize()
    time.sleep(BATCH_IDLE_GAP_S)      # the same short idle gap before every timed region: the board sits at its power cap
    l0 = idx.launch_count             # under these kernels, and without it the previous mode's heat sets this one's clocks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        idx.search_batch_device(Qd.data_ptr(), nq, k, ids_d.data_ptr(), sc_d.data_ptr(), nf_d.data_ptr())
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    launches = (idx.launch_count - l0) / steps
    idx.set_stream(None)
    return (ms, launches, ids_d.cpu().numpy().astype(np.uint64).reshape(nq, k), sc_d.cpu().numpy().reshape(nq, k),
            nf_d.cpu().numpy())


def batch_modes(fmt):
    """mode -> (name, dtype text, tensor passes issued); fmt = element format of the split (idx.batch_precision_active)."""
    f = "fp16" if fmt == 1 else "bf16"
    return {2: (f"{f}x3", f"3 x {f} split passes (q_hi.x_hi + q_lo.x_hi + q_hi.x_lo) -> f32 accumulate, exact f32 re-scoring", 3.0),
            3: (f"{f}x1", f"1 {f} pass as candidate filter -> exact f32 re-scoring (+ K2 for unproven queries)", 1.0),
            0: ("cascade", f"1 {f} pass -> {f}x3 only for queries the loose bound cannot prove -> K2; exact f32 re-scoring", 1.0)}


BATCH_MODES = batch_modes(0)


def batch_record(rows, dim, nq, k, mode, ms, launches, stats, fmt=0):
    peaks, src = load_peaks()
    peak = float(peaks.get("bf16_tflops", 1590.0))
    sustained = peaks.get("bf16_tflops_sustained")
    flop = 2.0 * nq * rows * dim
    ach = flop / (ms * 1e-3) / 1e12
    name, dtype, passes = batch_modes(fmt)[mode]
    return {"mode": name, "dtype": dtype, "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms,
            "gpu_launches_per_batch": launches,
            "roofline": {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                         "issued_passes": passes, "issued_frac": passes * ach / peak,
                         "frac_of_sustained": None if not sustained else ach / float(sustained),
                         "peak_source": src + " bf16_tflops (cuBLAS burst)", "traffic": None,
                         "note": "achieved = algorithmic 2*Q*N*d FLOP / device time per batch; issued = x passes of the split"},
            **stats}


def sub_batch(a, idx, hc, dev, stream, k2_ids, k2_sc):
    """BASELINE.json configs[2]: 10M x 384, batches of 1024 queries on the tensor cores (K3), both the north star's
    bf16x3 split and the automatic precision cascade.  Checked against the oracle (first queries) and bit-for-bit
    against the single-query kernel's results for the headline's query pool."""
    import torch
    nq, k = 1024, a.k
    Q = hc.Q1024 if hc is not None else make_queries(nq, a.dim, dev.index or 0)
    Qd = torch.from_numpy(Q).to(dev)
    out = {"workload": f"{a.rows}x{a.dim} fp32 corpus, batches of {nq} queries, exact top-{k} (BASELINE configs[2])",
           "metric": f"qps_batched_{nq}q_exact_top{k}_cosine_{a.rows}x{a.dim}_fp32", "modes": {}}
    # the three-pass split and the cascade in the default element format (fp16 halves on unit-norm rows), then the
    # three-pass split with bf16 halves (the north star's 3 x bf16 by name)
    for prec, mode in ((0, 0), (0, 2), (1, 2)):
        idx.set_batch_precision(prec)
        q0, f0 = idx.batch_stats()
        c0 = idx.batch_cascaded
        ms, launches, ids, sc, nf = batch_measure(idx, Qd, nq, k, mode, 6, stream)
        fmt = idx.batch_precision_active
        q1, f1 = idx.batch_stats()
        batches = max((q1 - q0) // nq, 1)
        fails = []
        if hc is not None:
            fails = [f for f in (hc.check(i, k, ids[i, :nf[i]], sc[i, :nf[i]]) for i in range(N_ORACLE_QUERIES)) if f]
        same_as_k2 = None
        if k2_ids is not None:
            m = min(len(k2_ids), a.queries, nq)
            same_as_k2 = bool(np.array_equal(ids[:m], k2_ids[:m]) and np.array_equal(sc[:m], k2_sc[:m]))
        stats = {"k2_fallback_queries_per_batch": (f1 - f0) / batches, "cascaded_queries_per_batch": (idx.batch_cascaded - c0) / batches,
                 "oracle_checked_queries": 0 if hc is None else N_ORACLE_QUERIES, "oracle_failures": fails,
                 "bit_identical_to_k2_on_first_queries": same_as_k2,
                 "verified": (not fails and same_as_k2 is not False) if (hc is not None or same_as_k2 is not None) else None}
        rec = batch_record(a.rows, a.dim, nq, k, mode, ms, launches, stats, fmt)
        out["modes"][rec["mode"]] = rec
    idx.set_batch_precision(0)
    idx.set_batch_mode(0)
    out["value"] = out["modes"]["cascade"]["value"]           # what sema_index_search_batch does by default
    out["value_is"] = "cascade (automatic mode, the call's default); the pure three-pass splits are listed beside it"
    out["timing"] = (f"per mode: 3 warm-up batches, {BATCH_IDLE_GAP_S} s idle, then 6 batches back to back between two CUDA events on the "
                     "launching stream")
    out["unit"] = "queries/s"
    out["verified"] = all(m["verified"] is not False for m in out["modes"].values())
    return out


def ingest_measure(dim, k, rows, B, dev_index, oracle_rows):
    """BASELINE.json configs[4]: the index grows to rows x dim by appending batches of B rows from pinned host memory
    (H2D + K1 on the ingest stream) while top-k queries run on the query stream; every query scans exactly the rows
    whose ingest had completed when it started (snapshot semantics).  The first `oracle_rows` rows are unique
    synthetic rows the oracle also holds, so every query whose snapshot lies inside them is re-derived by the oracle
    on exactly that snapshot; later batches cycle the last few host buffers (the host never holds 30 GB)."""
    import torch
    import sema_b200
    from oracle import c_oracle, oracle as O            # checker for the snapshot results
    nb = (rows + B - 1) // B
    n_unique = min(nb, max((oracle_rows + B - 1) // B, 3))
    c_oracle.use_all_cores()
    raw = c_oracle.synth(7, 0, n_unique * B, dim)        # un-normalised rows, as an embedder would hand them over
    pool = [torch.from_numpy(raw[i * B:(i + 1) * B]).pin_memory() for i in range(n_unique)]
    host = [t.numpy() for t in pool]
    Xo = c_oracle.normalize(raw)                         # what the index must hold for those rows (K1 is bit-exact)
    del raw
    Q = c_oracle.normalize(c_oracle.synth(8, 0, 64, dim))
    ids_h, sc_h = np.zeros(k, dtype=np.uint64), np.zeros(k, dtype=np.float32)
    idx = sema_b200.GpuIndex(dim, nb * B, device=dev_index)     # warm-up: one batch + a few queries, then start over
    idx.append(host[0], normalize=True)
    for i in range(5):
        idx.search_into(Q[i], k, ids_h, sc_h)
    idx.close()
    idx = sema_b200.GpuIndex(dim, nb * B, device=dev_index)
    total_rows = nb * B
    results = []                                          # (query index, snapshot, ids, scores) of the checkable queries
    snaps, nq = [], 0
    torch.cuda.synchronize()
    with ClockSampler(dev_index) as clk:
        t0 = time.perf_counter()
        for b in range(nb):                               # H2D + K1 per batch, all on the ingest stream
            hb = host[b] if b < n_unique else host[n_unique - 3 + (b % 3)]
            idx.append(hb, normalize=True, asynchronous=True)
        t_enq = time.perf_counter() - t0
        while True:                                       # K2 on the query stream, while the ingest runs
            v = idx.visible
            if v >= total_rows:
                break
            if v == 0:
                time.sleep(0.0002)
                continue
            nf = idx.search_into(Q[nq % 64], k, ids_h, sc_h)
            s = idx.last_snapshot
            snaps.append(s)
            if s <= n_unique * B:
                results.append((nq % 64, s, ids_h[:nf].copy(), sc_h[:nf].copy()))
            nq += 1
        idx.flush()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    ok_shape = idx.visible == total_rows and all(s % B == 0 and s > 0 for s in snaps) and snaps == sorted(snaps)
    # a final query on the complete index, and a stored row must find itself
    probe = idx.read_rows(total_rows - 1, 1)[0]
    r_ids, r_sc = idx.search(probe, 1)
    launches = idx.launch_count
    idx.close()
    fails, checked = [], 0
    for qi, s, ids, sc in results[:: max(len(results) // 24, 1)][:32]:      # bounded CPU time: <= 32 oracle scans
        r_i, r_s = c_oracle.scan(Xo[:s], Q[qi], k)
        checked += 1
        try:
            O.check_parity(ids, sc, r_i, r_s)
        except AssertionError as e:
            fails.append(f"snapshot {s}, query {qi}: {e}")
    scanned_bytes = float(sum(snaps)) * dim * 4
    return {
        "metric": f"streaming_ingest_rows_per_s_with_interleaved_top{k}_queries_{total_rows}x{dim}_fp32",
        "workload": f"append {nb} batches of {B} x {dim} fp32 rows from pinned host memory (normalise on the GPU) while "
                    f"top-{k} queries run back to back on the query stream (BASELINE configs[4])",
        "value": total_rows / dt, "unit": "rows/s", "rows": total_rows, "dim": dim, "k": k, "seconds": dt,
        "h2d_GBps": total_rows * dim * 4 / dt / 1e9, "queries": nq, "qps_during_ingest": nq / dt, "enqueue_s": t_enq,
        "mean_snapshot_rows": float(np.mean(snaps)) if snaps else 0.0, "query_scan_GBps": scanned_bytes / dt / 1e9,
        "snapshots_monotonic_and_batch_aligned": bool(ok_shape),
        "self_probe_ok": bool(len(r_ids) == 1 and int(r_ids[0]) % B == (total_rows - 1) % B and abs(float(r_sc[0]) - 1.0) < 1e-5),
        "oracle_checked_queries": checked, "oracle_checkable_queries": len(results), "oracle_rows": n_unique * B,
        "oracle_failures": fails, "verified": bool(ok_shape and not fails and checked > 0),
        "gpu_launches": int(launches), "clocks": clk.summary(),
    }


def sub_sharded_100m(a, dist, host_barrier, dev, stream, world, rank, local):
    """BASELINE.json configs[3]: 100M x 384 generated on the device per shard (row offset = shard base, so the corpus
    does not depend on the shard count), sharded over the N GPUs, fused exchange.  Checks: (i) fused == NCCL all-gather
    + K4; (ii) every hit's score re-derived by the oracle from the regenerated row; (iii) against the oracle's top-k of
    the 1M-row prefix: every global hit below row 1M is in it and nothing in it that beats the global k-th is missing."""
    import torch
    import sema_b200
    from sema_b200.sharded import ShardedSearcher, make_shard_group
    rows, k, dim = a.rows_sharded, a.k, a.dim
    per = (rows + world - 1) // world
    lo, hi = min(rank * per, rows), min((rank + 1) * per, rows)
    idx = sema_b200.GpuIndex(dim, max(hi - lo, 1), device=local)
    idx.set_row_base(lo)
    idx.append_synthetic(seed=1, row0=lo, n=hi - lo, normalize=True)
    Q = make_queries(a.queries, dim, local)
    group = make_shard_group(idx, dist)
    idx.set_stream(stream.cuda_stream)
    steps, warm = 20, 5
    Qd = torch.from_numpy(Q).to(dev)
    Qs = Qd[torch.arange(steps, device=dev) % a.queries].contiguous()
    ids_s = torch.zeros((steps, k), dtype=torch.int64, device=dev)
    sc_s = torch.zeros((steps, k), dtype=torch.float32, device=dev)
    nf_s = torch.zeros(steps, dtype=torch.int32, device=dev)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    group.search_stream_device(Qs.data_ptr(), warm, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(3):
        barrier()
        e0.record(stream)
        group.search_stream_device(Qs.data_ptr(), steps, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
        e1.record(stream)
        barrier()
        ms.append(e0.elapsed_time(e1) / steps)
    t = torch.tensor(ms, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, spread = median_spread(t.tolist())
    ids = ids_s.cpu().numpy().astype(np.uint64)
    sc = sc_s.cpu().numpy()
    # (i) fused == NCCL path
    sh = ShardedSearcher(idx, dist, k)
    same = True
    for i in range(4):
        n_ids, n_sc = sh.search(Q[i])
        same &= bool(np.array_equal(ids[i], n_ids) and np.array_equal(sc[i], n_sc))
    v = torch.tensor([int(same)], device=dev)
    dist.all_reduce(v, op=dist.ReduceOp.MIN)
    same = bool(v.item())
    rec = None
    if rank == 0:
        fails = []
        if not a.no_verify:
            from oracle import c_oracle, oracle as O
            c_oracle.use_all_cores()
            prefix = min(1_000_000, rows)
            Xp = c_oracle.normalize_inplace(c_oracle.synth(1, 0, prefix, dim))
            for i in range(4):
                hit_rows = np.stack([c_oracle.normalize(c_oracle.synth(1, int(r), 1, dim))[0] for r in ids[i]])
                o_ids, o_sc = c_oracle.scan(hit_rows, Q[i], k)            # oracle arithmetic on the regenerated hit rows
                try:
                    O.check_parity(ids[i], sc[i], ids[i][o_ids.astype(np.int64)], o_sc)
                except AssertionError as e:
                    fails.append(f"query {i} (scores / order of the hits): {e}")
                p_ids, p_sc = c_oracle.scan(Xp, Q[i], k)
                kth = float(sc[i][-1])
                in_prefix = set(int(r) for r in ids[i] if r < prefix)
                must = set(int(r) for r, s in zip(p_ids, p_sc) if s > kth + 1e-5)
                if not in_prefix <= set(int(r) for r in p_ids) or not must <= in_prefix:
                    fails.append(f"query {i}: result disagrees with the oracle on the {prefix}-row prefix")
        rec = {"metric": metric_name(rows, dim, k), "workload": f"{rows}x{dim} fp32 synthetic corpus row-sharded over {world} GPUs, "
               f"single-query exact top-{k}, fused exchange (BASELINE configs[3])", "value": 1e3 / ms_step, "unit": "queries/s",
               "ms_per_step": ms_step, "steps": steps, "repeats": spread, "rows_per_gpu": hi - lo,
               "roofline": hbm_roofline(hi - lo, dim, ms_step, "scan_topk_tma_kernel (K2) + fused exchange"),
               "fused_equals_nccl_path": same, "oracle_failures": fails,
               "result_checksum": [int(x) for x in ids[0]],
               "verified": bool(same and not fails) if not a.no_verify else same}
    barrier()
    group.close()
    idx.close()
    return rec


# ----------------------------------------------------------------------------- GPU arm: the headline
def run_ours(a):
    import torch

    import sema_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus != world and world == 1 and a.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torchrun (one process per GPU)")
    need_gpu()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    host_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        host_group = dist.new_group(backend="gloo")       # host-side barriers (no kernel spinning on the GPUs)

    # ---- corpus: rows [lo, hi) of the fixed synthetic corpus live on this GPU
    per = (a.rows + world - 1) // world
    lo, hi = min(rank * per, a.rows), min((rank + 1) * per, a.rows)
    idx = sema_b200.GpuIndex(a.dim, max(hi - lo, 1), device=local, growable=a.growable)
    idx.set_row_base(lo)
    idx.append_synthetic(seed=1, row0=lo, n=hi - lo, normalize=True)
    if a.variant >= 0:
        idx.set_scan_variant(a.variant)
    Q = make_queries(a.queries, a.dim, local)
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
    n_s = max(a.steps, a.warmup, a.queries, 1)
    Qs = Qd[torch.arange(n_s, device=dev) % a.queries].contiguous()      # query i of the stream = pool[i % pool]
    ids_s = torch.zeros((n_s, k), dtype=torch.int64, device=dev)
    sc_s = torch.zeros((n_s, k), dtype=torch.float32, device=dev)
    nf_s = torch.zeros(n_s, dtype=torch.int32, device=dev)
    if a.no_chain:
        idx.set_scan_variant(600)
    if a.no_chain or a.launch_per_query:
        idx.set_scan_variant(902)
    searcher = idx if world == 1 else group

    def run_device(nq):
        if nq <= 0:
            return
        if use_stream:
            searcher.search_stream_device(Qs.data_ptr(), nq, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
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

    # ---- reference results of the pool queries, one unchained call each, taken before anything is timed: the
    # timed (chained) stream must reproduce them bit for bit whatever --steps is
    pool_ids = np.zeros((a.queries, k), dtype=np.uint64)
    pool_sc = np.zeros((a.queries, k), dtype=np.float32)
    for i in range(a.queries):
        step_device(i)
        torch.cuda.synchronize()
        pool_ids[i], pool_sc[i] = ids_d.cpu().numpy().astype(np.uint64), sc_d.cpu().numpy()
    barrier()

    # ---- device-timed region: inputs resident in HBM, CUDA events on the launching stream.  The region of exactly
    # --steps steps is measured --repeats times; the median is the reported number, the spread is kept.
    run_device(a.warmup)
    barrier()
    l0 = idx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(a.repeats, 1)
    dev_ms_all = []
    with ClockSampler(local) as clk:
        for _ in range(reps):
            barrier()
            e0.record(stream)
            run_device(a.steps)
            e1.record(stream)
            barrier()
            dev_ms_all.append(e0.elapsed_time(e1))
        launches = (idx.launch_count - l0) // reps
        if use_stream:
            stream_ids = ids_s[:a.steps].cpu().numpy().astype(np.uint64)
            stream_sc = sc_s[:a.steps].cpu().numpy()

        # ---- end-to-end region: host query in, host results out, through the public call.  The clock sampler keeps
        # running but ten times less often: NVML queries take driver locks, and a host-driven loop of 2 ms calls is
        # exposed to that where a pre-enqueued device stream is not (on the power-capped boxes of this pool the five
        # repeats read 2.13 - 2.29 ms per call with a sample every 20 ms and 2.062 - 2.066 ms with one every 200 ms;
        # the spread of this region is reported in `repeats`)
        clk.period = 0.2
        e2e_all, e2e_lat, e2e_pipe_ms = [], None, None
        if a.staged_host_path:
            idx.set_scan_variant(500)
        if world == 1:
            idx.set_stream(None)
            e2e_ms, e2e_spread, e2e_lat = sync_e2e_bench(idx.search_ptr, Q, k, a.steps, reps, torch.cuda.synchronize)
            e2e_all = [x * a.steps for x in e2e_spread["all"]]
            e2e_pipe_ms = pipelined_e2e(idx, Q, a.steps, k, torch.cuda.synchronize) * a.steps
        elif group is not None:
            e2e_ms, e2e_spread, e2e_lat = sync_e2e_bench(group.search_ptr, Q, k, a.steps, reps, barrier)
            e2e_all = [x * a.steps for x in e2e_spread["all"]]
            e2e_pipe_ms = pipelined_e2e(group, Q, a.steps, k, barrier) * a.steps
        else:
            from sema_b200.sharded import ShardedSearcher
            sh = ShardedSearcher(idx, dist, k)
            for i in range(min(a.warmup, 5)):
                sh.search(Q[i % a.queries])
            for _ in range(reps):
                barrier()
                t0 = time.perf_counter()
                for i in range(a.steps):
                    sh.search(Q[i % a.queries])
                barrier()
                e2e_all.append((time.perf_counter() - t0) * 1e3)
    if dist is not None:
        t = torch.tensor(dev_ms_all + e2e_all + [e2e_pipe_ms or 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)               # max over ranks, repeat by repeat
        t = t.tolist()
        dev_ms_all, e2e_all, e2e_pipe_ms = t[:reps], t[reps:2 * reps], (t[2 * reps] or None)
    dev_ms, dev_spread = median_spread([x / a.steps for x in dev_ms_all])
    e2e_step_ms, e2e_spread = median_spread([x / a.steps for x in e2e_all])
    qps = 1e3 / dev_ms
    e2e_qps = 1e3 / e2e_step_ms

    # ---- N > 1: every rank's scan of its own shard WITHOUT the exchange (same stream of queries, local top-k only).
    # The group runs at the pace of its slowest GPU; this shows how much of the gap to N x the 1-GPU rate is the
    # spread between the N boards (the N = 1 run only ever measures GPU 0) and how much the exchange itself costs.
    per_rank_local = None
    if world > 1 and use_stream:
        idx.search_stream_device(Qs.data_ptr(), a.warmup, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
        loc = []
        for _ in range(reps):
            barrier()
            e0.record(stream)
            idx.search_stream_device(Qs.data_ptr(), a.steps, k, ids_s.data_ptr(), sc_s.data_ptr(), nf_s.data_ptr())
            e1.record(stream)
            torch.cuda.synchronize()
            loc.append(e0.elapsed_time(e1) / a.steps)
        t = torch.zeros(world, dtype=torch.float64, device=dev)
        t[rank] = sorted(loc)[len(loc) // 2]
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        per_rank_local = [round(x, 6) for x in t.tolist()]

    # ---- every result of the timed stream must equal the pre-computed result of its pool query (query i =
    # pool[i % pool]): any race between chained launches, or between a launch and the peer exchange, breaks this
    stream_consistent = None
    if use_stream:
        which = np.arange(a.steps) % a.queries
        stream_consistent = bool(np.array_equal(stream_ids, pool_ids[which]) and np.array_equal(stream_sc, pool_sc[which]))
        if dist is not None:
            c = torch.tensor([int(stream_consistent)], device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.MIN)
            stream_consistent = bool(c.item())

    # ---- single-process group (the drop-in handle: ONE process owns every GPU, like the reference's one
    # StorageManager): rank 0 drives all N GPUs through sema_shard_group_create_local while the other ranks'
    # processes sit idle in a host-side barrier
    single_proc = None
    if world > 1 and not a.no_extra:
        if rank == 0:
            try:
                single_proc = single_process_e2e(a, world, Q, pool_ids, pool_sc)
            except Exception as e:
                single_proc = {"unavailable": str(e)[:300]}
        dist.barrier(group=host_group)

    # ---- verification: against a fresh host-API search, against the NCCL path (N > 1), and against the oracle
    verified, oracle_fail, hc = None, None, None
    if not a.no_verify and world == 1:
        verified = True
        for j in range(min(a.queries, 8)):
            r_ids, r_sc = idx.search(Q[j], k)
            verified &= bool(np.array_equal(pool_ids[j], r_ids) and np.array_equal(pool_sc[j], r_sc))
    elif not a.no_verify:
        # multi-rank: the fused result must equal the NCCL all-gather + K4 result, and every global
        # hit that lives on this rank must be this rank's own local hit
        from sema_b200.sharded import ShardedSearcher
        sh = ShardedSearcher(idx, dist, k)
        verified = True
        for i in range(4):
            n_ids, n_sc = sh.search(Q[i])
            verified &= bool(np.array_equal(pool_ids[i], n_ids) and np.array_equal(pool_sc[i], n_sc))
            l_ids, l_sc = idx.search(Q[i], k)
            mine = (n_ids >= lo) & (n_ids < hi)
            verified &= bool(set(n_ids[mine].tolist()) <= set(l_ids.tolist()))
        v = torch.tensor([int(verified)], device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MIN)
        verified = bool(v.item())

    line = None
    if rank == 0:
        need_host = not a.no_verify or (world == 1 and not a.no_cpu_baseline)
        ram = host_ram_bytes()
        if need_host and (not ram or a.rows * a.dim * 4 < 0.6 * ram):
            hc = HostCorpus(a.rows, a.dim, max(a.queries, 1024))
            hc.Q1024, hc.Q = hc.Q, hc.Q[:a.queries]
        if hc is not None and not a.no_verify:
            # the ORACLE re-derives the top-k of the first pool queries over the whole corpus (its own rows, its own
            # arithmetic); the results of the timed stream were shown above to equal pool_ids / pool_sc bit for bit
            oracle_fail = [f for f in (hc.check(i, k, pool_ids[i], pool_sc[i])
                                        for i in range(min(N_ORACLE_QUERIES, a.queries))) if f]
            oracle_fail += [] if np.array_equal(Q, hc.Q) else ["K1-normalised queries differ from the oracle's bits"]
        shard_rows = hi - lo
        traffic = None
        try:   # DRAM bytes per launch from the committed `ncu --set full` capture of this exact workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if world == 1 and a.k == 10:
                traffic = tr.get(f"k2_single_{a.rows}x{a.dim}")
        except Exception:
            pass
        # one persistent launch for the whole region (k2_stream.cuh) when the library chose it: launches < steps
        persistent = bool(use_stream and launches < a.steps)
        kern = ("scan_stream_kernel (K2, TMA ring, one persistent launch per query stream)" if persistent
                else "scan_topk_tma_kernel (K2, TMA ring)" if (a.variant <= 0 and a.dim in (384, 768)) else "scan_topk_kernel (K2, register-fed)")
        tsrc = "profiles/r01_k2_scan_full_raw.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch = one pass over the corpus)"
        if persistent:
            try:   # the persistent launch's own capture: DRAM bytes of the launch divided by its passes
                t2 = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
                per_pass = t2.get(f"k2_stream_{a.rows}x{a.dim}_per_pass") if (world == 1 and a.k == 10) else None
                traffic = per_pass * a.steps if per_pass else None
                tsrc = (f"profiles/r02_k2_stream_full_raw.csv: ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of "
                        f"{t2.get('k2_stream_capture_passes')} passes, per pass ({per_pass} B), times this launch's {a.steps} passes")
            except Exception:
                traffic = None
        roof = hbm_roofline(shard_rows, a.dim, dev_ms, kern, traffic)
        roof["traffic_source"] = tsrc if traffic else None
        tail = "" if world == 1 else ", which includes the top-k exchange and the global merge"
        if persistent:
            roof["passes_per_launch"] = a.steps
            roof["algorithmic_bytes"] = int(shard_rows) * a.dim * 4 * a.steps      # per launch, like traffic
            roof["algorithmic_bytes_per_pass"] = int(shard_rows) * a.dim * 4
            roof["note"] = (f"ONE K2 launch scans the corpus once per step ({a.steps} passes per launch): achieved = "
                            "passes * rows_per_gpu*dim*4 bytes / launch duration = rows_per_gpu*dim*4 / device time per step" + tail)
        else:
            roof["note"] = "achieved = rows_per_gpu*dim*4 bytes / device time per step (one K2 launch per step" + tail + ")"
        all_ok = all(x is not False for x in (verified, stream_consistent)) and not oracle_fail
        line = {
            "metric": metric_name(a.rows, a.dim, a.k), "value": qps, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": shared_config(a),
            "setup": {
                "rows_per_gpu": shard_rows,
                "parallelism": "single GPU" if world == 1 else f"corpus row-sharded over {world} GPUs (one process each)",
                "exchange": exchange if world > 1 else None,
                "issue": ("one query stream of K queries (sema_index_search_stream_device / sema_shard_group_search_stream_device): "
                          + ("ONE persistent K2 launch for the stream — the producer warp streams query i+1's rows while a finisher warp "
                             "merges, exchanges and publishes query i (gpu_launches counts launches, not passes)" if persistent else
                             "one K2 launch per query, " + ("unchained" if a.no_chain else "consecutive launches chained with programmatic dependent launch")))
                         if use_stream else "one search call per query",
                "timing": "CUDA events on the launching stream, barrier + synchronize both sides, max over ranks; "
                          f"the region of exactly {a.steps} steps is timed {reps} times, value = median",
                "per_rank_local_scan_ms": per_rank_local,
                "exchange_cost_ms_per_step": None if not per_rank_local else dev_ms - max(per_rank_local),
                "per_rank_note": None if not per_rank_local else
                    "ms per step of each rank's shard scan alone (no exchange, same query stream); the group cannot beat the slowest rank",
            },
            "repeats": {"device_ms_per_step": dev_spread, "e2e_ms_per_step": e2e_spread},
            "roofline": roof,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": ((a.dim + 3) // 4) * 16,
                    "d2h_bytes_per_step": 8 + 12 * k, "ms_per_step": e2e_step_ms, "latency": e2e_lat,
                    "pipelined": None if not e2e_pipe_ms else {
                        "value": a.steps / (e2e_pipe_ms / 1e3), "unit": "queries/s", "ms_per_step": e2e_pipe_ms / a.steps,
                        "path": "same host buffers through the submit / collect form of the call, two searches in flight"},
                    "single_process": single_proc,
                    "path": ("sema_index_search" if world == 1 else "sema_shard_group_search (one process per GPU)" if group is not None else "sharded.ShardedSearcher.search")
                            + " with host buffers: the query travels in the kernel parameters (these bytes), K2 (+ exchange + merge) stores"
                              " the result block into mapped host memory (these bytes), the call polls its completion flag"},
            "gpu_launches": int(launches),
            "clocks": clk.summary(),
            "verified": bool(all_ok) if not a.no_verify else None,
            "verification": {
                "stream_equals_precomputed_pool_results": stream_consistent,
                "equals_fresh_host_search" if world == 1 else "fused_equals_nccl_allgather_k4": verified,
                "oracle_checked_queries": None if hc is None or a.no_verify else min(N_ORACLE_QUERIES, a.queries),
                "oracle_failures": oracle_fail,
                "oracle": None if hc is None else f"oracle/cpu_scan.c over its own copy of the whole {a.rows} x {a.dim} corpus ({hc.build_s:.1f} s to build on {hc.cores} cores)",
            },
        }

    # ---- the other BASELINE.json configs, as compact sub-records
    if not a.no_extra and a.workload == "single":
        extra = {}
        if world == 1:
            def attempt(name, fn):
                try:
                    extra[name] = fn()
                except Exception as e:                       # a sub-record never costs the headline
                    extra[name] = {"failed": f"{type(e).__name__}: {e}"[:400]}
            attempt("config1_10k_chunks_storage_manager", lambda: {kk: vv for kk, vv in config1_measure(a.k).items()
                                                                    if kk in ("metric", "value", "unit", "ms_per_step", "config", "latency_ms",
                                                                              "oracle_checked_queries", "oracle_failures", "verified")})
            attempt("1Mx384_single", lambda: sub_single_1m(a, hc, dev, stream))
            if a.dim % 64 == 0 and a.dim <= 768:
                attempt("batch_1024q", lambda: sub_batch(a, idx, hc, dev, stream, pool_ids, pool_sc))
            idx.close()
            idx = None
            attempt("ingest_10Mx768_k100", lambda: ingest_measure(768, 100, 10_000_000, a.ingest_batch, local, 1_000_000))
        else:
            idx.set_stream(None)
            if group is not None:
                group.close()
                group = None
            idx.close()
            idx = None
            try:
                rec = sub_sharded_100m(a, dist, host_group, dev, stream, world, rank, local)
            except Exception as e:
                rec = {"failed": f"{type(e).__name__}: {e}"[:400]}
            if rank == 0:
                extra[f"{a.rows_sharded // 1_000_000}Mx{a.dim}_sharded"] = rec
        if line is not None:
            line["configs"] = extra

    if rank == 0:
        if world == 1 and not a.no_cpu_baseline and hc is not None:
            try:
                line["cpu_baseline"] = cpu_baseline_record(hc, a, a.cpu_steps, 2)
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": None, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(group=host_group)
        dist.destroy_process_group()


def single_process_e2e(a, world, Q, pool_ids, pool_sc):
    """Rank 0 only: shards of the same corpus on every GPU of the box, one sema_shard_group_create_local handle,
    one synchronous host call per query (and the submit / collect form).  Results must equal the multi-process ones."""
    import torch
    import sema_b200
    per = (a.rows + world - 1) // world
    shards = []
    grp = None
    try:
        for g in range(world):
            lo, hi = min(g * per, a.rows), min((g + 1) * per, a.rows)
            s = sema_b200.GpuIndex(a.dim, max(hi - lo, 1), device=g)
            s.set_row_base(lo)
            s.append_synthetic(seed=1, row0=lo, n=hi - lo, normalize=True)
            shards.append(s)
        grp = sema_b200.ShardGroup.local(shards)

        def sync():
            for g in range(world):
                torch.cuda.synchronize(g)

        same = True
        for i in range(min(8, len(Q))):
            ids, sc = grp.search(Q[i], a.k)
            same &= bool(np.array_equal(ids, pool_ids[i]) and np.array_equal(sc, pool_sc[i]))
        ms, spread, lat = sync_e2e_bench(grp.search_ptr, Q, a.k, a.steps, max(a.repeats, 1), sync)
        pipe = pipelined_e2e(grp, Q, a.steps, a.k, sync)
        return {"value": 1e3 / ms, "unit": "queries/s", "ms_per_step": ms, "repeats": spread, "latency": lat,
                "pipelined": {"value": 1e3 / pipe, "unit": "queries/s", "ms_per_step": pipe},
                "equals_multi_process_results": same,
                "path": f"sema_shard_group_create_local over {world} GPUs in ONE process: one sema_shard_group_search call per query, "
                        "host buffers, worker-thread fan-out, fused peer exchange"}
    finally:
        if grp is not None:
            grp.close()
        for s in shards:
            s.close()
        torch.cuda.set_device(0)


def run_batch(a):
    """BASELINE.json configs[2] alone: 10M x 384, batched nq-query top-k on one B200 (kernel K3), one mode."""
    import torch

    import sema_b200

    need_gpu()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    k, nq = a.k, a.nq
    idx = sema_b200.GpuIndex(a.dim, a.rows, device=0)
    idx.append_synthetic(seed=1, row0=0, n=a.rows, normalize=True)
    if a.k3_cluster:
        idx.set_scan_variant(100 + a.k3_cluster)
    if a.k3_pair >= 0:
        idx.set_scan_variant(700 + a.k3_pair)
    idx.set_batch_precision(a.precision)
    Q = make_queries(nq, a.dim, 0)
    stream = torch.cuda.current_stream()
    Qd = torch.from_numpy(Q).to(dev)
    steps = a.steps
    with ClockSampler(0) as clk:
        ms, launches, ids, sc, nf = batch_measure(idx, Qd, nq, k, a.batch_mode, steps, stream)
        idx.set_batch_mode(a.batch_mode)
        for _ in range(2):
            idx.search_batch(Q, k)
        t0 = time.perf_counter()
        for _ in range(steps):
            ids_h, sc_h, nf_h = idx.search_batch(Q, k)
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
    served, fallbacks = idx.batch_stats()
    ok = True                                   # spot-check against the single-query kernel
    for i in (0, nq // 2, nq - 1):
        r_ids, r_sc = idx.search(Q[i], k)
        ok &= bool(np.array_equal(ids_h[i, :nf_h[i]], r_ids) and np.array_equal(sc_h[i, :nf_h[i]], r_sc))
    rec = batch_record(a.rows, a.dim, nq, k, a.batch_mode if a.batch_mode in BATCH_MODES else 2, ms, launches,
                       {"k3_queries": served, "k3_fallback_queries": fallbacks, "cascaded_queries": idx.batch_cascaded},
                       idx.batch_precision_active)
    line = {
        "metric": f"qps_batched_{nq}q_exact_top{k}_cosine_{a.rows}x{a.dim}_fp32", "value": rec["value"],
        "unit": "queries/s", "n_gpus": 1, "steps": steps, "warmup": 3, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": rec["dtype"], "data": "synthetic",
        "config": {"workload": f"{a.rows}x{a.dim} fp32 corpus, batches of {nq} queries, exact top-{k} (BASELINE configs[2])",
                   "batch_mode": rec["mode"], "k3_cluster": a.k3_cluster or "auto", "k3_pair": "default (single-CTA kernel)" if a.k3_pair < 0 else a.k3_pair,
                   "l2_flush": "none needed: each batch streams the 16-bit planes of the whole corpus"},
        "roofline": rec["roofline"], "batch": rec,
        "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": nq * a.dim * 4,
                "d2h_bytes_per_step": nq * (k * 12 + 4), "ms_per_step": e2e_ms,
                "path": "sema_index_search_batch (C ABI) with host buffers"},
        "gpu_launches": int(round(launches * steps)), "clocks": clk.summary(), "verified_against_k2": ok,
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
        peaks, _ = load_peaks()
        peak = float(peaks.get("bf16_tflops", 1590.0))
        achieved = 2.0 * nq * (hi - lo) * a.dim / (dev_ms * 1e-3) / 1e12        # per GPU
        line = {
            "metric": f"qps_batched_{nq}q_exact_top{k}_cosine_{a.rows}x{a.dim}_fp32", "value": nq / (dev_ms * 1e-3),
            "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": dev_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": batch_modes(idx.batch_precision_active).get(a.batch_mode, batch_modes(idx.batch_precision_active)[2])[1], "data": "synthetic",
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


def run_maintenance(a):
    """SURVEY.md 8(f) rows 1-2 measured: remove_file_chunks' device side (tombstones: src/storage/lance_indexer.rs:234-250),
    compaction, and the on-disk vector cache.  Every phase is followed by oracle checks of searches on the result."""
    import tempfile

    import torch

    import sema_b200
    need_gpu()
    torch.cuda.set_device(0)
    rows, dim, k = a.maint_rows, a.dim, a.k
    hc = HostCorpus(rows, dim, 16)
    rng = np.random.default_rng(5)
    peaks, _ = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rec = {}
    file_rows = np.arange(rows // 3, rows // 3 + 50, dtype=np.uint64)
    scattered = np.sort(rng.choice(rows, rows // 10, replace=False)).astype(np.uint64)
    with sema_b200.GpuIndex(dim, rows, device=0) as idx:

        def check(valid, X, tag):
            from oracle import oracle as O
            fails = []
            for qi in range(4):
                ids, sc = idx.search(hc.Q[qi], k)
                r_ids, r_sc = hc.c.scan(X, hc.Q[qi], k, 0, valid)      # the oracle over its own rows and validity bytes
                try:
                    O.check_parity(ids, sc, r_ids, r_sc)
                except AssertionError as e:
                    fails.append(f"{tag} query {qi}: {e}")
            return fails

        fails = []
        map_buf = np.full(rows, 7, dtype=np.uint64)
        ROUNDS = 3                         # the index is rebuilt and the same rows removed each round; phases report the median
        t_tomb = {"one_file_50_chunks": [], "10pct_of_rows": []}
        t_compact = []
        l0 = idx.launch_count
        for rnd in range(ROUNDS):
            if rnd:
                idx.compact_keep(np.zeros(len(idx), np.uint8))           # empty the index (every row dropped), refill
            idx.append_synthetic(seed=1, row0=0, n=rows, normalize=True)
            valid = np.ones(rows, np.uint8)
            # (1) a "file" of 50 consecutive chunks removed (the reference deletes by file_path), then 10 % of all rows
            for name, dead in (("one_file_50_chunks", file_rows), ("10pct_of_rows", scattered)):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                idx.tombstone(dead)
                t_tomb[name].append(time.perf_counter() - t0)
                valid[dead.astype(np.int64)] = 0
                if rnd == 0:
                    fails += check(valid, hc.X, "after tombstone " + name)
            # (2) compaction: plan on the device (prefix sums over the keep flags), ordered gather of the surviving rows
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            new_of_old = idx.compact(out=map_buf)        # the caller's own, already-touched buffer (as a Rust Vec would be)
            t_compact.append(time.perf_counter() - t0)
        for name, dead in (("one_file_50_chunks", file_rows), ("10pct_of_rows", scattered)):
            dt = float(np.median(t_tomb[name]))
            rec["tombstone_" + name] = {"rows": int(len(dead)), "seconds": dt, "rows_per_s": len(dead) / dt, "all_seconds": t_tomb[name],
                                        "note": "synchronous call: H2D of the row ids, NaN-poison rows (+ K3 planes when built), validity bytes"}
        dt = float(np.median(t_compact))
        live = int(valid.sum())
        moved = live - int(np.argmin(valid))                      # rows behind the first dropped row all move
        assert len(idx) == live
        keep = valid.astype(bool)
        Xc = hc.X[keep]
        ok_map = bool(np.array_equal(new_of_old[keep], np.arange(live, dtype=np.uint64)) and np.all(new_of_old[~keep] == np.uint64(2**64 - 1)))
        rec["compact"] = {"rows_before": rows, "rows_after": live, "rows_moved": moved, "seconds": dt, "all_seconds": t_compact,
                          "algorithmic_bytes": 2 * moved * dim * 4,
                          "roofline": {"bound": "hbm", "achieved": 2 * moved * dim * 4 / dt / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": 2 * moved * dim * 4 / dt / 1e9 / peak, "traffic": None,
                                       "note": "whole synchronous call — scratch allocation, plan kernels (prefix sums over the keep flags), "
                                               f"the {rows * 8 / 1e6:.0f} MB old->new map copied to pageable host memory (what the caller's chunk table "
                                               "needs; the largest phase of the call), gather kernels — against 1 read + 1 write of every moved row"},
                          "old_to_new_map_correct": ok_map}
        fails += check(None, Xc, "after compaction")
        fails += [] if ok_map else ["compaction: old -> new row map is wrong"]
        launches = (idx.launch_count - l0) // ROUNDS
    # (3) disk cache: save / load round trip of a smaller index (page-cache speed on this box, stated as such)
    srows = min(a.maint_save_rows, rows)
    with tempfile.TemporaryDirectory() as td, sema_b200.GpuIndex(dim, srows, device=0) as idx:
        idx.append_synthetic(seed=1, row0=0, n=srows, normalize=True)
        path = os.path.join(td, "vectors.semaidx")
        t0 = time.perf_counter()
        idx.save(path)
        ts = time.perf_counter() - t0
        size = os.path.getsize(path)
        t0 = time.perf_counter()
        idx2 = sema_b200.GpuIndex.load(path, device=0)
        tl = time.perf_counter() - t0
        try:
            same = True
            for qi in range(4):
                a1, b1 = idx.search(hc.Q[qi], k)
                a2, b2 = idx2.search(hc.Q[qi], k)
                same &= bool(np.array_equal(a1, a2) and np.array_equal(b1, b2))
                f = hc.check(qi, k, a2, b2, prefix=srows)
                fails += [f] if f else []
        finally:
            idx2.close()
        rec["disk_cache"] = {"rows": srows, "file_bytes": size, "save_seconds": ts, "save_GBps": size / ts / 1e9,
                             "load_seconds": tl, "load_GBps": size / tl / 1e9, "loaded_index_answers_bit_identically": same,
                             "note": "D2H + write / read + H2D through a temporary directory (page cache, not a disk benchmark)"}
        fails += [] if same else ["disk cache: loaded index answers differently"]
    line = {"metric": f"compaction_rows_per_s_{rows}x{dim}_fp32_10pct_dropped", "value": rec["compact"]["rows_moved"] / rec["compact"]["seconds"],
            "unit": "rows/s", "n_gpus": 1, "steps": 1, "warmup": 0, "ms_per_step": rec["compact"]["seconds"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{rows}x{dim} fp32 index: tombstone one file's 50 chunks, tombstone 10 % of the rows, compact, "
                                   f"then save / load a {srows}-row index (SURVEY.md 8(f) rows 1-2)", "k": k},
            "roofline": rec["compact"]["roofline"], "phases": rec, "gpu_launches": int(launches),
            "oracle_checked_queries": 4 * 4 + 4, "oracle_failures": fails, "verified": not fails}
    print(json.dumps(line), flush=True)


def run_ingest(a):
    """BASELINE.json configs[4] alone (see ingest_measure)."""
    import torch
    need_gpu()
    torch.cuda.set_device(0)
    rec = ingest_measure(a.dim, a.k, a.rows, a.ingest_batch, 0, 1_000_000)
    nb = rec["rows"] // a.ingest_batch
    line = {"metric": rec["metric"], "value": rec["value"], "unit": "rows/s", "n_gpus": 1, "steps": nb, "warmup": 1,
            "ms_per_step": rec["seconds"] * 1e3 / nb, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": rec["workload"], "rows": rec["rows"], "dim": a.dim, "k": a.k},
            "ingest": rec,
            "e2e": {"value": rec["value"], "unit": "rows/s", "h2d_bytes_per_step": a.ingest_batch * a.dim * 4, "d2h_bytes_per_step": 0,
                    "note": "plus one query in the kernel parameters and one result block per search"},
            "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"], "verified": rec["verified"]}
    print(json.dumps(line), flush=True)


def run_config1(a):
    print(json.dumps(config1_measure(a.k)), flush=True)


def config1_measure(k):
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
    with StorageManager(dim=corpus.DIM, capacity_rows=len(chunks) + 64, normalize=True,
                        embedder=lambda t: qvec.get(t)) as mgr:
        t0 = time.perf_counter()
        mgr.index_chunks(chunks, vectors=emb)
        t_index = time.perf_counter() - t0
        for q in queries[:10]:
            mgr.search(q, k)
        lat, all_hits = [], []
        for q in queries:
            t0 = time.perf_counter()
            hits = mgr.search(q, k)
            lat.append(time.perf_counter() - t0)
            all_hits.append(hits)
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
    from oracle import oracle as O
    mismatches = []
    for qi, hits_q in enumerate(all_hits):                      # every query: (chunk id, score) list vs the oracle's scan
        o_ids, o_sc = c_oracle.scan(X, Qn[qi], k)
        row_of = {c.id: r for r, c in enumerate(chunks)}
        try:
            O.check_parity(np.array([row_of[c.id] for c, _ in hits_q], dtype=np.uint64),
                           np.array([s_ for _, s_ in hits_q], dtype=np.float32), o_ids, o_sc)
        except AssertionError as e:
            mismatches.append(f"query {qi}: {str(e)[:120]}")
    ids_ok = not mismatches
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
        "oracle_checked_queries": len(all_hits), "oracle_failures": mismatches[:5], "verified": bool(ids_ok),
    }
    return line


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
    elif a.workload == "maintenance":
        run_maintenance(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()

"""CPU tests of the host-side mirror (csrc/host/storage.cpp) against the Python restatement of
the reference's caller-side code (oracle/engine.py): group_results_by_file
(src/tui/engine.rs:156-182) and the LIKE fallback predicate (src/storage/lance_indexer.rs:143-147)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import engine as E
from sema_b200 import _lib

pytestmark = pytest.mark.skipif(not os.path.exists(_lib.SO_PATH), reason="libsema_b200.so not built")


def test_store_header_and_binding_agree():
    from sema_b200 import storage
    src = open(os.path.join(os.path.dirname(_lib.HEADER_PATH), "sema_store.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = set(re.findall(r"\b(sema_[a-z0-9_]+)\s*\(", src)) - {"sema_embed_fn"}
    assert declared == set(storage.STORE_SIGNATURES)
    L = C.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name)


def _random_hits(rng, n, n_files, ties):
    paths = [f"dir/f{int(i)}.md" for i in rng.integers(0, n_files, n)]
    starts = rng.integers(1, 400, n).astype(np.uint64)
    scores = np.sort(rng.random(n).astype(np.float32))[::-1].copy()       # ranked: best first
    if ties:
        scores = np.round(scores, 1)
    return paths, starts, scores


@pytest.mark.parametrize("n,n_files,ties", [(0, 1, False), (1, 1, False), (50, 7, False), (50, 50, True), (50, 3, True),
                                            (200, 20, False)])
def test_group_results_by_file_matches_reference_logic(n, n_files, ties):
    from sema_b200.storage import group_results_by_file
    rng = np.random.default_rng(n * 31 + n_files)
    paths, starts, scores = _random_hits(rng, n, n_files, ties)
    got = group_results_by_file(paths, starts, scores)
    want = E.group_results_by_file([{"file_path": p, "start_line": int(s), "score": float(c), "i": i}
                                    for i, (p, s, c) in enumerate(zip(paths, starts, scores))])
    assert [(w["i"], w["total_matches_in_file"]) for w in want] == got
    assert sum(t for _, t in got) == n
    sc = [scores[i] for i, _ in got]
    assert all(a >= b for a, b in zip(sc, sc[1:]))                       # engine.rs:176-180


def test_group_results_keeps_lowest_start_line_and_counts():
    from sema_b200.storage import group_results_by_file
    paths = ["a", "b", "a", "a", "b"]
    starts = [30, 5, 10, 20, 1]
    scores = [0.9, 0.8, 0.7, 0.6, 0.5]
    # file a: representative = start_line 10 (score 0.7), 3 matches; file b: start_line 1 (0.5), 2 matches
    assert group_results_by_file(paths, starts, scores) == [(2, 3), (4, 2)]


def test_group_results_nan_scores_compare_equal():
    from sema_b200.storage import group_results_by_file
    got = group_results_by_file(["a", "b", "c"], [1, 1, 1], [float("nan"), 0.5, 0.7])
    want = E.group_results_by_file([{"file_path": p, "start_line": 1, "score": s, "i": i}
                                    for i, (p, s) in enumerate(zip("abc", [float("nan"), 0.5, 0.7]))])
    assert [g[0] for g in got] == [w["i"] for w in want]


@pytest.mark.parametrize("content,needle", [
    ("hello world", "lo w"), ("hello world", "xyz"), ("hello", ""), ("", ""), ("", "a"), ("100% sure", "0% s"),
    ("a_b", "a_b"), ("axb", "a_b"), ("ab", "a_b"), ("abcabc", "c%b"), ("abcabc", "c%a%c"), ("abc", "c%a"),
    ("Hello", "hello"), ("multi\nline text", "i\nl"), ("über straße", "r str"), ("aaa", "aaaa"), ("it's", "t's"),
])
def test_like_contains_matches_sql_like(content, needle):
    from sema_b200.storage import like_contains
    assert like_contains(content, needle) == E.like_contains(content, needle)


def test_chunker_restatement_properties():
    # src/storage/processor.rs:6-8, 31-85 — used only to build the config-1 corpus
    text = "\n".join(f"line {i} " + "x" * (i % 60) for i in range(400))
    cs = E.create_chunks("p.md", text)
    assert cs[0]["id"] == "p.md:0" and cs[0]["start_line"] == 1
    assert all(len(c["content"].encode()) <= E.CHUNK_SIZE for c in cs)
    assert all(c["content"].endswith("\n") or text.endswith(c["content"]) for c in cs)   # snapped to a newline
    assert all(c["end_line"] == c["start_line"] + c["content"].count("\n") for c in cs)
    assert [c["id"] for c in cs] == [f"p.md:{i}" for i in range(len(cs))]
    assert E.create_chunks("p.md", "short") == []                         # < MIN_CHUNK_SIZE
    for a, b in zip(cs, cs[1:]):                                          # overlapping windows
        assert b["start_line"] <= a["end_line"]


def test_bench_reference_arm_prints_one_contract_line():
    """bench.py --impl reference (the CPU search path on the host cores) runs without a GPU and prints one
    JSON line with the contract's keys; under torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0",
           "--rows", "20000"]
    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm must not become single-threaded because of it
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    # the whole corpus named in `config` is scanned, nothing is extrapolated, the step time is a measured one
    assert d["cpu_baseline"]["extrapolated"] is False and d["config"]["rows"] == 20000 and d["steps"] == 2
    assert abs(d["ms_per_step"] - 1e3 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["cpu_baseline"]["ms_per_scan_min"] <= d["ms_per_step"] <= d["cpu_baseline"]["ms_per_scan_max"] * 1.5
    sys.path.insert(0, root)
    import argparse
    import bench
    a = argparse.Namespace(rows=20000, dim=384, k=10, queries=64, gpus=1)
    assert d["config"] == bench.shared_config(a)                 # both arms state the workload identically
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    quiet = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=root, env=env)
    assert quiet.returncode == 0 and not [l for l in quiet.stdout.splitlines() if l.startswith("{")]

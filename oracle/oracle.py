"""oracle/oracle.py — NumPy restatement of Sema's vector-search hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sema_b200/`` may import this module;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs
do, as the checker.

PARITY UNPINNED: the reference (akshitsinha/sema) holds no tests, golden vectors
or fixtures for this path, and neither it nor its engine (lancedb 0.23.1 ->
lance 1.0.1 -> lance-index / lance-linalg 1.0.1, ``Cargo.lock:4258-4259,
3755-3756, 4028-4029, 4133-4134``; not vendored) can be built or imported here.
The functions below restate the behaviour at the reference's call sites:

* :func:`normalize`  — ``src/semantic/embeddings.rs:83-88``
* :func:`mean_pool`  — ``src/semantic/embeddings.rs:61-91``
* :func:`scan`       — ``src/storage/lance_indexer.rs:121-126`` (LanceDB flat
  exact KNN, default metric squared L2, ascending ``_distance``, ``limit`` rows,
  null vectors skipped) and its dot/cosine twin on unit rows
* :func:`merge`      — global top-k of per-shard top-k
* :func:`synth`      — the deterministic synthetic corpus of SURVEY.md §8(d)

The arithmetic mirrors ``oracle/cpu_scan.c`` operation for operation (16 partial
sums folded pairwise, no FMA contraction), so the two agree bit for bit; the
float64 "shadow" functions are the independent arithmetic used for tolerance
accounting.
"""
from __future__ import annotations

import numpy as np

METRIC_DOT = 0  # score = q.x, descending
METRIC_L2 = 1   # _distance = sum((q-x)^2), ascending — LanceDB default

_LANES = 16
_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def normalize(rows: np.ndarray) -> np.ndarray:
    """L2-normalise each row — ``src/semantic/embeddings.rs:83-88``.

    ``norm = sqrt(sum_j x_j*x_j)`` with a *sequential* f32 sum; every element is
    divided by ``norm`` iff ``norm > 0`` (a zero row stays zero).
    """
    x = np.array(rows, dtype=np.float32, copy=True, order="C")
    if x.ndim == 1:
        return normalize(x[None, :])[0]
    n, d = x.shape
    acc = np.zeros(n, dtype=np.float32)
    for j in range(d):  # sequential over j, like iter().map(|x| x*x).sum()
        acc = acc + x[:, j] * x[:, j]
    norm = np.sqrt(acc, dtype=np.float32)
    nz = norm > 0
    x[nz] = x[nz] / norm[nz, None]
    return x


def mean_pool(tokens: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """``mean_pool`` — ``src/semantic/embeddings.rs:61-91`` for n texts at once.

    tokens [n, seq, hidden] f32, mask [n, seq] f32.  ``pooled[j] += tokens[i][j] * mask[i]`` with
    i ascending (sequential f32, multiply then add), ``mask_sum`` likewise; divide by ``mask_sum``
    iff it is > 0; then :func:`normalize`.
    """
    tokens = np.asarray(tokens, dtype=np.float32)
    mask = np.asarray(mask, dtype=np.float32)
    n, seq, hidden = tokens.shape
    pooled = np.zeros((n, hidden), dtype=np.float32)
    mask_sum = np.zeros(n, dtype=np.float32)
    for i in range(seq):
        mask_sum = mask_sum + mask[:, i]
        pooled = pooled + tokens[:, i, :] * mask[:, i, None]
    nz = mask_sum > 0
    pooled[nz] = pooled[nz] / mask_sum[nz, None]
    return normalize(pooled)


def synth(seed: int, row0: int, n: int, d: int) -> np.ndarray:
    """Deterministic synthetic rows (un-normalised), identical to
    ``sema_oracle_synth`` (cpu_scan.c) and to the device generator."""
    with np.errstate(over="ignore"):
        rows = (np.arange(n, dtype=np.uint64) + np.uint64(row0))[:, None]
        cols = np.arange(d, dtype=np.uint64)[None, :]
        x = np.uint64(seed) * _M1 + rows * _M2 + cols * _M3 + np.uint64(1)
        x ^= x >> np.uint64(30)
        x *= _M2
        x ^= x >> np.uint64(27)
        x *= _M3
        x ^= x >> np.uint64(31)
        m = np.uint64(0xFFFF)
        s = (x & m) + ((x >> np.uint64(16)) & m) + ((x >> np.uint64(32)) & m) + (x >> np.uint64(48))
    return (s.astype(np.int64) - 131070).astype(np.float32)


def _fold16(p: np.ndarray) -> np.ndarray:
    a = p[:, :8] + p[:, 8:]
    b = a[:, :4] + a[:, 4:]
    return (b[:, 0] + b[:, 2]) + (b[:, 1] + b[:, 3])


def row_keys(X: np.ndarray, q: np.ndarray, metric: int = METRIC_DOT) -> np.ndarray:
    """f32 ranking key per row (larger is better): the dot product, or minus the
    squared L2 distance.  Same operation order as cpu_scan.c row_dot/row_l2sq."""
    X = np.asarray(X, dtype=np.float32)
    q = np.asarray(q, dtype=np.float32)
    n, d = X.shape
    p = np.zeros((n, _LANES), dtype=np.float32)
    j = 0
    while j + _LANES <= d:
        if metric == METRIC_L2:
            t = q[None, j:j + _LANES] - X[:, j:j + _LANES]
            p = p + t * t
        else:
            p = p + X[:, j:j + _LANES] * q[None, j:j + _LANES]
        j += _LANES
    s = _fold16(p)
    while j < d:
        if metric == METRIC_L2:
            t = q[j] - X[:, j]
            s = s + t * t
        else:
            s = s + X[:, j] * q[j]
        j += 1
    s = s.astype(np.float32)
    return (-s if metric == METRIC_L2 else s) + np.float32(0.0)


def row_keys_f64(X: np.ndarray, q: np.ndarray, metric: int = METRIC_DOT) -> np.ndarray:
    """float64 shadow of :func:`row_keys` (independent arithmetic)."""
    X = np.asarray(X, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    if metric == METRIC_L2:
        return -((q[None, :] - X) ** 2).sum(axis=1)
    return X @ q


def _rank(keys: np.ndarray, ids: np.ndarray, k: int):
    ok = ~np.isnan(keys)
    keys, ids = keys[ok], ids[ok]
    order = np.lexsort((ids, -keys))[:k]  # best key first, ties by lower id
    return ids[order], keys[order]


def scan(X, q, k: int, metric: int = METRIC_DOT, valid=None, id_base: int = 0,
         chunk: int = 1 << 16, f64: bool = False):
    """Flat exact k-NN — ``src/storage/lance_indexer.rs:121-126``.

    Returns ``(ids uint64[n_found], scores float32[n_found])`` best first.  For
    ``METRIC_L2`` the returned score is the squared distance (LanceDB's
    ``_distance``), ascending.  ``valid`` (one byte per row, 0 = null vector) rows
    are skipped (``lance_indexer.rs:41-45, 66-70``); ``k > N`` returns every valid
    row; an empty table returns nothing (``lance_indexer.rs:108-111``).
    """
    X = np.asarray(X)
    n = X.shape[0]
    best_ids = np.zeros(0, dtype=np.uint64)
    best_keys = np.zeros(0, dtype=np.float64 if f64 else np.float32)
    if k == 0 or n == 0:
        return best_ids, best_keys.astype(np.float32)
    fn = row_keys_f64 if f64 else row_keys
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        keys = fn(X[s:e], q, metric)
        ids = np.arange(s, e, dtype=np.uint64)
        if valid is not None:
            m = np.asarray(valid[s:e]).astype(bool)
            keys, ids = keys[m], ids[m]
        ids, keys = _rank(np.concatenate([best_keys, keys]), np.concatenate([best_ids, ids]), k)
        best_ids, best_keys = ids, keys
    scores = -best_keys + 0.0 if metric == METRIC_L2 else best_keys
    return best_ids + np.uint64(id_base), scores.astype(np.float64 if f64 else np.float32)


def merge(scores, ids, lens, k: int, metric: int = METRIC_DOT):
    """Global top-k of G ranked per-shard lists (scores[G,k], ids[G,k], lens[G])."""
    scores = np.asarray(scores, dtype=np.float32)
    ids = np.asarray(ids, dtype=np.uint64)
    ks, iis = [], []
    for g, ln in enumerate(lens):
        ks.append(scores[g, :ln])
        iis.append(ids[g, :ln])
    keys = np.concatenate(ks) if ks else np.zeros(0, np.float32)
    iid = np.concatenate(iis) if iis else np.zeros(0, np.uint64)
    if metric == METRIC_L2:
        keys = -keys
    out_ids, out_keys = _rank(keys, iid, k)
    return out_ids, ((-out_keys + np.float32(0.0)) if metric == METRIC_L2 else out_keys).astype(np.float32)


def l2sq_from_cosine(score):
    """On unit rows ``L2^2 = 2 - 2 cos`` — the bridge between LanceDB's default
    metric and the cosine score returned at the search boundary."""
    return 2.0 - 2.0 * np.asarray(score, dtype=np.float64)


def check_parity(ids, scores, ref_ids, ref_scores, tie_tol: float = 1e-5, rel_tol: float = 1e-5,
                 tail_scores=None):
    """BASELINE.json acceptance rule: identical top-k id set and order except for
    ties within ``tie_tol``; scores within ``rel_tol`` relative.

    ``ids``/``scores`` are the implementation under test, ``ref_*`` the oracle's.
    Position i may hold a different id only if the two scores involved are within
    ``tie_tol`` of each other.  Raises AssertionError with a description.
    """
    ids = np.asarray(ids, dtype=np.uint64)
    ref_ids = np.asarray(ref_ids, dtype=np.uint64)
    scores = np.asarray(scores, dtype=np.float64)
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    assert ids.shape == ref_ids.shape, f"n_found differs: {ids.shape} vs {ref_ids.shape}"
    denom = np.maximum(np.abs(ref_scores), 1e-30)
    rel = np.abs(scores - ref_scores) / denom
    bad = rel > rel_tol
    # an absolute floor for scores that are ~0 (relative error is meaningless there)
    bad &= np.abs(scores - ref_scores) > 1e-7
    assert not bad.any(), f"score mismatch at {np.nonzero(bad)[0][:5]}: {scores[bad][:5]} vs {ref_scores[bad][:5]}"
    diff = np.nonzero(ids != ref_ids)[0]
    for i in diff:
        # the id at position i must appear somewhere in the oracle list (or be a tie with the
        # oracle's boundary element) with a score within the tie tolerance
        j = np.nonzero(ref_ids == ids[i])[0]
        if len(j):
            assert abs(ref_scores[j[0]] - ref_scores[i]) <= tie_tol, (
                f"order differs at {i} beyond tie tolerance: id {ids[i]} (ref pos {j[0]})")
        else:
            assert abs(scores[i] - ref_scores[-1]) <= tie_tol, (
                f"id {ids[i]} at {i} is not in the oracle top-k and is not a boundary tie")

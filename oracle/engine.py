"""oracle/engine.py — Python restatement of the caller-side post-processing on the path.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  PARITY UNPINNED: the reference has no
tests for these functions either; they are restated line by line from the Rust source.

* :func:`group_results_by_file` — ``src/tui/engine.rs:156-182``
* :func:`like_contains`         — the ``content LIKE '%q%'`` fallback predicate of
  ``src/storage/lance_indexer.rs:143-147`` (DataFusion LIKE: case sensitive, ``%`` any run,
  ``_`` any single character)
* :func:`create_chunks`         — ``src/storage/processor.rs:31-85`` (only to build the
  config-1 synthetic corpus; chunking is outside the hot path)
"""
from __future__ import annotations

import re

CHUNK_SIZE, OVERLAP_SIZE, MIN_CHUNK_SIZE = 1000, 100, 50  # src/storage/processor.rs:6-8


def group_results_by_file(results):
    """results: list of dicts {file_path, start_line, score, ...} in rank order.
    Returns the grouped list (best score first), each with ``total_matches_in_file``.
    The reference iterates a HashMap (order unspecified); groups are taken here in
    first-appearance order, which only matters between groups of equal score."""
    groups = {}
    for r in results:                                     # engine.rs:159-164
        groups.setdefault(r["file_path"], []).append(r)
    grouped = []
    for g in groups.values():                             # engine.rs:167-174
        g = sorted(g, key=lambda r: r["start_line"])      # stable, like sort_by_key
        first = dict(g[0])
        first["total_matches_in_file"] = len(g)
        grouped.append(first)
    # engine.rs:176-180: stable sort, best score first, incomparable (NaN) = equal
    import functools

    def cmp(a, b):
        if b["score"] > a["score"]:
            return 1
        if b["score"] < a["score"]:
            return -1
        return 0
    return sorted(grouped, key=functools.cmp_to_key(cmp))


def like_contains(content: str, needle: str) -> bool:
    pat = "".join(".*" if ch == "%" else "." if ch == "_" else re.escape(ch) for ch in "%" + needle + "%")
    return re.fullmatch(pat, content, flags=re.S) is not None


def create_chunks(file_path: str, content: str):
    """``FileProcessor::create_chunks`` (src/storage/processor.rs:31-85) on UTF-8 bytes."""
    data = content.encode("utf-8")
    chunks = []
    if len(data) < MIN_CHUNK_SIZE:
        return chunks

    def is_boundary(i):
        return i == len(data) or (data[i] & 0xC0) != 0x80

    start, chunk_id = 0, 0
    while start < len(data):
        end = min(start + CHUNK_SIZE, len(data))
        safe_end = end
        while safe_end > start and not is_boundary(safe_end):
            safe_end -= 1
        if safe_end < len(data):
            nl = data.rfind(b"\n", start, safe_end)
            if nl != -1:
                safe_end = nl + 1
        piece = data[start:safe_end]
        if len(piece) >= MIN_CHUNK_SIZE or chunk_id == 0:
            start_line = data[:start].count(b"\n") + 1
            end_line = start_line + piece.count(b"\n")
            chunks.append({"id": f"{file_path}:{chunk_id}", "file_path": file_path, "start_line": start_line,
                           "end_line": end_line, "content": piece.decode("utf-8")})
            chunk_id += 1
        next_start = max(safe_end - OVERLAP_SIZE, 0)
        start = safe_end if next_start <= start else next_start
        if start >= len(data):
            break
    return chunks

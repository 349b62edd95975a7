"""oracle/corpus.py — the config-1 input (BASELINE.json configs[0]): a small synthetic markdown
corpus, chunked like the reference chunks files, embedded with a deterministic STAND-IN embedder.

TEST INFRASTRUCTURE ONLY.  The reference embeds with all-MiniLM-L6-v2 through ONNX Runtime
(src/semantic/embeddings.rs:14-58); neither the model nor the runtime is available offline, so
`embed` below is a seeded hashed-token random projection to 384-d — it has the right shape and
the normalise tail is still applied by the code under test, but it is NOT the reference's model
and says nothing about retrieval quality.
"""
from __future__ import annotations

import zlib

import numpy as np

from . import oracle as O
from .engine import create_chunks

DIM = 384  # src/storage/lance_indexer.rs:43, 76

_WORDS = ("vector index search query embedding chunk file storage engine table row column scan kernel "
          "memory bandwidth latency throughput cache tensor shard merge score cosine distance normal "
          "markdown heading paragraph list code build test error config path line crawl hash token model "
          "batch stream ingest delete update result rank limit filter group sort terminal interface").split()


def make_markdown_tree(n_files: int, seed: int = 0):
    """-> {path: text}; ~16 chunks of <= 1000 bytes per file."""
    rng = np.random.default_rng(seed)
    files = {}
    for f in range(n_files):
        lines = [f"# Document {f}", ""]
        for s in range(12):
            lines.append(f"## Section {s} of document {f}")
            for _ in range(6):
                k = int(rng.integers(8, 20))
                lines.append(" ".join(_WORDS[int(i)] for i in rng.integers(0, len(_WORDS), k)) + ".")
            lines.append("")
        files[f"docs/dir{f % 7}/file{f}.md"] = "\n".join(lines)
    return files


def chunk_tree(files):
    chunks = []
    for path, text in files.items():
        chunks.extend(create_chunks(path, text))
    return chunks


_token_cache: dict[str, np.ndarray] = {}


def _token_vec(tok: str) -> np.ndarray:
    v = _token_cache.get(tok)
    if v is None:
        v = O.synth(7, zlib.crc32(tok.encode("utf-8")), 1, DIM)[0] / np.float32(65536.0)
        _token_cache[tok] = v
    return v


def embed(text: str):
    """Stand-in for VectorStore::generate_embedding: un-normalised 384-d vector, or None for
    text without a single token (an "embedding failure")."""
    toks = [t for t in "".join(ch.lower() if ch.isalnum() else " " for ch in text).split() if t]
    if not toks:
        return None
    acc = np.zeros(DIM, dtype=np.float32)
    for t in toks:
        acc = acc + _token_vec(t)
    return acc

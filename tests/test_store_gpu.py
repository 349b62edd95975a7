"""GPU tests of the drop-in boundary: StorageManager / index_chunks / search / execute_search /
remove_file_chunks through include/sema_store.h, on BASELINE.json configs[0] — a ~10k-chunk
synthetic markdown corpus chunked like the reference (src/storage/processor.rs) and embedded
with a STAND-IN embedder (oracle/corpus.py: the reference's MiniLM model is not available
offline).  The oracle recomputes every ranking on the CPU."""
import numpy as np
import pytest

from oracle import corpus, engine as E, oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    from sema_b200.storage import Chunk, StorageManager
    files = corpus.make_markdown_tree(1050, seed=3)
    raw_chunks = corpus.chunk_tree(files)
    assert 9000 <= len(raw_chunks) <= 13000                   # "~10k chunks"
    chunks = [Chunk(c["id"], c["file_path"], c["start_line"], c["end_line"], c["content"]) for c in raw_chunks]
    emb = np.stack([corpus.embed(c.content) for c in chunks])
    mgr = StorageManager(dim=corpus.DIM, capacity_rows=len(chunks) + 64, normalize=True, embedder=corpus.embed)
    yield mgr, chunks, emb
    mgr.close()


def test_empty_store_returns_no_results():
    # missing table => Ok(vec![]) (src/storage/lance_indexer.rs:108-111)
    from sema_b200.storage import StorageManager
    with StorageManager(dim=corpus.DIM, capacity_rows=16, embedder=corpus.embed) as m:
        assert m.search("vector index", 10) == []
        assert m.execute_search("vector index") == []
        m.index_chunks([])                                     # empty slice is Ok (:31-33)
        assert len(m) == 0


def test_config1_index_and_search_top10(world):
    mgr, chunks, emb = world
    # index in two batches: one through the embedder callback, one with precomputed vectors
    half = len(chunks) // 2
    mgr.index_chunks(chunks[:half])
    mgr.index_chunks(chunks[half:], vectors=emb[half:])
    assert len(mgr) == len(chunks)
    X = O.normalize(emb)                                       # reference normalise tail on the CPU
    rng = np.random.default_rng(5)
    for _ in range(100):                                       # 100 seeded queries, top-10
        words = " ".join(corpus._WORDS[int(i)] for i in rng.integers(0, len(corpus._WORDS), 6))
        got = mgr.search(words, 10)
        q = O.normalize(corpus.embed(words))
        r_ids, r_sc = O.scan(X, q, 10)
        O.check_parity(np.array([_row_of(chunks, c) for c, _ in got], dtype=np.uint64),
                       np.array([s for _, s in got], dtype=np.float32), r_ids, r_sc)
        # the boundary type: (Chunk, f32) with the Chunk fields intact
        c0, s0 = got[0]
        assert c0 == chunks[int(r_ids[0])] and abs(s0 - r_sc[0]) <= 1e-5 * abs(r_sc[0])


_ROW_CACHE = {}


def _row_of(chunks, c):
    if not _ROW_CACHE:
        _ROW_CACHE.update({ch.id: i for i, ch in enumerate(chunks)})
    return _ROW_CACHE[c.id]


def test_search_limit_50_and_execute_search_grouping(world):
    mgr, chunks, emb = world
    X = O.normalize(emb)
    query = "tensor shard merge bandwidth"
    hits = mgr.search(query, 50)                               # SEARCH_RESULTS_LIMIT, src/tui/engine.rs:11
    q = O.normalize(corpus.embed(query))
    r_ids, r_sc = O.scan(X, q, 50)
    O.check_parity(np.array([_row_of(chunks, c) for c, _ in hits], dtype=np.uint64),
                   np.array([s for _, s in hits], dtype=np.float32), r_ids, r_sc)
    grouped = mgr.execute_search(query)
    want = E.group_results_by_file([{"file_path": chunks[int(i)].file_path, "start_line": chunks[int(i)].start_line,
                                     "score": float(s), "row": int(i)} for i, s in zip(r_ids, r_sc)])
    assert [g.chunk.id for g in grouped] == [chunks[w["row"]].id for w in want]
    assert [g.total_matches_in_file for g in grouped] == [w["total_matches_in_file"] for w in want]
    assert sum(g.total_matches_in_file for g in grouped) == 50
    sc = [g.score for g in grouped]
    assert all(a >= b for a, b in zip(sc, sc[1:]))             # real scores make the sort meaningful


def test_query_is_trimmed_and_keyword_route_is_out_of_scope(world):
    from sema_b200 import SemaError
    mgr, chunks, emb = world
    a = mgr.search("  kernel memory  \n", 5)
    b = mgr.search("kernel memory", 5)
    assert [c.id for c, _ in a] == [c.id for c, _ in b]
    assert mgr.search("'", 5) == []                            # bare "'" => Ok(empty), src/storage/mod.rs:115-120
    with pytest.raises(SemaError) as e:
        mgr.search("'keyword", 5)
    assert e.value.code == -5                                  # SEMA_ERR_UNSUPPORTED: Tantivy path not built


def test_failed_query_embedding_falls_back_to_like_scan(world):
    mgr, chunks, emb = world
    # the stand-in embedder fails on text without alphanumerics -> `content LIKE '%q%'`
    # (src/storage/lance_indexer.rs:143-162): first `limit` matching rows in table order, score 1.0
    hits = mgr.search("## ", 7)
    want = [c for c in chunks if "## " in c.content][:7]
    assert [c.id for c, _ in hits] == [c.id for c in want]
    assert all(s == 1.0 for _, s in hits)


def test_null_vectors_from_failed_chunk_embeddings():
    from sema_b200.storage import Chunk, StorageManager
    with StorageManager(dim=corpus.DIM, capacity_rows=64, embedder=corpus.embed) as m:
        cs = [Chunk(f"f.md:{i}", "f.md", i + 1, i + 1, t) for i, t in
              enumerate(["vector index search", "--- ***", "kernel memory bandwidth", "???"])]
        m.index_chunks(cs)                                     # chunks 1 and 3 fail to embed -> null vectors
        got = m.search("vector kernel", 10)
        assert sorted(c.id for c, _ in got) == ["f.md:0", "f.md:2"]


def test_remove_file_chunks(world):
    mgr, chunks, emb = world
    query = "storage engine table row column"
    before = mgr.search(query, 10)
    victim = before[0][0].file_path
    n_file = sum(1 for c in chunks if c.file_path == victim)
    assert mgr.remove_file_chunks(victim) == n_file            # src/storage/lance_indexer.rs:234-250
    assert mgr.remove_file_chunks(victim) == 0
    after = mgr.search(query, 10)
    assert all(c.file_path != victim for c, _ in after)
    X = O.normalize(emb)
    valid = np.array([c.file_path != victim for c in chunks], dtype=np.uint8)
    r_ids, r_sc = O.scan(X, O.normalize(corpus.embed(query)), 10, valid=valid)
    O.check_parity(np.array([_row_of(chunks, c) for c, _ in after], dtype=np.uint64),
                   np.array([s for _, s in after], dtype=np.float32), r_ids, r_sc)
    assert all(c.file_path != victim for c, _ in mgr.search("## ", 1000)[:50])   # LIKE scan skips deleted rows too


def test_store_compaction_keeps_null_chunks_and_drops_removed_files():
    from sema_b200.storage import Chunk, StorageManager
    with StorageManager(dim=corpus.DIM, capacity_rows=64, embedder=corpus.embed) as m:
        cs = [Chunk(f"f{i % 3}.md:{i}", f"f{i % 3}.md", i + 1, i + 1, t) for i, t in enumerate(
            ["vector index search", "--- ***", "kernel memory bandwidth", "tensor shard merge", "cache latency",
             "???", "storage engine table"])]
        m.index_chunks(cs)
        assert m.remove_file_chunks("f1.md") == 2              # chunks 1 and 4
        assert m.compact() == 5
        assert len(m) == 5
        ids = [m.chunk(r).id for r in range(5)]
        assert ids == ["f0.md:0", "f2.md:2", "f0.md:3", "f2.md:5", "f0.md:6"]
        got = m.search("tensor shard", 10)
        assert got[0][0].id == "f0.md:3"
        assert sorted(c.id for c, _ in got) == ["f0.md:0", "f0.md:3", "f0.md:6", "f2.md:2"]   # the null chunk never matches
        assert [c.id for c, _ in m.search("???", 10)] == ["f2.md:5"]                           # ... but LIKE still finds it


def test_store_without_normalisation_ranks_by_l2_like_the_reference(oracle_c):
    """normalize=False stores the caller's vectors as given.  The reference ranks by LanceDB's default
    squared-L2 `_distance` (src/storage/lance_indexer.rs:121-126), which differs from the dot-product
    order as soon as the norms differ: the store must follow the L2 order (and hand out 1 - d/2)."""
    from sema_b200.storage import Chunk, StorageManager
    rng = np.random.default_rng(11)
    n, d, k = 4000, corpus.DIM, 10
    X = rng.standard_normal((n, d)).astype(np.float32)
    X *= rng.uniform(0.2, 5.0, (n, 1)).astype(np.float32)                  # norms all over the place
    X[7] = 0.0                                                              # a zero-padded / zero vector
    chunks = [Chunk(f"f{i % 37}.md:{i}", f"f{i % 37}.md", i + 1, i + 2, f"chunk {i}") for i in range(n)]
    q = rng.standard_normal(d).astype(np.float32)
    with StorageManager(dim=d, capacity_rows=n + 8, normalize=False, embedder=lambda t: q) as mgr:
        mgr.index_chunks(chunks, vectors=X)
        got = mgr.search_vector(q, k)
        via_text = mgr.search("anything", k)
    l2_ids, l2_d = oracle_c.scan(X, q, k, metric=oracle_c.METRIC_L2)
    dot_ids, _ = oracle_c.scan(X, q, k, metric=oracle_c.METRIC_DOT)
    assert not np.array_equal(l2_ids, dot_ids)                              # the two metrics really disagree here
    got_rows = np.array([int(c.id.split(":")[1]) for c, _ in got], dtype=np.uint64)
    assert np.array_equal(got_rows, l2_ids)                                 # gaps between distances are >> 1e-5 here
    np.testing.assert_allclose([s for _, s in got], 1.0 - 0.5 * l2_d, rtol=1e-5, atol=1e-5)
    assert [c.id for c, _ in via_text] == [c.id for c, _ in got]


def _all_devices():
    from sema_b200 import _lib
    return list(range(min(int(_lib.lib().sema_device_count()), 8)))


def test_multi_gpu_store_equals_single_gpu_store_and_oracle():
    """sema_store_create_multi: ONE StorageManager handle in ONE process over every visible GPU (the reference is one
    process with one StorageManager).  Runs with however many GPUs the box shows — on a 1-GPU box it is the single-device
    store again; `gpurun --gpus N` exercises the contiguous row ranges, the fused peer exchange behind every search,
    per-shard tombstones, per-shard compaction and the table-order LIKE scan across shards."""
    from sema_b200.storage import Chunk, StorageManager
    devs = _all_devices()
    files = corpus.make_markdown_tree(260, seed=9)
    raw = corpus.chunk_tree(files)
    chunks = [Chunk(c["id"], c["file_path"], c["start_line"], c["end_line"], c["content"]) for c in raw]
    emb = np.stack([corpus.embed(c.content) for c in chunks])
    X = O.normalize(emb)
    n = len(chunks)
    row_of = {c.id: i for i, c in enumerate(chunks)}
    with StorageManager(dim=corpus.DIM, capacity_rows=n + 3, normalize=True, embedder=corpus.embed, devices=devs) as mgr, \
         StorageManager(dim=corpus.DIM, capacity_rows=n + 3, normalize=True, embedder=corpus.embed) as one:
        third = n // 3
        for m in (mgr, one):                                   # three batches: ranges fill in table order, batches straddle shards
            m.index_chunks(chunks[:third], vectors=emb[:third])
            m.index_chunks(chunks[third:2 * third])
            m.index_chunks(chunks[2 * third:], vectors=emb[2 * third:])
            assert len(m) == n
        rng = np.random.default_rng(21)
        for _ in range(40):
            words = " ".join(corpus._WORDS[int(i)] for i in rng.integers(0, len(corpus._WORDS), 5))
            got = mgr.search(words, 50)
            ref = one.search(words, 50)
            assert [c.id for c, _ in got] == [c.id for c, _ in ref]
            assert [s for _, s in got] == [s for _, s in ref]                  # same kernels, same bits
            r_ids, r_sc = O.scan(X, O.normalize(corpus.embed(words)), 50)
            O.check_parity(np.array([row_of[c.id] for c, _ in got], dtype=np.uint64),
                           np.array([s for _, s in got], dtype=np.float32), r_ids, r_sc)
        q = "storage engine table row column"
        assert [g.chunk.id for g in mgr.execute_search(q)] == [g.chunk.id for g in one.execute_search(q)]
        assert [c.id for c, _ in mgr.search("## ", 30)] == [c.id for c, _ in one.search("## ", 30)]   # LIKE: table order
        # remove the files of the first ten hits (they live on different shards), then compact
        victims = sorted({c.file_path for c, _ in mgr.search(q, 10)})
        for v in victims:
            assert mgr.remove_file_chunks(v) == one.remove_file_chunks(v) > 0
        after = mgr.search(q, 50)
        assert all(c.file_path not in victims for c, _ in after)
        assert [c.id for c, _ in after] == [c.id for c, _ in one.search(q, 50)]
        live = mgr.compact()
        assert live == one.compact() == sum(1 for c in chunks if c.file_path not in victims) == len(mgr)
        assert [c.id for c, _ in mgr.search(q, 50)] == [c.id for c, _ in after]
        assert [c.id for c, _ in mgr.search("## ", 30)] == [c.id for c, _ in one.search("## ", 30)]
        more = [Chunk(f"new.md:{i}", "new.md", i + 1, i + 1, "tensor shard merge bandwidth kernel") for i in range(5)]
        mgr.index_chunks(more)                                 # appends land in the first range with room
        assert any(c.file_path == "new.md" for c, _ in mgr.search("tensor shard merge bandwidth kernel", 10))


def test_multi_gpu_store_argument_checks():
    from sema_b200 import SemaError
    from sema_b200.storage import StorageManager
    devs = _all_devices()
    if len(devs) < 2:
        pytest.skip("needs two GPUs")
    with pytest.raises(SemaError):
        StorageManager(dim=corpus.DIM, capacity_rows=0, devices=devs)          # ranges need a capacity
    with pytest.raises(SemaError):
        StorageManager(dim=corpus.DIM, capacity_rows=100, devices=[0, 0])      # one shard per device

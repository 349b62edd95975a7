/*
 * sema_store.h — C ABI of the host-side mirror of Sema's storage boundary.
 *
 * The reference's search boundary is
 *   StorageManager::search(&mut self, query: &str, limit) -> Result<Vec<(Chunk, f32)>>
 *                                                     (src/storage/mod.rs:112-125)
 *   LanceIndexer::{index_chunks, search, remove_file_chunks}
 *                                                     (src/storage/lance_indexer.rs:30-163, 234-250)
 *   Engine::{execute_search, group_results_by_file}   (src/tui/engine.rs:102-182)
 * over the types Chunk / SearchResult (src/types/mod.rs:40-60).  The C++ classes behind this
 * header (sema_b200/csrc/host/storage.hpp) keep those names, argument meanings and error
 * behaviour; the vector column lives in a sema_index (include/sema_b200.h) instead of a Lance
 * table and the returned score is the real cosine instead of the constant 1.0
 * (src/storage/mod.rs:123).  Chunk ids are strings, so the GPU returns row indices and this
 * layer owns the row -> Chunk table.
 *
 * What is NOT here: the embedding model (src/semantic/embeddings.rs:1-82 — ONNX Runtime +
 * hub download; the caller supplies an embedder callback), the Tantivy keyword path ("'"
 * prefix, src/storage/text_indexer.rs), crawling and chunking.
 */
#ifndef SEMA_STORE_H
#define SEMA_STORE_H

#include <stddef.h>
#include <stdint.h>

#include "sema_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sema_store sema_store;

/* (Chunk, f32) of StorageManager::search, by row: the Chunk fields are read with sema_store_chunk */
typedef struct {
    uint64_t row;
    float score;
} sema_hit;

/* SearchResult (src/types/mod.rs:55-60) after group_results_by_file */
typedef struct {
    uint64_t row;                   /* the group's representative: its lowest start_line chunk */
    float score;                    /* that chunk's score */
    uint64_t total_matches_in_file; /* chunks of the same file among the hits */
} sema_search_result;

/* The query embedder: write `dim` floats for `text` into `out`, return 0; non-zero = embedding
 * failed (the reference then falls back to `content LIKE '%query%'`,
 * src/storage/lance_indexer.rs:143-162).  Stands in for VectorStore::generate_embedding
 * (src/semantic/embeddings.rs:26-58).  The output need not be normalised when the store was
 * created with normalize != 0. */
typedef int (*sema_embed_fn)(void *user, const char *text, float *out, uint32_t dim);

#define SEMA_SEARCH_RESULTS_LIMIT 50u /* src/tui/engine.rs:11 */

/* StorageManager::new / LanceIndexer::new (src/storage/mod.rs:19-29): normalize != 0 applies
 * the mean_pool normalise tail (src/semantic/embeddings.rs:83-88) to stored rows and queries
 * on the device — the reference's pipeline, where every vector is unit-norm and LanceDB's default
 * squared-L2 order (lance_indexer.rs:121-126) equals descending cosine; scores are then the cosine.
 * normalize == 0 stores the caller's vectors as given, whatever their norm: the store then ranks by
 * the literal squared-L2 `_distance` d like the reference does, and the score it returns is
 * 1 - d/2 (the cosine again whenever the vectors are unit-norm).  The vector index is growable (sema_index_create_growable): capacity_rows bounds
 * the address space only, 0 = no bound short of the 32-bit row ids — the reference's table has no
 * declared size either. */
int sema_store_create(int device, uint32_t dim, uint64_t capacity_rows, int normalize, sema_store **out);
/* The same store over several GPUs of the box — still ONE handle in ONE process, which is what the reference is
 * (one StorageManager owned by the Engine, src/storage/mod.rs:13-16, src/tui/engine.rs:30).  The table is dealt out
 * in contiguous ranges of ceil(capacity_rows / n_devices) rows per device (capacity_rows must be > 0), so table order
 * = shard order and the row ids a search reports are range base + row within the shard; every vector search is one
 * sema_shard_group_search over a single-process shard group (sema_b200.h: sema_shard_group_create_local): each GPU
 * scans its range, the shards exchange their top-k over NVLink inside the scan kernel, the call returns the global
 * top-k.  limit <= 128 for searches on such a store (the reference asks for 50).  n_devices == 1 is sema_store_create. */
int sema_store_create_multi(const int *devices, uint32_t n_devices, uint32_t dim, uint64_t capacity_rows, int normalize,
                            sema_store **out);
int sema_store_destroy(sema_store *st);
int sema_store_set_embedder(sema_store *st, sema_embed_fn fn, void *user);

/* LanceIndexer::index_chunks (src/storage/lance_indexer.rs:30-105) with the embeddings already
 * computed: n chunks, their n x dim vectors and validity (0 = the embedding failed -> null
 * vector, :66-70; may be NULL).  An empty slice is Ok (:31-33). */
int sema_store_index_chunks(sema_store *st, uint64_t n, const char *const *ids, const char *const *file_paths,
                            const uint64_t *start_lines, const uint64_t *end_lines,
                            const char *const *contents, const float *vectors, const uint8_t *valid);
/* Same, embedding each chunk's content through the embedder callback, sequentially (:59-73). */
int sema_store_index_chunks_embed(sema_store *st, uint64_t n, const char *const *ids,
                                  const char *const *file_paths, const uint64_t *start_lines,
                                  const uint64_t *end_lines, const char *const *contents);

/* The inner seam: nearest_to(query_embedding).limit(limit) (src/storage/lance_indexer.rs:121-126)
 * + score attachment (src/storage/mod.rs:122-123).  hits: `limit` entries, best first. */
int sema_store_search_vector(sema_store *st, const float *query_embedding, uint32_t limit, sema_hit *hits,
                             uint32_t *n_found);
/* StorageManager::search(query, limit) (src/storage/mod.rs:112-125): trims; a "'" prefix is the
 * keyword route (SEMA_ERR_UNSUPPORTED here; a bare "'" is Ok(empty) as in the reference);
 * otherwise embeds the query and runs the vector search; if the embedding fails, the
 * `content LIKE '%query%'` scan (score 1.0, table order). */
int sema_store_search(sema_store *st, const char *query, uint32_t limit, sema_hit *hits, uint32_t *n_found);
/* Engine::execute_search (src/tui/engine.rs:102-154): search(query, 50) -> SearchResult ->
 * group_results_by_file.  out: up to `cap` grouped results, best first. */
int sema_store_execute_search(sema_store *st, const char *query, sema_search_result *out, uint32_t cap,
                              uint32_t *n_out);
/* group_results_by_file alone (src/tui/engine.rs:156-182) over already ranked hits. */
int sema_store_group_results_by_file(sema_store *st, const sema_hit *hits, uint32_t n, sema_search_result *out,
                                     uint32_t cap, uint32_t *n_out);

/* LanceIndexer::remove_file_chunks (src/storage/lance_indexer.rs:234-250): every chunk whose
 * file_path equals `file_path` stops matching.  *removed (may be NULL) = rows deleted. */
int sema_store_remove_file_chunks(sema_store *st, const char *file_path, uint64_t *removed);

/* Compaction after deletions: chunks removed by sema_store_remove_file_chunks leave both the GPU
 * matrix and the row -> Chunk table; rows are renumbered (order kept).  *n_live (may be NULL). */
int sema_store_compact(sema_store *st, uint64_t *n_live);

/* extract_chunk_from_batch (src/storage/lance_indexer.rs:252-281): the Chunk of a row.  The
 * returned strings belong to the store and stay valid until it is destroyed. */
int sema_store_chunk(const sema_store *st, uint64_t row, const char **id, const char **file_path,
                     uint64_t *start_line, uint64_t *end_line, const char **content);
uint64_t sema_store_len(const sema_store *st);  /* chunks indexed (including removed ones) */
const char *sema_store_last_error(void);         /* thread-local text of the last sema_store_* failure */

/* Pure host helpers (no GPU, no store): the caller-side post-processing on plain arrays.
 * sema_group_results_by_file: n ranked hits given by (file_path, start_line, score); writes, per
 * group and best first, the index of the representative hit and the group size. */
int sema_group_results_by_file(uint32_t n, const char *const *file_paths, const uint64_t *start_lines,
                               const float *scores, uint32_t *rep_index, uint64_t *totals, uint32_t *n_groups);
/* SQL `content LIKE '%needle%'` (the fallback predicate of src/storage/lance_indexer.rs:143-147) */
int sema_like_contains(const char *content, const char *needle);
sema_index *sema_store_index(sema_store *st);    /* the underlying GPU index */

#ifdef __cplusplus
}
#endif
#endif /* SEMA_STORE_H */

/*
 * sema_b200.h — C ABI of the B200-native exact vector-search path for Sema.
 *
 * This is the drop-in boundary for ONE path of akshitsinha/sema: the flat exact
 * nearest-neighbour scan with top-k over the chunk-embedding index.  Every entry
 * point names the reference interface (file:line under the reference tree) it
 * replaces.  Plain pointers and sizes only; no C++/torch types; nothing throws or
 * aborts across this boundary.  All functions return SEMA_OK (0) or a negative
 * SEMA_ERR_* code; sema_last_error() returns a thread-local description.
 *
 * Threading: a handle may be used from different OS threads (the reference's tokio
 * runtime moves `&mut self` calls between workers, src/storage/mod.rs:112) but by
 * at most one thread at a time; every entry point re-binds the CUDA device.
 * There is NO CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef SEMA_B200_H
#define SEMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SEMA_OK 0
#define SEMA_ERR_INVALID (-1)     /* bad argument                               */
#define SEMA_ERR_CUDA (-2)        /* CUDA runtime/driver error                  */
#define SEMA_ERR_CAPACITY (-3)    /* append would exceed capacity_rows          */
#define SEMA_ERR_NOMEM (-4)       /* host or device allocation failed           */
#define SEMA_ERR_UNSUPPORTED (-5) /* shape outside what the kernels cover       */

/* Ranking metric.  The reference never sets distance_type() on its query
 * (src/storage/lance_indexer.rs:121-126), i.e. LanceDB's default squared L2,
 * ascending; on the unit-norm rows the embedder produces
 * (src/semantic/embeddings.rs:83-88) that order equals descending cosine. */
#define SEMA_METRIC_COSINE 0 /* score = q.x (cosine on unit rows), best = largest        */
#define SEMA_METRIC_L2 1     /* score = sum (q-x)^2 = LanceDB `_distance`, best = smallest */

#define SEMA_MAX_K 1024u   /* largest `limit`; k <= 128 is one fused pass, larger k = ceil(k/128) passes */
#define SEMA_MAX_INFLIGHT 8 /* host searches submitted and not yet collected, per handle          */
#define SEMA_MAX_DIM 8192u /* dim 384 / 768 have unrolled kernels; others use the generic kernel */

typedef struct sema_index sema_index;

/* ---- lifecycle -----------------------------------------------------------
 * Replaces LanceIndexer::new (src/storage/lance_indexer.rs:19-28) +
 * create_table("chunks") (:97-101).  The handle owns the row-major, 16-byte
 * aligned fp32 matrix in HBM (capacity_rows x round_up(dim,4)), its streams and
 * staging buffers.  metric is SEMA_METRIC_*. */
int sema_index_create(int device, uint32_t dim, uint64_t capacity_rows, int metric,
                      sema_index **out);
/* Growable variant — what the reference's table is (table.add appends without a declared size,
 * src/storage/lance_indexer.rs:92-95): max_rows only reserves ADDRESS SPACE for the matrix (and the
 * validity bytes, and K3's bf16 planes); physical HBM is mapped in 256 MB steps as rows are
 * appended (CUDA virtual memory management), so the matrix stays contiguous for the scan kernels,
 * never needs a realloc-and-copy, and an index that holds 10 k rows costs 256 MB whatever its
 * max_rows.  Everything else behaves like sema_index_create; sema_index_capacity() = max_rows. */
int sema_index_create_growable(int device, uint32_t dim, uint64_t max_rows, int metric,
                               sema_index **out);
int sema_index_destroy(sema_index *idx);

/* ---- ingest (kernel K1) --------------------------------------------------
 * Replaces the vector half of LanceIndexer::index_chunks
 * (src/storage/lance_indexer.rs:30-105): `rows` is the contiguous n x dim f32
 * values buffer of the FixedSizeList<Float32,dim> column (:75-76); `valid` is its
 * validity (one byte per row, 0 = null vector = failed embedding, :66-70) or NULL
 * for "all valid".  normalize != 0 applies the mean_pool tail
 * (src/semantic/embeddings.rs:83-88) on the device.  Null rows and rows holding a
 * non-finite value are never returned by a search.  On return the rows are
 * visible to every later search (table.add(..).await, :92-95); *first_row (may be
 * NULL) receives the local row index of rows[0]. */
int sema_index_append(sema_index *idx, const float *rows, uint64_t n, const uint8_t *valid,
                      int normalize, uint64_t *first_row);
/* Same, but returns once the copy + K1 are enqueued on the ingest stream; `rows`
 * and `valid` must stay untouched until sema_index_flush() (use pinned memory from
 * sema_host_alloc for a truly asynchronous copy).  Rows become visible to searches
 * that start after their ingest has completed on the device (snapshot semantics). */
int sema_index_append_async(sema_index *idx, const float *rows, uint64_t n,
                            const uint8_t *valid, int normalize, uint64_t *first_row);
int sema_index_flush(sema_index *idx);
/* rows already resident on this device (n x dim, dense). */
int sema_index_append_device(sema_index *idx, const float *rows_dev, uint64_t n,
                             const uint8_t *valid_dev, int normalize, uint64_t *first_row);
/* Benchmark/test corpus generated on the device: value(seed,row,col) of SURVEY.md
 * §8(d) (bit-identical to oracle/cpu_scan.c:sema_oracle_synth) for rows
 * [synth_row0, synth_row0+n), then K1. */
int sema_index_append_synthetic(sema_index *idx, uint64_t seed, uint64_t synth_row0, uint64_t n,
                                int normalize, uint64_t *first_row);

/* ---- the step before the path: mean pooling (kernel K0) ----------------------
 * Replaces mean_pool (src/semantic/embeddings.rs:61-91), which turns the embedder's
 * last_hidden_state into the vector that is stored (:27-59 -> lance_indexer.rs:63-71) or used as
 * the query (lance_indexer.rs:114-118): pooled[j] = sum_i tokens[i][j]*mask[i] / sum_i mask[i],
 * then the L2 normalise tail.  tokens: n x seq_len x dim fp32 (the reference runs n = 1,
 * seq_len = MAX_LENGTH = 256, :7); mask: n x seq_len fp32 (attention_mask_f32, :38-46).  All sums
 * keep the reference's order (i ascending, j ascending, multiply then add), so the output is
 * bit-identical to it.  skip_masked != 0 does not read the rows of tokens whose mask is exactly 0
 * (padding) — identical results for finite inputs, less HBM traffic.  out: n x dim dense.
 * The _device variant enqueues on the query stream and does not synchronise. */
int sema_mean_pool(sema_index *idx, const float *tokens, const float *mask, uint64_t n,
                   uint32_t seq_len, int skip_masked, float *out);
int sema_mean_pool_device(sema_index *idx, const float *tokens_dev, const float *mask_dev, uint64_t n,
                          uint32_t seq_len, int skip_masked, float *out_dev);
/* mean_pool fused with the append: K0 writes the pooled, normalised rows straight into their final
 * place in the matrix (no intermediate n x dim buffer), then the column bookkeeping of
 * sema_index_append_device runs in place.  valid_dev: nullable validity bytes (0 = the embedding
 * failed, lance_indexer.rs:66-70).  Rows are visible when the call returns. */
int sema_index_append_pooled_device(sema_index *idx, const float *tokens_dev, const float *mask_dev,
                                    uint64_t n, uint32_t seq_len, const uint8_t *valid_dev,
                                    int skip_masked, uint64_t *first_row);

/* Replaces remove_file_chunks' `table.delete(predicate)`
 * (src/storage/lance_indexer.rs:234-250): the listed local rows stop matching. */
int sema_index_tombstone(sema_index *idx, const uint64_t *rows, uint64_t n);

/* Compaction after deletions (LanceDB's `optimize`/compact step for the delete at
 * src/storage/lance_indexer.rs:234-250): live rows move down, keeping their order, dead rows
 * (null / tombstoned) disappear and size shrinks.  new_row_of_old (may be NULL) receives, for every
 * old row, its new index or UINT64_MAX if it was dropped, so the caller can remap its row -> Chunk
 * table; *n_live (may be NULL) the new size.  The plan (new positions by prefix sums over the keep
 * flags, gather list, map) is built on the device and the rows move by an ordered gather queued on
 * the query stream; the call is synchronous.  Its device scratch (about 6 bytes per row, 14 with the
 * map, plus one bounce chunk of 65 536 rows while rows have to move by less than a chunk) stays with
 * the handle for the next call.  A failure after rows have started to move leaves the handle refusing
 * further work (destroy and rebuild it). */
int sema_index_compact(sema_index *idx, uint64_t *new_row_of_old, uint64_t *n_live);
/* Same with an explicit keep mask (one byte per row, 0 = drop): kept rows keep their state, so a
 * null-vector row that is kept stays a (never matching) null row. */
int sema_index_compact_keep(sema_index *idx, const uint8_t *keep, uint64_t *new_row_of_old, uint64_t *n_live);

/* On-disk cache of the vector column: the raw rows x dim fp32 matrix as stored
 * (already normalised) plus the validity bytes, so a restart re-uploads instead of re-embedding.
 * File: 64-byte header {"SEMAIDX1", u32 dim, i32 metric, u64 n_rows, u32 version = 1, u32 byte-order
 * tag 0x01020304, zero padding}, n_rows validity bytes, n_rows*dim floats (little-endian fp32, dense).
 * sema_index_load checks the header against the file (version, byte order, field ranges, and
 * size == 64 + n_rows + n_rows*dim*4) before it allocates anything: a corrupt or truncated file is
 * SEMA_ERR_INVALID.  Neither call holds more than a 64 k-row staging chunk on the host. */
int sema_index_save(sema_index *idx, const char *path);
int sema_index_load(const char *path, int device, uint64_t capacity_rows, sema_index **out);

/* ---- search (kernel K2; K3 for batches) -----------------------------------
 * Replaces table.query().nearest_to(q)?.limit(k).execute()
 * (src/storage/lance_indexer.rs:121-126).  q: dim floats (unit-norm for cosine).
 * Out (caller-allocated, k entries): row_ids = row_base + local row, best first;
 * scores = cosine (descending) or squared-L2 `_distance` (ascending).  Exact ties
 * rank the lower row id first.  *n_found = min(k, visible valid rows); an empty
 * index is SEMA_OK with *n_found = 0 (:108-111).
 * Cost of a call with dim 384 / 768 and k <= 128: ONE kernel launch and no copies — the query
 * travels in the kernel's parameters and the last block stores the result block and a completion
 * flag into mapped host memory, which this call polls. */
int sema_index_search(sema_index *idx, const float *q, uint32_t k, uint64_t *row_ids,
                      float *scores, uint32_t *n_found);
/* Asynchronous form of sema_index_search for callers that keep several queries in flight (a search
 * service rather than the reference's one-query-at-a-time TUI): submit returns as soon as the scan
 * is enqueued (q is consumed before it returns) and hands out a ticket; collect waits for that
 * ticket and fills the caller's buffers exactly like sema_index_search.  Up to SEMA_MAX_INFLIGHT
 * tickets may be outstanding per handle; they complete in submission order and may be collected in
 * any order.  Consecutive submitted scans are chained on the device (programmatic dependent launch),
 * so with two or more in flight the handle serves host queries at the query-stream rate.  Shapes
 * outside the fast path (dim other than 384 / 768, k > 128) run synchronously inside submit.  sema_index_search(q) == submit(q) + collect. */
int sema_index_search_submit(sema_index *idx, const float *q, uint32_t k, uint64_t *ticket);
int sema_index_search_collect(sema_index *idx, uint64_t ticket, uint64_t *row_ids, float *scores,
                              uint32_t *n_found);
/* nq queries (Q: nq x dim row-major); outputs nq x k row-major, n_found[nq].  With
 * dim % 64 == 0, dim <= 768, k <= 100, nq >= 4 and either the cosine metric or the L2 metric over
 * rows of (nearly) constant norm — the reference's case: unit-norm rows under LanceDB's default
 * squared-L2 `_distance` — this runs kernel K3 (tcgen05
 * tensor cores, 16-bit split precision — see sema_index_set_batch_mode / _set_batch_precision — with exact fp32
 * re-scoring of the candidates; needs a second dim*4 bytes per row of HBM for the 16-bit hi/lo planes,
 * built on first use); otherwise, or if that memory cannot be had, K2 runs once per query.
 * Results are identical either way. */
int sema_index_search_batch(sema_index *idx, const float *Q, uint32_t nq, uint32_t k,
                            uint64_t *row_ids, float *scores, uint32_t *n_found);
/* Same with queries and results resident on the device (Q_dev: nq x dim dense): no host copies of
 * queries or results.  The call still SYNCHRONISES the query stream once per tensor-core stage: each
 * stage ends with a per-query exactness flag that the host reads back (4 bytes per query) to decide
 * which queries go on to the next stage of the cascade / to K2.  On return the work of the last
 * stage may still be running on the query stream; results are ordered after it on that stream. */
int sema_index_search_batch_device(sema_index *idx, const float *Q_dev, uint32_t nq, uint32_t k,
                                   uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev);
/* A query stream: nq independent single-query searches (kernel K2, one HBM pass per query — the
 * reference's own pattern of one nearest_to() call per query, lance_indexer.rs:121-126) issued back
 * to back on the query stream.  With dim 384 / 768 and k <= 128 the stream runs at the HBM rate
 * without a per-query launch / merge gap, in one of two forms the library picks by corpus size:
 * ONE persistent (cooperative) launch for the whole stream — a TMA producer warp keeps streaming
 * rows across query boundaries while a finisher warp merges, exchanges (shard groups) and publishes
 * the query just scanned — when one scan is long against that finisher's work (k <= 16: about 100 k
 * rows of dim 384 and up; k <= 64: about 600 k rows; larger k: about 800 k rows), otherwise one launch per query with
 * consecutive launches chained by programmatic dependent launch (query i+1 starts scanning on the
 * SMs query i has left while i's last block still merges).  Both forms give bit-identical results.
 * If the device cannot hold the persistent grid at once (SM partitioning), the launch is refused by
 * the driver and the stream falls back to the chained form.  Same layouts and results as
 * sema_index_search_batch_device; all rows the stream sees are one snapshot. */
int sema_index_search_stream_device(sema_index *idx, const float *Q_dev, uint32_t nq, uint32_t k,
                                    uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev);
/* mode 0 = automatic: when the shape allows and nq >= 4, a precision cascade on the tensor cores —
 * a single 16-bit pass (q_hi . x_hi) as a coarse candidate filter (a third of the tensor work), then the
 * three-pass error-compensated split (q_hi.x_hi + q_lo.x_hi + q_hi.x_lo) for the queries whose exactness
 * proof failed under the looser single-pass error bound, then K2 for the few neither can prove (exact
 * ties beyond the candidate list);
 * 1 = always K2 per query; 2 = K3 three-pass split only (+ K2 fallback); 3 = K3 single pass only (+ K2
 * fallback).  Every path ends in the same exact fp32 re-scoring: results are identical.  Other
 * values only query.  Returns the mode now active. */
int sema_index_set_batch_mode(sema_index *idx, int mode);
/* Element format of the split: 0 = automatic (default) — fp16 halves (11 significant bits: per-stage error
 * bounds 1.15e-3 / 2.0e-4 of |q||x| for one / three passes) whenever every stored element is at most 1024 in
 * magnitude (max |x|^2 <= 2^20; queries of any magnitude are scaled by a power of two inside the kernel),
 * bf16 halves (8 significant bits, fp32's range: 8.5e-3 / 2.5e-4) otherwise; 1 = bf16 always (the 3 x bf16 split
 * by name).  Either way the tensor-core pass only selects candidates; results are identical.  The planes are
 * re-split (one read + one write of the corpus) when the active format changes.  Other values only query.
 * Returns the preference now set; _active returns the format the planes are in (0 bf16, 1 fp16, -1 none built yet). */
int sema_index_set_batch_precision(sema_index *idx, int prec);
int sema_index_batch_precision_active(const sema_index *idx);
/* queries served by K3 so far; how many of them were re-run through K2 because exactness could not
 * be proven from the candidate lists (heavy ties / duplicates); how many went from the single-pass
 * stage to the bf16x3 stage in automatic mode.  Any pointer may be NULL. */
int sema_index_batch_stats(const sema_index *idx, uint64_t *k3_queries, uint64_t *k3_fallbacks,
                           uint64_t *k3_cascaded);

/* ---- device-resident variants (no host copies, no synchronisation) --------
 * Used for kernel-only timing and by the sharded path.  q_dev: dim floats on the
 * device.  keys_dev: k packed 64-bit ranking keys, best first, 0 = empty slot:
 *   key = ordered_u32(rank value) << 32 | (0xFFFFFFFF - global_row_id)
 * so that a plain unsigned compare orders by score then by lower row id.  Work is
 * enqueued on the index's query stream (see sema_index_set_stream). */
int sema_index_search_keys_device(sema_index *idx, const float *q_dev, uint32_t k,
                                  uint64_t *keys_dev);
/* Same search with the decoded result left on the device: ids_dev/scores_dev hold k
 * entries, n_found_dev one uint32 (for callers whose embedder already runs on the GPU). */
int sema_index_search_device(sema_index *idx, const float *q_dev, uint32_t k, uint64_t *ids_dev,
                             float *scores_dev, uint32_t *n_found_dev);
/* Kernel K4: merge n_lists ranked key lists of k entries each (e.g. the allgathered
 * per-shard results) into the global top-k and decode it.  All pointers are device
 * pointers; ids_dev/scores_dev hold k entries, n_found_dev one uint32. */
int sema_topk_merge_device(sema_index *idx, const uint64_t *keys_dev, uint32_t n_lists,
                           uint32_t k, uint64_t *ids_dev, float *scores_dev,
                           uint32_t *n_found_dev);

/* Batched forms of the two calls above, for the corpus-sharded batched search (SURVEY.md §8(e):
 * every shard runs the batch on its rows — K3 when the shape allows — and contributes nq x k packed
 * keys to one all-gather; kernel K4 then merges per query).  keys_dev of the first: [nq][k];
 * keys_dev of the second: [n_lists][nq][k] (the all-gathered buffer), k <= 128. */
int sema_index_search_batch_keys_device(sema_index *idx, const float *Q_dev, uint32_t nq, uint32_t k,
                                        uint64_t *keys_dev);
int sema_topk_merge_batch_device(sema_index *idx, const uint64_t *keys_dev, uint32_t n_lists,
                                 uint32_t nq, uint32_t k, uint64_t *ids_dev, float *scores_dev,
                                 uint32_t *n_found_dev);

/* ---- corpus-sharded group: scan + exchange + merge in ONE kernel per rank -------------
 * One process per GPU; every rank holds a contiguous row range (sema_index_set_row_base).  The
 * last block of K2 stores the shard's top-k keys straight into every rank's exchange buffer over
 * NVLink (peer stores on cudaIpc-mapped memory), raises a flag, waits for all ranks' flags and
 * merges the world*k keys itself: no NCCL call and no second launch on the query path.  Setup:
 * create -> local_handle -> (all-gather the 64-byte handles with any transport) -> connect.
 * Every rank must then issue the same sequence of group searches (SPMD); a missing rank is
 * reported as an error after ~2 s instead of hanging.  k <= 128. */
typedef struct sema_shard_group sema_shard_group;
#define SEMA_IPC_HANDLE_BYTES 64
#define SEMA_MAX_SHARDS 16
int sema_shard_group_create(sema_index *idx, uint32_t world, uint32_t rank, sema_shard_group **out);
/* Single-process form — what a drop-in for the reference needs: Sema is ONE process that owns ONE
 * StorageManager (src/main.rs:9, src/storage/mod.rs:13-16, src/tui/engine.rs:30), so all the GPUs of
 * the box have to hang off one handle.  shards[r] is an index created on its own device (one shard
 * per GPU; set each shard's sema_index_set_row_base so that ids are global; ingest into the shards
 * with the ordinary append calls).  Peer access between the devices replaces the IPC handles, so
 * there is no handle exchange and no connect step.  sema_shard_group_search / _submit / _collect on
 * such a group take ONE host call: a worker thread per shard launches that shard's fused scan +
 * exchange + merge kernel (all N launches leave the host together), every GPU ends up with the
 * global top-k and the call returns shard 0's copy.  While the group exists, search its shards only
 * through the group.  The device-resident variants (_search_device, _search_stream_device) return
 * SEMA_ERR_UNSUPPORTED on a local group: a device-resident query lives on one GPU. */
int sema_shard_group_create_local(sema_index *const *shards, uint32_t n_shards, sema_shard_group **out);
int sema_shard_group_local_handle(sema_shard_group *g, void *handle_out /* 64 bytes */);
int sema_shard_group_connect(sema_shard_group *g, const void *handles /* world x 64 bytes, by rank */);
int sema_shard_group_search(sema_shard_group *g, const float *q, uint32_t k, uint64_t *row_ids,
                            float *scores, uint32_t *n_found);
/* submit / collect forms of sema_shard_group_search (same contract as sema_index_search_submit;
 * every rank submits the same queries in the same order) */
int sema_shard_group_search_submit(sema_shard_group *g, const float *q, uint32_t k, uint64_t *ticket);
int sema_shard_group_search_collect(sema_shard_group *g, uint64_t ticket, uint64_t *row_ids, float *scores,
                                    uint32_t *n_found);
int sema_shard_group_search_device(sema_shard_group *g, const float *q_dev, uint32_t k,
                                   uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev);
/* nq group searches issued back to back (Q_dev: nq x dim, results nq x k), in the two forms of
 * sema_index_search_stream_device: one persistent launch per rank whose finisher warp runs the peer
 * exchange of query i while the rest of the GPU scans query i+1, or launches chained with programmatic
 * dependent launch — either way a rank's next scan overlaps the wait for its peers' keys.  Ranks may
 * mix the two forms (both use the same sequence numbers), results are identical. */
int sema_shard_group_search_stream_device(sema_shard_group *g, const float *Q_dev, uint32_t nq, uint32_t k,
                                          uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev);
int sema_shard_group_destroy(sema_shard_group *g);

/* ---- properties ------------------------------------------------------------ */
/* Shard offset of row 0.  HARD LIMIT: global row ids are 32-bit inside the ranking keys the kernels
 * exchange (key = ordered score << 32 | ~id), so row_base + capacity must stay below 2^32 - 1
 * (4.29 billion rows per logical index; a 100M-row corpus uses 2 % of it); a row_base that would let
 * an id reach 2^32 - 1 is refused with SEMA_ERR_INVALID here, and an append that would do so with
 * SEMA_ERR_CAPACITY.  The uint64_t types at this boundary are the reference-facing width only. */
int sema_index_set_row_base(sema_index *idx, uint64_t row_base);
/* external != 0: run searches on the caller's cudaStream_t `cuda_stream` (0 = the CUDA
 * default stream); external == 0: back to the handle's own non-blocking query stream. */
int sema_index_set_stream(sema_index *idx, void *cuda_stream, int external);
uint64_t sema_index_size(const sema_index *idx);     /* rows appended (visible + in flight) */
uint64_t sema_index_visible(sema_index *idx);        /* rows a search starting now would scan */
uint64_t sema_index_capacity(const sema_index *idx);
uint32_t sema_index_dim(const sema_index *idx);
int sema_index_device(const sema_index *idx);
/* on != 0: sema_index_search / _search_batch first apply the mean_pool normalise tail
 * (src/semantic/embeddings.rs:83-88) to the query on the device, as the reference's embedder does
 * for queries and rows alike: kernel K1 for batches and the staged path, the same arithmetic in
 * K2's registers on the host-query path (identical bits either way).  Returns the setting now active (on < 0 = query). */
int sema_index_set_normalize_queries(sema_index *idx, int on);
/* snapshot (visible rows) the most recent search on this handle scanned */
uint64_t sema_index_last_snapshot(const sema_index *idx);
/* copy stored (normalised) rows back to the host: out = n x dim floats */
int sema_index_read_rows(sema_index *idx, uint64_t first_row, uint64_t n, float *out);
/* Tuning knob for measurements.  NO setting changes a result: every value selects among kernels / launch shapes
 * that produce identical output (the work-skipping timing probes of earlier versions are compiled out of this
 * library; they exist only in the separate probe build made by scripts/build_probe.sh and are refused here).
 * 0 = default K2 kernel (TMA bulk-copy ring for dim 384 / 768), 1 / 2 / 3 = the register-fed K2 kernel with
 * 4 / 2 / 8 rows per warp batch; 100 + c = K3 cluster size c of the single-CTA kernels (0 = automatic);
 * 200 / 201 = K3 two / one query tiles per CTA in the single-CTA single-pass kernel; 300 / 308 = K3 epilogue
 * with / without its group early-out; 400 / 401 = K3 single-pass candidate lists of 32 / 16 for k <= 10;
 * 500 / 501 = host searches staged through H2D + D2H copies / query by kernel parameter + results to mapped
 * host memory (default); 600 / 601 = query streams unchained / chained (default); 700 / 701 = K3 single-pass
 * stage on the single-CTA kernel (default) / on CTA pairs (tcgen05 cta_group::2); 800 + d = K3 producer
 * prefetches into L2 d stages ahead (default 0: measured no gain); 900 / 901 / 902 = query streams as one
 * persistent launch when a scan is long enough (default) / whenever the shape allows / never (always one launch
 * per query); 1100 / 1101 = a K3 stage as one launch / as
 * two concurrent launches, clusters of 4 plus clusters of 2 on the SMs those leave free (default); 1200 + w = row
 * weight of a 4-cluster partition in that split, 0.70 + w / 100 (0 = built-in 1.05); negative = query.
 * Returns the value set, or -1 for a value this build does not have. */
int sema_index_set_scan_variant(sema_index *idx, int variant);
/* number of kernels this handle has launched so far */
uint64_t sema_index_launch_count(const sema_index *idx);

/* ---- misc -------------------------------------------------------------------- */
int sema_host_alloc(void **out, size_t bytes); /* pinned host memory */
int sema_host_free(void *p);
int sema_device_count(void);
const char *sema_last_error(void);
const char *sema_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SEMA_B200_H */

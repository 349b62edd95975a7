// storage.hpp — host-side mirror of Sema's storage / search boundary (C++17).
//
// Same names, argument meanings and error behaviour as the reference's Rust types:
//   Chunk, SearchResult            src/types/mod.rs:40-47, 55-60
//   GpuVectorIndexer               LanceIndexer        src/storage/lance_indexer.rs:14-163, 234-281
//   StorageManager                 StorageManager      src/storage/mod.rs:13-132 (vector route)
//   group_results_by_file          Engine::group_results_by_file   src/tui/engine.rs:156-182
//   execute_search                 Engine::execute_search          src/tui/engine.rs:102-154
// The vector column lives in HBM behind the C ABI of include/sema_b200.h; this layer owns the
// row -> Chunk table (ids are strings, src/storage/processor.rs:62) and the (Chunk, score)
// pairing.  Errors are reported as Status {code, message} instead of anyhow::Result.
#pragma once
#include <cstdint>
#include <functional>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "../../../include/sema_b200.h"

namespace sema_host {

struct Status {
    int code = SEMA_OK;
    std::string message;
    bool ok() const { return code == SEMA_OK; }
    static Status Ok() { return {}; }
    static Status Err(int c, std::string m) { return {c, std::move(m)}; }
};

// src/types/mod.rs:40-47
struct Chunk {
    std::string id;         // "{file_path}:{chunk_idx}" (src/storage/processor.rs:62)
    std::string file_path;  // PathBuf in the reference
    size_t start_line = 0;
    size_t end_line = 0;
    std::string content;
};

// src/types/mod.rs:55-60
struct SearchResult {
    Chunk chunk;
    float score = 0.0f;
    size_t total_matches_in_file = 0;
    uint64_t row = 0;  // not in the reference: the GPU row the chunk came from
};

// VectorStore::generate_embedding (src/semantic/embeddings.rs:26-58): nullopt = embedding failed
using Embedder = std::function<std::optional<std::vector<float>>(const std::string &)>;

constexpr size_t SEARCH_RESULTS_LIMIT = 50;  // src/tui/engine.rs:11

// src/tui/engine.rs:156-182.  Groups by file_path, keeps each file's lowest-start_line chunk,
// records the group size, then orders by score, best first.  The reference groups through a
// HashMap (iteration order unspecified) and stable-sorts; here groups are visited in first-
// appearance (i.e. rank) order, which fixes the order of equal-score groups deterministically.
std::vector<SearchResult> group_results_by_file(std::vector<SearchResult> results);

// SQL `content LIKE '%<needle>%'` as DataFusion evaluates it for the fallback at
// src/storage/lance_indexer.rs:143-147: case sensitive, '%' = any run, '_' = any one byte.
bool like_contains(const std::string &content, const std::string &needle);

// The `chunks` table: vector column on the GPU, the other five columns on the host.
class GpuVectorIndexer {
public:
    GpuVectorIndexer() = default;
    ~GpuVectorIndexer();
    GpuVectorIndexer(const GpuVectorIndexer &) = delete;
    GpuVectorIndexer &operator=(const GpuVectorIndexer &) = delete;

    // LanceIndexer::new (src/storage/lance_indexer.rs:19-28)
    Status open(int device, uint32_t dim, uint64_t capacity_rows, bool normalize);
    // The same table over several GPUs of the box, still ONE handle in ONE process like the reference's
    // StorageManager (src/storage/mod.rs:13-16): the rows are dealt out in contiguous ranges of
    // ceil(capacity_rows / n) rows per device (table order = shard order), every search is one
    // sema_shard_group_search over a single-process shard group (sema_shard_group_create_local: each GPU scans its
    // range, the shards exchange their top-k over NVLink inside the scan kernel, the call returns the global
    // top-k).  capacity_rows must be given (> 0) when n_devices > 1; limit <= 128 for searches then.
    Status open_multi(const int *devices, uint32_t n_devices, uint32_t dim, uint64_t capacity_rows, bool normalize);

    // index_chunks (:30-105).  `vectors`: chunks.size() x dim; valid[i] == 0 marks a failed
    // embedding (null vector).  Empty input is Ok (:31-33).
    Status index_chunks(const std::vector<Chunk> &chunks, const float *vectors, const uint8_t *valid);
    // index_chunks embedding each chunk's content, sequentially (:59-73)
    Status index_chunks(const std::vector<Chunk> &chunks, const Embedder &embed);

    // the vector branch of search (:121-141): rows in rank order with their real scores
    Status search(const float *query_embedding, size_t limit, std::vector<std::pair<Chunk, float>> *out,
                  std::vector<uint64_t> *rows = nullptr);
    // the `content LIKE '%query%'` fallback (:143-162): first `limit` live rows in table order
    Status search_like(const std::string &query, size_t limit, std::vector<std::pair<Chunk, float>> *out,
                       std::vector<uint64_t> *rows = nullptr);
    // remove_file_chunks (:234-250)
    Status remove_file_chunks(const std::string &file_path, uint64_t *removed = nullptr);

    // compaction after deletions: drops removed rows on the GPU and renumbers the row -> Chunk
    // table to match (row ids returned by later searches refer to the compacted table)
    Status compact(uint64_t *n_live = nullptr);

    // row = the id a search reports: shard base + row within the shard (one shard: simply the table row)
    const Chunk *chunk(uint64_t row) const;
    uint64_t len() const;                                  // chunks in the table (all shards)
    uint32_t dim() const { return dim_; }
    uint32_t n_shards() const { return (uint32_t)shards_.size(); }
    sema_index *index() { return shards_.empty() ? nullptr : shards_[0].idx; }   // shard 0's index (the only one on one GPU)

private:
    struct Shard {
        sema_index *idx = nullptr;
        uint64_t base = 0;            // global id of its row 0 (sema_index_set_row_base)
        uint64_t cap = 0;             // rows it may hold
        std::vector<Chunk> chunks;    // row -> Chunk (extract_chunk_from_batch, :252-281)
        std::vector<uint8_t> live;    // 0 after remove_file_chunks
    };
    bool locate(uint64_t row, size_t *shard, uint64_t *local) const;
    std::vector<Shard> shards_;
    sema_shard_group *group_ = nullptr;   // single-process shard group over the shards (more than one GPU only)
    uint64_t per_ = 0;                    // rows per shard range (more than one GPU only)
    uint32_t dim_ = 0;
    bool normalize_ = false;
    int metric_ = SEMA_METRIC_COSINE;   // SEMA_METRIC_L2 when the caller's vectors are stored as given (normalize = false)
    std::vector<uint64_t> ids_buf_;
    std::vector<float> sc_buf_;
};

// src/storage/mod.rs:13-132 — the vector route of the façade.
class StorageManager {
public:
    Status open(int device, uint32_t dim, uint64_t capacity_rows, bool normalize)
    {
        return lance_indexer.open(device, dim, capacity_rows, normalize);
    }
    Status open_multi(const int *devices, uint32_t n_devices, uint32_t dim, uint64_t capacity_rows, bool normalize)
    {
        return lance_indexer.open_multi(devices, n_devices, dim, capacity_rows, normalize);
    }
    void set_embedder(Embedder e) { embedder_ = std::move(e); }

    // index_chunks (src/storage/mod.rs:96-110): a failing vector index only warns there; here the
    // error is returned to the caller as well.
    Status index_chunks(const std::vector<Chunk> &chunks) { return lance_indexer.index_chunks(chunks, embedder_); }

    // search (src/storage/mod.rs:112-125)
    Status search(const std::string &query, size_t limit, std::vector<std::pair<Chunk, float>> *out,
                  std::vector<uint64_t> *rows = nullptr);

    // Engine::execute_search (src/tui/engine.rs:102-154): search(query, 50) + grouping
    Status execute_search(const std::string &query, std::vector<SearchResult> *out);

    GpuVectorIndexer lance_indexer;  // the reference's field name

private:
    Embedder embedder_;
};

}  // namespace sema_host

// storage.cpp — host-side mirror of Sema's storage boundary + its C ABI (include/sema_store.h).
#include "storage.hpp"

#include <algorithm>
#include <cstring>
#include <new>
#include <unordered_map>

#include "../../../include/sema_store.h"

namespace sema_host {

namespace {
Status from_rc(int rc)
{
    if (rc == SEMA_OK) return Status::Ok();
    return Status::Err(rc, sema_last_error());
}

std::string trim(const std::string &s)
{
    // Rust str::trim: Unicode White_Space; ASCII subset + NBSP-free approximation is enough
    // for queries typed into the TUI search box
    size_t a = 0, b = s.size();
    auto ws = [](unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); };
    while (a < b && ws((unsigned char)s[a])) ++a;
    while (b > a && ws((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}
}  // namespace

// ------------------------------------------------------------------ group_results_by_file
std::vector<SearchResult> group_results_by_file(std::vector<SearchResult> results)
{
    // src/tui/engine.rs:157-164: file_groups.entry(file_path).or_default().push(result)
    std::unordered_map<std::string, size_t> slot;
    std::vector<std::vector<SearchResult>> groups;
    for (auto &r : results) {
        auto it = slot.find(r.chunk.file_path);
        if (it == slot.end()) {
            it = slot.emplace(r.chunk.file_path, groups.size()).first;
            groups.emplace_back();
        }
        groups[it->second].push_back(std::move(r));
    }
    // :166-174: sort_by_key(start_line) (stable), keep the first, record the group size
    std::vector<SearchResult> grouped;
    grouped.reserve(groups.size());
    for (auto &g : groups) {
        std::stable_sort(g.begin(), g.end(),
                         [](const SearchResult &a, const SearchResult &b) { return a.chunk.start_line < b.chunk.start_line; });
        const size_t total = g.size();
        SearchResult first = std::move(g.front());
        first.total_matches_in_file = total;
        grouped.push_back(std::move(first));
    }
    // :176-180: sort_by(|a, b| b.score.partial_cmp(&a.score).unwrap_or(Equal)) — stable, NaN = equal
    std::stable_sort(grouped.begin(), grouped.end(), [](const SearchResult &a, const SearchResult &b) {
        return a.score > b.score;  // false for NaN operands, like unwrap_or(Equal)
    });
    return grouped;
}

// ------------------------------------------------------------------ LIKE '%needle%'
namespace {
// classic wildcard match of `pat` against the whole of `s`: '%' any run, '_' one byte
bool like_match(const char *s, size_t n, const char *pat, size_t m)
{
    size_t i = 0, j = 0, star_j = std::string::npos, star_i = 0;
    while (i < n) {
        if (j < m && (pat[j] == '_' || pat[j] == s[i]) && pat[j] != '%') {
            ++i; ++j;
        } else if (j < m && pat[j] == '%') {
            star_j = j++;
            star_i = i;
        } else if (star_j != std::string::npos) {
            j = star_j + 1;
            i = ++star_i;
        } else {
            return false;
        }
    }
    while (j < m && pat[j] == '%') ++j;
    return j == m;
}
}  // namespace

bool like_contains(const std::string &content, const std::string &needle)
{
    const std::string pat = "%" + needle + "%";
    return like_match(content.data(), content.size(), pat.data(), pat.size());
}

// ------------------------------------------------------------------ GpuVectorIndexer
GpuVectorIndexer::~GpuVectorIndexer()
{
    if (group_) sema_shard_group_destroy(group_);
    for (Shard &sh : shards_)
        if (sh.idx) sema_index_destroy(sh.idx);
}

Status GpuVectorIndexer::open(int device, uint32_t dim, uint64_t capacity_rows, bool normalize)
{
    return open_multi(&device, 1, dim, capacity_rows, normalize);
}

Status GpuVectorIndexer::open_multi(const int *devices, uint32_t n_devices, uint32_t dim, uint64_t capacity_rows, bool normalize)
{
    if (!shards_.empty()) return Status::Err(SEMA_ERR_INVALID, "indexer already open");
    if (!devices || n_devices < 1 || n_devices > SEMA_MAX_SHARDS) return Status::Err(SEMA_ERR_INVALID, "between 1 and SEMA_MAX_SHARDS devices");
    if (n_devices > 1 && capacity_rows == 0)
        return Status::Err(SEMA_ERR_INVALID, "capacity_rows must be given to split the table over several GPUs");
    // like the reference's table, the index grows as chunks arrive: capacity_rows bounds the address
    // space only (0 = as many rows as 32-bit row ids allow); a driver without virtual memory
    // management gets the fixed-capacity index instead
    const uint64_t max_rows = capacity_rows ? capacity_rows : 0xfffffffeull;
    const uint64_t per = n_devices == 1 ? max_rows : (capacity_rows + n_devices - 1) / n_devices;
    if (n_devices > 1 && per * n_devices >= 0xfffffffeull) return Status::Err(SEMA_ERR_INVALID, "capacity_rows above the 32-bit row-id range");
    // Ranking metric.  The reference ranks by LanceDB's default squared-L2 `_distance`
    // (lance_indexer.rs:121-126 sets no distance_type).  With normalize (the reference's own
    // pipeline: every row and query passes through the mean_pool normalise tail) all vectors are
    // unit-norm, L2^2 = 2 - 2 cos, and the dot-product kernel gives the same order and the cosine
    // directly.  Without it the caller's vectors may have any norm, where only the literal L2
    // metric ranks like the reference: the index is then created with SEMA_METRIC_L2 and the
    // score handed out is 1 - d/2 (= the cosine whenever the vectors do happen to be unit-norm).
    metric_ = normalize ? SEMA_METRIC_COSINE : SEMA_METRIC_L2;
    Status err = Status::Ok();
    shards_.resize(n_devices);
    for (uint32_t g = 0; g < n_devices && err.ok(); ++g) {
        Shard &sh = shards_[g];
        sh.base = n_devices == 1 ? 0 : (uint64_t)g * per;
        sh.cap = per;
        int rc = sema_index_create_growable(devices[g], dim, per, metric_, &sh.idx);
        if (rc == SEMA_ERR_UNSUPPORTED && capacity_rows) rc = sema_index_create(devices[g], dim, per, metric_, &sh.idx);
        if (rc == SEMA_OK && sh.base) rc = sema_index_set_row_base(sh.idx, sh.base);
        if (rc) { err = from_rc(rc); break; }
        sema_index_set_normalize_queries(sh.idx, normalize ? 1 : 0);
    }
    if (err.ok() && n_devices > 1) {
        std::vector<sema_index *> ptrs;
        for (Shard &sh : shards_) ptrs.push_back(sh.idx);
        const int rc = sema_shard_group_create_local(ptrs.data(), n_devices, &group_);
        if (rc) err = from_rc(rc);
    }
    if (!err.ok()) {
        for (Shard &sh : shards_)
            if (sh.idx) sema_index_destroy(sh.idx);
        shards_.clear();
        return err;
    }
    per_ = per;
    dim_ = dim;
    normalize_ = normalize;
    return Status::Ok();
}

bool GpuVectorIndexer::locate(uint64_t row, size_t *shard, uint64_t *local) const
{
    if (shards_.empty()) return false;
    const size_t g = shards_.size() == 1 ? 0 : (size_t)(row / per_);
    if (g >= shards_.size() || row < shards_[g].base) return false;
    const uint64_t l = row - shards_[g].base;
    if (l >= shards_[g].chunks.size()) return false;
    *shard = g;
    *local = l;
    return true;
}

const Chunk *GpuVectorIndexer::chunk(uint64_t row) const
{
    size_t g;
    uint64_t l;
    return locate(row, &g, &l) ? &shards_[g].chunks[l] : nullptr;
}

uint64_t GpuVectorIndexer::len() const
{
    uint64_t n = 0;
    for (const Shard &sh : shards_) n += sh.chunks.size();
    return n;
}

Status GpuVectorIndexer::index_chunks(const std::vector<Chunk> &chunks, const float *vectors, const uint8_t *valid)
{
    if (chunks.empty()) return Status::Ok();  // lance_indexer.rs:31-33
    if (shards_.empty()) return Status::Err(SEMA_ERR_INVALID, "indexer not open");
    // table order = shard order: fill the first shard that still has room, then the next one
    size_t done = 0;
    for (Shard &sh : shards_) {
        if (done == chunks.size()) break;
        const uint64_t room = sh.cap > sh.chunks.size() ? sh.cap - sh.chunks.size() : 0;
        if (room == 0) continue;
        const size_t m = (size_t)std::min<uint64_t>(room, chunks.size() - done);
        uint64_t first = 0;
        int rc = sema_index_append(sh.idx, vectors + done * dim_, m, valid ? valid + done : nullptr, normalize_ ? 1 : 0, &first);
        if (rc) return from_rc(rc);
        if (first != sh.chunks.size()) return Status::Err(SEMA_ERR_INVALID, "row table out of step with the GPU index");
        sh.chunks.insert(sh.chunks.end(), chunks.begin() + (std::ptrdiff_t)done, chunks.begin() + (std::ptrdiff_t)(done + m));
        sh.live.insert(sh.live.end(), m, 1);
        done += m;
    }
    if (done != chunks.size()) return Status::Err(SEMA_ERR_CAPACITY, "the table is full (capacity_rows)");
    return Status::Ok();
}

Status GpuVectorIndexer::index_chunks(const std::vector<Chunk> &chunks, const Embedder &embed)
{
    if (chunks.empty()) return Status::Ok();
    if (!embed) return Status::Err(SEMA_ERR_INVALID, "no embedder set");
    // lance_indexer.rs:59-73: one embedder, chunks embedded sequentially; a failure is a null vector
    std::vector<float> vec((size_t)chunks.size() * dim_, 0.0f);
    std::vector<uint8_t> valid(chunks.size(), 0);
    for (size_t i = 0; i < chunks.size(); ++i) {
        auto e = embed(chunks[i].content);
        if (e && e->size() == dim_) {
            std::memcpy(&vec[i * dim_], e->data(), dim_ * sizeof(float));
            valid[i] = 1;
        }
    }
    return index_chunks(chunks, vec.data(), valid.data());
}

Status GpuVectorIndexer::search(const float *q, size_t limit, std::vector<std::pair<Chunk, float>> *out,
                                std::vector<uint64_t> *rows)
{
    out->clear();
    if (rows) rows->clear();
    if (shards_.empty()) return Status::Ok();  // no table yet => Ok(empty), lance_indexer.rs:108-111
    if (limit > SEMA_MAX_K) return Status::Err(SEMA_ERR_INVALID, "limit above SEMA_MAX_K");
    ids_buf_.resize(limit ? limit : 1);
    sc_buf_.resize(limit ? limit : 1);
    uint32_t nf = 0;
    int rc;
    if (group_) {
        if (limit == 0) return Status::Ok();
        rc = sema_shard_group_search(group_, q, (uint32_t)limit, ids_buf_.data(), sc_buf_.data(), &nf);
    } else {
        rc = sema_index_search(shards_[0].idx, q, (uint32_t)limit, ids_buf_.data(), sc_buf_.data(), &nf);
    }
    if (rc) return from_rc(rc);
    out->reserve(nf);
    for (uint32_t i = 0; i < nf; ++i) {  // lance_indexer.rs:131-138: rows -> Chunk, in rank order
        const uint64_t row = ids_buf_[i];
        const Chunk *c = chunk(row);
        if (!c) return Status::Err(SEMA_ERR_INVALID, "GPU returned a row outside the chunk table");
        // the real score (mod.rs:123 attaches 1.0): the cosine, or 1 - d/2 under the literal L2 metric
        const float score = metric_ == SEMA_METRIC_L2 ? 1.0f - 0.5f * sc_buf_[i] : sc_buf_[i];
        out->emplace_back(*c, score);
        if (rows) rows->push_back(row);
    }
    return Status::Ok();
}

Status GpuVectorIndexer::search_like(const std::string &query, size_t limit, std::vector<std::pair<Chunk, float>> *out,
                                     std::vector<uint64_t> *rows)
{
    out->clear();
    if (rows) rows->clear();
    for (const Shard &sh : shards_)
        for (uint64_t r = 0; r < sh.chunks.size() && out->size() < limit; ++r) {
            if (!sh.live[r]) continue;
            if (like_contains(sh.chunks[r].content, query)) {
                out->emplace_back(sh.chunks[r], 1.0f);  // mod.rs:123
                if (rows) rows->push_back(sh.base + r);
            }
        }
    return Status::Ok();
}

Status GpuVectorIndexer::remove_file_chunks(const std::string &file_path, uint64_t *removed)
{
    if (removed) *removed = 0;
    uint64_t total = 0;
    for (Shard &sh : shards_) {  // no table => nothing to delete (lance_indexer.rs:235)
        std::vector<uint64_t> dead;
        for (uint64_t r = 0; r < sh.chunks.size(); ++r)
            if (sh.live[r] && sh.chunks[r].file_path == file_path) dead.push_back(r);
        if (dead.empty()) continue;
        int rc = sema_index_tombstone(sh.idx, dead.data(), dead.size());
        if (rc) return from_rc(rc);
        for (uint64_t r : dead) sh.live[r] = 0;
        total += dead.size();
    }
    if (removed) *removed = total;
    return Status::Ok();
}

Status GpuVectorIndexer::compact(uint64_t *n_live)
{
    uint64_t live_total = 0;
    for (Shard &sh : shards_) {
        if (sh.chunks.empty()) continue;
        std::vector<uint64_t> map(sh.chunks.size());
        uint64_t live = 0;
        // keep every chunk that was not removed — also those whose embedding failed: they never match a
        // vector query but the reference's LIKE fallback still finds them
        int rc = sema_index_compact_keep(sh.idx, sh.live.data(), map.data(), &live);
        if (rc) return from_rc(rc);
        std::vector<Chunk> kept;
        kept.reserve(live);
        for (uint64_t r = 0; r < sh.chunks.size(); ++r)
            if (map[r] != ~0ull) kept.push_back(std::move(sh.chunks[r]));   // the map is order preserving
        if (kept.size() != live) return Status::Err(SEMA_ERR_INVALID, "compaction map out of step with the chunk table");
        sh.chunks = std::move(kept);
        sh.live.assign(sh.chunks.size(), 1);
        live_total += live;
    }
    if (n_live) *n_live = live_total;
    return Status::Ok();
}

// ------------------------------------------------------------------ StorageManager
Status StorageManager::search(const std::string &query_in, size_t limit, std::vector<std::pair<Chunk, float>> *out,
                              std::vector<uint64_t> *rows)
{
    out->clear();
    if (rows) rows->clear();
    const std::string query = trim(query_in);  // mod.rs:113
    if (!query.empty() && query[0] == '\'') {  // mod.rs:115-120: the Tantivy keyword route
        if (query.size() == 1) return Status::Ok();
        return Status::Err(SEMA_ERR_UNSUPPORTED,
                           "keyword search (\"'\" prefix, src/storage/text_indexer.rs) is outside this path");
    }
    if (lance_indexer.len() == 0) return Status::Ok();  // no table => Ok(empty), lance_indexer.rs:108-111
    std::optional<std::vector<float>> emb;
    if (embedder_) emb = embedder_(query);  // lance_indexer.rs:113-118
    if (emb && emb->size() == lance_indexer.dim())
        return lance_indexer.search(emb->data(), limit, out, rows);
    return lance_indexer.search_like(query, limit, out, rows);  // lance_indexer.rs:143-162
}

Status StorageManager::execute_search(const std::string &query, std::vector<SearchResult> *out)
{
    out->clear();
    std::vector<std::pair<Chunk, float>> hits;
    std::vector<uint64_t> rows;
    Status st = search(query, SEARCH_RESULTS_LIMIT, &hits, &rows);  // engine.rs:125-126
    if (!st.ok()) return st;  // engine.rs:147-149: surfaces as "Search failed: {e}"
    std::vector<SearchResult> results;
    results.reserve(hits.size());
    for (size_t i = 0; i < hits.size(); ++i) {  // engine.rs:128-135
        SearchResult r;
        r.chunk = std::move(hits[i].first);
        r.score = hits[i].second;
        r.total_matches_in_file = 1;
        r.row = rows[i];
        results.push_back(std::move(r));
    }
    *out = group_results_by_file(std::move(results));  // engine.rs:137
    return Status::Ok();
}

}  // namespace sema_host

// ====================================================================== C ABI
using namespace sema_host;

struct sema_store {
    StorageManager mgr;
    sema_embed_fn fn = nullptr;
    void *user = nullptr;
};

namespace {
thread_local std::string g_store_err;
int store_fail(const Status &st)
{
    g_store_err = st.message;
    return st.code;
}
int store_fail(int code, const char *msg)
{
    g_store_err = msg;
    return code;
}
std::vector<Chunk> make_chunks(uint64_t n, const char *const *ids, const char *const *paths, const uint64_t *sl,
                               const uint64_t *el, const char *const *contents)
{
    std::vector<Chunk> v(n);
    for (uint64_t i = 0; i < n; ++i) {
        v[i].id = ids[i];
        v[i].file_path = paths[i];
        v[i].start_line = (size_t)sl[i];
        v[i].end_line = (size_t)el[i];
        v[i].content = contents[i];
    }
    return v;
}
}  // namespace

extern "C" {

const char *sema_store_last_error(void) { return g_store_err.c_str(); }

int sema_store_create(int device, uint32_t dim, uint64_t capacity_rows, int normalize, sema_store **out)
{
    if (!out) return store_fail(SEMA_ERR_INVALID, "null out");
    *out = nullptr;
    sema_store *st = new (std::nothrow) sema_store();
    if (!st) return store_fail(SEMA_ERR_NOMEM, "host allocation failed");
    Status s = st->mgr.open(device, dim, capacity_rows, normalize != 0);
    if (!s.ok()) {
        delete st;
        return store_fail(s);
    }
    *out = st;
    return SEMA_OK;
}

int sema_store_create_multi(const int *devices, uint32_t n_devices, uint32_t dim, uint64_t capacity_rows, int normalize,
                            sema_store **out)
{
    if (!out) return store_fail(SEMA_ERR_INVALID, "null out");
    *out = nullptr;
    sema_store *st = new (std::nothrow) sema_store();
    if (!st) return store_fail(SEMA_ERR_NOMEM, "host allocation failed");
    Status s = st->mgr.open_multi(devices, n_devices, dim, capacity_rows, normalize != 0);
    if (!s.ok()) {
        delete st;
        return store_fail(s);
    }
    *out = st;
    return SEMA_OK;
}

int sema_store_destroy(sema_store *st)
{
    delete st;
    return SEMA_OK;
}

int sema_store_set_embedder(sema_store *st, sema_embed_fn fn, void *user)
{
    if (!st) return store_fail(SEMA_ERR_INVALID, "null store");
    st->fn = fn;
    st->user = user;
    if (!fn) {
        st->mgr.set_embedder(nullptr);
        return SEMA_OK;
    }
    const uint32_t dim = st->mgr.lance_indexer.dim();
    st->mgr.set_embedder([st, dim](const std::string &text) -> std::optional<std::vector<float>> {
        std::vector<float> v(dim);
        if (st->fn(st->user, text.c_str(), v.data(), dim) != 0) return std::nullopt;
        return v;
    });
    return SEMA_OK;
}

int sema_store_index_chunks(sema_store *st, uint64_t n, const char *const *ids, const char *const *file_paths,
                            const uint64_t *start_lines, const uint64_t *end_lines, const char *const *contents,
                            const float *vectors, const uint8_t *valid)
{
    if (!st) return store_fail(SEMA_ERR_INVALID, "null store");
    if (n == 0) return SEMA_OK;
    if (!ids || !file_paths || !start_lines || !end_lines || !contents || !vectors)
        return store_fail(SEMA_ERR_INVALID, "null column");
    Status s = st->mgr.lance_indexer.index_chunks(make_chunks(n, ids, file_paths, start_lines, end_lines, contents),
                                                  vectors, valid);
    return s.ok() ? SEMA_OK : store_fail(s);
}

int sema_store_index_chunks_embed(sema_store *st, uint64_t n, const char *const *ids, const char *const *file_paths,
                                  const uint64_t *start_lines, const uint64_t *end_lines,
                                  const char *const *contents)
{
    if (!st) return store_fail(SEMA_ERR_INVALID, "null store");
    if (n == 0) return SEMA_OK;
    if (!ids || !file_paths || !start_lines || !end_lines || !contents) return store_fail(SEMA_ERR_INVALID, "null column");
    Status s = st->mgr.index_chunks(make_chunks(n, ids, file_paths, start_lines, end_lines, contents));
    return s.ok() ? SEMA_OK : store_fail(s);
}

static int emit_hits(const std::vector<std::pair<Chunk, float>> &hits, const std::vector<uint64_t> &rows,
                     sema_hit *out, uint32_t *n_found)
{
    for (size_t i = 0; i < hits.size(); ++i) {
        out[i].row = rows[i];
        out[i].score = hits[i].second;
    }
    *n_found = (uint32_t)hits.size();
    return SEMA_OK;
}

int sema_store_search_vector(sema_store *st, const float *q, uint32_t limit, sema_hit *hits, uint32_t *n_found)
{
    if (!st || !q || !n_found || (limit && !hits)) return store_fail(SEMA_ERR_INVALID, "null argument");
    std::vector<std::pair<Chunk, float>> res;
    std::vector<uint64_t> rows;
    Status s = st->mgr.lance_indexer.search(q, limit, &res, &rows);
    if (!s.ok()) return store_fail(s);
    return emit_hits(res, rows, hits, n_found);
}

int sema_store_search(sema_store *st, const char *query, uint32_t limit, sema_hit *hits, uint32_t *n_found)
{
    if (!st || !query || !n_found || (limit && !hits)) return store_fail(SEMA_ERR_INVALID, "null argument");
    std::vector<std::pair<Chunk, float>> res;
    std::vector<uint64_t> rows;
    Status s = st->mgr.search(query, limit, &res, &rows);
    if (!s.ok()) return store_fail(s);
    return emit_hits(res, rows, hits, n_found);
}

static int emit_grouped(const std::vector<SearchResult> &g, sema_search_result *out, uint32_t cap, uint32_t *n_out)
{
    const size_t n = g.size() < cap ? g.size() : cap;
    for (size_t i = 0; i < n; ++i) {
        out[i].row = g[i].row;
        out[i].score = g[i].score;
        out[i].total_matches_in_file = g[i].total_matches_in_file;
    }
    *n_out = (uint32_t)n;
    return SEMA_OK;
}

int sema_store_execute_search(sema_store *st, const char *query, sema_search_result *out, uint32_t cap,
                              uint32_t *n_out)
{
    if (!st || !query || !n_out || (cap && !out)) return store_fail(SEMA_ERR_INVALID, "null argument");
    std::vector<SearchResult> g;
    Status s = st->mgr.execute_search(query, &g);
    if (!s.ok()) return store_fail(s);
    return emit_grouped(g, out, cap, n_out);
}

int sema_store_group_results_by_file(sema_store *st, const sema_hit *hits, uint32_t n, sema_search_result *out,
                                     uint32_t cap, uint32_t *n_out)
{
    if (!st || !n_out || (n && !hits) || (cap && !out)) return store_fail(SEMA_ERR_INVALID, "null argument");
    std::vector<SearchResult> rs;
    rs.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        const Chunk *c = st->mgr.lance_indexer.chunk(hits[i].row);
        if (!c) return store_fail(SEMA_ERR_INVALID, "hit row outside the chunk table");
        SearchResult r;
        r.chunk = *c;
        r.score = hits[i].score;
        r.total_matches_in_file = 1;
        r.row = hits[i].row;
        rs.push_back(std::move(r));
    }
    return emit_grouped(group_results_by_file(std::move(rs)), out, cap, n_out);
}

int sema_store_remove_file_chunks(sema_store *st, const char *file_path, uint64_t *removed)
{
    if (!st || !file_path) return store_fail(SEMA_ERR_INVALID, "null argument");
    Status s = st->mgr.lance_indexer.remove_file_chunks(file_path, removed);
    return s.ok() ? SEMA_OK : store_fail(s);
}

int sema_store_compact(sema_store *st, uint64_t *n_live)
{
    if (!st) return store_fail(SEMA_ERR_INVALID, "null store");
    Status s = st->mgr.lance_indexer.compact(n_live);
    return s.ok() ? SEMA_OK : store_fail(s);
}

int sema_store_chunk(const sema_store *st, uint64_t row, const char **id, const char **file_path,
                     uint64_t *start_line, uint64_t *end_line, const char **content)
{
    if (!st) return store_fail(SEMA_ERR_INVALID, "null store");
    const Chunk *c = st->mgr.lance_indexer.chunk(row);
    if (!c) return store_fail(SEMA_ERR_INVALID, "row outside the chunk table");
    if (id) *id = c->id.c_str();
    if (file_path) *file_path = c->file_path.c_str();
    if (start_line) *start_line = c->start_line;
    if (end_line) *end_line = c->end_line;
    if (content) *content = c->content.c_str();
    return SEMA_OK;
}

int sema_group_results_by_file(uint32_t n, const char *const *file_paths, const uint64_t *start_lines,
                               const float *scores, uint32_t *rep_index, uint64_t *totals, uint32_t *n_groups)
{
    if (!n_groups || (n && (!file_paths || !start_lines || !scores || !rep_index || !totals)))
        return store_fail(SEMA_ERR_INVALID, "null argument");
    std::vector<SearchResult> rs(n);
    for (uint32_t i = 0; i < n; ++i) {
        rs[i].chunk.file_path = file_paths[i];
        rs[i].chunk.start_line = (size_t)start_lines[i];
        rs[i].score = scores[i];
        rs[i].total_matches_in_file = 1;
        rs[i].row = i;
    }
    const std::vector<SearchResult> g = group_results_by_file(std::move(rs));
    for (size_t i = 0; i < g.size(); ++i) {
        rep_index[i] = (uint32_t)g[i].row;
        totals[i] = g[i].total_matches_in_file;
    }
    *n_groups = (uint32_t)g.size();
    return SEMA_OK;
}

int sema_like_contains(const char *content, const char *needle)
{
    if (!content || !needle) return 0;
    return like_contains(content, needle) ? 1 : 0;
}

uint64_t sema_store_len(const sema_store *st) { return st ? st->mgr.lance_indexer.len() : 0; }
sema_index *sema_store_index(sema_store *st) { return st ? st->mgr.lance_indexer.index() : nullptr; }

}  // extern "C"

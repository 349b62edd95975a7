// api_shard.cu — corpus-sharded group: scan + peer exchange + merge in one kernel per rank.
#include "index_impl.cuh"
#include "k2_scan.cuh"

using namespace sema;
using namespace sema_impl;

extern "C" {

// ---- corpus-sharded group with the fused peer exchange ---------------------------------
struct sema_shard_group {
    sema_index *idx = nullptr;
    uint32_t world = 1, rank = 0;
    uint64_t *xbuf = nullptr;                  // own exchange buffer (cudaMalloc: IPC-exportable)
    uint64_t *peer[XCHG_MAX_WORLD] = {nullptr};
    bool opened[XCHG_MAX_WORLD] = {false};
    bool connected = false;
    uint64_t seq = 0;
};

static size_t xbuf_bytes(uint32_t world) { return ((size_t)2 * world * XCHG_KEYS + (size_t)2 * world) * sizeof(uint64_t); }

int sema_shard_group_create(sema_index *idx, uint32_t world, uint32_t rank, sema_shard_group **out)
{
    if (!idx || !out) return fail(SEMA_ERR_INVALID, "null argument");
    *out = nullptr;
    if (world < 1 || world > (uint32_t)XCHG_MAX_WORLD || rank >= world)
        return fail(SEMA_ERR_INVALID, "world %u / rank %u outside [1, %d]", world, rank, XCHG_MAX_WORLD);
    CK(cudaSetDevice(idx->device));
    sema_shard_group *g = new (std::nothrow) sema_shard_group();
    if (!g) return fail(SEMA_ERR_NOMEM, "host allocation failed");
    g->idx = idx;
    g->world = world;
    g->rank = rank;
    cudaError_t e = cudaMalloc(&g->xbuf, xbuf_bytes(world));
    if (e == cudaSuccess) e = cudaMemset(g->xbuf, 0, xbuf_bytes(world));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(g->xbuf);
        delete g;
        return fail(SEMA_ERR_CUDA, "exchange buffer: %s", cudaGetErrorString(e));
    }
    g->peer[rank] = g->xbuf;
    g->connected = world == 1;
    *out = g;
    return SEMA_OK;
}

int sema_shard_group_local_handle(sema_shard_group *g, void *handle_out)
{
    if (!g || !handle_out) return fail(SEMA_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SEMA_IPC_HANDLE_BYTES, "IPC handle size");
    CK(cudaSetDevice(g->idx->device));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, g->xbuf));
    memcpy(handle_out, &h, sizeof h);
    return SEMA_OK;
}

int sema_shard_group_connect(sema_shard_group *g, const void *handles)
{
    if (!g || !handles) return fail(SEMA_ERR_INVALID, "null argument");
    CK(cudaSetDevice(g->idx->device));
    for (uint32_t r = 0; r < g->world; ++r) {
        if (r == g->rank || g->opened[r]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const unsigned char *>(handles) + (size_t)r * SEMA_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        g->peer[r] = static_cast<uint64_t *>(p);
        g->opened[r] = true;
    }
    g->connected = true;
    return SEMA_OK;
}

static void group_exchange(const sema_shard_group *g, Exchange &x)
{
    memset(&x, 0, sizeof x);
    for (uint32_t r = 0; r < g->world; ++r) x.peer[r] = g->peer[r];
    x.world = g->world;
    x.rank = g->rank;
}

// An empty shard still takes part in the exchange: with n = 0 the kernel scans nothing (one block)
// and goes straight to publish / wait / merge.
static int group_check(sema_shard_group *g, uint32_t k)
{
    if (!g->connected) return fail(SEMA_ERR_INVALID, "shard group not connected");
    if (k == 0 || k > (uint32_t)K_PASS) return fail(SEMA_ERR_UNSUPPORTED, "fused shard search covers 1 <= k <= %d", K_PASS);
    return SEMA_OK;
}

int sema_shard_group_search_stream_device(sema_shard_group *g, const float *Q_dev, uint32_t nq, uint32_t k,
                                          uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev)
{
    if (!g || !Q_dev || !ids_dev || !scores_dev || !n_found_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (nq == 0) return fail(SEMA_ERR_INVALID, "nq = 0");
    int rc = group_check(g, k);
    if (rc) return rc;
    sema_index *s = g->idx;
    CK(cudaSetDevice(s->device));
    rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    Exchange x;
    group_exchange(g, x);
    return scan_stream(s, Q_dev, nq, (uint32_t)n, k, ids_dev, scores_dev, n_found_dev, &x, &g->seq);
}

int sema_shard_group_search_device(sema_shard_group *g, const float *q_dev, uint32_t k, uint64_t *ids_dev,
                                   float *scores_dev, uint32_t *n_found_dev)
{
    return sema_shard_group_search_stream_device(g, q_dev, 1, k, ids_dev, scores_dev, n_found_dev);
}

int sema_shard_group_search(sema_shard_group *g, const float *q, uint32_t k, uint64_t *row_ids, float *scores,
                            uint32_t *n_found)
{
    if (!g || !q || !row_ids || !scores || !n_found) return fail(SEMA_ERR_INVALID, "null argument");
    int rc = group_check(g, k);
    if (rc) return rc;
    sema_index *s = g->idx;
    CK(cudaSetDevice(s->device));
    rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    const bool fast = host_query_ok(s, k);
    // refuse before the sequence number moves: a rank that skips a search would leave its peers waiting
    if (fast && s->slots[(s->host_seq + 1) % RES_SLOTS].ticket != 0)
        return fail(SEMA_ERR_INVALID, "%d searches are already in flight on this handle: collect one first", RES_SLOTS);
    Exchange x;
    group_exchange(g, x);
    x.seq = ++g->seq;
    if (fast) return host_query_run(s, q, (uint32_t)n, k, &x, row_ids, scores, n_found);
    return staged_query_run(s, q, (uint32_t)n, k, &x, row_ids, scores, n_found);
}

int sema_shard_group_search_submit(sema_shard_group *g, const float *q, uint32_t k, uint64_t *ticket)
{
    if (!g || !q || !ticket) return fail(SEMA_ERR_INVALID, "null argument");
    int rc = group_check(g, k);
    if (rc) return rc;
    sema_index *s = g->idx;
    CK(cudaSetDevice(s->device));
    rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    rc = slot_claim(s, k, ticket);
    if (rc) return rc;
    sema_index::Slot &sl = s->slots[*ticket % RES_SLOTS];
    if (host_query_ok(s, k)) {
        Exchange x;
        group_exchange(g, x);
        x.seq = ++g->seq;
        rc = host_query_launch(s, q, (uint32_t)n, k, &x, *ticket);
    } else {
        sl.ticket = 0;                                  // the synchronous call claims its own slot if it needs one
        sl.ids.resize(k);
        sl.sc.resize(k);
        rc = sema_shard_group_search(g, q, k, sl.ids.data(), sl.sc.data(), &sl.nf);
        sl.ticket = *ticket;
        sl.sync_done = true;
    }
    if (rc) sl.ticket = 0;
    return rc;
}

int sema_shard_group_search_collect(sema_shard_group *g, uint64_t ticket, uint64_t *row_ids, float *scores,
                                    uint32_t *n_found)
{
    if (!g || !row_ids || !scores || !n_found) return fail(SEMA_ERR_INVALID, "null argument");
    CK(cudaSetDevice(g->idx->device));
    return slot_collect(g->idx, ticket, row_ids, scores, n_found);
}

int sema_shard_group_destroy(sema_shard_group *g)
{
    if (!g) return SEMA_OK;
    cudaSetDevice(g->idx->device);
    cudaStreamSynchronize(g->idx->stream);
    for (uint32_t r = 0; r < g->world; ++r)
        if (g->opened[r]) cudaIpcCloseMemHandle(g->peer[r]);
    cudaFree(g->xbuf);
    cudaGetLastError();
    delete g;
    return SEMA_OK;
}

}  // extern "C"

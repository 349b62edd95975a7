// api_shard.cu — corpus-sharded group: scan + peer exchange + merge in one kernel per rank.
#include "index_impl.cuh"
#include "k2_scan.cuh"

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

using namespace sema;
using namespace sema_impl;

// ---- fan-out workers of the single-process group ------------------------------------------------
// One host thread per shard, bound to that shard's device.  A group call posts one command; every
// worker runs it against its own shard (all N kernel launches leave the host within ~1 us of each
// other instead of one after the other: the peers wait for the slowest rank inside the kernel, so
// launch skew is search latency), then the caller waits for rank 0's result.  Workers spin briefly
// for the next command (a search service issues them back to back) and then sleep on a condition
// variable, so an idle group costs no CPU.
namespace {

enum : int { CMD_NONE = 0, CMD_SUBMIT = 1, CMD_SEARCH = 2, CMD_COLLECT = 3, CMD_EXIT = 4 };

struct LocalShared {
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> posted{0};        // commands posted so far
    std::atomic<int> sleepers{0};
    int cmd = CMD_NONE;
    const float *q = nullptr;
    uint32_t k = 0;
    uint64_t ticket = 0;                    // CMD_COLLECT
};

struct LocalWorker {
    std::thread th;
    std::atomic<uint64_t> done{0};          // commands this worker has finished
    int rc = SEMA_OK;
    char err[256] = "";
    uint64_t ticket = 0;                    // CMD_SUBMIT: the ticket this shard issued
    uint32_t nf = 0;                        // CMD_SEARCH / CMD_COLLECT: this shard's copy of the (global) result
    std::vector<uint64_t> ids;
    std::vector<float> sc;
};

inline void cpu_relax()
{
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}

}  // namespace

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
    // single-process group (sema_shard_group_create_local): this object is the front; member[r] is an
    // ordinary group object for shard r (world = n_local, rank = r, peers wired by direct peer access)
    bool local = false;
    uint32_t n_local = 0;
    sema_shard_group *member[XCHG_MAX_WORLD] = {nullptr};
    LocalShared *shared = nullptr;
    LocalWorker *worker[XCHG_MAX_WORLD] = {nullptr};
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

}  // extern "C"

namespace {

void group_exchange(const sema_shard_group *g, Exchange &x);
int group_check(sema_shard_group *g, uint32_t k);

// what one shard does for one command (runs on that shard's worker thread, device already bound)
void local_run(sema_shard_group *front, uint32_t r, int cmd, const float *q, uint32_t k, uint64_t collect_ticket)
{
    sema_shard_group *m = front->member[r];
    LocalWorker *w = front->worker[r];
    sema_index *s = m->idx;
    w->rc = SEMA_OK;
    if (cmd == CMD_COLLECT) {
        w->rc = slot_collect(s, collect_ticket, w->ids.data(), w->sc.data(), &w->nf);
    } else {
        int rc = poll_ingest(s, false);
        const uint64_t n = s->n_visible;
        s->last_snapshot = n;
        Exchange x;
        group_exchange(m, x);
        if (rc == SEMA_OK && cmd == CMD_SUBMIT) {
            rc = slot_claim(s, k, &w->ticket);
            if (rc == SEMA_OK) {
                x.seq = ++m->seq;
                rc = host_query_launch(s, q, (uint32_t)n, k, &x, w->ticket);
                if (rc) s->slots[w->ticket % RES_SLOTS].ticket = 0;
            }
        } else if (rc == SEMA_OK) {           // CMD_SEARCH: shapes outside the one-launch path, whole search per shard
            x.seq = ++m->seq;
            rc = staged_query_run(s, q, (uint32_t)n, k, &x, w->ids.data(), w->sc.data(), &w->nf);
        }
        w->rc = rc;
    }
    if (w->rc) snprintf(w->err, sizeof w->err, "shard %u: %s", r, sema_last_error());
}

void local_worker_main(sema_shard_group *front, uint32_t r)
{
    LocalShared *sh = front->shared;
    LocalWorker *w = front->worker[r];
    cudaSetDevice(front->member[r]->idx->device);
    uint64_t seen = 0;
    for (;;) {
        // wait for command seen + 1: spin first (back-to-back searches), then sleep
        uint32_t spins = 0;
        while (sh->posted.load(std::memory_order_acquire) == seen) {
            if (++spins < 20000) { cpu_relax(); continue; }
            std::unique_lock<std::mutex> lk(sh->mu);
            sh->sleepers.fetch_add(1, std::memory_order_seq_cst);
            sh->cv.wait(lk, [&] { return sh->posted.load(std::memory_order_acquire) != seen; });
            sh->sleepers.fetch_sub(1, std::memory_order_seq_cst);
        }
        ++seen;
        const int cmd = sh->cmd;
        if (cmd == CMD_EXIT) { w->done.store(seen, std::memory_order_release); return; }
        local_run(front, r, cmd, sh->q, sh->k, sh->ticket);
        w->done.store(seen, std::memory_order_release);
    }
}

// post one command to every worker and wait until all of them have run it; first failure wins
int local_post(sema_shard_group *g, int cmd, const float *q, uint32_t k, uint64_t ticket)
{
    LocalShared *sh = g->shared;
    sh->cmd = cmd;
    sh->q = q;
    sh->k = k;
    sh->ticket = ticket;
    const uint64_t seq = sh->posted.load(std::memory_order_relaxed) + 1;
    sh->posted.store(seq, std::memory_order_seq_cst);
    if (sh->sleepers.load(std::memory_order_seq_cst) > 0) {
        std::lock_guard<std::mutex> lk(sh->mu);
        sh->cv.notify_all();
    }
    for (uint32_t r = 0; r < g->n_local; ++r)
        while (g->worker[r]->done.load(std::memory_order_acquire) != seq) cpu_relax();
    for (uint32_t r = 0; r < g->n_local; ++r)
        if (g->worker[r]->rc) return fail(g->worker[r]->rc, "%s", g->worker[r]->err);
    return SEMA_OK;
}

void local_teardown(sema_shard_group *g)
{
    if (g->shared) {
        bool any = false;
        for (uint32_t r = 0; r < g->n_local; ++r) any = any || (g->worker[r] && g->worker[r]->th.joinable());
        if (any) {
            LocalShared *sh = g->shared;
            sh->cmd = CMD_EXIT;
            sh->posted.store(sh->posted.load() + 1, std::memory_order_seq_cst);
            { std::lock_guard<std::mutex> lk(sh->mu); sh->cv.notify_all(); }
            for (uint32_t r = 0; r < g->n_local; ++r)
                if (g->worker[r] && g->worker[r]->th.joinable()) g->worker[r]->th.join();
        }
    }
    for (uint32_t r = 0; r < g->n_local; ++r) {
        delete g->worker[r];
        if (g->member[r]) {
            cudaSetDevice(g->member[r]->idx->device);
            cudaStreamSynchronize(g->member[r]->idx->stream);
            cudaFree(g->member[r]->xbuf);
            delete g->member[r];
        }
    }
    delete g->shared;
    cudaGetLastError();
}

}  // namespace

extern "C" {

int sema_shard_group_create_local(sema_index *const *shards, uint32_t n_shards, sema_shard_group **out)
{
    if (!shards || !out) return fail(SEMA_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_shards < 1 || n_shards > (uint32_t)XCHG_MAX_WORLD)
        return fail(SEMA_ERR_INVALID, "n_shards %u outside [1, %d]", n_shards, XCHG_MAX_WORLD);
    for (uint32_t r = 0; r < n_shards; ++r) {
        if (!shards[r]) return fail(SEMA_ERR_INVALID, "shard %u is null", r);
        if (shards[r]->dim != shards[0]->dim || shards[r]->metric != shards[0]->metric)
            return fail(SEMA_ERR_INVALID, "shard %u differs from shard 0 in dim or metric", r);
        // the fused exchange needs all N kernels running at once: one shard per device
        for (uint32_t o = 0; o < r; ++o)
            if (shards[o]->device == shards[r]->device)
                return fail(SEMA_ERR_INVALID, "shards %u and %u share device %d: a local group holds one shard per GPU", o, r,
                            shards[r]->device);
    }
    // every device must reach every other device's exchange buffer (NVLink / NVSwitch peer access)
    for (uint32_t a = 0; a < n_shards; ++a) {
        CK(cudaSetDevice(shards[a]->device));
        for (uint32_t b = 0; b < n_shards; ++b) {
            if (a == b) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, shards[a]->device, shards[b]->device));
            if (!can) return fail(SEMA_ERR_UNSUPPORTED, "device %d cannot access device %d", shards[a]->device, shards[b]->device);
            cudaError_t e = cudaDeviceEnablePeerAccess(shards[b]->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return fail(SEMA_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", shards[a]->device,
                                                   shards[b]->device, cudaGetErrorString(e));
        }
    }
    sema_shard_group *g = new (std::nothrow) sema_shard_group();
    if (!g) return fail(SEMA_ERR_NOMEM, "host allocation failed");
    g->local = true;
    g->n_local = n_shards;
    g->world = n_shards;
    g->idx = shards[0];
    g->connected = true;
    int rc = SEMA_OK;
    for (uint32_t r = 0; r < n_shards && rc == SEMA_OK; ++r) {
        rc = sema_shard_group_create(shards[r], n_shards, r, &g->member[r]);
        if (rc == SEMA_OK) g->member[r]->connected = true;
    }
    if (rc == SEMA_OK) {
        for (uint32_t r = 0; r < n_shards; ++r)
            for (uint32_t o = 0; o < n_shards; ++o) g->member[r]->peer[o] = g->member[o]->xbuf;   // same address space: no IPC
        g->shared = new (std::nothrow) LocalShared();
        if (!g->shared) rc = fail(SEMA_ERR_NOMEM, "host allocation failed");
    }
    for (uint32_t r = 0; r < n_shards && rc == SEMA_OK; ++r) {
        g->worker[r] = new (std::nothrow) LocalWorker();
        if (!g->worker[r]) { rc = fail(SEMA_ERR_NOMEM, "host allocation failed"); break; }
        try {
            g->worker[r]->ids.resize(SEMA_MAX_K);
            g->worker[r]->sc.resize(SEMA_MAX_K);
            g->worker[r]->th = std::thread(local_worker_main, g, r);
        } catch (const std::exception &e) {
            rc = fail(SEMA_ERR_NOMEM, "cannot start the worker of shard %u: %s", r, e.what());
        }
    }
    if (rc != SEMA_OK) {
        local_teardown(g);
        delete g;
        return rc;
    }
    *out = g;
    return SEMA_OK;
}

static void group_exchange_impl(const sema_shard_group *g, Exchange &x)
{
    memset(&x, 0, sizeof x);
    for (uint32_t r = 0; r < g->world; ++r) x.peer[r] = g->peer[r];
    x.world = g->world;
    x.rank = g->rank;
}
}  // extern "C"
namespace {
void group_exchange(const sema_shard_group *g, Exchange &x) { group_exchange_impl(g, x); }
}
extern "C" {

// An empty shard still takes part in the exchange: with n = 0 the kernel scans nothing (one block)
// and goes straight to publish / wait / merge.
static int group_check_impl(sema_shard_group *g, uint32_t k)
{
    if (!g->connected) return fail(SEMA_ERR_INVALID, "shard group not connected");
    if (k == 0 || k > (uint32_t)K_PASS) return fail(SEMA_ERR_UNSUPPORTED, "fused shard search covers 1 <= k <= %d", K_PASS);
    return SEMA_OK;
}
}  // extern "C"
namespace {
int group_check(sema_shard_group *g, uint32_t k) { return group_check_impl(g, k); }

// ---- the single-process group's searches: fan out through the workers, answer from shard 0 ----
int local_search(sema_shard_group *g, const float *q, uint32_t k, uint64_t *row_ids, float *scores, uint32_t *n_found)
{
    bool fast = true;
    for (uint32_t r = 0; r < g->n_local; ++r) fast = fast && host_query_ok(g->member[r]->idx, k);
    if (!fast) {
        int rc = local_post(g, CMD_SEARCH, q, k, 0);
        if (rc) return rc;
        LocalWorker *w = g->worker[0];
        *n_found = w->nf;
        memcpy(row_ids, w->ids.data(), w->nf * sizeof(uint64_t));
        memcpy(scores, w->sc.data(), w->nf * sizeof(float));
        return SEMA_OK;
    }
    int rc = local_post(g, CMD_SUBMIT, q, k, 0);
    if (rc) return rc;
    // every shard holds the global result; the caller gets shard 0's, the others only free their slots
    rc = slot_collect(g->member[0]->idx, g->worker[0]->ticket, row_ids, scores, n_found);
    for (uint32_t r = 1; r < g->n_local; ++r) {
        LocalWorker *w = g->worker[r];
        const int rr = slot_collect(g->member[r]->idx, w->ticket, w->ids.data(), w->sc.data(), &w->nf);
        if (rc == SEMA_OK) rc = rr;
    }
    return rc;
}

int local_submit(sema_shard_group *g, const float *q, uint32_t k, uint64_t *ticket)
{
    for (uint32_t r = 0; r < g->n_local; ++r)
        if (!host_query_ok(g->member[r]->idx, k))
            return fail(SEMA_ERR_UNSUPPORTED, "submit / collect on a local group covers dim 384 / 768 and k <= %d", K_PASS);
    int rc = local_post(g, CMD_SUBMIT, q, k, 0);
    if (rc) return rc;
    for (uint32_t r = 1; r < g->n_local; ++r)
        if (g->worker[r]->ticket != g->worker[0]->ticket)
            return fail(SEMA_ERR_INVALID, "shard %u is out of step (its handle was searched outside the group)", r);
    *ticket = g->worker[0]->ticket;
    return SEMA_OK;
}

int local_collect(sema_shard_group *g, uint64_t ticket, uint64_t *row_ids, float *scores, uint32_t *n_found)
{
    int rc = slot_collect(g->member[0]->idx, ticket, row_ids, scores, n_found);
    for (uint32_t r = 1; r < g->n_local; ++r) {
        LocalWorker *w = g->worker[r];
        const int rr = slot_collect(g->member[r]->idx, ticket, w->ids.data(), w->sc.data(), &w->nf);
        if (rc == SEMA_OK) rc = rr;
    }
    return rc;
}
}  // namespace
extern "C" {

int sema_shard_group_search_stream_device(sema_shard_group *g, const float *Q_dev, uint32_t nq, uint32_t k,
                                          uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev)
{
    if (!g || !Q_dev || !ids_dev || !scores_dev || !n_found_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (nq == 0) return fail(SEMA_ERR_INVALID, "nq = 0");
    if (g->local) return fail(SEMA_ERR_UNSUPPORTED, "device-resident queries live on one GPU: a single-process group takes host queries "
                                                     "(sema_shard_group_search / _submit / _collect)");
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
    if (g->local) return local_search(g, q, k, row_ids, scores, n_found);
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
    if (g->local) return local_submit(g, q, k, ticket);
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
    if (g->local) return local_collect(g, ticket, row_ids, scores, n_found);
    CK(cudaSetDevice(g->idx->device));
    return slot_collect(g->idx, ticket, row_ids, scores, n_found);
}

int sema_shard_group_destroy(sema_shard_group *g)
{
    if (!g) return SEMA_OK;
    if (g->local) {
        local_teardown(g);
        delete g;
        return SEMA_OK;
    }
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

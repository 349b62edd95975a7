// api_search.cu — K2 (scan + fused top-k) and K4 (merge) dispatch; single-query entry points.
#include "index_impl.cuh"
#include "k2_scan.cuh"
#include "k4_merge.cuh"
#include "k2_scan_tma.cuh"
#include "k2_stream.cuh"

using namespace sema;
using namespace sema_impl;

namespace {

// ---- K2 dispatch ---------------------------------------------------------------
struct ScanArgs {
    const float *q_dev;
    uint32_t n, k;
    const uint64_t *bound;
    uint64_t *out_keys;
    uint64_t *res_ids;
    float *res_scores;
    uint32_t *res_nfound;
    const Exchange *x = nullptr;   // sharded mode: fused peer exchange (single pass only)
    unsigned flags = 0;            // SCAN_CHAINED, SCAN_HOST_QUERY (index_impl.cuh)
    uint64_t host_ticket = 0;      // SCAN_HOST_QUERY: value stored to the slot's flag when the results are in place
};

void fill_params(sema_index *s, const ScanArgs &a, ScanParams &p)
{
    p.X = reinterpret_cast<const float4 *>(s->X);
    p.q = a.q_dev;
    p.partials = s->partials;
    p.ticket = s->ticket;
    p.bound = a.bound;
    p.out_keys = a.out_keys;
    p.res_ids = a.res_ids;
    p.res_scores = a.res_scores;
    p.res_nfound = a.res_nfound;
    p.n = a.n;
    p.ld4 = s->ld / 4;
    p.k = a.k;
    p.row_base = s->row_base;
    p.work_ctr = nullptr;
    p.host_flag = nullptr;
    p.host_seq = 0;
    p.normalize_query = 0;
    if (a.x) p.x = *a.x;
    else memset(&p.x, 0, sizeof p.x);
}

template <int NV, int R, int M, int METRIC>
int run_scan(sema_index *s, const ScanArgs &a)
{
    auto kern = scan_topk_kernel<NV, R, M, METRIC>;
    static int occ[64] = {0};  // per device
    const size_t dyn = NV > 0 ? 0 : (size_t)s->ld * sizeof(float);
    int &o = occ[s->device & 63];
    if (o == 0 || NV == 0) {
        int t = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&t, kern, SCAN_THREADS, dyn));
        if (t < 1) return fail(SEMA_ERR_UNSUPPORTED, "scan kernel does not fit on an SM");
        o = t > MAX_BLOCKS_PER_SM ? MAX_BLOCKS_PER_SM : t;
    }
    const uint32_t nb = (a.n + R - 1) / R;
    uint32_t grid = (uint32_t)(s->num_sms * o);
    const uint32_t need = (nb + SCAN_WARPS - 1) / SCAN_WARPS;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    ScanParams p;
    fill_params(s, a, p);
    kern<<<grid, SCAN_THREADS, dyn, s->stream>>>(p);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

// rows reach the SM through a TMA bulk-copy ring (k2_scan_tma.cuh): the default for dim 384 / 768.
// QP: the query travels as a kernel parameter (a.q_dev is then a HOST pointer to dim floats) and the
// results / completion flag go to the handle's mapped host buffer.
template <int NV, int M, int METRIC, bool QP, int CW>
int run_scan_tma_cw(sema_index *s, const ScanArgs &a)
{
    auto kern = scan_topk_tma_kernel<NV, M, METRIC, QP, CW>;
    static bool attr_set[64] = {false};
    if (!attr_set[s->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem_bytes<NV, CW>()));
        attr_set[s->device & 63] = true;
    }
    const uint32_t n_tiles = (a.n + tma_tile_rows<NV, CW>() - 1) / tma_tile_rows<NV, CW>();
    uint32_t grid = (uint32_t)s->num_sms * (CW == 8 ? 1u : 2u);
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    ScanParams p;
    fill_params(s, a, p);
    p.work_ctr = s->ticket + 2 + (s->scan_seq++ & 1);   // alternate: a chained launch overlaps its predecessor
    QueryArg<QP ? NV : 0> qa;
    if constexpr (QP) {
        static_assert(sizeof(ScanParams) + sizeof(qa) <= 4096, "kernel parameter space");
        memcpy(qa.v, a.q_dev, s->dim * sizeof(float));   // ld == dim for these shapes
        p.q = nullptr;
        p.host_flag = reinterpret_cast<uint64_t *>(s->res_map_dev + (a.host_ticket % RES_SLOTS) * RES_SLOT_BYTES + RES_MAP_FLAG_OFF);
        p.host_seq = a.host_ticket;
        p.normalize_query = s->normalize_queries ? 1u : 0u;
    } else {
        qa.unused = 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(tma_threads<CW>());
    cfg.dynamicSmemBytes = tma_smem_bytes<NV, CW>();
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    // only a launch that directly follows another TMA scan of the same call may overlap it; `bound`
    // (multi-pass) is read at kernel start and is produced by the predecessor, so it never chains
    cfg.numAttrs = ((a.flags & SCAN_CHAINED) && !a.bound) ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, p, qa));
    s->launches++;
    return SEMA_OK;
}

// Two other shapes of the same kernel were measured and dropped (the template still takes them): CW = 4 (two CTAs per
// SM, 24 KB stages: 7.35 TB/s against 7.61 at 10 M rows — smaller copies) and CW = 9 (two CTAs per SM, two 48 KB stages
// each: equal at 1 M rows, slower below) — letting a successor launch's block move in early does not shorten a stream.
template <int NV, int M, int METRIC, bool QP>
int run_scan_tma(sema_index *s, const ScanArgs &a)
{
    return run_scan_tma_cw<NV, M, METRIC, QP, 8>(s, a);
}

template <int NV, int METRIC>
int scan_tma_m(sema_index *s, const ScanArgs &a)
{
    if (a.flags & SCAN_HOST_QUERY) {
        if (a.k <= 32) return run_scan_tma<NV, 1, METRIC, true>(s, a);
        if (a.k <= 64) return run_scan_tma<NV, 2, METRIC, true>(s, a);
        return run_scan_tma<NV, 4, METRIC, true>(s, a);
    }
    if (a.k <= 32) return run_scan_tma<NV, 1, METRIC, false>(s, a);
    if (a.k <= 64) return run_scan_tma<NV, 2, METRIC, false>(s, a);
    return run_scan_tma<NV, 4, METRIC, false>(s, a);
}

template <int NV, int R, int METRIC>
int scan_m(sema_index *s, const ScanArgs &a)
{
    if (a.k <= 32) return run_scan<NV, R, 1, METRIC>(s, a);
    if (a.k <= 64) return run_scan<NV, R, 2, METRIC>(s, a);
    return run_scan<NV, R, 4, METRIC>(s, a);
}

template <int METRIC>
int scan_shape(sema_index *s, const ScanArgs &a)
{
    const uint32_t ld4 = s->ld / 4;
    // variants: 0 = default (TMA ring), 1 / 2 / 3 = LDG kernel with R = 4 / 2 / 8 rows per warp batch
    if (ld4 == 96) {
        switch (s->variant) {
            case 1: return scan_m<3, 4, METRIC>(s, a);
            case 2: return scan_m<3, 2, METRIC>(s, a);
            case 3: return scan_m<3, 8, METRIC>(s, a);
            default: return scan_tma_m<3, METRIC>(s, a);
        }
    }
    if (ld4 == 192) {
        switch (s->variant) {
            case 1: return scan_m<6, 4, METRIC>(s, a);
            case 2: return scan_m<6, 2, METRIC>(s, a);
            default: return scan_tma_m<6, METRIC>(s, a);
        }
    }
    return scan_m<0, 4, METRIC>(s, a);
}

// one fused pass (k <= K_PASS)
int scan_pass(sema_index *s, const ScanArgs &a)
{
    return s->metric == SEMA_METRIC_L2 ? scan_shape<METRIC_L2>(s, a) : scan_shape<METRIC_COSINE>(s, a);
}

// ---- K2 for a whole query stream in ONE persistent launch (k2_stream.cuh) ----
constexpr int STREAM_REFUSED = 1;   // internal: the cooperative launch was refused, nothing was launched
struct StreamArgs {
    const float *Q;      // nq x ld floats, device, 16-byte aligned
    uint32_t nq, n, k;
    uint64_t *ids;
    float *sc;
    uint32_t *nf;
    const Exchange *x;   // x->seq = sequence number of query 0
};

template <int NV, int M, int METRIC>
int run_stream(sema_index *s, const StreamArgs &a)
{
    auto kern = scan_stream_kernel<NV, M, METRIC>;
    static bool attr_set[64] = {false};
    if (!attr_set[s->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, stream_smem_bytes<NV, M>()));
        attr_set[s->device & 63] = true;
    }
    const uint32_t n_tiles = (a.n + stream_tile_rows<NV>() - 1) / stream_tile_rows<NV>();
    uint32_t grid = (uint32_t)s->num_sms;
    if (grid > n_tiles) grid = n_tiles;
    if (grid < 1) grid = 1;
    static_assert(2 * 32 * 4 <= MAX_BLOCKS_PER_SM * K_PASS, "two parities of block lists fit the partials buffer");
    StreamParams p;
    p.X = reinterpret_cast<const float4 *>(s->X);
    p.Q = a.Q;
    p.partials = s->partials;
    p.work_ctr = reinterpret_cast<unsigned long long *>(s->stream_ctl);
    p.done = s->stream_ctl + 2;
    p.fault = s->stream_ctl + 3;
    p.ticket = s->stream_ctl + 4;
    p.res_ids = a.ids;
    p.res_scores = a.sc;
    p.res_nfound = a.nf;
    p.n = a.n;
    p.ld4 = s->ld / 4;
    p.k = a.k;
    p.row_base = s->row_base;
    p.nq = a.nq;
    if (a.x) p.x = *a.x;
    else memset(&p.x, 0, sizeof p.x);
    CK(cudaMemsetAsync(s->stream_ctl, 0, STREAM_CTL_WORDS * sizeof(unsigned int), s->stream));
    // The blocks wait for each other (done, tickets), so all of them must be resident at once: a cooperative launch
    // guarantees that or is refused (SM partitioning, MPS limits) — then the caller falls back to one launch per query.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(STREAM_THREADS);
    cfg.dynamicSmemBytes = stream_smem_bytes<NV, M>();
    cfg.stream = s->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorLaunchOutOfResources || e == cudaErrorNotSupported) {
        (void)cudaGetLastError();
        return STREAM_REFUSED;
    }
    CK(e);
    s->launches++;
    return SEMA_OK;
}

template <int NV, int METRIC>
int stream_m(sema_index *s, const StreamArgs &a)
{
    if (a.k <= 32) return run_stream<NV, 1, METRIC>(s, a);
    if (a.k <= 64) return run_stream<NV, 2, METRIC>(s, a);
    return run_stream<NV, 4, METRIC>(s, a);
}

template <int METRIC>
int stream_shape(sema_index *s, const StreamArgs &a)
{
    return s->ld / 4 == 96 ? stream_m<3, METRIC>(s, a) : stream_m<6, METRIC>(s, a);
}

template <int METRIC>
int merge_pass_m(sema_index *s, const uint64_t *keys, uint32_t total, uint32_t k,
                 const uint64_t *bound, uint64_t *out_keys, uint64_t *ids, float *sc, uint32_t *nf)
{
    if (k <= 32)
        merge_topk_kernel<1, METRIC><<<1, MERGE_THREADS, 0, s->stream>>>(keys, total, (int)k, bound, out_keys, ids, sc, nf);
    else if (k <= 64)
        merge_topk_kernel<2, METRIC><<<1, MERGE_THREADS, 0, s->stream>>>(keys, total, (int)k, bound, out_keys, ids, sc, nf);
    else
        merge_topk_kernel<4, METRIC><<<1, MERGE_THREADS, 0, s->stream>>>(keys, total, (int)k, bound, out_keys, ids, sc, nf);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

int decode(sema_index *s, const uint64_t *keys, uint32_t k, uint64_t *ids, float *sc, uint32_t *nf)
{
    if (s->metric == SEMA_METRIC_L2)
        decode_kernel<METRIC_L2><<<1, 256, 0, s->stream>>>(keys, (int)k, ids, sc, nf);
    else
        decode_kernel<METRIC_COSINE><<<1, 256, 0, s->stream>>>(keys, (int)k, ids, sc, nf);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

// Full selection of the best k (any k <= SEMA_MAX_K) for one query that is already on
// the device.  Exactly one of {out_keys} / {res_*} may be null.
}  // namespace

namespace sema_impl {

int scan_query(sema_index *s, const float *q_dev, uint32_t n, uint32_t k, uint64_t *out_keys,
               uint64_t *res_ids, float *res_scores, uint32_t *res_nfound, const Exchange *x, unsigned flags,
               uint64_t host_ticket)
{
    SEMA_NVTX("sema.K2.scan");
    if (k <= K_PASS) {
        ScanArgs a{q_dev, n, k, nullptr, out_keys ? out_keys : s->keys_dev, res_ids, res_scores, res_nfound, x, flags, host_ticket};
        return scan_pass(s, a);
    }
    if (flags & SCAN_HOST_QUERY) return fail(SEMA_ERR_UNSUPPORTED, "the host-query path covers k <= %d", K_PASS);
    if (x) return fail(SEMA_ERR_UNSUPPORTED, "the fused shard exchange covers k <= %d", K_PASS);
    uint64_t *keys = out_keys ? out_keys : s->keys_dev;
    for (uint32_t done = 0; done < k; done += K_PASS) {
        const uint32_t kp = (k - done) < (uint32_t)K_PASS ? (k - done) : (uint32_t)K_PASS;
        ScanArgs a{q_dev, n, kp, done ? keys + done - 1 : nullptr, keys + done, nullptr, nullptr, nullptr};
        int rc = scan_pass(s, a);
        if (rc) return rc;
    }
    if (res_ids) return decode(s, keys, k, res_ids, res_scores, res_nfound);
    return SEMA_OK;
}

// One persistent launch for the whole stream?  Shapes of the TMA kernel only; by default only when one scan is long
// against the finisher warp's per-query work (block merge + the last block's merge over all blocks, one warp: ~19 us
// per query for k <= 16, 50 - 100 us for larger k), so short scans keep the chained launches that finish a query on
// eight warps.
bool stream_kernel_ok(const sema_index *s, uint32_t nq, uint32_t n, uint32_t k)
{
    const uint32_t ld4 = s->ld / 4;
    if (s->stream_mode == 2 || s->variant != 0 || nq < 2 || k < 1 || k > (uint32_t)K_PASS) return false;
    if (s->ld != s->dim || (ld4 != 96 && ld4 != 192)) return false;
    if (s->stream_mode == 1) return true;
    const uint64_t floats = (uint64_t)n * s->ld;
    // measured crossovers (scripts/gpu_r2_stream_thr.sh, dim 384): the finisher chain costs ~19 us per query for k <= 16
    // (persistent wins from 100 k rows: 22.9 against 24.9 us), ~50 - 100 us for the bitonic merges of k = 17 .. 64
    // (loses at 300 k rows, wins at 600 k: 122 against 126 us) and more for k = 65 .. 128 (600 k rows: +1 %, 1 M: +1 - 5 %)
    return floats >= (k <= 16 ? 38000000ull : k <= 64 ? 230000000ull : 300000000ull);
}

int stream_kernel_launch(sema_index *s, const float *Q, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
                         uint32_t *nf_d, const Exchange *x)
{
    SEMA_NVTX("sema.K2.stream");
    const StreamArgs a{Q, nq, n, k, ids_d, sc_d, nf_d, x};
    const int rc = s->metric == SEMA_METRIC_L2 ? stream_shape<METRIC_L2>(s, a) : stream_shape<METRIC_COSINE>(s, a);
    if (rc == STREAM_REFUSED) {
        s->stream_mode = 2;      // this device / context cannot hold the whole grid at once: do not try again
        return STREAM_FALLBACK;
    }
    return rc;
}

bool host_query_ok(const sema_index *s, uint32_t k)
{
    const uint32_t ld4 = s->ld / 4;
    return s->host_path && s->variant == 0 && k >= 1 && k <= (uint32_t)K_PASS &&
           s->ld == s->dim && (ld4 == 96 || ld4 == 192) && s->res_map != nullptr;
}

// A host-query launch reads nothing but its parameters and rows [0, n) of X, and everything on this
// stream that writes those rows synchronises before it returns (tombstones, compaction), so the
// launch may always overlap the tail of whatever kernel precedes it: it is chained (PDL) whenever
// chaining is enabled.  Its own merge / exchange / result phase still waits for the predecessor.
int host_query_launch(sema_index *s, const float *q_host, uint32_t n, uint32_t k, const Exchange *x, uint64_t ticket)
{
    SEMA_NVTX("sema.K2.scan(host query)");
    unsigned char *slot = s->res_map_dev + (ticket % RES_SLOTS) * RES_SLOT_BYTES;
    uint64_t *ids_m = reinterpret_cast<uint64_t *>(slot + 8);
    float *sc_m = reinterpret_cast<float *>(slot + 8 + 8 * (size_t)k);
    return scan_query(s, q_host, n, k, nullptr, ids_m, sc_m, reinterpret_cast<uint32_t *>(slot), x,
                      SCAN_HOST_QUERY | (s->chain ? SCAN_CHAINED : 0u), ticket);
}

int host_query_wait(sema_index *s, uint64_t ticket, uint32_t k, uint64_t *row_ids, float *scores, uint32_t *n_found)
{
    // Poll the flag the last block stores after the results.  The stream is queried now and then so
    // that a failed launch / faulting kernel surfaces as an error instead of an endless wait.
    const unsigned char *slot = s->res_map + (ticket % RES_SLOTS) * RES_SLOT_BYTES;
    volatile const uint64_t *flag = reinterpret_cast<volatile const uint64_t *>(slot + RES_MAP_FLAG_OFF);
    static const uint32_t query_mask = [] {       // SEMA_POLL_QUERY_SHIFT: stream query every 2^shift polls (measurement knob)
        const char *e = getenv("SEMA_POLL_QUERY_SHIFT");
        const int sh = e ? atoi(e) : 14;
        return (uint32_t)((1u << (sh < 4 ? 4 : (sh > 30 ? 30 : sh))) - 1u);
    }();
    for (uint32_t spin = 1;; ++spin) {
        if (*flag == ticket) break;
        if ((spin & query_mask) == 0) {
            const cudaError_t e = cudaStreamQuery(s->stream);
            if (e == cudaSuccess) {
                if (*flag == ticket) break;
                return fail(SEMA_ERR_CUDA, "scan finished without publishing its result");
            }
            if (e != cudaErrorNotReady)
                return fail(SEMA_ERR_CUDA, "scan failed: %s", cudaGetErrorString(e));
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    __atomic_thread_fence(__ATOMIC_ACQUIRE);
    const uint32_t nf = *reinterpret_cast<const uint32_t *>(slot);
    if (nf == 0xffffffffu) return fail(SEMA_ERR_CUDA, "shard exchange timed out: a rank did not take part in the search");
    *n_found = nf;
    memcpy(row_ids, slot + 8, nf * sizeof(uint64_t));
    memcpy(scores, slot + 8 + 8 * (size_t)k, nf * sizeof(float));
    return SEMA_OK;
}

// The general host-buffer path (any dim, any k <= SEMA_MAX_K, optional K1 on the query): the query
// is staged through pinned memory and copied to the device, the result block comes back with a D2H
// copy, the stream is synchronised.
int staged_query_run(sema_index *s, const float *q_host, uint32_t n, uint32_t k, const Exchange *x, uint64_t *row_ids,
                     float *scores, uint32_t *n_found)
{
    memcpy(s->q_pin, q_host, s->dim * sizeof(float));
    CK(cudaMemcpyAsync(s->q_dev, s->q_pin, s->ld * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    if (s->normalize_queries) {
        int rc = normalize_queries_dev(s, s->q_dev, s->ld, 1);
        if (rc) return rc;
    }
    uint64_t *ids_d = reinterpret_cast<uint64_t *>(s->res_dev + 8);
    float *sc_d = reinterpret_cast<float *>(s->res_dev + 8 + 8 * (size_t)k);
    int rc = scan_query(s, s->q_dev, n, k, nullptr, ids_d, sc_d, reinterpret_cast<uint32_t *>(s->res_dev), x);
    if (rc) return rc;
    const size_t bytes = 8 + 12 * (size_t)k;
    CK(cudaMemcpyAsync(s->res_pin, s->res_dev, bytes, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const uint32_t nf = *reinterpret_cast<uint32_t *>(s->res_pin);
    if (nf == 0xffffffffu) return fail(SEMA_ERR_CUDA, "shard exchange timed out: a rank did not take part in the search");
    *n_found = nf;
    memcpy(row_ids, s->res_pin + 8, nf * sizeof(uint64_t));
    memcpy(scores, s->res_pin + 8 + 8 * (size_t)k, nf * sizeof(float));
    return SEMA_OK;
}

int slot_claim(sema_index *s, uint32_t k, uint64_t *ticket)
{
    const uint64_t t = s->host_seq + 1;
    sema_index::Slot &sl = s->slots[t % RES_SLOTS];
    if (sl.ticket != 0)
        return fail(SEMA_ERR_INVALID, "%d searches are already in flight on this handle: collect ticket %llu first",
                    RES_SLOTS, (unsigned long long)sl.ticket);
    s->host_seq = t;
    sl.ticket = t;
    sl.k = k;
    sl.sync_done = false;
    sl.nf = 0;
    *ticket = t;
    return SEMA_OK;
}

int slot_collect(sema_index *s, uint64_t ticket, uint64_t *row_ids, float *scores, uint32_t *n_found)
{
    sema_index::Slot &sl = s->slots[ticket % RES_SLOTS];
    if (ticket == 0 || sl.ticket != ticket) return fail(SEMA_ERR_INVALID, "ticket %llu is not outstanding", (unsigned long long)ticket);
    int rc = SEMA_OK;
    if (sl.sync_done) {
        *n_found = sl.nf;
        memcpy(row_ids, sl.ids.data(), sl.nf * sizeof(uint64_t));
        memcpy(scores, sl.sc.data(), sl.nf * sizeof(float));
    } else {
        rc = host_query_wait(s, ticket, sl.k, row_ids, scores, n_found);
    }
    sl.ticket = 0;      // the slot is free again, whatever the outcome
    return rc;
}

int host_query_run(sema_index *s, const float *q_host, uint32_t n, uint32_t k, const Exchange *x, uint64_t *row_ids,
                   float *scores, uint32_t *n_found)
{
    uint64_t ticket = 0;
    int rc = slot_claim(s, k, &ticket);
    if (rc) return rc;
    rc = host_query_launch(s, q_host, n, k, x, ticket);
    if (rc) {
        s->slots[ticket % RES_SLOTS].ticket = 0;
        return rc;
    }
    return slot_collect(s, ticket, row_ids, scores, n_found);
}

}  // namespace sema_impl

extern "C" {

int sema_index_search(sema_index *s, const float *q, uint32_t k, uint64_t *row_ids, float *scores,
                      uint32_t *n_found)
{
    if (!s || !q || !n_found) return fail(SEMA_ERR_INVALID, "null argument");
    if (k > SEMA_MAX_K) return fail(SEMA_ERR_INVALID, "k %u > SEMA_MAX_K %u", k, SEMA_MAX_K);
    if (k && (!row_ids || !scores)) return fail(SEMA_ERR_INVALID, "null output");
    CK(cudaSetDevice(s->device));
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    *n_found = 0;
    if (k == 0 || n == 0) return SEMA_OK;
    if (host_query_ok(s, k)) return host_query_run(s, q, (uint32_t)n, k, nullptr, row_ids, scores, n_found);
    return staged_query_run(s, q, (uint32_t)n, k, nullptr, row_ids, scores, n_found);
}

int sema_index_search_submit(sema_index *s, const float *q, uint32_t k, uint64_t *ticket)
{
    if (!s || !q || !ticket) return fail(SEMA_ERR_INVALID, "null argument");
    if (k > SEMA_MAX_K) return fail(SEMA_ERR_INVALID, "k %u > SEMA_MAX_K %u", k, SEMA_MAX_K);
    CK(cudaSetDevice(s->device));
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    rc = slot_claim(s, k, ticket);
    if (rc) return rc;
    sema_index::Slot &sl = s->slots[*ticket % RES_SLOTS];
    if (k != 0 && n != 0 && host_query_ok(s, k)) {
        s->last_snapshot = n;
        rc = host_query_launch(s, q, (uint32_t)n, k, nullptr, *ticket);
    } else {
        // outside the fast path (or nothing to scan): answer now, hand the result out at collect
        sl.ids.resize(k ? k : 1);
        sl.sc.resize(k ? k : 1);
        rc = sema_index_search(s, q, k, sl.ids.data(), sl.sc.data(), &sl.nf);
        sl.sync_done = true;
    }
    if (rc) sl.ticket = 0;
    return rc;
}

int sema_index_search_collect(sema_index *s, uint64_t ticket, uint64_t *row_ids, float *scores, uint32_t *n_found)
{
    if (!s || !n_found) return fail(SEMA_ERR_INVALID, "null argument");
    sema_index::Slot &sl = s->slots[ticket % RES_SLOTS];
    if (ticket != 0 && sl.ticket == ticket && sl.k && (!row_ids || !scores)) return fail(SEMA_ERR_INVALID, "null output");
    CK(cudaSetDevice(s->device));
    return slot_collect(s, ticket, row_ids, scores, n_found);
}

int sema_index_search_keys_device(sema_index *s, const float *q_dev, uint32_t k, uint64_t *keys_dev)
{
    if (!s || !q_dev || !keys_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (k == 0 || k > SEMA_MAX_K) return fail(SEMA_ERR_INVALID, "k %u outside [1, %u]", k, SEMA_MAX_K);
    CK(cudaSetDevice(s->device));
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    if (n == 0) {
        CK(cudaMemsetAsync(keys_dev, 0, k * sizeof(uint64_t), s->stream));
        return SEMA_OK;
    }
    const float *qd = q_dev;
    if (s->ld != s->dim || (reinterpret_cast<uintptr_t>(q_dev) & 15)) {
        CK(cudaMemcpyAsync(s->q_dev, q_dev, s->dim * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        qd = s->q_dev;
    }
    return scan_query(s, qd, (uint32_t)n, k, keys_dev, nullptr, nullptr, nullptr);
}

int sema_index_search_device(sema_index *s, const float *q_dev, uint32_t k, uint64_t *ids_dev,
                             float *scores_dev, uint32_t *n_found_dev)
{
    if (!s || !q_dev || !ids_dev || !scores_dev || !n_found_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (k == 0 || k > SEMA_MAX_K) return fail(SEMA_ERR_INVALID, "k %u outside [1, %u]", k, SEMA_MAX_K);
    CK(cudaSetDevice(s->device));
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    if (n == 0) {
        CK(cudaMemsetAsync(n_found_dev, 0, sizeof(uint32_t), s->stream));
        return SEMA_OK;
    }
    const float *qd = q_dev;
    if (s->ld != s->dim || (reinterpret_cast<uintptr_t>(q_dev) & 15)) {
        CK(cudaMemcpyAsync(s->q_dev, q_dev, s->dim * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        qd = s->q_dev;
    }
    return scan_query(s, qd, (uint32_t)n, k, nullptr, ids_dev, scores_dev, n_found_dev);
}

int sema_topk_merge_device(sema_index *s, const uint64_t *keys_dev, uint32_t n_lists, uint32_t k,
                           uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev)
{
    SEMA_NVTX("sema.K4.merge");
    if (!s || !keys_dev || !ids_dev || !scores_dev || !n_found_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (k == 0 || k > SEMA_MAX_K) return fail(SEMA_ERR_INVALID, "k %u outside [1, %u]", k, SEMA_MAX_K);
    if ((uint64_t)n_lists * k > 0x7fffffffull) return fail(SEMA_ERR_INVALID, "too many candidates");
    CK(cudaSetDevice(s->device));
    const uint32_t total = n_lists * k;
    const bool l2 = s->metric == SEMA_METRIC_L2;
    if (k <= (uint32_t)K_PASS)
        return l2 ? merge_pass_m<METRIC_L2>(s, keys_dev, total, k, nullptr, nullptr, ids_dev, scores_dev, n_found_dev)
                  : merge_pass_m<METRIC_COSINE>(s, keys_dev, total, k, nullptr, nullptr, ids_dev, scores_dev, n_found_dev);
    uint64_t *keys = s->keys_dev;
    for (uint32_t done = 0; done < k; done += K_PASS) {
        const uint32_t kp = (k - done) < (uint32_t)K_PASS ? (k - done) : (uint32_t)K_PASS;
        const uint64_t *bound = done ? keys + done - 1 : nullptr;
        int rc = l2 ? merge_pass_m<METRIC_L2>(s, keys_dev, total, kp, bound, keys + done, nullptr, nullptr, nullptr)
                    : merge_pass_m<METRIC_COSINE>(s, keys_dev, total, kp, bound, keys + done, nullptr, nullptr, nullptr);
        if (rc) return rc;
    }
    return decode(s, keys, k, ids_dev, scores_dev, n_found_dev);
}

}  // extern "C"

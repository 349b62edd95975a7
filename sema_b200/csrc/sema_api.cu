// sema_api.cu — host side of the C ABI declared in include/sema_b200.h.
//
// Owns the HBM layout (row-major fp32, row stride ld = round_up(dim,4) floats, base
// 256-byte aligned by cudaMalloc, so every row is 16-byte aligned), the streams and
// the small staging buffers; dispatches to K1 (ingest), K2 (scan + top-k) and K4
// (merge).  No CPU fallback exists: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <deque>
#include <vector>
#include <mutex>
#include <new>

#include "../../include/sema_b200.h"
#include "k1_ingest.cuh"
#include "k2_scan.cuh"
#include "k4_merge.cuh"
#include "k3_batch.cuh"

using namespace sema;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(SEMA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

struct Pending {
    cudaEvent_t ev;
    uint64_t rows_after;
};

constexpr int K_PASS = 128;  // keys one fused pass can select (WarpTopK<4>)
constexpr int MAX_BLOCKS_PER_SM = 8;

}  // namespace

struct sema_index {
    int device = 0;
    uint32_t dim = 0, ld = 0;
    uint64_t capacity = 0;
    int metric = 0;
    int num_sms = 0;
    float *X = nullptr;
    uint8_t *valid = nullptr;
    uint64_t n_rows = 0;     // appended (enqueued)
    uint64_t n_visible = 0;  // ingest completed on the device
    uint64_t last_snapshot = 0;
    uint32_t row_base = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr, ingest_stream = nullptr;
    float *q_dev = nullptr;           // ld floats
    float *q_pin = nullptr;           // pinned, ld floats
    uint64_t *partials = nullptr;     // num_sms * MAX_BLOCKS_PER_SM * 128 keys
    unsigned int *ticket = nullptr;
    uint64_t *keys_dev = nullptr;     // SEMA_MAX_K keys (multi-pass scratch)
    unsigned char *res_dev = nullptr; // [n_found u32, pad][ids u64 K][scores f32 K]
    unsigned char *res_pin = nullptr;
    // batch buffers, grown on demand
    float *Q_dev = nullptr;
    uint64_t *bids_dev = nullptr;
    float *bsc_dev = nullptr;
    uint32_t *bnf_dev = nullptr;
    size_t batch_cap_q = 0, batch_cap_res = 0;
    uint64_t *tomb_dev = nullptr;
    size_t tomb_cap = 0;
    std::deque<Pending> pending;
    int variant = 0;
    uint64_t launches = 0;
    // ---- K3 (batched tensor-core path) state
    float *max_norm2 = nullptr;         // device: max squared row norm (written by K1)
    unsigned char *planes = nullptr;    // pre-tiled bf16 hi/lo planes, built lazily
    uint64_t planes_rows = 0;           // rows [0, planes_rows) are reflected in the planes
    bool planes_failed = false;         // allocation failed once: stay on the K2 loop
    float *Qpad_dev = nullptr;
    uint32_t *cand_rows = nullptr;
    float *cand_thr = nullptr;
    uint32_t *flags_dev = nullptr, *flags_pin = nullptr;
    size_t qpad_cap = 0, cand_cap = 0, thr_cap = 0, flags_cap = 0;
    int batch_mode = 0;                 // 0 auto, 1 always the K2 loop, 2 K3 bf16x3 whenever the shape allows, 3 K3 single bf16 pass
    int k3_cluster = 0;                 // 0 auto, else forced cluster size (1, 2, 4) — tuning
    int k3_qt = 0;                      // 0 auto, 1 = one query tile per CTA even in the single-pass mode — tuning
    int normalize_queries = 0;          // apply K1 to host queries before scanning
    unsigned char *qscratch = nullptr;  // [valid byte x MAXQ pad][float max_norm2 scratch]
    uint64_t k3_queries = 0, k3_fallbacks = 0;
};

namespace {

int poll_ingest(sema_index *s, bool wait)
{
    while (!s->pending.empty()) {
        Pending &p = s->pending.front();
        cudaError_t e = wait ? cudaEventSynchronize(p.ev) : cudaEventQuery(p.ev);
        if (e == cudaErrorNotReady) break;
        if (e != cudaSuccess) return fail(SEMA_ERR_CUDA, "ingest event: %s", cudaGetErrorString(e));
        s->n_visible = p.rows_after;
        cudaEventDestroy(p.ev);
        s->pending.pop_front();
    }
    return SEMA_OK;
}

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
};

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
    if (a.x) p.x = *a.x;
    else memset(&p.x, 0, sizeof p.x);
    kern<<<grid, SCAN_THREADS, dyn, s->stream>>>(p);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
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
    if (ld4 == 96) {
        switch (s->variant) {
            case 1: return scan_m<3, 4, METRIC>(s, a);
            case 2: return scan_m<3, 2, METRIC>(s, a);
            default: return scan_m<3, 8, METRIC>(s, a);   // R = 8: best or tied on every box measured
        }
    }
    if (ld4 == 192) {
        switch (s->variant) {
            case 1: return scan_m<6, 4, METRIC>(s, a);
            default: return scan_m<6, 2, METRIC>(s, a);   // 24 float4 per lane per 4 rows is too many registers
        }
    }
    return scan_m<0, 4, METRIC>(s, a);
}

// one fused pass (k <= K_PASS)
int scan_pass(sema_index *s, const ScanArgs &a)
{
    return s->metric == SEMA_METRIC_L2 ? scan_shape<METRIC_L2>(s, a) : scan_shape<METRIC_COSINE>(s, a);
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
int scan_query(sema_index *s, const float *q_dev, uint32_t n, uint32_t k, uint64_t *out_keys,
               uint64_t *res_ids, float *res_scores, uint32_t *res_nfound, const Exchange *x = nullptr)
{
    if (k <= K_PASS) {
        ScanArgs a{q_dev, n, k, nullptr, out_keys ? out_keys : s->keys_dev, res_ids, res_scores, res_nfound, x};
        return scan_pass(s, a);
    }
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

int ensure(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return SEMA_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    CK(cudaMalloc(p, need));
    *cap = need;
    return SEMA_OK;
}

int launch_ingest(sema_index *s, const float *src, uint64_t src_ld, uint64_t first, uint64_t n,
                  const uint8_t *valid_in, int normalize, bool vec4)
{
    float *dst = s->X + first * s->ld;
    uint64_t warps_needed = n;
    uint64_t blocks = (warps_needed * 32 + INGEST_THREADS - 1) / INGEST_THREADS;
    const uint64_t maxb = (uint64_t)s->num_sms * 16;
    if (blocks > maxb) blocks = maxb;
    if (blocks < 1) blocks = 1;
    if (vec4)
        ingest_kernel<true><<<(unsigned)blocks, INGEST_THREADS, 0, s->ingest_stream>>>(
            src, src_ld, dst, s->ld, s->dim, n, valid_in, s->valid + first, normalize, s->max_norm2);
    else
        ingest_kernel<false><<<(unsigned)blocks, INGEST_THREADS, 0, s->ingest_stream>>>(
            src, src_ld, dst, s->ld, s->dim, n, valid_in, s->valid + first, normalize, s->max_norm2);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

int publish(sema_index *s, uint64_t n)
{
    s->n_rows += n;
    Pending p;
    CK(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming));
    CK(cudaEventRecord(p.ev, s->ingest_stream));
    p.rows_after = s->n_rows;
    s->pending.push_back(p);
    return SEMA_OK;
}


// ---- K3 dispatch -----------------------------------------------------------------
constexpr uint32_t K3_MAX_K = 100;
constexpr float K3_ERR_REL = 2.5e-4f;  // >= 3*2^-16 (dropped split terms) + fp32 accumulation over 3*dim terms
constexpr float K3_ERR_REL_1PASS = 8.5e-3f;  // >= 2*2^-8 + 2^-16 (both operands rounded to bf16) + accumulation

// passes the batch would run with: bf16x3 up to dim 384 (TMEM holds q_hi and q_lo), the single-pass
// filter up to dim 768 (q_hi only) or when asked for (batch mode 3); 0 = K3 cannot serve this shape
int k3_passes(const sema_index *s, uint32_t k)
{
    if (s->metric != SEMA_METRIC_COSINE || s->dim % k3::BLOCK_K != 0 || k > K3_MAX_K || s->planes_failed) return 0;
    if (s->dim <= (uint32_t)k3::MAX_DIM) return s->batch_mode == 3 ? 1 : 3;
    if (s->dim <= (uint32_t)k3::MAX_DIM_1PASS) return 1;
    return 0;
}

// Bring the bf16 planes up to date with rows [0, n).  Tombstones invalidate from their row on.
int k3_sync_planes(sema_index *s, uint64_t n)
{
    if (!s->planes) {
        const uint64_t tiles = (s->capacity + k3::TILE_N - 1) / k3::TILE_N;
        cudaError_t e = cudaMalloc(&s->planes, (size_t)(tiles ? tiles : 1) * k3::tile_bytes((int)s->dim));
        if (e != cudaSuccess) {
            cudaGetLastError();
            s->planes = nullptr;
            s->planes_failed = true;  // not an error: the K2 loop serves the batch instead
            return SEMA_ERR_NOMEM;
        }
        s->planes_rows = 0;
    }
    if (s->planes_rows >= n) return SEMA_OK;
    const uint64_t begin = (s->planes_rows / k3::TILE_N) * k3::TILE_N;            // re-tile the partial last tile
    const uint64_t end = ((n + k3::TILE_N - 1) / k3::TILE_N) * k3::TILE_N;
    const uint64_t work = (end - begin) * (s->dim / 8);
    uint64_t blocks = (work + 255) / 256;
    if (blocks > (uint64_t)s->num_sms * 32) blocks = (uint64_t)s->num_sms * 32;
    k3::split_planes_kernel<<<(unsigned)blocks, 256, 0, s->stream>>>(s->X, s->ld, s->dim, begin, end, n, s->planes);
    CK(cudaGetLastError());
    s->launches++;
    s->planes_rows = n;
    return SEMA_OK;
}

// q_ctas = CTAs along the query axis (each owns QT query tiles)
template <int KC, int C, int PASSES, int QT>
int k3_launch_scan_c(sema_index *s, const k3::Params &p, uint32_t q_ctas)
{
    auto kern = k3::batch_scan_kernel<KC, C, PASSES, QT>;
    static bool attr_set[64] = {false};
    if (!attr_set[s->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::Smem<KC, PASSES, QT>::TOTAL));
        attr_set[s->device & 63] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(q_ctas, p.parts, 1);
    cfg.blockDim = dim3(k3::THREADS, 1, 1);
    cfg.dynamicSmemBytes = k3::Smem<KC, PASSES, QT>::TOTAL;
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, p));
    s->launches++;
    return SEMA_OK;
}

// how many clusters of C CTAs of this kernel can be resident at once (cached per device)
template <int KC, int C>
int k3_max_clusters(sema_index *s, int *out)
{
    static int cached[64] = {0};
    int &v = cached[s->device & 63];
    if (v == 0) {
        auto kern = k3::batch_scan_kernel<KC, C, 3, 1>;   // the single-pass kernels use no more shared memory
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::Smem<KC, 3>::TOTAL));
        if (C == 1) {
            v = s->num_sms;
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(C, (unsigned)s->num_sms, 1);
            cfg.blockDim = dim3(k3::THREADS, 1, 1);
            cfg.dynamicSmemBytes = k3::Smem<KC, 3>::TOTAL;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = C;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n = 0;
            CK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
            v = n > 0 ? n : 1;
        }
    }
    *out = v;
    return SEMA_OK;
}

template <int KC, int PASSES, int QT>
int k3_launch_scan_p(sema_index *s, const k3::Params &p, uint32_t q_ctas, int c)
{
    return c == 4 ? k3_launch_scan_c<KC, 4, PASSES, QT>(s, p, q_ctas)
         : c == 2 ? k3_launch_scan_c<KC, 2, PASSES, QT>(s, p, q_ctas) : k3_launch_scan_c<KC, 1, PASSES, QT>(s, p, q_ctas);
}
template <int KC>
int k3_launch_scan(sema_index *s, const k3::Params &p, uint32_t q_ctas, int c, int passes, int qt)
{
    if (passes == 3) return k3_launch_scan_p<KC, 3, 1>(s, p, q_ctas, c);
    if constexpr (KC <= 64) {
        if (qt == 2) return k3_launch_scan_p<KC, 1, 2>(s, p, q_ctas, c);
    }
    return k3_launch_scan_p<KC, 1, 1>(s, p, q_ctas, c);
}
template <int KC>
int k3_clusters(sema_index *s, int c, int *out)
{
    return c == 4 ? k3_max_clusters<KC, 4>(s, out) : c == 2 ? k3_max_clusters<KC, 2>(s, out) : k3_max_clusters<KC, 1>(s, out);
}

// Qd: nq x dim dense on the device.  Results: device arrays [nq*k], [nq*k], [nq].
int k3_batch(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
             uint32_t *nf_d)
{
    const uint32_t kc = k <= 16 ? 32 : (k <= 48 ? 64 : 128);
    const int passes = k3_passes(s, k);
    const uint32_t n_tiles = (n + k3::TILE_N - 1) / k3::TILE_N;
    const uint32_t q_tiles_all = (nq + k3::TILE_Q - 1) / k3::TILE_Q;
    int rc;
    // QT query tiles per CTA (2 in the single-pass mode when there are at least 2 tiles), clusters of
    // csize CTAs along the query axis: the query-tile count is padded to a multiple of QT*csize
    // (Qpad rows beyond nq are zero queries whose results are never read).
    // (two tiles need 2*dim/2 + 128 TMEM columns and two candidate lists in shared memory)
    const int qt_per_cta = (passes == 1 && q_tiles_all >= 2 && s->k3_qt != 1 && s->dim <= 384 && kc <= 64) ? 2 : 1;
    const uint32_t q_ctas_all = (q_tiles_all + qt_per_cta - 1) / qt_per_cta;
    // measured on 10M x 384 x 1024q: bf16x3 is fastest with clusters of 4 (128 SMs, higher clocks under the
    // power cap), the single-pass filter with clusters of 2 (144 SMs; it is bound by L2->SM delivery)
    const int cpref = passes == 1 ? 2 : 4;
    const int csize = s->k3_cluster > 0 ? s->k3_cluster : (q_ctas_all >= (uint32_t)cpref ? cpref : (q_ctas_all >= 2 ? 2 : 1));
    const uint32_t q_ctas_pad = ((q_ctas_all + csize - 1) / csize) * csize;
    const size_t qpad_rows = (size_t)q_ctas_pad * qt_per_cta * k3::TILE_Q;
    rc = ensure(reinterpret_cast<void **>(&s->Qpad_dev), &s->qpad_cap, qpad_rows * s->dim * sizeof(float));
    if (rc) return rc;
    if (s->flags_cap < nq) {
        cudaFree(s->flags_dev); cudaFreeHost(s->flags_pin);
        s->flags_dev = nullptr; s->flags_pin = nullptr; s->flags_cap = 0;
        CK(cudaMalloc(&s->flags_dev, (size_t)nq * sizeof(uint32_t)));
        CK(cudaHostAlloc(&s->flags_pin, (size_t)nq * sizeof(uint32_t), cudaHostAllocPortable));
        s->flags_cap = nq;
    }
    CK(cudaMemsetAsync(s->Qpad_dev, 0, qpad_rows * s->dim * sizeof(float), s->stream));
    CK(cudaMemcpyAsync(s->Qpad_dev, Qd, (size_t)nq * s->dim * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));

    int max_clusters = 1;
    rc = kc == 32 ? k3_clusters<32>(s, csize, &max_clusters) : kc == 64 ? k3_clusters<64>(s, csize, &max_clusters) : k3_clusters<128>(s, csize, &max_clusters);
    if (rc) return rc;
    const uint32_t groups_all = q_ctas_pad / csize;             // clusters along the query axis
    // at most max_clusters cluster columns per launch; the clusters left over become row partitions
    for (uint32_t g0 = 0; g0 < groups_all; g0 += (uint32_t)max_clusters) {
        const uint32_t groups = (groups_all - g0) < (uint32_t)max_clusters ? (groups_all - g0) : (uint32_t)max_clusters;
        const uint32_t q_ctas = groups * csize;                 // CTAs along the query axis in this launch
        const uint32_t qt0 = g0 * csize * qt_per_cta;           // first query tile of this launch
        const uint32_t q_tiles = q_ctas * qt_per_cta;
        uint32_t parts = (uint32_t)max_clusters / groups;
        if (parts > n_tiles) parts = n_tiles;
        if (parts < 1) parts = 1;
        const size_t nqp = (size_t)q_tiles * k3::TILE_Q;
        rc = ensure(reinterpret_cast<void **>(&s->cand_rows), &s->cand_cap, nqp * parts * kc * sizeof(uint32_t));
        if (rc) return rc;
        rc = ensure(reinterpret_cast<void **>(&s->cand_thr), &s->thr_cap, nqp * parts * sizeof(float));
        if (rc) return rc;
        k3::Params p;
        p.planes = s->planes;
        p.Q = s->Qpad_dev + (size_t)qt0 * k3::TILE_Q * s->dim;
        p.cand_rows = s->cand_rows;
        p.cand_thr = s->cand_thr;
        p.n_rows = n;
        p.n_tiles = n_tiles;
        p.parts = parts;
        p.dim = s->dim;
        rc = kc == 32 ? k3_launch_scan<32>(s, p, q_ctas, csize, passes, qt_per_cta) : kc == 64 ? k3_launch_scan<64>(s, p, q_ctas, csize, passes, qt_per_cta)
                      : k3_launch_scan<128>(s, p, q_ctas, csize, passes, qt_per_cta);
        if (rc) return rc;
        const uint32_t q_first = qt0 * k3::TILE_Q;
        if (q_first >= nq) break;
        const uint32_t q_cnt = (nq - q_first) < (uint32_t)nqp ? (nq - q_first) : (uint32_t)nqp;
        k3::RescoreParams r;
        r.X = reinterpret_cast<const float4 *>(s->X);
        r.Q = p.Q;
        r.cand_rows = s->cand_rows;
        r.cand_thr = s->cand_thr;
        r.res_ids = ids_d + (size_t)q_first * k;
        r.res_scores = sc_d + (size_t)q_first * k;
        r.res_nfound = nf_d + q_first;
        r.flags = s->flags_dev + q_first;
        r.ld4 = s->ld / 4;
        r.dim = s->dim;
        r.k = k;
        r.parts = parts;
        r.kc = kc;
        r.row_base = s->row_base;
        r.max_norm2 = s->max_norm2;
        r.err_rel = passes == 1 ? K3_ERR_REL_1PASS : K3_ERR_REL;
        if (k <= 32) k3::rescore_kernel<1><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
        else if (k <= 64) k3::rescore_kernel<2><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
        else k3::rescore_kernel<4><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
        CK(cudaGetLastError());
        s->launches++;
    }
    // queries whose exactness could not be proven (heavy ties / near-duplicates) go through K2
    CK(cudaMemcpyAsync(s->flags_pin, s->flags_dev, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->k3_queries += nq;
    for (uint32_t i = 0; i < nq; ++i) {
        if (!s->flags_pin[i]) continue;
        s->k3_fallbacks++;
        rc = scan_query(s, s->Qpad_dev + (size_t)i * s->dim, n, k, nullptr, ids_d + (size_t)i * k, sc_d + (size_t)i * k, nf_d + i);
        if (rc) return rc;
    }
    return SEMA_OK;
}

// Batched search with the queries already on the device (nq x dim dense).
int batch_core(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
               uint32_t *nf_d)
{
    const bool want_k3 = s->batch_mode >= 2 || (s->batch_mode == 0 && nq >= 4);
    if (want_k3 && k3_passes(s, k) != 0) {
        int rc = k3_sync_planes(s, n);
        if (rc == SEMA_OK) return k3_batch(s, Qd, nq, n, k, ids_d, sc_d, nf_d);
        if (rc != SEMA_ERR_NOMEM) return rc;
    }
    // K2 once per query (still one HBM pass per query)
    const float *Qp = Qd;
    if (s->ld != s->dim) {
        int rc = ensure(reinterpret_cast<void **>(&s->Qpad_dev), &s->qpad_cap, (size_t)nq * s->ld * sizeof(float));
        if (rc) return rc;
        CK(cudaMemsetAsync(s->Qpad_dev, 0, (size_t)nq * s->ld * sizeof(float), s->stream));
        CK(cudaMemcpy2DAsync(s->Qpad_dev, s->ld * sizeof(float), Qd, s->dim * sizeof(float), s->dim * sizeof(float),
                             nq, cudaMemcpyDeviceToDevice, s->stream));
        Qp = s->Qpad_dev;
    }
    for (uint32_t i = 0; i < nq; ++i) {
        int rc = scan_query(s, Qp + (size_t)i * s->ld, n, k, nullptr, ids_d + (size_t)i * k, sc_d + (size_t)i * k, nf_d + i);
        if (rc) return rc;
    }
    return SEMA_OK;
}

// K1 on query vectors in place (nq rows of `stride` floats on the device, on the query stream)
int normalize_queries_dev(sema_index *s, float *q, uint64_t stride, uint32_t nq)
{
    for (uint32_t done = 0; done < nq; done += 65536) {
        const uint32_t m = (nq - done) < 65536u ? (nq - done) : 65536u;
        uint64_t blocks = ((uint64_t)m * 32 + INGEST_THREADS - 1) / INGEST_THREADS;
        if (blocks > (uint64_t)s->num_sms * 16) blocks = (uint64_t)s->num_sms * 16;
        float *base = q + (size_t)done * stride;
        // generic (scalar) kernel: src == dst, same stride; pad columns [dim, stride) are rewritten as zeros
        ingest_kernel<false><<<(unsigned)blocks, INGEST_THREADS, 0, s->stream>>>(
            base, stride, base, (uint32_t)stride, s->dim, m, nullptr, s->qscratch,
            1, reinterpret_cast<float *>(s->qscratch + 65536));
        CK(cudaGetLastError());
        s->launches++;
    }
    return SEMA_OK;
}

int check_append(sema_index *s, uint64_t n)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (s->n_rows + n > s->capacity)
        return fail(SEMA_ERR_CAPACITY, "append of %llu rows exceeds capacity %llu (size %llu)",
                    (unsigned long long)n, (unsigned long long)s->capacity, (unsigned long long)s->n_rows);
    if ((uint64_t)s->row_base + s->n_rows + n > 0xfffffffeull)
        return fail(SEMA_ERR_CAPACITY, "global row ids must stay below 2^32-1");
    return SEMA_OK;
}

}  // namespace

extern "C" {

const char *sema_last_error(void) { return g_err; }
const char *sema_version(void) { return "sema_b200 0.1 (sm_100a)"; }

int sema_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int sema_host_alloc(void **out, size_t bytes)
{
    if (!out) return fail(SEMA_ERR_INVALID, "null out");
    CK(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return SEMA_OK;
}

int sema_host_free(void *p)
{
    if (p) CK(cudaFreeHost(p));
    return SEMA_OK;
}

int sema_index_create(int device, uint32_t dim, uint64_t capacity_rows, int metric, sema_index **out)
{
    if (!out) return fail(SEMA_ERR_INVALID, "null out");
    *out = nullptr;
    if (dim == 0 || dim > SEMA_MAX_DIM) return fail(SEMA_ERR_INVALID, "dim %u outside [1, %u]", dim, SEMA_MAX_DIM);
    if (metric != SEMA_METRIC_COSINE && metric != SEMA_METRIC_L2) return fail(SEMA_ERR_INVALID, "unknown metric %d", metric);
    if (capacity_rows > 0xfffffffeull) return fail(SEMA_ERR_INVALID, "capacity_rows must be < 2^32-1");
    int ndev = sema_device_count();
    if (ndev == 0) return fail(SEMA_ERR_CUDA, "no CUDA device available (there is no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(SEMA_ERR_INVALID, "device %d outside [0, %d)", device, ndev);
    CK(cudaSetDevice(device));
    sema_index *s = new (std::nothrow) sema_index();
    if (!s) return fail(SEMA_ERR_NOMEM, "host allocation failed");
    s->device = device;
    s->dim = dim;
    s->ld = (dim + 3u) & ~3u;
    s->capacity = capacity_rows;
    s->metric = metric;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete s; return fail(SEMA_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e)); }
    s->num_sms = prop.multiProcessorCount;
    const size_t xbytes = (size_t)(capacity_rows ? capacity_rows : 1) * s->ld * sizeof(float);
    const size_t res_bytes = 8 + (size_t)SEMA_MAX_K * 12;
#define CKD(call)                                                                           \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            int code_ = e_ == cudaErrorMemoryAllocation ? SEMA_ERR_NOMEM : SEMA_ERR_CUDA;   \
            fail(code_, "%s failed: %s", #call, cudaGetErrorString(e_));                    \
            cudaGetLastError();                                                             \
            sema_index_destroy(s);                                                          \
            return code_;                                                                   \
        }                                                                                   \
    } while (0)
    CKD(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    CKD(cudaStreamCreateWithFlags(&s->ingest_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    CKD(cudaMalloc(&s->X, xbytes));
    CKD(cudaMalloc(&s->valid, capacity_rows ? capacity_rows : 1));
    CKD(cudaMalloc(&s->q_dev, s->ld * sizeof(float)));
    CKD(cudaMemset(s->q_dev, 0, s->ld * sizeof(float)));
    CKD(cudaHostAlloc(&s->q_pin, s->ld * sizeof(float), cudaHostAllocPortable));
    memset(s->q_pin, 0, s->ld * sizeof(float));
    CKD(cudaMalloc(&s->partials, (size_t)s->num_sms * MAX_BLOCKS_PER_SM * K_PASS * sizeof(uint64_t)));
    CKD(cudaMalloc(&s->ticket, 2 * sizeof(unsigned int)));
    CKD(cudaMemset(s->ticket, 0, 2 * sizeof(unsigned int)));
    CKD(cudaMalloc(&s->keys_dev, SEMA_MAX_K * sizeof(uint64_t)));
    CKD(cudaMalloc(&s->max_norm2, sizeof(float)));
    CKD(cudaMalloc(&s->qscratch, 65536 + 16));
    CKD(cudaMemset(s->max_norm2, 0, sizeof(float)));
    CKD(cudaMalloc(&s->res_dev, res_bytes));
    CKD(cudaHostAlloc(&s->res_pin, res_bytes, cudaHostAllocPortable));
    CKD(cudaDeviceSynchronize());
#undef CKD
    *out = s;
    return SEMA_OK;
}

int sema_index_destroy(sema_index *s)
{
    if (!s) return SEMA_OK;
    cudaSetDevice(s->device);
    if (s->own_stream) cudaStreamSynchronize(s->own_stream);
    if (s->ingest_stream) cudaStreamSynchronize(s->ingest_stream);
    for (auto &p : s->pending) cudaEventDestroy(p.ev);
    cudaFree(s->X); cudaFree(s->valid); cudaFree(s->q_dev); cudaFreeHost(s->q_pin);
    cudaFree(s->partials); cudaFree(s->ticket); cudaFree(s->keys_dev); cudaFree(s->res_dev);
    cudaFreeHost(s->res_pin); cudaFree(s->Q_dev); cudaFree(s->bids_dev); cudaFree(s->bsc_dev);
    cudaFree(s->bnf_dev); cudaFree(s->tomb_dev);
    cudaFree(s->qscratch); cudaFree(s->max_norm2); cudaFree(s->planes); cudaFree(s->Qpad_dev); cudaFree(s->cand_rows);
    cudaFree(s->cand_thr); cudaFree(s->flags_dev); cudaFreeHost(s->flags_pin);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->ingest_stream) cudaStreamDestroy(s->ingest_stream);
    cudaGetLastError();
    delete s;
    return SEMA_OK;
}

int sema_index_append_async(sema_index *s, const float *rows, uint64_t n, const uint8_t *valid,
                            int normalize, uint64_t *first_row)
{
    int rc = check_append(s, n);
    if (rc) return rc;
    if (n && !rows) return fail(SEMA_ERR_INVALID, "null rows");
    CK(cudaSetDevice(s->device));
    const uint64_t first = s->n_rows;
    if (first_row) *first_row = first;
    if (n == 0) return SEMA_OK;
    float *dst = s->X + first * s->ld;
    // host rows land directly in their final place; K1 then normalises in place
    if (s->ld == s->dim)
        CK(cudaMemcpyAsync(dst, rows, n * s->dim * sizeof(float), cudaMemcpyHostToDevice, s->ingest_stream));
    else
        CK(cudaMemcpy2DAsync(dst, s->ld * sizeof(float), rows, s->dim * sizeof(float),
                             s->dim * sizeof(float), n, cudaMemcpyHostToDevice, s->ingest_stream));
    const uint8_t *vin = nullptr;
    if (valid) {
        CK(cudaMemcpyAsync(s->valid + first, valid, n, cudaMemcpyHostToDevice, s->ingest_stream));
        vin = s->valid + first;
    }
    rc = launch_ingest(s, dst, s->ld, first, n, vin, normalize, s->ld == s->dim);
    if (rc) return rc;
    return publish(s, n);
}

int sema_index_flush(sema_index *s)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    CK(cudaSetDevice(s->device));
    CK(cudaStreamSynchronize(s->ingest_stream));
    return poll_ingest(s, true);
}

int sema_index_append(sema_index *s, const float *rows, uint64_t n, const uint8_t *valid,
                      int normalize, uint64_t *first_row)
{
    int rc = sema_index_append_async(s, rows, n, valid, normalize, first_row);
    if (rc) return rc;
    return sema_index_flush(s);
}

int sema_index_append_device(sema_index *s, const float *rows_dev, uint64_t n,
                             const uint8_t *valid_dev, int normalize, uint64_t *first_row)
{
    int rc = check_append(s, n);
    if (rc) return rc;
    if (n && !rows_dev) return fail(SEMA_ERR_INVALID, "null rows");
    CK(cudaSetDevice(s->device));
    const uint64_t first = s->n_rows;
    if (first_row) *first_row = first;
    if (n == 0) return SEMA_OK;
    const bool vec4 = s->ld == s->dim && (reinterpret_cast<uintptr_t>(rows_dev) & 15) == 0;
    rc = launch_ingest(s, rows_dev, s->dim, first, n, valid_dev, normalize, vec4);
    if (rc) return rc;
    rc = publish(s, n);
    if (rc) return rc;
    return sema_index_flush(s);
}

int sema_index_append_synthetic(sema_index *s, uint64_t seed, uint64_t synth_row0, uint64_t n,
                                int normalize, uint64_t *first_row)
{
    int rc = check_append(s, n);
    if (rc) return rc;
    CK(cudaSetDevice(s->device));
    const uint64_t first = s->n_rows;
    if (first_row) *first_row = first;
    if (n == 0) return SEMA_OK;
    float *dst = s->X + first * s->ld;
    synth_kernel<<<s->num_sms * 16, INGEST_THREADS, 0, s->ingest_stream>>>(dst, s->ld, s->dim, seed, synth_row0, n);
    CK(cudaGetLastError());
    s->launches++;
    rc = launch_ingest(s, dst, s->ld, first, n, nullptr, normalize, s->ld == s->dim);
    if (rc) return rc;
    rc = publish(s, n);
    if (rc) return rc;
    return sema_index_flush(s);
}

int sema_index_tombstone(sema_index *s, const uint64_t *rows, uint64_t n)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (n && !rows) return fail(SEMA_ERR_INVALID, "null rows");
    if (n == 0) return SEMA_OK;
    int rc = sema_index_flush(s);
    if (rc) return rc;
    for (uint64_t i = 0; i < n; ++i)
        if (rows[i] >= s->n_rows)
            return fail(SEMA_ERR_INVALID, "tombstone row %llu >= size %llu", (unsigned long long)rows[i], (unsigned long long)s->n_rows);
    rc = ensure(reinterpret_cast<void **>(&s->tomb_dev), &s->tomb_cap, n * sizeof(uint64_t));
    if (rc) return rc;
    CK(cudaMemcpyAsync(s->tomb_dev, rows, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    uint64_t blocks = (n * 32 + INGEST_THREADS - 1) / INGEST_THREADS;
    if (blocks > (uint64_t)s->num_sms * 16) blocks = (uint64_t)s->num_sms * 16;
    tombstone_kernel<<<(unsigned)blocks, INGEST_THREADS, 0, s->stream>>>(s->X, s->ld, s->tomb_dev, n, s->n_rows, s->valid);
    CK(cudaGetLastError());
    s->launches++;
    CK(cudaStreamSynchronize(s->stream));
    // the bf16 planes of K3 must forget the dead rows: re-tile from the first one on
    uint64_t lowest = s->planes_rows;
    for (uint64_t i = 0; i < n; ++i) if (rows[i] < lowest) lowest = rows[i];
    s->planes_rows = lowest;
    return SEMA_OK;
}


int sema_index_compact(sema_index *s, uint64_t *new_row_of_old, uint64_t *n_live_out)
{
    return sema_index_compact_keep(s, nullptr, new_row_of_old, n_live_out);
}

int sema_index_compact_keep(sema_index *s, const uint8_t *keep, uint64_t *new_row_of_old, uint64_t *n_live_out)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    int rc = sema_index_flush(s);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));
    const uint64_t n = s->n_rows;
    std::vector<uint8_t> valid(n ? n : 1);
    if (n) CK(cudaMemcpy(valid.data(), s->valid, n, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> src;   // src[new] = old, ascending
    std::vector<uint8_t> new_valid;
    src.reserve(n);
    new_valid.reserve(n);
    for (uint64_t r = 0; r < n; ++r) {
        if (keep ? keep[r] != 0 : valid[r] != 0) {
            new_valid.push_back(valid[r]);
            if (new_row_of_old) new_row_of_old[r] = src.size();
            src.push_back((uint32_t)r);
        } else if (new_row_of_old) {
            new_row_of_old[r] = ~0ull;
        }
    }
    const uint64_t live = src.size();
    if (n_live_out) *n_live_out = live;
    if (live == n) return SEMA_OK;  // nothing to drop
    // gather through a bounce buffer, chunk by chunk in ascending order: a chunk's destination
    // [j*C, (j+1)*C) never overlaps a later chunk's sources (src[i] >= i)
    const uint64_t C = 1u << 16;
    float *tmp = nullptr;
    uint32_t *src_dev = nullptr;
    cudaError_t e = cudaMalloc(&tmp, C * s->ld * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&src_dev, C * sizeof(uint32_t));
    if (e != cudaSuccess) {
        cudaFree(tmp); cudaFree(src_dev); cudaGetLastError();
        return fail(SEMA_ERR_NOMEM, "compaction scratch: %s", cudaGetErrorString(e));
    }
    uint64_t first_moved = 0;
    while (first_moved < live && src[first_moved] == first_moved) ++first_moved;   // untouched prefix
    for (uint64_t at = first_moved; at < live; at += C) {
        const uint64_t m = (live - at) < C ? (live - at) : C;
        CK(cudaMemcpyAsync(src_dev, src.data() + at, m * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
        uint64_t blocks = (m * 32 + INGEST_THREADS - 1) / INGEST_THREADS;
        if (blocks > (uint64_t)s->num_sms * 16) blocks = (uint64_t)s->num_sms * 16;
        gather_rows_kernel<<<(unsigned)blocks, INGEST_THREADS, 0, s->stream>>>(s->X, s->ld, src_dev, m, tmp);
        CK(cudaGetLastError());
        s->launches++;
        CK(cudaMemcpyAsync(s->X + at * s->ld, tmp, m * s->ld * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        CK(cudaStreamSynchronize(s->stream));   // src_dev / tmp are reused by the next chunk
    }
    if (live) CK(cudaMemcpyAsync(s->valid, new_valid.data(), live, cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    cudaFree(tmp);
    cudaFree(src_dev);
    s->n_rows = live;
    s->n_visible = live;
    if (s->planes_rows > first_moved) s->planes_rows = first_moved;   // K3 planes: re-tile from the first moved row
    return SEMA_OK;
}

namespace {
struct SemaFileHeader {
    char magic[8];
    uint32_t dim;
    int32_t metric;
    uint64_t n_rows;
    unsigned char pad[40];
};
static_assert(sizeof(SemaFileHeader) == 64, "header is 64 bytes");
}  // namespace

int sema_index_save(sema_index *s, const char *path)
{
    if (!s || !path) return fail(SEMA_ERR_INVALID, "null argument");
    int rc = sema_index_flush(s);
    if (rc) return rc;
    CK(cudaStreamSynchronize(s->stream));
    FILE *f = fopen(path, "wb");
    if (!f) return fail(SEMA_ERR_INVALID, "cannot open %s for writing", path);
    SemaFileHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "SEMAIDX1", 8);
    h.dim = s->dim;
    h.metric = s->metric;
    h.n_rows = s->n_rows;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    const uint64_t n = s->n_rows;
    std::vector<uint8_t> valid(n ? n : 1);
    if (ok && n) {
        cudaError_t e = cudaMemcpy(valid.data(), s->valid, n, cudaMemcpyDeviceToHost);
        ok = e == cudaSuccess && fwrite(valid.data(), 1, n, f) == n;
    }
    const uint64_t C = 1u << 16;
    float *pin = nullptr;
    if (ok && n) ok = cudaHostAlloc(&pin, C * s->dim * sizeof(float), cudaHostAllocDefault) == cudaSuccess;
    for (uint64_t at = 0; ok && at < n; at += C) {
        const uint64_t m = (n - at) < C ? (n - at) : C;
        cudaError_t e = cudaMemcpy2D(pin, s->dim * sizeof(float), s->X + at * s->ld, s->ld * sizeof(float),
                                     s->dim * sizeof(float), m, cudaMemcpyDeviceToHost);
        ok = e == cudaSuccess && fwrite(pin, sizeof(float), m * s->dim, f) == m * s->dim;
    }
    if (pin) cudaFreeHost(pin);
    ok = (fclose(f) == 0) && ok;
    cudaGetLastError();
    return ok ? SEMA_OK : fail(SEMA_ERR_CUDA, "writing %s failed", path);
}

int sema_index_load(const char *path, int device, uint64_t capacity_rows, sema_index **out)
{
    if (!path || !out) return fail(SEMA_ERR_INVALID, "null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return fail(SEMA_ERR_INVALID, "cannot open %s", path);
    SemaFileHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "SEMAIDX1", 8) != 0) {
        fclose(f);
        return fail(SEMA_ERR_INVALID, "%s is not a sema index file", path);
    }
    if (capacity_rows < h.n_rows) capacity_rows = h.n_rows;
    sema_index *s = nullptr;
    int rc = sema_index_create(device, h.dim, capacity_rows, h.metric, &s);
    if (rc) { fclose(f); return rc; }
    const uint64_t n = h.n_rows, C = 1u << 16;
    std::vector<uint8_t> valid(n ? n : 1);
    bool ok = n == 0 || fread(valid.data(), 1, n, f) == n;
    float *pin = nullptr;
    if (ok && n) ok = cudaHostAlloc(&pin, C * h.dim * sizeof(float), cudaHostAllocDefault) == cudaSuccess;
    for (uint64_t at = 0; ok && at < n; at += C) {
        const uint64_t m = (n - at) < C ? (n - at) : C;
        ok = fread(pin, sizeof(float), m * h.dim, f) == m * h.dim;
        if (ok) {
            rc = sema_index_append(s, pin, m, valid.data() + at, /*normalize=*/0, nullptr);   // rows are stored normalised
            ok = rc == SEMA_OK;
        }
    }
    if (pin) cudaFreeHost(pin);
    fclose(f);
    if (!ok) {
        sema_index_destroy(s);
        return rc ? rc : fail(SEMA_ERR_INVALID, "%s is truncated", path);
    }
    *out = s;
    return SEMA_OK;
}

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
    memcpy(s->q_pin, q, s->dim * sizeof(float));
    CK(cudaMemcpyAsync(s->q_dev, s->q_pin, s->ld * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    if (s->normalize_queries) {
        rc = normalize_queries_dev(s, s->q_dev, s->ld, 1);
        if (rc) return rc;
    }
    uint64_t *ids_d = reinterpret_cast<uint64_t *>(s->res_dev + 8);
    float *sc_d = reinterpret_cast<float *>(s->res_dev + 8 + 8 * (size_t)k);
    rc = scan_query(s, s->q_dev, (uint32_t)n, k, nullptr, ids_d, sc_d, reinterpret_cast<uint32_t *>(s->res_dev));
    if (rc) return rc;
    const size_t bytes = 8 + 12 * (size_t)k;
    CK(cudaMemcpyAsync(s->res_pin, s->res_dev, bytes, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const uint32_t nf = *reinterpret_cast<uint32_t *>(s->res_pin);
    *n_found = nf;
    memcpy(row_ids, s->res_pin + 8, nf * sizeof(uint64_t));
    memcpy(scores, s->res_pin + 8 + 8 * (size_t)k, nf * sizeof(float));
    return SEMA_OK;
}

int sema_index_search_batch(sema_index *s, const float *Q, uint32_t nq, uint32_t k,
                            uint64_t *row_ids, float *scores, uint32_t *n_found)
{
    if (!s || !n_found) return fail(SEMA_ERR_INVALID, "null argument");
    if (nq && !Q) return fail(SEMA_ERR_INVALID, "null queries");
    if (k > SEMA_MAX_K) return fail(SEMA_ERR_INVALID, "k %u > SEMA_MAX_K %u", k, SEMA_MAX_K);
    if (k && nq && (!row_ids || !scores)) return fail(SEMA_ERR_INVALID, "null output");
    CK(cudaSetDevice(s->device));
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    for (uint32_t i = 0; i < nq; ++i) n_found[i] = 0;
    if (k == 0 || n == 0 || nq == 0) return SEMA_OK;
    rc = ensure(reinterpret_cast<void **>(&s->Q_dev), &s->batch_cap_q, (size_t)nq * s->dim * sizeof(float));
    if (rc) return rc;
    if (s->batch_cap_res < (size_t)nq * k) {
        cudaFree(s->bids_dev); cudaFree(s->bsc_dev); cudaFree(s->bnf_dev);
        s->bids_dev = nullptr; s->bsc_dev = nullptr; s->bnf_dev = nullptr; s->batch_cap_res = 0;
        CK(cudaMalloc(&s->bids_dev, (size_t)nq * k * sizeof(uint64_t)));
        CK(cudaMalloc(&s->bsc_dev, (size_t)nq * k * sizeof(float)));
        CK(cudaMalloc(&s->bnf_dev, (size_t)nq * sizeof(uint32_t)));
        s->batch_cap_res = (size_t)nq * k;
    }
    CK(cudaMemcpyAsync(s->Q_dev, Q, (size_t)nq * s->dim * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    if (s->normalize_queries) {
        rc = normalize_queries_dev(s, s->Q_dev, s->dim, nq);
        if (rc) return rc;
    }
    rc = batch_core(s, s->Q_dev, nq, (uint32_t)n, k, s->bids_dev, s->bsc_dev, s->bnf_dev);
    if (rc) return rc;
    CK(cudaMemcpyAsync(row_ids, s->bids_dev, (size_t)nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(scores, s->bsc_dev, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaMemcpyAsync(n_found, s->bnf_dev, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SEMA_OK;
}

int sema_index_search_batch_device(sema_index *s, const float *Q_dev, uint32_t nq, uint32_t k,
                                   uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev)
{
    if (!s || !Q_dev || !ids_dev || !scores_dev || !n_found_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (k == 0 || k > SEMA_MAX_K || nq == 0) return fail(SEMA_ERR_INVALID, "k %u / nq %u out of range", k, nq);
    CK(cudaSetDevice(s->device));
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    if (n == 0) {
        CK(cudaMemsetAsync(n_found_dev, 0, (size_t)nq * sizeof(uint32_t), s->stream));
        return SEMA_OK;
    }
    return batch_core(s, Q_dev, nq, (uint32_t)n, k, ids_dev, scores_dev, n_found_dev);
}

int sema_index_set_normalize_queries(sema_index *s, int on)
{
    if (!s) return -1;
    if (on >= 0) s->normalize_queries = on ? 1 : 0;
    return s->normalize_queries;
}

int sema_index_set_batch_mode(sema_index *s, int mode)
{
    if (!s) return -1;
    if (mode >= 0 && mode <= 3) s->batch_mode = mode;
    return s->batch_mode;
}

int sema_index_batch_stats(const sema_index *s, uint64_t *k3_queries, uint64_t *k3_fallbacks)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (k3_queries) *k3_queries = s->k3_queries;
    if (k3_fallbacks) *k3_fallbacks = s->k3_fallbacks;
    return SEMA_OK;
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

int sema_index_set_row_base(sema_index *s, uint64_t row_base)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (row_base + s->capacity > 0xfffffffeull) return fail(SEMA_ERR_INVALID, "row_base + capacity must be < 2^32-1");
    s->row_base = (uint32_t)row_base;
    return SEMA_OK;
}

int sema_index_set_stream(sema_index *s, void *cuda_stream, int external)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    s->stream = external ? reinterpret_cast<cudaStream_t>(cuda_stream) : s->own_stream;
    return SEMA_OK;
}

uint64_t sema_index_size(const sema_index *s) { return s ? s->n_rows : 0; }
uint64_t sema_index_visible(sema_index *s)
{
    if (!s) return 0;
    cudaSetDevice(s->device);
    poll_ingest(s, false);
    return s->n_visible;
}
uint64_t sema_index_capacity(const sema_index *s) { return s ? s->capacity : 0; }
uint32_t sema_index_dim(const sema_index *s) { return s ? s->dim : 0; }
int sema_index_device(const sema_index *s) { return s ? s->device : -1; }
uint64_t sema_index_last_snapshot(const sema_index *s) { return s ? s->last_snapshot : 0; }
uint64_t sema_index_launch_count(const sema_index *s) { return s ? s->launches : 0; }

int sema_index_set_scan_variant(sema_index *s, int variant)
{
    if (!s) return -1;
    if (variant >= 200) { s->k3_qt = variant - 200; return variant; }        // 200 = auto, 201 = one query tile per CTA
    if (variant >= 100) { s->k3_cluster = variant - 100; return variant; }   // 100 = auto, 101/102/104 = K3 cluster size
    if (variant >= 0) s->variant = variant;
    return s->variant;
}

int sema_index_read_rows(sema_index *s, uint64_t first_row, uint64_t n, float *out)
{
    if (!s || (n && !out)) return fail(SEMA_ERR_INVALID, "null argument");
    int rc = sema_index_flush(s);
    if (rc) return rc;
    if (first_row + n > s->n_rows) return fail(SEMA_ERR_INVALID, "rows [%llu, %llu) outside size %llu",
                                               (unsigned long long)first_row, (unsigned long long)(first_row + n), (unsigned long long)s->n_rows);
    if (n == 0) return SEMA_OK;
    CK(cudaStreamSynchronize(s->stream));
    CK(cudaMemcpy2D(out, s->dim * sizeof(float), s->X + first_row * s->ld, s->ld * sizeof(float),
                    s->dim * sizeof(float), n, cudaMemcpyDeviceToHost));
    return SEMA_OK;
}

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

static int group_scan(sema_shard_group *g, const float *q_dev, uint32_t k, uint64_t *ids_dev, float *scores_dev,
                      uint32_t *nf_dev)
{
    sema_index *s = g->idx;
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    Exchange x;
    memset(&x, 0, sizeof x);
    for (uint32_t r = 0; r < g->world; ++r) x.peer[r] = g->peer[r];
    x.world = g->world;
    x.rank = g->rank;
    x.seq = ++g->seq;
    // an empty shard still takes part in the exchange: with n = 0 the kernel scans nothing
    // (one block) and goes straight to publish / wait / merge
    return scan_query(s, q_dev, (uint32_t)n, k, nullptr, ids_dev, scores_dev, nf_dev, &x);
}

int sema_shard_group_search_device(sema_shard_group *g, const float *q_dev, uint32_t k, uint64_t *ids_dev,
                                   float *scores_dev, uint32_t *n_found_dev)
{
    if (!g || !q_dev || !ids_dev || !scores_dev || !n_found_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (!g->connected) return fail(SEMA_ERR_INVALID, "shard group not connected");
    if (k == 0 || k > (uint32_t)K_PASS) return fail(SEMA_ERR_UNSUPPORTED, "fused shard search covers 1 <= k <= %d", K_PASS);
    sema_index *s = g->idx;
    CK(cudaSetDevice(s->device));
    const float *qd = q_dev;
    if (s->ld != s->dim || (reinterpret_cast<uintptr_t>(q_dev) & 15)) {
        CK(cudaMemcpyAsync(s->q_dev, q_dev, s->dim * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        qd = s->q_dev;
    }
    return group_scan(g, qd, k, ids_dev, scores_dev, n_found_dev);
}

int sema_shard_group_search(sema_shard_group *g, const float *q, uint32_t k, uint64_t *row_ids, float *scores,
                            uint32_t *n_found)
{
    if (!g || !q || !row_ids || !scores || !n_found) return fail(SEMA_ERR_INVALID, "null argument");
    if (!g->connected) return fail(SEMA_ERR_INVALID, "shard group not connected");
    if (k == 0 || k > (uint32_t)K_PASS) return fail(SEMA_ERR_UNSUPPORTED, "fused shard search covers 1 <= k <= %d", K_PASS);
    sema_index *s = g->idx;
    CK(cudaSetDevice(s->device));
    memcpy(s->q_pin, q, s->dim * sizeof(float));
    CK(cudaMemcpyAsync(s->q_dev, s->q_pin, s->ld * sizeof(float), cudaMemcpyHostToDevice, s->stream));
    if (s->normalize_queries) {
        int rc = normalize_queries_dev(s, s->q_dev, s->ld, 1);
        if (rc) return rc;
    }
    uint64_t *ids_d = reinterpret_cast<uint64_t *>(s->res_dev + 8);
    float *sc_d = reinterpret_cast<float *>(s->res_dev + 8 + 8 * (size_t)k);
    int rc = group_scan(g, s->q_dev, k, ids_d, sc_d, reinterpret_cast<uint32_t *>(s->res_dev));
    if (rc) return rc;
    const size_t bytes = 8 + 12 * (size_t)k;
    CK(cudaMemcpyAsync(s->res_pin, s->res_dev, bytes, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    const uint32_t nf = *reinterpret_cast<uint32_t *>(s->res_pin);
    if (nf == 0xffffffffu) return fail(SEMA_ERR_CUDA, "shard exchange timed out: a rank did not take part in search %llu",
                                       (unsigned long long)g->seq);
    *n_found = nf;
    memcpy(row_ids, s->res_pin + 8, nf * sizeof(uint64_t));
    memcpy(scores, s->res_pin + 8 + 8 * (size_t)k, nf * sizeof(float));
    return SEMA_OK;
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

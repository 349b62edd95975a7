// api_batch.cu — K3: 16-bit hi/lo planes (fp16 or bf16), cluster launches (mixed cluster sizes, CTA-pair variant),
// exact re-scoring, cascade and K2 fallback; batched entry points.
#include "index_impl.cuh"
#include "k3_batch.cuh"
#include "k3_pair.cuh"
#include "k4_merge.cuh"

using namespace sema;
using namespace sema_impl;

namespace {

// ---- K3 dispatch -----------------------------------------------------------------
constexpr uint32_t K3_MAX_K = 100;
// |tensor-core score - exact score| <= err_rel * |q| * max|x| + err_abs * |q|   (unit roundoff u = 2^-8 bf16, 2^-11 fp16)
//   3 passes (hi.hi + lo.hi + hi.lo): 3 u^2 (dropped lo.lo term and the residuals of the two-term splits) + fp32
//   accumulation over 3*dim <= 1152 products (allowance 2e-4);  1 pass: 2u + u^2 (both operands rounded) + accumulation
//   over dim <= 768 products (allowance 1.4e-4).
//   fp16 only: elements below 2^-14 are subnormal and carry an ABSOLUTE error <= 2^-25 per element and term instead
//   of a relative one: <= 3 * 2^-25 * sqrt(dim) * |q| over a row (queries are scaled into [0.5, 1), where the same
//   term is <= 2^-24 * sqrt(dim) * |q||x| and is part of err_rel's slack).
constexpr float K3_ERR_REL_BF16X3 = 2.5e-4f;
constexpr float K3_ERR_REL_BF16X1 = 8.5e-3f;
constexpr float K3_ERR_REL_FP16X3 = 2.0e-4f;
constexpr float K3_ERR_REL_FP16X1 = 1.15e-3f;
inline float k3_err_rel(int fmt, int passes)
{
    if (fmt == k3::FMT_FP16) return passes == 1 ? K3_ERR_REL_FP16X1 : K3_ERR_REL_FP16X3;
    return passes == 1 ? K3_ERR_REL_BF16X1 : K3_ERR_REL_BF16X3;
}
inline float k3_err_abs(int fmt, uint32_t dim)
{
    return fmt == k3::FMT_FP16 ? 3.0f * 2.98023224e-8f * sqrtf((float)dim) : 0.0f;
}

// passes the batch would run with: bf16x3 up to dim 384 (TMEM holds q_hi and q_lo), the single-pass
// filter up to dim 768 (q_hi only) or when asked for (batch mode 3); 0 = K3 cannot serve this shape
int k3_passes(const sema_index *s, uint32_t k)
{
    if (s->dim % k3::BLOCK_K != 0 || k > K3_MAX_K || s->planes_failed) return 0;
    if (s->metric == SEMA_METRIC_L2 && !s->l2_norms_constant) return 0;   // see k3_check_norms
    if (s->dim <= (uint32_t)k3::MAX_DIM) return s->batch_mode == 3 ? 1 : 3;
    if (s->dim <= (uint32_t)k3::MAX_DIM_1PASS) return 1;
    return 0;
}

// L2 metric: K3 selects candidates by dot product, which ranks like the distance only when the row
// norms are (nearly) constant — the reference's case, every stored vector is unit-norm
// (src/semantic/embeddings.rs:83-88).  K1 tracks [max, min] |x|^2; anything else stays on K2.
int k3_check_norms(sema_index *s)
{
    if (s->metric != SEMA_METRIC_L2) return SEMA_OK;
    float mm[2] = {0.0f, 0.0f};
    CK(cudaMemcpyAsync(mm, s->max_norm2, sizeof mm, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    s->l2_norms_constant = mm[1] > 0.0f && mm[0] < __builtin_inff() && (mm[0] - mm[1]) <= 1e-3f * mm[0];
    return SEMA_OK;
}

// Element format of the planes.  fp16 (11 significant bits) gives every stage an 8x tighter error bound than bf16
// at the same tensor-core cost, but its range is narrow: it is used when every stored element is <= 1024 in
// magnitude (max |x|^2 <= 2^20 — the reference's unit-norm rows by a wide margin); otherwise bf16.
int k3_choose_fmt(sema_index *s, int *fmt)
{
    if (s->k3_prec == 1) { *fmt = k3::FMT_BF16; return SEMA_OK; }
    float mm[2] = {0.0f, 0.0f};
    CK(cudaMemcpyAsync(mm, s->max_norm2, sizeof mm, cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    *fmt = (mm[0] <= 1048576.0f) ? k3::FMT_FP16 : k3::FMT_BF16;
    return SEMA_OK;
}

// Bring the 16-bit planes up to date with rows [0, n).  Tombstones are poisoned in place (k3_poison_rows).
int k3_sync_planes(sema_index *s, uint64_t n)
{
    SEMA_NVTX("sema.K3.split_planes");
    if (!s->planes || s->planes_rows < n || s->planes_prec != s->k3_prec) {   // new rows (or a new preference): (re)decide the format
        int fmt = k3::FMT_BF16;
        int rc = k3_choose_fmt(s, &fmt);
        if (rc) return rc;
        if (fmt != s->planes_fmt) { s->planes_fmt = fmt; s->planes_rows = 0; }   // re-split everything in the other format
        s->planes_prec = s->k3_prec;
    }
    if (!s->planes) {
        const uint64_t tiles = (s->capacity + k3::TILE_N - 1) / k3::TILE_N;
        const size_t bytes = (size_t)(tiles ? tiles : 1) * k3::tile_bytes((int)s->dim);
        if (s->growable) {
            if (growbuf_reserve(s->gPlanes, s->device, bytes, (size_t)256 << 20) != SEMA_OK) {
                s->planes_failed = true;
                return SEMA_ERR_NOMEM;
            }
            s->planes = reinterpret_cast<unsigned char *>(s->gPlanes.base);
        } else {
            cudaError_t e = cudaMalloc(&s->planes, bytes);
            if (e != cudaSuccess) {
                cudaGetLastError();
                s->planes = nullptr;
                s->planes_failed = true;  // not an error: the K2 loop serves the batch instead
                return SEMA_ERR_NOMEM;
            }
        }
        s->planes_rows = 0;
    }
    if (s->growable) {   // physical memory for the tiles that exist
        const uint64_t tiles_n = (n + k3::TILE_N - 1) / k3::TILE_N;
        int rc = growbuf_commit(s->gPlanes, (size_t)tiles_n * k3::tile_bytes((int)s->dim));
        if (rc == SEMA_ERR_NOMEM) { s->planes_failed = true; return SEMA_ERR_NOMEM; }
        if (rc) return rc;
    }
    if (s->planes_rows >= n) return SEMA_OK;
    const uint64_t begin = (s->planes_rows / k3::TILE_N) * k3::TILE_N;            // re-tile the partial last tile
    const uint64_t end = ((n + k3::TILE_N - 1) / k3::TILE_N) * k3::TILE_N;
    const uint64_t work = (end - begin) * (s->dim / 8);
    uint64_t blocks = (work + 255) / 256;
    if (blocks > (uint64_t)s->num_sms * 32) blocks = (uint64_t)s->num_sms * 32;
    k3::split_planes_kernel<<<(unsigned)blocks, 256, 0, s->stream>>>(s->X, s->ld, s->dim, begin, end, n, s->planes, s->planes_fmt);
    CK(cudaGetLastError());
    s->launches++;
    s->planes_rows = n;
    return SEMA_OK;
}

// q_ctas = CTAs along the query axis (each owns QT query tiles)
template <int KC, int C, int PASSES, int QT>
int k3_launch_scan_c(sema_index *s, const k3::Params &p, uint32_t q_ctas, uint32_t grid_parts, cudaStream_t stream)
{
    auto kern = k3::batch_scan_kernel<KC, C, PASSES, QT>;
    static bool attr_set[64] = {false};
    if (!attr_set[s->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, k3::Smem<KC, PASSES, QT>::TOTAL));
        attr_set[s->device & 63] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(q_ctas, grid_parts, 1);
    cfg.blockDim = dim3(k3::THREADS, 1, 1);
    cfg.dynamicSmemBytes = k3::Smem<KC, PASSES, QT>::TOTAL;
    cfg.stream = stream;
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

// ---- CTA-pair form of the single-pass stage (k3_pair.cuh): grid (query tiles, parts), clusters of 2 CTAs
inline int k3_pair_smem(int kc, uint32_t dim)
{
    const int need = k3::pair_smem_bytes(kc, (int)dim);
    return need < 116 * 1024 ? 116 * 1024 : need;      // never two CTAs per SM: each allocates all 512 TMEM columns
}
// The planes as the pair kernel's 3-D tensor: 8-byte elements, {256 (one 2 KB row), 4 (rows of an 8 KB hi block),
// groups of 16 KB (the hi|lo pair of one k-block of one tile)}; box = the hi blocks of pair_kps(dim) consecutive
// k-blocks.  Encoded once per planes allocation (the growable index reserves its whole address range up front).
int k3_pair_tmap(sema_index *s)
{
    if (s->tmap_base == s->planes) return SEMA_OK;
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            return fail(SEMA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
        }
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    const uint64_t tiles = (s->capacity + k3::TILE_N - 1) / k3::TILE_N;
    const cuuint64_t gdim[3] = {256, 4, (tiles ? tiles : 1) * (s->dim / k3::BLOCK_K)};
    const cuuint64_t gstride[2] = {2048, (cuuint64_t)k3::STAGE_BYTES};
    const cuuint32_t box[3] = {256, 4, (cuuint32_t)k3::pair_kps((int)s->dim)};
    const cuuint32_t estride[3] = {1, 1, 1};
    CUresult r = encode(&s->planes_tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, s->planes, gdim, gstride, box, estride,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SEMA_ERR_CUDA, "cuTensorMapEncodeTiled failed for the K3 planes (CUresult %d)", (int)r);
    s->tmap_base = s->planes;
    return SEMA_OK;
}

template <int KC>
int k3_launch_pair(sema_index *s, const k3::Params &p, uint32_t q_tiles)
{
    int trc = k3_pair_tmap(s);
    if (trc) return trc;
    auto kern = k3::pair_scan_kernel<KC>;
    static bool attr_set[64] = {false};
    if (!attr_set[s->device & 63]) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[s->device & 63] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(q_tiles, p.parts, 1);      // the pair kernel always runs a stage as ONE launch: grid.y = all partitions
    cfg.blockDim = dim3(k3::PAIR_THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)k3_pair_smem(KC, p.dim);
    cfg.stream = s->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, p, s->planes_tmap));
    s->launches++;
    return SEMA_OK;
}
template <int KC>
int k3_pair_clusters(sema_index *s, int *out)
{
    static int cached[64] = {0};
    int &v = cached[s->device & 63];
    if (v == 0) {
        auto kern = k3::pair_scan_kernel<KC>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2, (unsigned)s->num_sms, 1);
        cfg.blockDim = dim3(k3::PAIR_THREADS, 1, 1);
        cfg.dynamicSmemBytes = 227 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        CK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
        v = n > 0 ? n : 1;
    }
    *out = v;
    return SEMA_OK;
}

template <int KC, int PASSES, int QT>
int k3_launch_scan_p(sema_index *s, const k3::Params &p, uint32_t q_ctas, int c, uint32_t grid_parts, cudaStream_t stream)
{
    return c == 4 ? k3_launch_scan_c<KC, 4, PASSES, QT>(s, p, q_ctas, grid_parts, stream)
         : c == 2 ? k3_launch_scan_c<KC, 2, PASSES, QT>(s, p, q_ctas, grid_parts, stream)
                  : k3_launch_scan_c<KC, 1, PASSES, QT>(s, p, q_ctas, grid_parts, stream);
}
template <int KC>
int k3_launch_scan(sema_index *s, const k3::Params &p, uint32_t q_ctas, int c, int passes, int qt, uint32_t grid_parts,
                   cudaStream_t stream)
{
    if (passes == 3) return k3_launch_scan_p<KC, 3, 1>(s, p, q_ctas, c, grid_parts, stream);
    if constexpr (KC <= 64) {
        if (qt == 2) return k3_launch_scan_p<KC, 1, 2>(s, p, q_ctas, c, grid_parts, stream);
    }
    return k3_launch_scan_p<KC, 1, 1>(s, p, q_ctas, c, grid_parts, stream);
}
int k3_launch_scan_kc(sema_index *s, uint32_t kc, const k3::Params &p, uint32_t q_ctas, int c, int passes, int qt,
                      uint32_t grid_parts, cudaStream_t stream)
{
    return kc == 16 ? k3_launch_scan<16>(s, p, q_ctas, c, passes, qt, grid_parts, stream)
         : kc == 32 ? k3_launch_scan<32>(s, p, q_ctas, c, passes, qt, grid_parts, stream)
         : kc == 64 ? k3_launch_scan<64>(s, p, q_ctas, c, passes, qt, grid_parts, stream)
                    : k3_launch_scan<128>(s, p, q_ctas, c, passes, qt, grid_parts, stream);
}
int k3_clusters_kc(sema_index *s, uint32_t kc, int c, int *out);   // below
template <int KC>
int k3_clusters(sema_index *s, int c, int *out)
{
    return c == 4 ? k3_max_clusters<KC, 4>(s, out) : c == 2 ? k3_max_clusters<KC, 2>(s, out) : k3_max_clusters<KC, 1>(s, out);
}
int k3_clusters_kc(sema_index *s, uint32_t kc, int c, int *out)
{
    return kc == 16 ? k3_clusters<16>(s, c, out) : kc == 32 ? k3_clusters<32>(s, c, out) : kc == 64 ? k3_clusters<64>(s, c, out)
                    : k3_clusters<128>(s, c, out);
}

// Qd: nq x dim dense on the device.  Results: device arrays [nq*k], [nq*k], [nq].
// One K3 stage over nq device-resident queries: tensor-core scan with `passes` bf16 passes, exact fp32
// re-scoring, exactness proof.  On return (stream synchronised) the results are in ids_d / sc_d / nf_d
// and s->flags_pin[i] != 0 marks the queries whose proof failed; s->Qpad_dev holds the queries.
int k3_stage(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, int passes, uint64_t *ids_d,
             float *sc_d, uint32_t *nf_d)
{
    SEMA_NVTX("sema.K3.stage(scan+rescore)");
    // candidates kept per (query, row partition): the list minimum is the admission threshold and every
    // insertion re-scans the list, so the list is only as long as the exactness proof needs
    // Single-pass stage: SHORT lists.  Insertions per list grow like KC ln(rows / KC) and each one re-scans the list, so
    // the epilogue's work grows like KC^2: with KC = 128 a k = 50 batch took 17.8 ms against 5.4 ms at k = 10.  A
    // partition only needs its list to reach below the global k-th score, and with ~37 partitions it holds k / 37 of the
    // top k on average, so 16 (k <= 10), 32 (k <= 50) or 64 entries prove every query unless the neighbours cluster in
    // one partition (rows of one document are adjacent) — those queries fail the proof and the cascade re-runs them in
    // the three-pass stage, whose lists are long (32 / 64 / 128).
    const uint32_t kc = passes == 1 ? ((k <= 10 && s->k3_kc16 != 0) ? 16 : (k <= 50 ? 32 : 64))
                                    : (k <= 16 ? 32 : (k <= 48 ? 64 : 128));
    const uint32_t n_tiles = (n + k3::TILE_N - 1) / k3::TILE_N;
    const uint32_t q_tiles_all = (nq + k3::TILE_Q - 1) / k3::TILE_Q;
    int rc;
    // QT query tiles per CTA (2 in the single-pass mode when there are at least 2 tiles), clusters of
    // csize CTAs along the query axis: the query-tile count is padded to a multiple of QT*csize
    // (Qpad rows beyond nq are zero queries whose results are never read).
    // (two tiles need 2*dim/2 + 128 TMEM columns and two candidate lists in shared memory)
    // The single-pass stage runs as CTA pairs (k3_pair.cuh: tcgen05 cta_group::2, M = 256 x N = 128) whenever there
    // are at least two query tiles and the queries plus two 128-column accumulators fit TMEM; pp = pairs per cluster.
    const bool pair = passes == 1 && s->k3_pair != 0 && q_tiles_all >= 2 && s->dim <= (uint32_t)k3::PAIR_MAX_DIM && kc <= 64 &&
                      k3::pair_stages((int)kc, (int)s->dim) >= 2 * k3::pair_spt((int)s->dim);
    const int qt_per_cta = pair ? 1 : ((passes == 1 && q_tiles_all >= 2 && s->k3_qt != 1 && s->dim <= 384 && kc <= 64) ? 2 : 1);
    const uint32_t q_ctas_all = (q_tiles_all + qt_per_cta - 1) / qt_per_cta;
    // measured on 10M x 384 x 1024q (alternating A/B, same idle gap before every measurement): clusters of 2 are
    // fastest for both stages — single pass 5.64 ms (clusters of 4: 5.83, no cluster: 6.10), three passes 14.85 ms
    // (clusters of 4: 15.29 on 132 of the 148 SMs, no cluster: 17.4)
    const int cpref = 2;
    const int csize = pair ? 2
                           : (s->k3_cluster > 0 ? s->k3_cluster : (q_ctas_all >= (uint32_t)cpref ? cpref : (q_ctas_all >= 2 ? 2 : 1)));
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
    if (pair)
        rc = kc == 16 ? k3_pair_clusters<16>(s, &max_clusters) : kc == 32 ? k3_pair_clusters<32>(s, &max_clusters) : k3_pair_clusters<64>(s, &max_clusters);
    else
        rc = k3_clusters_kc(s, kc, csize, &max_clusters);
    if (rc) return rc;
    const uint32_t groups_all = q_ctas_pad / csize;             // clusters along the query axis

    // Mixed cluster sizes.  Clusters of 4 are the most efficient shape per SM (one fetch from L2 serves four CTAs and no
    // two clusters stream the same rows), but only 33 of them fit the GPCs of a B200 (132 SMs); clusters of 2 reach all
    // 148 SMs.  So a stage is issued as TWO concurrent launches: clusters of 4 over as many row partitions as fit, and —
    // on a second stream — clusters of 2 on the SMs the first launch leaves free (GPC remainders), each launch with its
    // own slice of the rows (a 4-cluster partition gets 1.05x the rows of a 2-cluster one; measured flat between 1.01 and 1.09).
    // 10M x 384 x 1024q, alternating A/B: single pass 5.67 -> 5.61 ms, three passes 15.12 -> 14.55 ms.
    uint32_t mixA = 0, mixB = 0;
    if (!pair && s->k3_mixed != 0 && s->k3_cluster == 0 && csize == 2 && q_ctas_pad % 4 == 0 && groups_all <= (uint32_t)max_clusters) {
        int max4 = 0;
        rc = k3_clusters_kc(s, kc, 4, &max4);
        if (rc) return rc;
        const uint32_t gA = q_ctas_pad / 4, gB = q_ctas_pad / 2;
        mixA = (uint32_t)max4 / gA;
        const uint32_t left = (uint32_t)s->num_sms > mixA * gA * 4 ? (uint32_t)s->num_sms - mixA * gA * 4 : 0;
        mixB = (left / 2) / gB;
        if (mixA == 0 || mixB == 0 || n_tiles < 16 * (mixA + mixB)) mixA = mixB = 0;
    }
    if (mixA && (!s->aux_stream)) {
        CK(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    }
    // at most max_clusters cluster columns per launch; the clusters left over become row partitions
    for (uint32_t g0 = 0; g0 < groups_all; g0 += (uint32_t)max_clusters) {
        const uint32_t groups = (groups_all - g0) < (uint32_t)max_clusters ? (groups_all - g0) : (uint32_t)max_clusters;
        const uint32_t q_ctas = groups * csize;                 // CTAs along the query axis in this launch
        const uint32_t qt0 = g0 * csize * qt_per_cta;           // first query tile of this launch
        const uint32_t q_tiles = q_ctas * qt_per_cta;
        uint32_t parts = mixA ? mixA + mixB : (uint32_t)max_clusters / groups;
        if (parts > n_tiles) parts = n_tiles;
        if (parts < 1) parts = 1;
        const size_t nqp = (size_t)q_tiles * k3::TILE_Q;
        rc = ensure(reinterpret_cast<void **>(&s->cand_rows), &s->cand_cap, nqp * parts * kc * sizeof(uint32_t));
        if (rc) return rc;
        rc = ensure(reinterpret_cast<void **>(&s->cand_thr), &s->thr_cap, nqp * parts * sizeof(float));
        if (rc) return rc;
        rc = ensure(reinterpret_cast<void **>(&s->cand_sc), &s->cand_sc_cap, nqp * parts * kc * sizeof(float));
        if (rc) return rc;
        k3::Params p;
        p.planes = s->planes;
        p.Q = s->Qpad_dev + (size_t)qt0 * k3::TILE_Q * s->dim;
        p.cand_rows = s->cand_rows;
        p.cand_sc = s->cand_sc;
        p.cand_thr = s->cand_thr;
        p.n_rows = n;
        p.n_tiles = n_tiles;
        p.parts = parts;
        p.part_base = 0;
        p.tile_first = 0;
        p.tile_per = (n_tiles + parts - 1) / parts;
        p.tile_end = n_tiles;
        p.dim = s->dim;
        p.fmt = (uint32_t)s->planes_fmt;
        p.prefetch = (uint32_t)s->k3_prefetch;
        p.debug = (uint32_t)s->k3_debug;
        if (pair) {
            const uint32_t n_super = (n_tiles + 1) / 2;
            p.tile_per = 2 * ((n_super + parts - 1) / parts);   // whole super-tiles (two 64-row tiles) per partition
            rc = kc == 16 ? k3_launch_pair<16>(s, p, q_ctas) : kc == 32 ? k3_launch_pair<32>(s, p, q_ctas) : k3_launch_pair<64>(s, p, q_ctas);
        } else if (mixA) {
            const double wA = s->k3_mix_w > 0 ? 0.70 + s->k3_mix_w / 100.0 : 1.05;   // rows of a 4-cluster partition per row of a 2-cluster partition
            uint32_t perA = (uint32_t)((double)n_tiles * wA / (wA * mixA + mixB) + 0.999);
            if ((uint64_t)perA * mixA > n_tiles) perA = n_tiles / mixA;
            const uint32_t restB = n_tiles - perA * mixA;
            k3::Params pa = p, pb = p;
            pa.tile_per = perA;
            pa.tile_end = perA * mixA;
            pb.part_base = mixA;
            pb.tile_first = perA * mixA;
            pb.tile_per = (restB + mixB - 1) / mixB;
            CK(cudaEventRecord(s->ev_fork, s->stream));          // the second launch also needs the padded queries
            CK(cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0));
            rc = k3_launch_scan_kc(s, kc, pa, q_ctas, 4, passes, qt_per_cta, mixA, s->stream);
            if (rc) return rc;
            rc = k3_launch_scan_kc(s, kc, pb, q_ctas, 2, passes, qt_per_cta, mixB, s->aux_stream);
            if (rc) return rc;
            CK(cudaEventRecord(s->ev_join, s->aux_stream));
            CK(cudaStreamWaitEvent(s->stream, s->ev_join, 0));   // the re-scoring reads both launches' candidate lists
        } else {
            rc = k3_launch_scan_kc(s, kc, p, q_ctas, csize, passes, qt_per_cta, parts, s->stream);
        }
        if (rc) return rc;
        const uint32_t q_first = qt0 * k3::TILE_Q;
        if (q_first >= nq) break;
        const uint32_t q_cnt = (nq - q_first) < (uint32_t)nqp ? (nq - q_first) : (uint32_t)nqp;
        k3::RescoreParams r;
        r.X = reinterpret_cast<const float4 *>(s->X);
        r.Q = p.Q;
        r.cand_rows = s->cand_rows;
        r.cand_sc = s->cand_sc;
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
        r.err_rel = k3_err_rel(s->planes_fmt, passes);
        r.err_abs = k3_err_abs(s->planes_fmt, s->dim);
        if (s->metric == SEMA_METRIC_L2) {
            if (k <= 32) k3::rescore_kernel<1, METRIC_L2><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
            else if (k <= 64) k3::rescore_kernel<2, METRIC_L2><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
            else k3::rescore_kernel<4, METRIC_L2><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
        } else {
            if (k <= 32) k3::rescore_kernel<1, METRIC_COSINE><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
            else if (k <= 64) k3::rescore_kernel<2, METRIC_COSINE><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
            else k3::rescore_kernel<4, METRIC_COSINE><<<q_cnt, SCAN_THREADS, 0, s->stream>>>(r);
        }
        CK(cudaGetLastError());
        s->launches++;
    }
    CK(cudaMemcpyAsync(s->flags_pin, s->flags_dev, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return SEMA_OK;
}

// K3 for a batch.  Modes 2 / 3 run one stage (bf16x3 / single pass) and send the queries whose
// exactness proof failed through K2.  Automatic mode is a precision cascade: the cheap single-pass
// filter first (a third of the tensor work); the queries it cannot prove (corpora with dense
// neighbourhoods: its error bound is 34x looser) are gathered and re-run with the bf16x3 split; what
// even that cannot prove (exact ties beyond the candidate list) goes through K2.  Every path ends in
// the same fp32 re-scoring, so the results are identical.
int k3_batch(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
             uint32_t *nf_d)
{
    SEMA_NVTX("sema.K3.batch");
    if (reinterpret_cast<uintptr_t>(Qd) & 15) {   // the kernels read queries as float4
        int rc0 = ensure(reinterpret_cast<void **>(&s->q_aligned), &s->q_aligned_cap, (size_t)nq * s->dim * sizeof(float));
        if (rc0) return rc0;
        CK(cudaMemcpyAsync(s->q_aligned, Qd, (size_t)nq * s->dim * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        Qd = s->q_aligned;
    }
    const bool can3 = s->dim <= (uint32_t)k3::MAX_DIM;
    const bool cascade = s->batch_mode == 0 && can3;
    const int passes = cascade ? 1 : k3_passes(s, k);
    int rc = k3_stage(s, Qd, nq, n, k, passes, ids_d, sc_d, nf_d);
    if (rc) return rc;
    s->k3_queries += nq;
    std::vector<uint32_t> open_q;   // queries still without a proven result
    for (uint32_t i = 0; i < nq; ++i)
        if (s->flags_pin[i]) open_q.push_back(i);
    if (cascade && open_q.size() >= 4) {
        s->k3_cascaded += open_q.size();
        const uint32_t nsub = (uint32_t)open_q.size();
        if ((size_t)nsub * 2 > nq) {
            // most of the batch failed the loose bound: run the whole batch with the tight one
            rc = k3_stage(s, Qd, nq, n, k, 3, ids_d, sc_d, nf_d);
            if (rc) return rc;
            open_q.clear();
            for (uint32_t i = 0; i < nq; ++i)
                if (s->flags_pin[i]) open_q.push_back(i);
        } else {
            rc = ensure(reinterpret_cast<void **>(&s->sub_q), &s->sub_q_cap, (size_t)nsub * s->dim * sizeof(float));
            if (rc) return rc;
            rc = ensure(reinterpret_cast<void **>(&s->sub_ids), &s->sub_ids_cap, (size_t)nsub * k * sizeof(uint64_t));
            if (rc) return rc;
            rc = ensure(reinterpret_cast<void **>(&s->sub_sc), &s->sub_sc_cap, (size_t)nsub * k * sizeof(float));
            if (rc) return rc;
            rc = ensure(reinterpret_cast<void **>(&s->sub_nf), &s->sub_nf_cap, (size_t)nsub * sizeof(uint32_t));
            if (rc) return rc;
            // one list upload + one gather kernel (was: a small device-to-device copy per query)
            rc = ensure(reinterpret_cast<void **>(&s->sub_idx), &s->sub_idx_cap, (size_t)nsub * sizeof(uint32_t));
            if (rc) return rc;
            CK(cudaMemcpyAsync(s->sub_idx, open_q.data(), (size_t)nsub * sizeof(uint32_t), cudaMemcpyHostToDevice, s->stream));
            {
                const uint64_t work = (uint64_t)nsub * (s->dim / 4);
                const unsigned blocks = (unsigned)((work + 255) / 256 < 1024 ? (work + 255) / 256 : 1024);
                k3::gather_queries_kernel<<<blocks, 256, 0, s->stream>>>(reinterpret_cast<const float4 *>(Qd), s->sub_idx, nsub,
                                                                         s->dim / 4, reinterpret_cast<float4 *>(s->sub_q));
                CK(cudaGetLastError());
                s->launches++;
            }
            rc = k3_stage(s, s->sub_q, nsub, n, k, 3, s->sub_ids, s->sub_sc, s->sub_nf);
            if (rc) return rc;
            // proven sub-batch results go back to their places in one kernel (flags_dev holds the sub-batch's flags)
            k3::scatter_results_kernel<<<nsub, 128, 0, s->stream>>>(s->sub_ids, s->sub_sc, s->sub_nf, s->sub_idx, s->flags_dev, k,
                                                                    ids_d, sc_d, nf_d);
            CK(cudaGetLastError());
            s->launches++;
            std::vector<uint32_t> still;
            for (uint32_t j = 0; j < nsub; ++j)
                if (s->flags_pin[j]) still.push_back(open_q[j]);
            open_q.swap(still);
        }
    }
    // whatever no tensor-core stage could prove (heavy exact ties, near-duplicates) goes through K2
    for (uint32_t i : open_q) {
        s->k3_fallbacks++;
        rc = scan_query(s, Qd + (size_t)i * s->dim, n, k, nullptr, ids_d + (size_t)i * k, sc_d + (size_t)i * k, nf_d + i);
        if (rc) return rc;
    }
    return SEMA_OK;
}

}  // namespace

namespace sema_impl {

int k3_poison_rows(sema_index *s, const uint64_t *rows_dev, uint64_t n)
{
    if (!s->planes || s->planes_rows == 0 || n == 0) return SEMA_OK;
    const uint64_t work = n * (s->dim / 8);
    uint64_t blocks = (work + 255) / 256;
    if (blocks > (uint64_t)s->num_sms * 8) blocks = (uint64_t)s->num_sms * 8;
    k3::poison_planes_kernel<<<(unsigned)blocks, 256, 0, s->stream>>>(s->planes, s->dim, rows_dev, n, s->planes_rows);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

// Batched search with the queries already on the device (nq x dim dense).
int batch_core(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
               uint32_t *nf_d)
{
    const bool want_k3 = s->batch_mode >= 2 || (s->batch_mode == 0 && nq >= 4);
    if (want_k3) {
        int rc = k3_check_norms(s);
        if (rc) return rc;
    }
    if (want_k3 && k3_passes(s, k) != 0) {
        int rc = k3_sync_planes(s, n);
        if (rc == SEMA_OK) return k3_batch(s, Qd, nq, n, k, ids_d, sc_d, nf_d);
        if (rc != SEMA_ERR_NOMEM) return rc;
    }
    return scan_stream(s, Qd, nq, n, k, ids_d, sc_d, nf_d, nullptr, nullptr);
}

// K2 once per query (one HBM pass each), issued back to back.  With the TMA kernel and k <= 128
// launch i+1 is chained to launch i (programmatic dependent launch): its scan starts on the SMs
// launch i has left while i's last block still merges (and, sharded, exchanges with the peers).
// x / seq: optional shard exchange; *seq is advanced once per query.
int scan_stream(sema_index *s, const float *Qd, uint32_t nq, uint32_t n, uint32_t k, uint64_t *ids_d, float *sc_d,
                uint32_t *nf_d, sema::Exchange *x, uint64_t *seq)
{
    const float *Qp = Qd;
    if (s->ld != s->dim || (reinterpret_cast<uintptr_t>(Qd) & 15)) {
        int rc = ensure(reinterpret_cast<void **>(&s->Qpad_dev), &s->qpad_cap, (size_t)nq * s->ld * sizeof(float));
        if (rc) return rc;
        CK(cudaMemsetAsync(s->Qpad_dev, 0, (size_t)nq * s->ld * sizeof(float), s->stream));
        CK(cudaMemcpy2DAsync(s->Qpad_dev, s->ld * sizeof(float), Qd, s->dim * sizeof(float), s->dim * sizeof(float),
                             nq, cudaMemcpyDeviceToDevice, s->stream));
        Qp = s->Qpad_dev;
    }
    if (stream_kernel_ok(s, nq, n, k)) {
        if (x) x->seq = *seq + 1;  // query i exchanges under sequence number *seq + 1 + i, as the launch-per-query form does
        const int rc = stream_kernel_launch(s, Qp, nq, n, k, ids_d, sc_d, nf_d, x);
        if (rc != STREAM_FALLBACK) {
            if (x && rc == SEMA_OK) *seq += nq;
            return rc;
        }
    }
    for (uint32_t i = 0; i < nq; ++i) {
        if (x) x->seq = ++*seq;
        const unsigned flags = (i > 0 && s->chain) ? SCAN_CHAINED : 0u;
        int rc = scan_query(s, Qp + (size_t)i * s->ld, n, k, nullptr, ids_d + (size_t)i * k, sc_d + (size_t)i * k, nf_d + i,
                            x, flags);
        if (rc) return rc;
    }
    return SEMA_OK;
}

}  // namespace sema_impl

namespace {

// the handle's own device buffers for a batch result: ids / scores [nq][k], n_found [nq]
int ensure_batch_results(sema_index *s, uint32_t nq, uint32_t k)
{
    if (s->batch_cap_res < (size_t)nq * k) {
        cudaFree(s->bids_dev); cudaFree(s->bsc_dev);
        s->bids_dev = nullptr; s->bsc_dev = nullptr; s->batch_cap_res = 0;
        CK(cudaMalloc(&s->bids_dev, (size_t)nq * k * sizeof(uint64_t)));
        CK(cudaMalloc(&s->bsc_dev, (size_t)nq * k * sizeof(float)));
        s->batch_cap_res = (size_t)nq * k;
    }
    if (s->batch_cap_nf < nq) {
        cudaFree(s->bnf_dev);
        s->bnf_dev = nullptr; s->batch_cap_nf = 0;
        CK(cudaMalloc(&s->bnf_dev, (size_t)nq * sizeof(uint32_t)));
        s->batch_cap_nf = nq;
    }
    return SEMA_OK;
}

}  // namespace

extern "C" {

int sema_index_search_batch_keys_device(sema_index *s, const float *Q_dev, uint32_t nq, uint32_t k, uint64_t *keys_dev)
{
    if (!s || !Q_dev || !keys_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (k == 0 || k > SEMA_MAX_K || nq == 0) return fail(SEMA_ERR_INVALID, "k %u / nq %u out of range", k, nq);
    CK(cudaSetDevice(s->device));
    int rc = poll_ingest(s, false);
    if (rc) return rc;
    const uint64_t n = s->n_visible;
    s->last_snapshot = n;
    if (n == 0) {
        CK(cudaMemsetAsync(keys_dev, 0, (size_t)nq * k * sizeof(uint64_t), s->stream));
        return SEMA_OK;
    }
    rc = ensure_batch_results(s, nq, k);
    if (rc) return rc;
    rc = batch_core(s, Q_dev, nq, (uint32_t)n, k, s->bids_dev, s->bsc_dev, s->bnf_dev);
    if (rc) return rc;
    const size_t total = (size_t)nq * k;
    const unsigned blocks = (unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    if (s->metric == SEMA_METRIC_L2)
        pack_keys_kernel<METRIC_L2><<<blocks, 256, 0, s->stream>>>(s->bids_dev, s->bsc_dev, s->bnf_dev, nq, k, keys_dev);
    else
        pack_keys_kernel<METRIC_COSINE><<<blocks, 256, 0, s->stream>>>(s->bids_dev, s->bsc_dev, s->bnf_dev, nq, k, keys_dev);
    CK(cudaGetLastError());
    s->launches++;
    return SEMA_OK;
}

int sema_topk_merge_batch_device(sema_index *s, const uint64_t *keys_dev, uint32_t n_lists, uint32_t nq, uint32_t k,
                                 uint64_t *ids_dev, float *scores_dev, uint32_t *n_found_dev)
{
    SEMA_NVTX("sema.K4.merge_batch");
    if (!s || !keys_dev || !ids_dev || !scores_dev || !n_found_dev) return fail(SEMA_ERR_INVALID, "null argument");
    if (k == 0 || nq == 0 || n_lists == 0) return fail(SEMA_ERR_INVALID, "k %u / nq %u / n_lists %u out of range", k, nq, n_lists);
    if (k > (uint32_t)K_PASS) return fail(SEMA_ERR_UNSUPPORTED, "the batched merge covers k <= %d", K_PASS);
    if ((uint64_t)n_lists * k > 0x7fffffffull) return fail(SEMA_ERR_INVALID, "too many candidates");
    CK(cudaSetDevice(s->device));
    const bool l2 = s->metric == SEMA_METRIC_L2;
#define SEMA_MERGE_BATCH(M)                                                                                             \
    do {                                                                                                                \
        if (l2) merge_topk_batch_kernel<M, METRIC_L2><<<nq, MERGE_THREADS, 0, s->stream>>>(keys_dev, n_lists, nq, (int)k, ids_dev, scores_dev, n_found_dev); \
        else merge_topk_batch_kernel<M, METRIC_COSINE><<<nq, MERGE_THREADS, 0, s->stream>>>(keys_dev, n_lists, nq, (int)k, ids_dev, scores_dev, n_found_dev); \
    } while (0)
    if (k <= 32) SEMA_MERGE_BATCH(1);
    else if (k <= 64) SEMA_MERGE_BATCH(2);
    else SEMA_MERGE_BATCH(4);
#undef SEMA_MERGE_BATCH
    CK(cudaGetLastError());
    s->launches++;
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
    rc = ensure_batch_results(s, nq, k);
    if (rc) return rc;
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

int sema_index_search_stream_device(sema_index *s, const float *Q_dev, uint32_t nq, uint32_t k,
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
    return scan_stream(s, Q_dev, nq, (uint32_t)n, k, ids_dev, scores_dev, n_found_dev, nullptr, nullptr);
}

int sema_index_set_batch_mode(sema_index *s, int mode)
{
    if (!s) return -1;
    if (mode >= 0 && mode <= 3) s->batch_mode = mode;
    return s->batch_mode;
}

int sema_index_batch_stats(const sema_index *s, uint64_t *k3_queries, uint64_t *k3_fallbacks, uint64_t *k3_cascaded)
{
    if (!s) return fail(SEMA_ERR_INVALID, "null index");
    if (k3_queries) *k3_queries = s->k3_queries;
    if (k3_fallbacks) *k3_fallbacks = s->k3_fallbacks;
    if (k3_cascaded) *k3_cascaded = s->k3_cascaded;
    return SEMA_OK;
}

}  // extern "C"

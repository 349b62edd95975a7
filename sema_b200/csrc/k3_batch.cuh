// k3_batch.cuh — kernel K3: batched multi-query scan on the tcgen05 tensor cores.
//
// The reference issues table.query().nearest_to(q)?.limit(k) once per query
// (src/storage/lance_indexer.rs:121-126); Q queries against the same table are a dense
// contraction S = Q . X^T (Q x N x d), so this path runs it on the 5th-gen tensor
// cores with error-compensated split precision and a fused per-query selection:
//
//   x = x_hi + x_lo, q = q_hi + q_lo (16-bit halves, round-to-nearest residuals: fp16 when every stored element is
//   <= 1024 in magnitude — unit roundoff u = 2^-11 — else bf16, u = 2^-8; Params::fmt)
//   S ~= q_hi.x_hi + q_lo.x_hi + q_hi.x_lo     (3 x kind::f16 UMMA, fp32 accumulate in TMEM)
//   |S - q.x| <= 3 u^2 |q||x| + fp32 accumulation error;  one pass (q_hi.x_hi): <= (2u + u^2) |q||x| + accumulation
//
// The tensor-core pass only *selects* candidates (a short list per query and row partition);
// rescore_kernel then recomputes the surviving candidates' scores in plain fp32 with K2's
// arithmetic, ranks them exactly and proves, per query, that no row outside the candidate
// lists can reach the k-th exact score (else the query is flagged and the host sends it to the
// next stage of the cascade / through K2).  Results are therefore exactly K2's.
//
// Layout.  A CTA owns one tile of 128 queries — the UMMA A operand, kept resident in
// TMEM for the CTA's lifetime (row m <-> TMEM lane m, 2 halves per 32-bit column: q_hi in
// columns [0,192), q_lo in [192,384) for d = 384) — and streams a contiguous range of
// 64-row corpus tiles (the B operand) through a shared-memory ring with 1-D bulk async
// copies (TMA, cp.async.bulk + mbarrier complete_tx).  The corpus planes are stored in
// HBM already in the UMMA canonical K-major no-swizzle core-matrix order, so a stage is
// one contiguous 16 KB copy and needs no tensor map: per 64-row tile, per 64-wide k
// block: [hi | lo][k-chunk of 8 elements (8)][row group (8)][row in group (8)][8 halves].
// Two 128x64 fp32 accumulators (TMEM columns [384,448) and [448,512)) double-buffer the
// MMA against the epilogue.  Each epilogue thread owns one query (= one TMEM lane): it
// reads its 64 scores with tcgen05.ld and keeps a private candidate list in shared memory.
//
// Warp roles (352 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer 0,
// warps 2..5 = epilogue set 0 (TMEM lane quarter = warp % 4), warp 6 = MMA issuer 1, warps 7..10 = epilogue set 1
// (two query tiles per CTA only: each set then drains one tile's accumulator, so a drain — tcgen05.ld of 32 KB,
// mask pass, insertions: ~0.75 us — has two accumulator periods of MMA time to hide in instead of one).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "k2_scan.cuh"
#include "ptx.cuh"

namespace sema {
namespace k3 {

using namespace ::sema::ptx;

constexpr int TILE_Q = 128;            // queries per CTA (UMMA M)
constexpr int TILE_N = 64;             // corpus rows per accumulator tile (UMMA N)
constexpr int BLOCK_K = 64;            // k elements per pipeline stage
constexpr int UMMA_K = 16;             // k per tcgen05.mma (bf16)
constexpr int STAGE_PLANE_BYTES = TILE_N * BLOCK_K * 2;   // 8 KB: one plane of one stage
constexpr int STAGE_BYTES = 2 * STAGE_PLANE_BYTES;        // hi + lo = 16 KB
constexpr int THREADS = 352;            // 11 warps: see the roles at batch_scan_kernel
constexpr int EPI_THREADS = 128;
constexpr uint32_t TMEM_COLS = 512;
constexpr int MAX_DIM = 384;           // bf16x3: q_hi + q_lo need dim TMEM columns; 128 are the accumulators
constexpr int MAX_DIM_1PASS = 768;     // single pass: q_hi alone needs dim/2 columns

// 16-bit element format of the planes and of the TMEM-resident queries (Params::fmt, idesc a/b format)
constexpr int FMT_BF16 = 0;            // 8 significant bits, fp32's exponent range: any finite corpus
constexpr int FMT_FP16 = 1;            // 11 significant bits: 8x tighter error bounds; needs |x_i| <= 1024 (queries are
                                       // scaled by a power of two in the kernel, rows are checked by the host)

// How the kernels wait on their mbarriers (A/B switch; see ptx.cuh for the three forms)
#if defined(SEMA_K3_WAIT_SPIN)
#define K3_WAIT(bar, par) mbar_wait_spin(bar, par)
#define K3_WAIT_EPI(bar, par) mbar_wait_spin(bar, par, 32)
#elif defined(SEMA_K3_WAIT_NOHINT)
#define K3_WAIT(bar, par) mbar_wait_cluster(bar, par)
#define K3_WAIT_EPI(bar, par) mbar_wait_cluster(bar, par)
#else
#define K3_WAIT(bar, par) mbar_wait(bar, par)
#define K3_WAIT_EPI(bar, par) mbar_wait(bar, par)
#endif

// Timing probes (Params::debug values that skip work and therefore give WRONG results) exist only in
// builds made with -DSEMA_K3_PROBES (scripts/build_probe.sh); the shipped library ignores them.
#ifdef SEMA_K3_PROBES
constexpr bool PROBES = true;
#else
constexpr bool PROBES = false;
#endif

// bytes of the pre-tiled planes per 64-row tile
__host__ __device__ constexpr size_t tile_bytes(int dim) { return (size_t)TILE_N * dim * 4; }

// (generic mbarrier / bulk-copy wrappers live in ptx.cuh)
// slice of a stage delivered to the same shared-memory offset of every CTA in cta_mask; each
// destination CTA's mbarrier (same offset) receives the complete_tx for the bytes it got
__device__ __forceinline__ void bulk_g2s_multicast(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                                   uint16_t cta_mask)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;\n\t}"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// arrive on the mbarrier at this offset in every CTA of cta_mask once the prior MMAs retire
__device__ __forceinline__ void umma_commit_multicast(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// pull `bytes` (a multiple of 16) at `src` into L2 without a destination in shared memory
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, uint32_t bytes)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e cp.async.bulk.prefetch.L2.global [%0], %1;\n\t}" ::"l"(src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]   (M=128, N=64, K=16, bf16 x bf16 -> f32).  Called by a
// converged warp; one elected lane issues.
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// K-major, no swizzle: core matrix = 8 rows x 16 B contiguous (128 B).
//   LBO = byte distance between the two 16-byte k-chunks of one K=16 step
//   SBO = byte distance between consecutive 8-row groups
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

// instruction descriptor: c = f32, a = b = bf16 (format 1) or fp16 (format 0), both K-major, dense, M x N
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int m, int n)
{
    const uint32_t ab = fmt == FMT_FP16 ? 0u : 1u;
    return (1u << 4) | (ab << 7) | (ab << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// two consecutive elements rounded to the 16-bit format (round to nearest even), packed low | high
__device__ __forceinline__ uint32_t pack16(int fmt, float lo_elem, float hi_elem)
{
    uint32_t a, b;
    if (fmt == FMT_FP16) {
        a = (uint32_t)__half_as_ushort(__float2half_rn(lo_elem));
        b = (uint32_t)__half_as_ushort(__float2half_rn(hi_elem));
    } else {
        a = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo_elem));
        b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi_elem));
    }
    return a | (b << 16);
}
__device__ __forceinline__ float round16(int fmt, float x)
{
    return fmt == FMT_FP16 ? __half2float(__float2half_rn(x)) : __bfloat162float(__float2bfloat16_rn(x));
}
// power of two s with max|q_i| * s in [0.5, 1): keeps a query of any magnitude inside fp16's range (exact: only
// the exponent changes); 1 for a zero or non-finite query
__device__ __forceinline__ float query_scale(float amax)
{
    if (!(amax > 0.0f) || !(amax < INFINITY)) return 1.0f;
    int e;
    frexpf(amax, &e);                       // amax = f * 2^e, f in [0.5, 1)
    e = e > 120 ? 120 : (e < -120 ? -120 : e);
    return ldexpf(1.0f, -e);
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// three-input maximum (one instruction on sm_100: FMNMX3); a NaN operand is ignored
__device__ __forceinline__ float fmax3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// ---------------------------------------------------------------- candidate lists
// A thread's candidate list lives in shared memory as column m of [KC][128] arrays.  Lists of 16 entries are kept
// unordered and re-scanned for their minimum after every insertion (16 loads); longer lists are binary MIN-HEAPS once
// they are full, so replacing the minimum costs 2 log2(KC) loads instead of KC.
template <int KC>
__device__ __forceinline__ void heap_sift_down(float *lsc, uint32_t *lrow, int m, int i, float v, uint32_t r)
{
    while (true) {
        int c = 2 * i + 1;
        if (c >= KC) break;
        float a = lsc[c * TILE_Q + m];
        if (c + 1 < KC) {
            const float b = lsc[(c + 1) * TILE_Q + m];
            if (b < a) { a = b; ++c; }
        }
        if (!(a < v)) break;
        lsc[i * TILE_Q + m] = a;
        lrow[i * TILE_Q + m] = lrow[c * TILE_Q + m];
        i = c;
    }
    lsc[i * TILE_Q + m] = v;
    lrow[i * TILE_Q + m] = r;
}
// One survivor (score v above the admission threshold, corpus row r) enters the list; returns the new threshold
// (-inf until the list is full).  cnt / min_pos are the thread's list state (min_pos only used by the KC = 16 form).
template <int KC>
__device__ __forceinline__ float list_insert(float *lsc, uint32_t *lrow, int m, float v, uint32_t r, int &cnt, int &min_pos, float thr)
{
#ifdef SEMA_K3_HEAP16
    constexpr int LINEAR_MAX = 8;
#else
    constexpr int LINEAR_MAX = 16;
#endif
    if constexpr (KC <= LINEAR_MAX) {
        const int slot = cnt < KC ? cnt : min_pos;
        lsc[slot * TILE_Q + m] = v;
        lrow[slot * TILE_Q + m] = r;
        if (cnt < KC) ++cnt;
        if (cnt == KC) {  // (re)locate the minimum: it is the admission threshold
            float mn = lsc[m];
            int mp = 0;
#pragma unroll 8
            for (int i = 1; i < KC; ++i) {
                const float sv = lsc[i * TILE_Q + m];
                if (sv < mn) { mn = sv; mp = i; }
            }
            min_pos = mp;
            return mn;
        }
        return thr;
    } else {
        if (cnt < KC) {
            lsc[cnt * TILE_Q + m] = v;
            lrow[cnt * TILE_Q + m] = r;
            if (++cnt < KC) return thr;
            for (int i = KC / 2 - 1; i >= 0; --i)          // the list has just filled up: heapify once
                heap_sift_down<KC>(lsc, lrow, m, i, lsc[i * TILE_Q + m], lrow[i * TILE_Q + m]);
        } else {
            heap_sift_down<KC>(lsc, lrow, m, 0, v, r);      // v replaces the minimum at the root
        }
        return lsc[m];
    }
}

// ---------------------------------------------------------------- plane builder
// X (fp32, row stride ld) rows [row_begin, row_end) -> pre-tiled bf16 hi/lo planes.
// One thread per (row, 8-element k-chunk); consecutive threads write consecutive 16 B.
__global__ void __launch_bounds__(256)
split_planes_kernel(const float *X, uint32_t ld, uint32_t dim, uint64_t row_begin, uint64_t row_end,
                    uint64_t n_valid_rows, unsigned char *planes, int fmt)
{
    const uint32_t chunks = dim / 8;
    const uint64_t total = (row_end - row_begin) * chunks;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        // i -> (tile-relative) ordering: r%8 fastest, then r/8 (within tile), then chunk, then tile
        const uint64_t rows_span = row_end - row_begin;  // multiple of TILE_N by construction
        (void)rows_span;
        const uint64_t t = i / ((uint64_t)TILE_N * chunks);
        const uint32_t w = (uint32_t)(i - t * (uint64_t)TILE_N * chunks);
        const uint32_t c = w / TILE_N;
        const uint32_t r = w - c * TILE_N;
        const uint64_t row = row_begin + t * TILE_N + r;
        float v[8];
        if (row < n_valid_rows) {
            const float4 a = *reinterpret_cast<const float4 *>(X + row * ld + c * 8);
            const float4 b = *reinterpret_cast<const float4 *>(X + row * ld + c * 8 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = 0.0f;  // padding rows of the last tile
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float h0 = round16(fmt, v[2 * e]), h1 = round16(fmt, v[2 * e + 1]);
            hi[e] = pack16(fmt, v[2 * e], v[2 * e + 1]);
            lo[e] = pack16(fmt, v[2 * e] - h0, v[2 * e + 1] - h1);
        }
        const uint64_t tile = (row_begin / TILE_N) + t;
        unsigned char *base = planes + tile * tile_bytes((int)dim) + (size_t)(c / 8) * STAGE_BYTES +
                              (size_t)(c % 8) * (TILE_N * 16) + (size_t)(r / 8) * 128 + (size_t)(r % 8) * 16;
        *reinterpret_cast<uint4 *>(base) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(base + STAGE_PLANE_BYTES) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// Tombstones (sema_index_tombstone): the listed local rows become NaN in both planes, in place — the
// same addresses split_planes_kernel writes for them — so deleting a file's chunks costs 2 x 16 bytes
// per row and 8 columns instead of a re-tiling of everything behind the first dead row.  A NaN score
// never passes the epilogue's `score > threshold`, exactly like a NaN row of X in K2.
__global__ void __launch_bounds__(256)
poison_planes_kernel(unsigned char *planes, uint32_t dim, const uint64_t *rows, uint64_t n, uint64_t planes_rows)
{
    const uint32_t chunks = dim / 8;
    const uint64_t total = n * chunks;
    const uint4 nan8 = make_uint4(0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u, 0x7fc07fc0u);   // 0x7fc0 is a quiet NaN as bf16 AND as fp16
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = rows[i / chunks];
        if (row >= planes_rows) continue;              // not tiled yet: split_planes_kernel will read the NaN row of X
        const uint32_t c = (uint32_t)(i % chunks), r = (uint32_t)(row % TILE_N);
        unsigned char *base = planes + (row / TILE_N) * tile_bytes((int)dim) + (size_t)(c / 8) * STAGE_BYTES +
                              (size_t)(c % 8) * (TILE_N * 16) + (size_t)(r / 8) * 128 + (size_t)(r % 8) * 16;
        *reinterpret_cast<uint4 *>(base) = nan8;
        *reinterpret_cast<uint4 *>(base + STAGE_PLANE_BYTES) = nan8;
    }
}

// ---------------------------------------------------------------- the batched scan
struct Params {
    const unsigned char *planes;  // pre-tiled hi/lo planes
    const float *Q;               // q_tiles*128 x dim fp32 (zero padded rows)
    uint32_t *cand_rows;          // [q_tiles*128][parts][KC]  local row ids (0xFFFFFFFF = empty)
    float *cand_sc;               // [q_tiles*128][parts][KC]  their tensor-core scores, in the query's own scale (-inf = empty)
    float *cand_thr;              // [q_tiles*128][parts]      lowest approx score kept (-inf if list not full)
    uint32_t n_rows;              // visible rows
    uint32_t n_tiles;             // ceil(n_rows / 64)
    uint32_t parts;               // row partitions of the whole stage = candidate lists per query (all launches of a stage)
    uint32_t part_base;           // this launch's first partition (a stage may be two concurrent launches with different
                                  // cluster sizes that together fill every SM: see k3_stage)
    uint32_t tile_first, tile_per, tile_end;   // this launch's partitions: tiles [tile_first + j*tile_per, ... + tile_per) clipped to tile_end
    uint32_t dim;
    uint32_t fmt;                 // FMT_BF16 / FMT_FP16: element format of the planes (and of the queries staged in TMEM)
    uint32_t prefetch;            // L2 prefetch distance of the producer, in pipeline stages (0 = none)
    uint32_t debug;               // 8 = no group early-out (correct results, A/B baseline); other values are timing probes
                                  // that only exist in -DSEMA_K3_PROBES builds
};

// PASSES = 3: bf16x3 split (hi.hi + lo.hi + hi.lo), a stage holds the hi and lo plane blocks.
// PASSES = 1: single bf16 pass as a coarser candidate filter (only the hi plane block is fetched;
//             the exactness proof then uses the 1-pass error bound and falls back more readily).
// QT = query tiles (of 128) per CTA.  The single-pass mode leaves room in TMEM for a second query
// tile (2 x dim/2 columns): both tiles consume the same shared-memory stages, which halves the
// operand bytes delivered per MMA (the L2->SM delivery rate is what bounds that mode), and their
// two accumulators ping-pong between the MMA and the epilogue.
template <int KC, int PASSES, int QT = 1>
struct Smem {
    static_assert(QT == 1 || PASSES == 1, "two query tiles per CTA only fit TMEM in the single-pass mode");
    static constexpr int STAGE = PASSES == 3 ? STAGE_BYTES : STAGE_PLANE_BYTES;
    static constexpr int LIST_BYTES = QT * KC * TILE_Q * 8;     // scores f32 + rows u32, per query tile
    static constexpr int BAR_BYTES = 1024;
    static constexpr int STAGES_RAW = (227 * 1024 - LIST_BYTES - BAR_BYTES) / STAGE;
    static constexpr int STAGES = STAGES_RAW > 24 ? 24 : STAGES_RAW;
    static constexpr int TOTAL = STAGES * STAGE + LIST_BYTES + BAR_BYTES;
};

// C = CTAs per cluster (QT and PASSES: see Smem).  The C CTAs of a cluster hold C different query tiles and stream the
// same corpus tiles: every stage is fetched once per cluster (each CTA issues 1/C of it) and
// multicast into all C shared memories, cutting the L2->SM operand traffic C-fold.
template <int KC, int C, int PASSES, int QT>
__global__ void __launch_bounds__(THREADS, 1)
batch_scan_kernel(const Params p)
{
    using S = Smem<KC, PASSES, QT>;
    constexpr int STAGES = S::STAGES;
    constexpr int STAGE = S::STAGE;          // bytes fetched per k-block
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ring = smem;
    float *list_sc = reinterpret_cast<float *>(smem + STAGES * STAGE);              // [QT][KC][128]
    uint32_t *list_row = reinterpret_cast<uint32_t *>(list_sc + QT * KC * TILE_Q);  // [QT][KC][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE + S::LIST_BYTES);
    uint64_t *full = bars;                    // [STAGES]  TMA -> MMA
    uint64_t *empty = bars + STAGES;          // [STAGES]  MMA -> TMA
    uint64_t *acc_full = bars + 2 * STAGES;   // [2]       MMA -> epilogue
    uint64_t *acc_empty = acc_full + 2;       // [2]       epilogue -> MMA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    // warp index through a shuffle: the compiler then KNOWS it is warp-uniform, so the role branches below are uniform
    // control flow and the MMA issue loop keeps descriptors / TMEM addresses in uniform registers (UIADD3 + UTCHMMA
    // instead of ~15 instructions with four R2UR per MMA, which made one issuing warp slower than the tensor pipe)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t qt = blockIdx.x, part = p.part_base + blockIdx.y;   // this CTA's query tiles: qt*QT .. qt*QT + QT-1
    const uint32_t kblocks = p.dim / BLOCK_K;            // stages per tile
    const uint32_t acols = p.dim / 2;                    // TMEM columns per query plane
    // contiguous tile range of this row partition
    const uint32_t t0 = min(p.tile_first + blockIdx.y * p.tile_per, p.tile_end), t1 = min(t0 + p.tile_per, p.tile_end);

    if (threadIdx.x == 0) {
        // a stage is released by one issuer per CTA of the cluster (QT = 1) or by both (QT = 2)
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], C * QT); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], EPI_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (C > 1) cluster_sync_all();           // every CTA's barriers exist before any remote arrive
    // The CTA owns all 512 TMEM columns (1 CTA/SM), so the allocation starts at lane 0, column 0;
    // using the constant keeps every UMMA operand address in uniform registers.
    if (*tmem_slot != 0) __trap();
    constexpr uint32_t tmem = 0;
    // TMEM columns: [query planes: hi (+ lo) for bf16x3, or QT hi planes][2 accumulators of 64]
    const uint32_t acc_col = PASSES == 3 ? 2 * acols : QT * acols;
    const uint32_t crank = C > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t cmask = (uint16_t)((1u << C) - 1u);
    constexpr uint32_t SLICE = STAGE / C;

    // ---- epilogue warps stage the query tile(s) into TMEM (A operand): row m <-> lane m.  Each query is scaled by a
    // power of two so that its largest element lies in [0.5, 1) (exact; keeps any query inside fp16's range);
    // the scores of its TMEM lane carry the same factor, and the published threshold is divided by it again.
    float qscale0 = 1.0f, qscale1 = 1.0f;
    // epilogue set of this warp: 0 = warps 2..5, 1 = warps 7..10 (active with two query tiles per CTA), -1 = none
    const int eset = (warp >= 2 && warp <= 5) ? 0 : ((QT == 2 && warp >= 7) ? 1 : -1);
    if (eset >= 0) {
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
        const int fmt = (int)p.fmt;
#pragma unroll 1
        for (int qi = (QT == 2 ? eset : 0); qi < (QT == 2 ? eset + 1 : 1); ++qi) {   // each set stages (and later drains) its own tile
            const float *q = p.Q + (((size_t)qt * QT + qi) * TILE_Q + m) * p.dim;
            float amax = 0.0f;
            for (uint32_t c = 0; c < p.dim; c += 4) {
                const float4 f = *reinterpret_cast<const float4 *>(q + c);
                amax = fmaxf(fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y))), fmaxf(fabsf(f.z), fabsf(f.w)));
            }
            const float sc = query_scale(amax);
            if (qi) qscale1 = sc; else qscale0 = sc;
            for (uint32_t c = 0; c < p.dim; c += 16) {
                float v[16];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 f = *reinterpret_cast<const float4 *>(q + c + 4 * e);
                    v[4 * e] = f.x * sc; v[4 * e + 1] = f.y * sc; v[4 * e + 2] = f.z * sc; v[4 * e + 3] = f.w * sc;
                }
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    hi[e] = pack16(fmt, v[2 * e], v[2 * e + 1]);
                    lo[e] = pack16(fmt, v[2 * e] - round16(fmt, v[2 * e]), v[2 * e + 1] - round16(fmt, v[2 * e + 1]));
                }
                tmem_st8(lane_addr + qi * acols + c / 2, hi);
                if (PASSES == 3) tmem_st8(lane_addr + acols + c / 2, lo);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp == 0) {
        // ===== TMA producer: one contiguous bulk copy per stage (whole warp converged, one elected
        // lane issues) =====
        {
            uint32_t stage = 0, phase = 0;
            // L2 prefetch `p.prefetch` stages ahead of the copies: the shared-memory ring bounds the bytes in flight
            // to the SM (192 KB), which at DRAM latency caps the operand stream below what the tensor pipe eats; a
            // prefetch pulls the lines into L2 early without occupying a stage, so the real copy is an L2 hit
            const uint32_t pf = p.prefetch;
            const uint32_t pf_tiles = pf / kblocks, pf_kb = pf % kblocks;
            for (uint32_t t = t0; t < t1; ++t) {
                const unsigned char *src = p.planes + (size_t)t * tile_bytes((int)p.dim);
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    if (pf) {
                        uint32_t tp = t + pf_tiles, kp = kb + pf_kb;
                        if (kp >= kblocks) { kp -= kblocks; ++tp; }
                        if (tp < t1)
                            bulk_prefetch_l2(p.planes + (size_t)tp * tile_bytes((int)p.dim) + (size_t)kp * STAGE_BYTES + crank * SLICE, SLICE);
                    }
                    K3_WAIT(&empty[stage], phase ^ 1);       // all C consumers released this stage
                    mbar_expect_tx(&full[stage], STAGE);
                    // the planes hold [hi | lo] per k-block; a 1-pass stage fetches the hi half only
                    if (C == 1)
                        bulk_g2s(ring + stage * STAGE, src + (size_t)kb * STAGE_BYTES, STAGE, &full[stage]);
                    else
                        bulk_g2s_multicast(ring + stage * STAGE + crank * SLICE,
                                           src + (size_t)kb * STAGE_BYTES + crank * SLICE, SLICE, &full[stage], cmask);
                    if (++stage == (uint32_t)STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 6) {
        // ===== MMA issuers (two warps: one elected lane of one warp cannot issue a 128x64x16 UMMA
        // every 32 cycles, two can).  Issuer w owns accumulator buffer w: with QT = 1 it takes the
        // corpus tiles of parity w, with QT = 2 it takes query tile w of every corpus tile (both
        // issuers then consume every stage, which is released after both commits).  Per k-step 3
        // UMMAs (hi.hi, lo.hi, hi.lo) or 1 (single pass).  The warp stays converged so that
        // descriptors and TMEM addresses live in uniform registers. =====
        const uint32_t w = warp == 1 ? 0u : 1u;
        const uint32_t idesc = make_idesc((int)p.fmt, TILE_Q, TILE_N);
        // Two issuers need the ring to hold two whole tiles when they work on different tiles
        // (an mbarrier waiter may be at most one phase ahead); otherwise issuer 0 works alone.
        const bool dual = QT == 2 || 2 * kblocks <= (uint32_t)STAGES;
        const uint32_t a_base = tmem + (QT == 2 ? w * acols : 0u);
        const uint32_t tstep = (QT == 1 && dual) ? 2u : 1u;
        uint32_t st = 0, ph = 0, it = 0;
        if (QT == 1 && dual && w == 1) st = kblocks;   // the first tile belongs to issuer 0
        if (dual || w == 0) {
            for (uint32_t t = t0 + ((QT == 1 && dual) ? w : 0u); t < t1; t += tstep, ++it) {
                const uint32_t buf = dual ? w : (it & 1);
                const uint32_t use = dual ? it : (it >> 1);
                const uint32_t d_tmem = tmem + acc_col + buf * TILE_N;
                K3_WAIT(&acc_empty[buf], (use & 1) ^ 1);
                tc_fence_after();
                for (uint32_t kb = 0; kb < kblocks; ++kb) {
                    K3_WAIT(&full[st], ph);
                    tc_fence_after();
                    const uint32_t sb = smem_u32(ring + st * STAGE);
#pragma unroll
                    for (int j = 0; j < BLOCK_K / UMMA_K; ++j) {
                        const uint32_t a_hi = a_base + kb * (BLOCK_K / 2) + j * (UMMA_K / 2);
                        const uint32_t a_lo = a_hi + acols;
                        const uint64_t b_hi = make_b_desc(sb + j * 2 * (TILE_N * 16), TILE_N * 16, 128);
                        const uint64_t b_lo = make_b_desc(sb + STAGE_PLANE_BYTES + j * 2 * (TILE_N * 16), TILE_N * 16, 128);
                        umma_ts(d_tmem, a_hi, b_hi, idesc, (kb | j) != 0);
                        if (PASSES == 3) {
                            umma_ts(d_tmem, a_lo, b_hi, idesc, 1);
                            umma_ts(d_tmem, a_hi, b_lo, idesc, 1);
                        }
                    }
                    if (C == 1) umma_commit(&empty[st]);   // frees the smem stage when the MMAs retire
                    else umma_commit_multicast(&empty[st], cmask);   // ... in every CTA of the cluster
                    if (++st == (uint32_t)STAGES) { st = 0; ph ^= 1; }
                }
                umma_commit(&acc_full[buf]);           // accumulator ready for the epilogue
                if (QT == 1 && dual) {                 // skip the other issuer's tile
                    st += kblocks;
                    if (st >= (uint32_t)STAGES) { st -= STAGES; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (eset >= 0) {
        // ===== epilogue: thread m owns query m of its set's query tile (QT = 2: set s <-> tile s <-> accumulator s;
        // QT = 1: set 0 alone drains both accumulators alternately); private candidate lists in shared memory =====
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + acc_col;
        float thr0 = -INFINITY, thr1 = -INFINITY;   // lowest score kept once the list is full
        int cnt0 = 0, cnt1 = 0, mp0 = 0, mp1 = 0;
        uint32_t it = 0;
        for (uint32_t t = t0; t < t1; ++t, ++it) {
#pragma unroll 1
            for (int qi = (QT == 2 ? eset : 0); qi < (QT == 2 ? eset + 1 : 1); ++qi) {
                const uint32_t buf = QT == 1 ? (it & 1) : (uint32_t)qi;
                const uint32_t use = QT == 1 ? (it >> 1) : it;
                float thr = qi ? thr1 : thr0;
                int cnt = qi ? cnt1 : cnt0, min_pos = qi ? mp1 : mp0;
                float *lsc = list_sc + qi * KC * TILE_Q;
                uint32_t *lrow = list_row + qi * KC * TILE_Q;
                K3_WAIT_EPI(&acc_full[buf], use & 1);
                tc_fence_after();
                uint32_t r[2][32];
                tmem_ld32(lane_addr + buf * TILE_N, r[0]);
                tmem_ld32(lane_addr + buf * TILE_N + 32, r[1]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);   // MMA may overwrite this accumulator
                const uint32_t row0 = t * TILE_N;
                if (PROBES && p.debug == 1) continue;             // probe: no scan at all
                // Pass 0: the maximum of each group of 16 scores (8 three-input maxima per group).  Once the
                // lists have warmed up almost no group holds a score above the admission threshold, and a
                // group nobody in the warp needs is skipped: the epilogue's instruction issue is energy the
                // power-capped tensor pipe does not get.  (Padding rows of the last tile can only make a
                // group look interesting, never hide a survivor; NaN scores are ignored by max and by >.)
                // Pass 1 (per group, warp-uniform): which of my scores beat the admission threshold?
                uint32_t mask[2] = {0u, 0u};
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int h = g >> 1, c0 = (g & 1) * 16;
                    bool scan = true;
                    if (!(p.debug & 8)) {                            // debug 8: every group is scanned (A/B baseline)
                        float mx = fmaxf(__uint_as_float(r[h][c0]), __uint_as_float(r[h][c0 + 1]));
#pragma unroll
                        for (int c = 2; c < 16; c += 2)
                            mx = fmax3(mx, __uint_as_float(r[h][c0 + c]), __uint_as_float(r[h][c0 + c + 1]));
                        scan = __any_sync(FULL, mx > thr);
                    }
                    if (scan) {
#pragma unroll
                        for (int c = 0; c < 16; ++c)
                            mask[h] |= (__uint_as_float(r[h][c0 + c]) > thr) ? (1u << (c0 + c)) : 0u;
                    }
                }
                const uint32_t live = p.n_rows - row0;           // rows of this tile that exist (>= 1)
                if (live < 32) { mask[0] &= (1u << live) - 1u; mask[1] = 0u; }
                else if (live < 64) mask[1] &= (1u << (live - 32)) - 1u;
                if (PROBES && p.debug == 2) { if (mask[0] | mask[1]) thr = fmaxf(thr, -1e30f); continue; }   // probe: mask pass only
                // Pass 2 (rare, not unrolled: keeps the instruction footprint small): insert them.
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t mk = mask[h];
#pragma unroll 1
                    while (mk) {
                        const int c = __ffs(mk) - 1;
                        mk &= mk - 1;
                        uint32_t bits = 0;
#pragma unroll
                        for (int e = 0; e < 32; ++e) bits = (e == c) ? r[h][e] : bits;   // register select
                        const float v = __uint_as_float(bits);
                        if (!(v > thr)) continue;                 // the threshold may have risen meanwhile
                        thr = list_insert<KC>(lsc, lrow, m, v, row0 + h * 32 + c, cnt, min_pos, thr);
                    }
                }
                if (qi) { thr1 = thr; cnt1 = cnt; mp1 = min_pos; }
                else { thr0 = thr; cnt0 = cnt; mp0 = min_pos; }
            }
        }
        // publish this partition's candidates for the CTA's queries
#pragma unroll 1
        for (int qi = (QT == 2 ? eset : 0); qi < (QT == 2 ? eset + 1 : 1); ++qi) {
            const float thr = qi ? thr1 : thr0;
            const int cnt = qi ? cnt1 : cnt0;
            const uint32_t *lrow = list_row + qi * KC * TILE_Q;
            const size_t q = ((size_t)qt * QT + qi) * TILE_Q + m;
            uint32_t *out = p.cand_rows + (q * p.parts + part) * KC;
            float *out_sc = p.cand_sc + (q * p.parts + part) * KC;
            const float *lsc = list_sc + qi * KC * TILE_Q;
            const float unscale = 1.0f / (qi ? qscale1 : qscale0);          // a power of two: exact
            for (int i = 0; i < KC; ++i) {
                out[i] = i < cnt ? lrow[i * TILE_Q + m] : 0xffffffffu;
                out_sc[i] = i < cnt ? lsc[i * TILE_Q + m] * unscale : -INFINITY;
            }
            p.cand_thr[q * p.parts + part] = (cnt == KC) ? thr * unscale : -INFINITY;   // back to q's own scale (exact)
        }
    }

    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();           // no CTA leaves while peers may still multicast into it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

// ---------------------------------------------------------------- exact fp32 rescoring
struct RescoreParams {
    const float4 *X;           // fp32 matrix, row stride ld4
    const float *Q;            // padded queries, row stride dim
    const uint32_t *cand_rows; // [q][parts*KC]
    const float *cand_sc;      // [q][parts*KC] tensor-core scores of the candidates
    const float *cand_thr;     // [q][parts]
    uint64_t *res_ids;         // [nq][k]
    float *res_scores;         // [nq][k]
    uint32_t *res_nfound;      // [nq]
    uint32_t *flags;           // [nq] 1 = exactness not proven, re-run through K2
    uint32_t ld4, dim, k, parts, kc, row_base;
    const float *max_norm2;    // device: [max, min] squared row norm seen by K1 (bounds |x|)
    float err_rel;             // |tensor-core score - exact score| <= err_rel * |q| * max|x| + err_abs * |q|
    float err_abs;
};
// METRIC_L2: the tensor-core pass still selects by dot product (the planes hold x, the queries q);
// the candidates are re-scored with K2's squared-L2 arithmetic and ranked by ascending distance.
// |q - x|^2 = |q|^2 + |x|^2 - 2 q.x, so a row outside partition P's list (dot <= thr_P + err) has
// distance >= |q|^2 + min|x|^2 - 2 (thr_P + err): the proof needs the k-th exact distance below that.
// With unit-norm rows (the reference's case) the two rankings coincide and the proof is as tight as
// for the cosine metric; the host only takes this path when the row norms are (nearly) constant.

// One block per query: every candidate is re-scored with K2's fp32 arithmetic (lane-strided
// float4 FMAs + butterfly), ranked by (score, lower row id) and emitted.
template <int M, int METRIC>
__global__ void __launch_bounds__(SCAN_THREADS)
rescore_kernel(const RescoreParams p)
{
    __shared__ uint64_t sm_keys[SCAN_WARPS * 32 * M];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x;
    const int k = (int)p.k;
    const uint32_t nv = p.dim / 4;  // float4 per row (dim % 64 == 0)
    const float4 *qp = reinterpret_cast<const float4 *>(p.Q + (size_t)q * p.dim);
    const uint32_t total = p.parts * p.kc;
    __shared__ float sm_cut;
    // |q|, computed on the query scaled into [0.5, 1) by a power of two so that |q|^2 can neither overflow nor
    // vanish (the L2 proof below needs the plain |q|^2, which is as representable as the distances themselves)
    float amax = 0.0f;
    for (uint32_t v = lane; v < nv; v += 32) {
        const float4 f = qp[v];
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y))), fmaxf(fabsf(f.z), fabsf(f.w)));
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(FULL, amax, d));
    const float qs = query_scale(amax);
    float qq = 0.0f, qqs = 0.0f;
    for (uint32_t v = lane; v < nv; v += 32) {
        const float4 f = qp[v];
        qq = accum4<METRIC_COSINE>(qq, f, f);
        const float4 g = make_float4(f.x * qs, f.y * qs, f.z * qs, f.w * qs);
        qqs = accum4<METRIC_COSINE>(qqs, g, g);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) { qq += __shfl_xor_sync(FULL, qq, d); qqs += __shfl_xor_sync(FULL, qqs, d); }
    const float qnorm = sqrtf(qqs) / qs;
    const float err_bound = qnorm * (p.err_rel * sqrtf(p.max_norm2[0]) + p.err_abs);

    // Which candidates are worth an exact re-scoring?  Let T be the k-th best TENSOR-CORE score among all candidates:
    // k candidates have exact score >= T - err, so a candidate whose tensor-core score is below T - 2 err has an exact
    // score below the exact k-th best and cannot be in the result (L2 metric over rows whose |x|^2 differ by at most
    // mx - mn: below T - 2 err - (mx - mn) / 2).  That leaves ~k of the parts * KC candidates (592 for config 3).
    WarpTopK<M> top;
    top.init();
    for (uint32_t i0 = warp * 32; i0 < total; i0 += SCAN_WARPS * 32) {
        const uint32_t i = i0 + lane;
        const float a = i < total ? p.cand_sc[(size_t)q * total + i] : -INFINITY;
        top.offer(make_key(a, i), a > -INFINITY, lane, k);
    }
    block_merge<M, SCAN_WARPS>(top, sm_keys, warp, lane, k);
    if (warp == 0) {
        const int kj0 = (k - 1) >> 5, kl0 = (k - 1) & 31;
        uint64_t kt = 0;
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j == kj0) kt = top.v[j];
        kt = __shfl_sync(FULL, kt, kl0);
        if (lane == 0) {
            float cut = -INFINITY;
            if (kt != 0) {
                const float t = key_rank(kt);
                cut = t - 2.0f * err_bound - 1e-6f * fabsf(t) - 1e-30f;
                if (METRIC == METRIC_L2) cut -= 0.5f * (p.max_norm2[0] - p.max_norm2[1]) * 1.000001f;
            }
            sm_cut = cut;
        }
    }
    __syncthreads();
    const float cut = sm_cut;
    top.init();
    for (uint32_t i = warp; i < total; i += SCAN_WARPS) {
        const uint32_t row = p.cand_rows[(size_t)q * total + i];
        if (row == 0xffffffffu) continue;  // warp-uniform
        if (!(p.cand_sc[(size_t)q * total + i] >= cut)) continue;   // warp-uniform: cannot reach the exact top-k
        const float4 *xp = p.X + (size_t)row * p.ld4;
        float acc = 0.0f;
        for (uint32_t v = lane; v < nv; v += 32) acc = accum4<METRIC>(acc, xp[v], qp[v]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
        const uint64_t key = make_key(METRIC == METRIC_L2 ? -acc : acc, p.row_base + row);
        top.offer(key, lane == 0 && acc == acc, lane, k);
    }
    block_merge<M, SCAN_WARPS>(top, sm_keys, warp, lane, k);
    if (warp == 0) {
        emit_results<M, METRIC>(top, k, nullptr, p.res_ids + (size_t)q * k, p.res_scores + (size_t)q * k,
                                p.res_nfound + q, lane);
        // exactness proof: rows outside partition P's list have approx score <= thr_P, hence
        // exact score <= thr_P + err_bound; the k-th exact score must beat that for every P.
        float worst = -INFINITY;
        for (uint32_t j = lane; j < p.parts; j += 32) worst = fmaxf(worst, p.cand_thr[(size_t)q * p.parts + j]);
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) worst = fmaxf(worst, __shfl_xor_sync(FULL, worst, d));
        const int kj = (k - 1) >> 5, kl = (k - 1) & 31;
        uint64_t kk = 0;
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j == kj) kk = top.v[j];
        kk = __shfl_sync(FULL, kk, kl);
        if (lane == 0) {
            bool proven = true;
            if (worst > -INFINITY) {
                if (METRIC == METRIC_L2)   // key_rank = -distance; slack for the fp32 evaluation of the bound itself
                    proven = (kk != 0) && (-key_rank(kk) < (qq + p.max_norm2[1] - 2.0f * (worst + err_bound)) * (1.0f - 1e-6f) - 1e-6f);
                else
                    proven = (kk != 0) && (key_rank(kk) > worst + err_bound);
            }
            p.flags[q] = proven ? 0u : 1u;
        }
    }
}

// ---------------------------------------------------------------- cascade plumbing
// dst[j] = Q[idx[j]] (rows of dim4 float4): the queries the single-pass stage could not prove, made dense
__global__ void __launch_bounds__(256)
gather_queries_kernel(const float4 *Q, const uint32_t *idx, uint32_t nsub, uint32_t dim4, float4 *dst)
{
    const uint64_t total = (uint64_t)nsub * dim4;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t j = (uint32_t)(i / dim4), c = (uint32_t)(i - (uint64_t)j * dim4);
        dst[i] = Q[(size_t)idx[j] * dim4 + c];
    }
}
// block j: if the sub-batch proved query j (flags[j] == 0), its k results replace those of query idx[j]
__global__ void __launch_bounds__(128)
scatter_results_kernel(const uint64_t *sub_ids, const float *sub_sc, const uint32_t *sub_nf, const uint32_t *idx,
                       const uint32_t *flags, uint32_t k, uint64_t *ids, float *sc, uint32_t *nf)
{
    const uint32_t j = blockIdx.x;
    if (flags[j] != 0) return;
    const uint32_t q = idx[j];
    for (uint32_t e = threadIdx.x; e < k; e += blockDim.x) {
        ids[(size_t)q * k + e] = sub_ids[(size_t)j * k + e];
        sc[(size_t)q * k + e] = sub_sc[(size_t)j * k + e];
    }
    if (threadIdx.x == 0) nf[q] = sub_nf[j];
}

}  // namespace k3
}  // namespace sema

// k3_pair.cuh — kernel K3, CTA-pair form of the single-pass candidate filter (tcgen05 cta_group::2).
//
// Same job as batch_scan_kernel<KC, C, 1, QT> (k3_batch.cuh): one 16-bit pass q_hi . x_hi over the hi plane of
// the corpus, fused per-query candidate selection, exactness decided afterwards by rescore_kernel.  What changes
// is the shape of the tensor-core work.  Two CTAs on the two SMs of a TPC form a pair and issue ONE
// tcgen05.mma.cta_group::2 of M = 256 x N = 128 x K = 16 where the single-CTA kernel issues four 128 x 64 x 16:
//
//   * A (queries): each CTA keeps its own tile of 128 queries resident in its own TMEM (row m <-> lane m,
//     columns [0, dim/2)), exactly as before; the pair covers 256 queries.
//   * B (corpus): each CTA streams the hi plane of ONE 64-row tile per step into its own shared memory — the
//     leader (even cluster rank) tile 2u, its peer tile 2u+1 — and the hardware reads both halves, so per SM the
//     shared-memory operand traffic per flop is halved and the planes keep their 64-row tiling.
//   * D: 128 lanes x 128 fp32 columns per CTA (columns [0,64) = rows of tile 2u, [64,128) = rows of tile 2u+1),
//     double-buffered in TMEM columns [dim/2, dim/2 + 256): the query tile is re-read from TMEM once per 128
//     corpus rows instead of once per 64, one elected thread of the LEADER issues for both SMs (one instruction per
//     64 tensor cycles instead of two issuer warps racing to launch one every 32), and the epilogue takes half as
//     many accumulator hand-offs per score.
//
// Feeding the pair.  Both CTAs' producers fetch with cp.async.bulk.tensor...cta_group::2, whose completion may be
// signalled on the OTHER CTA's mbarrier: every copy of a stage, the leader's and the peer's, completes its bytes on
// the LEADER's full[s] (the leader arms it with expect_tx of both halves), so the MMA warp learns about both halves
// with no software hop.  (The 1-D cp.async.bulk can only signal a barrier in the CTA it writes to; a forwarding warp
// in the peer doing a remote mbarrier.arrive per stage cost ~0.5 us each and made the whole kernel 3x slower.)  The
// planes are described by one 3-D tensor map over 8-byte elements — {256 (a 2 KB row), 4 (rows of one 8 KB hi block),
// 16 KB groups (a k-block's hi|lo pair)} — so a box {256, 4, kps} is the hi blocks of kps consecutive k-blocks of a
// tile, delivered as kps contiguous 8 KB blocks: the canonical no-swizzle K-major operand, as before.
// A pipeline stage is kps = 2 k-blocks (16 KB per CTA; 1 when dim/64 is odd): fine enough that 12-13 stages fit and
// the loop "commit -> producer wakes -> copy -> data lands" (~3 us) is covered, coarse enough that the producer's
// fixed costs (measured with scripts/probe/tma_ingest.cu: ~86 ns per stage hand-off + ~47 ns per copy instruction)
// stay 2x below the tensor pipe's demand of one k-block per 0.13 us.
//   leader -> peer : one multicast tcgen05.commit on done[s] per stage; in each CTA the producer (stage s may be
//                    refilled) and, for a super-tile's last stage, the epilogue (accumulator ready) wait on it.
//   peer -> leader : one remote arrive per super-tile on acc_free_peer[b] by the peer's warp 1 (its epilogue has
//                    drained accumulator b) — off the critical path, the accumulators are double-buffered.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + (leader) MMA issuer / (peer)
// accumulator-free forwarder, warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
#pragma once
#include "k3_batch.cuh"

namespace sema {
namespace k3 {

// how the pair kernel waits on barriers that are completed from the other SM / by asynchronous units
#ifdef SEMA_PAIR_TRYWAIT
#define PAIR_WAIT(bar, par) mbar_wait_cluster(bar, par)
#define PAIR_WAIT_EPI(bar, par) mbar_wait_cluster(bar, par)
#else
#define PAIR_WAIT(bar, par) mbar_wait_spin(bar, par)
#define PAIR_WAIT_EPI(bar, par) mbar_wait_spin(bar, par, 32)
#endif
constexpr int PAIR_THREADS = 192;
constexpr int PAIR_N = 128;            // corpus rows per accumulator (UMMA N): 64 from each CTA of the pair
constexpr int PAIR_MAX_DIM = 512;      // dim/2 query columns + 2 x 128 accumulator columns <= 512
constexpr int PAIR_MAX_STAGES = 24;
constexpr int PAIR_BAR_BYTES = 512;

// k-blocks per pipeline stage, stages per super-tile, ring depth (a multiple of the stages per super-tile)
__host__ __device__ constexpr int pair_kps(int dim) { return (dim / BLOCK_K) % 2 == 0 ? 2 : 1; }
__host__ __device__ constexpr int pair_stage_bytes(int dim) { return pair_kps(dim) * STAGE_PLANE_BYTES; }
__host__ __device__ constexpr int pair_spt(int dim) { return (dim / BLOCK_K) / pair_kps(dim); }
__host__ __device__ constexpr int pair_stages(int kc, int dim)
{
    int n = (227 * 1024 - kc * TILE_Q * 8 - PAIR_BAR_BYTES) / pair_stage_bytes(dim);
    n = n > PAIR_MAX_STAGES ? PAIR_MAX_STAGES : n;
    return n / pair_spt(dim) * pair_spt(dim);
}
__host__ __device__ constexpr int pair_smem_bytes(int kc, int dim)
{
    return pair_stages(kc, dim) * pair_stage_bytes(dim) + kc * TILE_Q * 8 + PAIR_BAR_BYTES;
}

// 3-D tiled bulk tensor copy global -> this CTA's shared memory whose completion bytes are counted on an mbarrier
// given as a shared::cluster address (the leader's, for both CTAs of the pair)
__device__ __forceinline__ void tma_load_3d_pair(void *dst, const CUtensorMap *tmap, uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t mbar_cluster_addr)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n\t}"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(mbar_cluster_addr)
                 : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(const void *smem_ptr, uint32_t cta_rank)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(smem_ptr)), "r"(cta_rank));
    return ra;
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta_rank)
{
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(cta_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem] over the CTA pair: M = 256 (128 query lanes in each CTA's TMEM), N = 128 (64 rows
// from each CTA's shared memory, same descriptor offset in both), K = 16.  Leader only; converged warp, one lane issues.
__device__ __forceinline__ void umma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair_multicast(uint64_t *bar, uint16_t cta_mask)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}

__device__ __forceinline__ void umma_commit_pair_local(uint64_t *bar)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
                 : "memory");
}

// KC = candidates kept per (query, row partition).  grid = (query tiles padded to an even count, row partitions),
// cluster = (2, 1, 1): the pair.  tmap: the planes as {256 x u64, 4, 16 KB groups}, box {256, 4, pair_kps(dim)}.
template <int KC>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
pair_scan_kernel(const Params p, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t kblocks = p.dim / BLOCK_K;
    const uint32_t kps = (uint32_t)pair_kps((int)p.dim);              // k-blocks per stage
    const uint32_t spt = kblocks / kps;                               // stages per super-tile
    const uint32_t stage_bytes = kps * STAGE_PLANE_BYTES;
    const uint32_t nstages = (uint32_t)pair_stages(KC, (int)p.dim);   // a multiple of spt
    unsigned char *ring = smem;
    float *list_sc = reinterpret_cast<float *>(smem + nstages * stage_bytes);  // [KC][128]
    uint32_t *list_row = reinterpret_cast<uint32_t *>(list_sc + KC * TILE_Q);  // [KC][128]
    uint64_t *bars = reinterpret_cast<uint64_t *>(list_row + KC * TILE_Q);
    uint64_t *full = bars;                          // [stages]  leader only: both CTAs' copies -> MMA
    uint64_t *done = bars + PAIR_MAX_STAGES;        // [stages]  MMA -> producer and epilogue of BOTH CTAs
    uint64_t *acc_free = done + PAIR_MAX_STAGES;    // [2]       this CTA's epilogue -> leader's MMA / peer's forwarder
    uint64_t *acc_free_peer = acc_free + 2;         // [2]       leader only: the peer's forwarder -> MMA
    uint64_t *dummy = acc_free_peer + 2;            // probe builds: target of extra commits, never waited on
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dummy + 1);

    // warp index through a shuffle: the compiler then KNOWS it is warp-uniform, so the role branches below are uniform
    // control flow and the MMA issue loop keeps descriptors / TMEM addresses in uniform registers (UIADD3 + UTCHMMA
    // instead of ~15 instructions with four R2UR per MMA, which made one issuing warp slower than the tensor pipe)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t half = crank & 1u;                    // 0 = leader of the pair
    const uint32_t leader_rank = crank & ~1u;
    const uint32_t qt = blockIdx.x, part = p.part_base + blockIdx.y;
    const uint32_t acols = p.dim / 2;
    // row partition in units of super-tiles (2 consecutive 64-row tiles = one accumulator)
    // (tile_first, tile_per and tile_end are even for this kernel: the host rounds them to super-tiles)
    const uint32_t n_super = (p.tile_end + 1) / 2;
    const uint32_t u0 = min(p.tile_first / 2 + blockIdx.y * (p.tile_per / 2), n_super), u1 = min(u0 + p.tile_per / 2, n_super);

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nstages; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_free[b], EPI_THREADS / 32); mbar_init(&acc_free_peer[b], 1); }
        mbar_init(dummy, 1u << 19);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();                       // both CTAs' barriers exist before any remote arrive / multicast commit / peer copy
    if (warp == 1) {                          // same warp in both CTAs of the pair
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0) __trap();            // 1 CTA/SM owns all 512 columns: the allocation starts at 0
    constexpr uint32_t tmem = 0;
    const uint32_t acc_col = acols;           // [queries: dim/2 columns][2 accumulators of 128]

    // ---- epilogue warps stage this CTA's query tile into its TMEM (see batch_scan_kernel)
    float qscale = 1.0f;
    if (warp >= 2) {
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
        const int fmt = (int)p.fmt;
        const float *q = p.Q + ((size_t)qt * TILE_Q + m) * p.dim;
        float amax = 0.0f;
        for (uint32_t c = 0; c < p.dim; c += 4) {
            const float4 f = *reinterpret_cast<const float4 *>(q + c);
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y))), fmaxf(fabsf(f.z), fabsf(f.w)));
        }
        qscale = query_scale(amax);
        for (uint32_t c = 0; c < p.dim; c += 16) {
            uint32_t hi[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 f = *reinterpret_cast<const float4 *>(q + c + 4 * e);
                hi[2 * e] = pack16(fmt, f.x * qscale, f.y * qscale);
                hi[2 * e + 1] = pack16(fmt, f.z * qscale, f.w * qscale);
            }
            tmem_st8(lane_addr + c / 2, hi);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // the leader's MMAs read the peer's TMEM-resident queries too
    tc_fence_after();

    if (warp == 0) {
        // ===== TMA producer (both CTAs): this CTA's half of every super-tile = the hi plane of tile 2u + half =====
        uint32_t stage = 0, phase = 0;
        for (uint32_t u = u0; u < u1; ++u) {
            uint32_t t = 2 * u + half;
            if (t >= p.n_tiles) t = p.n_tiles - 1;       // odd tile count: the peer re-reads the last tile, masked in the epilogue
            for (uint32_t j = 0; j < spt; ++j) {
                PAIR_WAIT(&done[stage], phase ^ 1);            // the MMAs that read this stage's previous occupant have retired
                if (PROBES && (p.debug & 32768)) {                      // probe: no loads at all (the MMAs read stale shared memory)
                    if (half == 0 && lane == 0) mbar_arrive(&full[stage]);
                    __syncwarp();
                } else {
                if (half == 0) mbar_expect_tx(&full[stage], 2 * stage_bytes);   // my copy + the peer's
                tma_load_3d_pair(ring + stage * stage_bytes, &tmap, 0u, 0u, t * kblocks + j * kps,
                                 mapa_u32(&full[stage], leader_rank));
                }
                if (++stage == nstages) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1 && half == 0) {
        // ===== MMA issuer (leader): dim/16 UMMAs of 256 x 128 x 16 per super-tile, one commit per stage =====
        const uint32_t idesc = make_idesc((int)p.fmt, 2 * TILE_Q, PAIR_N);
        uint32_t st = 0, ph = 0, it = 0;
        for (uint32_t u = u0; u < u1; ++u, ++it) {
            const uint32_t buf = it & 1, use = it >> 1;
            const uint32_t d_tmem = tmem + acc_col + buf * PAIR_N;
            PAIR_WAIT(&acc_free[buf], (use & 1) ^ 1);          // my epilogue has drained this accumulator ...
            PAIR_WAIT(&acc_free_peer[buf], (use & 1) ^ 1);     // ... and so has the peer's
            tc_fence_after();
            for (uint32_t j = 0; j < spt; ++j) {
                PAIR_WAIT(&full[st], ph);                      // both halves of the stage have landed
                tc_fence_after();
                const uint32_t sb = smem_u32(ring + st * stage_bytes);
                if (!(PROBES && (p.debug & 32))) {
                    for (uint32_t kk = 0; kk < kps; ++kk) {
                        const uint32_t kb = j * kps + kk;
#pragma unroll
                        for (int i = 0; i < BLOCK_K / UMMA_K; ++i) {
                            const uint32_t a = tmem + kb * (BLOCK_K / 2) + i * (UMMA_K / 2);
                            const uint64_t b = make_b_desc(sb + kk * STAGE_PLANE_BYTES + i * 2 * (TILE_N * 16), TILE_N * 16, 128);
                            umma_ts_pair(d_tmem, a, b, idesc, (kb | i) != 0);
                        }
                    }
                }
                if (PROBES) {                                          // probe: what does one more commit per stage cost?
                    for (uint32_t x = 0; x < ((p.debug >> 9) & 7u); ++x) umma_commit_pair_multicast(dummy, (uint16_t)3);
                    for (uint32_t x = 0; x < ((p.debug >> 12) & 7u); ++x) umma_commit_pair_local(dummy);
                }
                if (PROBES && (p.debug & 128)) {                       // probe (with 32: no MMAs to wait for): plain arrives instead of the commit
                    if (lane == 0) { mbar_arrive(&done[st]); mbar_arrive_remote(&done[st], leader_rank + 1); }
                    __syncwarp();
                } else
                umma_commit_pair_multicast(&done[st], (uint16_t)3);    // both CTAs: stage free (last stage: accumulator ready)
                if (++st == nstages) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== peer: tell the leader's MMA warp when my epilogue has drained an accumulator buffer =====
        uint32_t it = 0;
        for (uint32_t u = u0; u < u1; ++u, ++it) {
            const uint32_t buf = it & 1, use = it >> 1;
            mbar_wait_cluster(&acc_free[buf], use & 1);
            if (lane == 0) mbar_arrive_remote(&acc_free_peer[buf], leader_rank);
            __syncwarp();
        }
    } else {
        // ===== epilogue: thread m owns query m of this CTA's tile; private candidate list in shared memory =====
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16) + acc_col;
        float thr = -INFINITY;                 // lowest score kept once the list is full (scaled like the scores)
        int cnt = 0, min_pos = 0;
        uint32_t it = 0, st = 0, ph = 0;       // st = first stage of the super-tile; its last stage carries "accumulator ready"
        for (uint32_t u = u0; u < u1; ++u, ++it) {
            const uint32_t buf = it & 1;
            PAIR_WAIT_EPI(&done[st + spt - 1], ph);            // nstages is a multiple of spt: no wrap inside a super-tile
            st += spt;
            if (st == nstages) { st = 0; ph ^= 1; }
            tc_fence_after();
            uint32_t r[4][32];
            if (!(PROBES && (p.debug & 16))) {
#pragma unroll
                for (int h = 0; h < 4; ++h) tmem_ld32(lane_addr + buf * PAIR_N + h * 32, r[h]);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            } else {
#pragma unroll
                for (int h = 0; h < 4; ++h)
#pragma unroll
                    for (int c = 0; c < 32; ++c) r[h][c] = 0xff800000u;   // -inf: nothing survives
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_free[buf]);            // local: the MMA warp / forwarder of this CTA waits on it
            const uint32_t row0 = u * PAIR_N;  // column c <-> corpus row row0 + c
            if (PROBES && (p.debug & 1)) continue;                 // probe: no scan at all
            // Pass 0/1: per group of 16 scores, the maximum first; a group is compared element by element only if
            // some lane of the warp has a score above its admission threshold there (see batch_scan_kernel)
            uint32_t mask[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const int h = g >> 1, c0 = (g & 1) * 16;
                bool scan = true;
                if (!(p.debug & 8)) {
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
            const uint32_t live = p.n_rows - row0;                // rows of this super-tile that exist (>= 1)
            if (live < (uint32_t)PAIR_N) {
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const uint32_t lo = (uint32_t)h * 32u;
                    mask[h] &= live <= lo ? 0u : (live - lo >= 32u ? 0xffffffffu : ((1u << (live - lo)) - 1u));
                }
            }
            if (PROBES && (p.debug & 2)) { if (mask[0] | mask[1] | mask[2] | mask[3]) thr = fmaxf(thr, -1e30f); continue; }
            // Pass 2 (rare, not unrolled): insert the survivors
#pragma unroll
            for (int h = 0; h < 4; ++h) {
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
                    thr = list_insert<KC>(list_sc, list_row, m, v, row0 + h * 32 + c, cnt, min_pos, thr);
                }
            }
        }
        // publish this partition's candidates for the CTA's queries
        const size_t q = (size_t)qt * TILE_Q + m;
        uint32_t *out = p.cand_rows + (q * p.parts + part) * KC;
        float *out_sc = p.cand_sc + (q * p.parts + part) * KC;
        const float unscale = 1.0f / qscale;                                // a power of two: exact
        for (int i = 0; i < KC; ++i) {
            out[i] = i < cnt ? list_row[i * TILE_Q + m] : 0xffffffffu;
            out_sc[i] = i < cnt ? list_sc[i * TILE_Q + m] * unscale : -INFINITY;
        }
        p.cand_thr[q * p.parts + part] = (cnt == KC) ? thr * unscale : -INFINITY;
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                       // neither CTA leaves while its pair may still touch it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
    }
}

}  // namespace k3
}  // namespace sema

// k2_scan_tma.cuh — K2 variant fed by 1-D bulk async copies (TMA) through a shared-memory ring.
//
// Same contract and the same selection / merge machinery as scan_topk_kernel (k2_scan.cuh); only
// the way rows reach the FMAs differs: one persistent CTA per SM, warp 0 streams TILE_ROWS-row
// tiles (contiguous: ld == dim) into a ring of TMA_STAGES stages with cp.async.bulk + mbarrier
// complete_tx, 8 consumer warps read them back with conflict-free LDS.128.  Memory-level
// parallelism no longer depends on registers or occupancy: TMA_STAGES * tile bytes are in flight
// per SM.  Measured equal to the LDG kernel on large corpora (both sit at the HBM ceiling) and
// ~3 % faster around 1M rows; default for dim 384 and 768 (sema_index_set_scan_variant selects).
// Tiles are claimed dynamically and consecutive launches can be chained with programmatic
// dependent launch (query streams: sema_index_search_stream_device), see the kernel comment.
#pragma once
#include "k2_scan.cuh"
#include "ptx.cuh"

namespace sema {

// CW = consumer warps per CTA.  CW = 8: one CTA per SM, stages of 48 KB (32 rows at dim 384, 16 at dim 768), 192 KB
// in flight.  CW = 4: TWO CTAs per SM with stages of 24 KB each (the same 192 KB in flight per SM): when launches
// follow each other in a query stream, a block of the next launch becomes resident as soon as ONE of the SM's two
// blocks has merged and exited, so an SM is never entirely idle between launches (each launch's ramp and per-block
// merge used to idle its SM for ~5 us: 2.5 % of a 0.26 ms shard scan).
// CW = 9 is the third shape: 8 consumer warps and 48 KB stages like CW = 8, but only TWO stages per CTA and two CTAs per
// SM (the same 4 x 48 KB in flight per SM, and a successor launch's block can move in when one of the two exits).
template <int CW>
__host__ __device__ constexpr int tma_warps() { return CW == 9 ? 8 : CW; }
template <int CW>
__host__ __device__ constexpr int tma_stages() { return CW == 9 ? 2 : 4; }
template <int CW>
__host__ __device__ constexpr int tma_threads() { return (tma_warps<CW>() + 1) * 32; }
template <int NV, int CW>
__host__ __device__ constexpr int tma_tile_rows() { return (NV <= 3 ? 4 : 2) * tma_warps<CW>(); }   // rows per warp and tile: 4 (dim 384) / 2 (dim 768)
template <int NV, int CW>
__host__ __device__ constexpr int tma_stage_bytes() { return tma_tile_rows<NV, CW>() * NV * 32 * 16; }
template <int NV, int CW>
__host__ __device__ constexpr int tma_smem_bytes() { return tma_stages<CW>() * tma_stage_bytes<NV, CW>() + 256; }

// The query itself as a kernel parameter (host-query entry points): it travels with the launch,
// no H2D copy precedes the kernel.  QueryArg<0> is the placeholder of the device-pointer variant.
template <int NV>
struct QueryArg { float4 v[NV * 32]; };
template <>
struct QueryArg<0> { uint32_t unused; };

constexpr uint32_t TMA_NO_TILE = 0xffffffffu;

// Programmatic dependent launch: a launch carrying the programmatic-stream-serialization attribute
// may start once every block of its predecessor has executed launch_dependents (or exited);
// griddepcontrol.wait then blocks until the predecessor has completed and its writes are visible.
// Both are no-ops in a launch without the attribute / without a dependent.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Tiles are scheduled dynamically: a block's first tile is blockIdx.x, every further one is claimed
// with an atomicAdd on *p.work_ctr (two claims in flight per block, so the ~1 us round trip never
// stalls the ring).  The tile index reaches the consumers through tile_of[stage], published before
// the producer's arrive.expect_tx; TMA_NO_TILE marks the end.  With back-to-back launches chained
// by programmatic dependent launch, a block that starts late (its SM was still running the
// predecessor's last-block merge / shard exchange) simply claims fewer tiles.
template <int NV, int M, int METRIC, bool QP, int CW>
__global__ void __launch_bounds__(tma_threads<CW>(), CW == 8 ? 1 : 2)
scan_topk_tma_kernel(const __grid_constant__ ScanParams p, const __grid_constant__ QueryArg<QP ? NV : 0> qa)
{
    using namespace ptx;
    constexpr int TMA_CONSUMER_WARPS = tma_warps<CW>();
    constexpr int TMA_STAGES = tma_stages<CW>();
    constexpr int TMA_TILE_ROWS = tma_tile_rows<NV, CW>();
    constexpr int R = TMA_TILE_ROWS / TMA_CONSUMER_WARPS;   // rows per warp and tile (4 or 2)
    constexpr int STAGE = tma_stage_bytes<NV, CW>();
    extern __shared__ __align__(128) unsigned char tsm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(tsm + TMA_STAGES * STAGE);
    uint64_t *empty = full + TMA_STAGES;
    __shared__ uint64_t sm_keys[TMA_CONSUMER_WARPS * 32 * M];
    __shared__ uint32_t tile_of[TMA_STAGES];
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = p.n;
    const int k = (int)p.k;
    const uint32_t n_tiles = (n + TMA_TILE_ROWS - 1) / TMA_TILE_ROWS;

    // the producer's first claim goes out before anything else: its latency overlaps the set-up
    uint32_t c0 = blockIdx.x, c1 = 0;
    if (warp == TMA_CONSUMER_WARPS && lane == 0) c1 = atomicAdd(p.work_ctr, 1u) + gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TMA_CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    WarpTopK<M> top;
    top.init();

    if (warp == TMA_CONSUMER_WARPS) {
        // ===== producer =====
        uint32_t stage = 0, phase = 0;
        auto issue = [&](uint32_t t) {
            const uint32_t rows = min((uint32_t)TMA_TILE_ROWS, n - t * TMA_TILE_ROWS);
            const uint32_t bytes = rows * NV * 32 * 16;
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) tile_of[stage] = t;
            __syncwarp();
            mbar_expect_tx(&full[stage], bytes);
            bulk_g2s(tsm + stage * STAGE, p.X + (size_t)t * TMA_TILE_ROWS * p.ld4, bytes, &full[stage]);
            if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
        };
        for (;;) {
            uint32_t t = __shfl_sync(FULL, c0, 0);
            if (t >= n_tiles) break;
            if (lane == 0) c0 = atomicAdd(p.work_ctr, 1u) + gridDim.x;
            issue(t);
            t = __shfl_sync(FULL, c1, 0);
            if (t >= n_tiles) break;
            if (lane == 0) c1 = atomicAdd(p.work_ctr, 1u) + gridDim.x;
            issue(t);
        }
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) {
            tile_of[stage] = TMA_NO_TILE;
            mbar_arrive(&full[stage]);     // completes the phase: no bytes expected
        }
        __syncwarp();
    } else {
        // ===== consumers: warp w owns rows w*R .. w*R+R-1 of every tile =====
        const uint64_t bound = p.bound ? *p.bound : ~0ull;
        float4 qv[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            if constexpr (QP) qv[v] = qa.v[v * 32 + lane];
            else qv[v] = reinterpret_cast<const float4 *>(p.q)[v * 32 + lane];
        }
        if constexpr (QP) {
            if (p.normalize_query) {
                // K1's rule and order, which is the reference's (src/semantic/embeddings.rs:84): a sequential
                // fold over j = 0 .. dim-1, multiply then add.  The query sits in the kernel parameters
                // (constant bank, uniform reads), so every thread simply walks it in order — ~1 us, hidden
                // behind the first TMA round trip; a non-finite query becomes NaN (never matches)
                float ss = 0.0f;
#pragma unroll 4
                for (int j = 0; j < NV * 32; ++j) {
                    const float4 x = qa.v[j];
                    ss = __fadd_rn(ss, __fmul_rn(x.x, x.x)); ss = __fadd_rn(ss, __fmul_rn(x.y, x.y));
                    ss = __fadd_rn(ss, __fmul_rn(x.z, x.z)); ss = __fadd_rn(ss, __fmul_rn(x.w, x.w));
                }
                const bool ok = ss <= 3.402823466e+38f;
                const float norm = sqrtf(ss), qnan = __uint_as_float(0x7fc00000u);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    if (!ok) qv[v] = make_float4(qnan, qnan, qnan, qnan);
                    else if (norm > 0.0f) { qv[v].x = qv[v].x / norm; qv[v].y = qv[v].y / norm; qv[v].z = qv[v].z / norm; qv[v].w = qv[v].w / norm; }
                }
            }
        }
        const int my_r = row_of_lane<R>(lane);
        const bool rep = (lane & (32 / R - 1)) == 0;
        uint32_t stage = 0, phase = 0;
        for (;;) {
            mbar_wait(&full[stage], phase);
            const uint32_t t = *reinterpret_cast<volatile uint32_t *>(&tile_of[stage]);
            if (t == TMA_NO_TILE) break;
            const float4 *tile = reinterpret_cast<const float4 *>(tsm + stage * STAGE) + (size_t)warp * R * NV * 32 + lane;
            float acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r] = 0.0f;
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[r] = accum4<METRIC>(acc[r], tile[(r * NV + v) * 32], qv[v]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);   // this warp is done with the stage
            const float s = reduce_rows<R>(acc, lane);
            const uint32_t row = t * TMA_TILE_ROWS + warp * R + my_r;
            const float rank = (METRIC == METRIC_L2) ? -s : s;
            const uint64_t key = make_key(rank, p.row_base + row);
            top.offer(key, rep && row < n && s == s && key < bound, lane, k);   // rows past n read stale smem: masked
            if (++stage == TMA_STAGES) { stage = 0; phase ^= 1; }
        }
    }

    // Everything below touches state shared with the previous launch on the stream (partials,
    // ticket, result buffers, exchange slots): wait for it to have completed, then let the next
    // launch start scanning on the SMs this grid's blocks leave.
    griddep_wait();
    griddep_launch_dependents();

    // ---- block merge over the 8 consumer warps (the producer holds no list), last-block merge,
    // ---- optional shard exchange, result emission — shared with scan_topk_kernel
    finish_topk<M, METRIC, TMA_CONSUMER_WARPS + 1, TMA_CONSUMER_WARPS>(top, p, sm_keys, &is_last, warp, lane);
}

}  // namespace sema

// k2_stream.cuh — K2 for a query STREAM: one persistent launch scans the corpus once per query, back to back.
//
// Same scan, selection and result as scan_topk_tma_kernel (k2_scan_tma.cuh) — the reference's flat exact KNN,
// src/storage/lance_indexer.rs:121-126, once per query — but nq queries share ONE launch.  A chained (PDL) stream of
// nq launches still idles every SM for ~5 us per query: a block of launch i+1 cannot become resident until launch i's
// block on that SM has merged its lists and exited, and then has to refill its ring (2.5 % of a
// 0.26 ms shard scan at 8 GPUs).  Here nothing ever drains:
//
//   * warp 8, the TMA producer, claims work items w = q * n_tiles + t from ONE monotonic counter and keeps the
//     4 x 48 KB ring full across query boundaries;
//   * warps 0-7, the consumers, notice the query index of a tile change, park their sorted top-k lists in a
//     double-buffered shared-memory slot (32*M keys each), fetch the next query into registers and go on scanning —
//     a ~1 us pause the ring absorbs;
//   * warp 9, the FINISHER, does everything finish_topk does, off the scan's critical path: block merge of the eight
//     parked lists, block list to global, arrival ticket, and in the last block to arrive the merge over all blocks,
//     the fused shard exchange (Exchange, k2_scan.cuh) and the result — while the other nine warps scan query q+1.
//
// Ordering.  A block's finisher handles its queries strictly in order, so "every block has posted q" implies every
// block — the one that was last for q-1 included — is done with q-1: queries complete in order, `done` counts them,
// and a rank publishes query q to its peers only after it has consumed q-1 (the exchange slots alternate by parity,
// exactly as with chained launches).  Block lists and tickets alternate by query parity too; a finisher posts query q
// only once q-2 is complete (done >= q-1), so a fast block can never overwrite lists the merge of q-2 still reads.
// The host zeroes the control words (work counter, done, tickets, fault) on the stream before every launch.  Every
// wait on global memory is bounded: a wait that gives up raises `fault`, after which the finishers only keep the
// shared-memory handshake with their consumers going and mark the remaining queries failed (n_found = 0xffffffff).
#pragma once
#include "k2_scan_tma.cuh"

namespace sema {

constexpr int STREAM_CONSUMER_WARPS = 8;
constexpr int STREAM_THREADS = (STREAM_CONSUMER_WARPS + 2) * 32;
constexpr int STREAM_STAGES = 4;
constexpr int STREAM_CTL_BYTES = 256;    // barriers + per-stage tile / query index

template <int NV>
__host__ __device__ constexpr int stream_tile_rows() { return tma_tile_rows<NV, 8>(); }
template <int NV>
__host__ __device__ constexpr int stream_stage_bytes() { return tma_stage_bytes<NV, 8>(); }
template <int NV, int M>
__host__ __device__ constexpr int stream_smem_bytes()
{
    return STREAM_STAGES * stream_stage_bytes<NV>() + STREAM_CTL_BYTES + 2 * STREAM_CONSUMER_WARPS * 32 * M * 8;
}

struct StreamParams {
    const float4 *X;              // row-major, row stride ld4 float4 (== row length for these shapes)
    const float *Q;               // nq x (ld4 * 4) floats, 16-byte aligned
    uint64_t *partials;           // [2][gridDim.x][32*M] keys
    unsigned long long *work_ctr; // monotonic work-item counter (first item of a block is static: blockIdx.x)
    unsigned int *done;           // queries of this launch completely finished
    unsigned int *ticket;         // [2] arrival counters, by query parity
    unsigned int *fault;          // set when a wait gave up (cannot happen unless a block died): everyone stops posting
    uint64_t *res_ids;            // [nq][k]
    float *res_scores;            // [nq][k]
    uint32_t *res_nfound;         // [nq]
    uint32_t n, ld4, k, row_base, nq;
    Exchange x;                   // x.seq = sequence number of query 0; query i uses x.seq + i
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// nlists sorted lists (first k entries each, list stride 32*M keys) -> one warp's top-k, by selection rounds over
// chunks of 32 * SELECT_C keys; the next chunk's loads are in flight while the current one is selected from
template <int M>
__device__ __forceinline__ void stream_select_lists(WarpTopK<M> &top, const uint64_t *lists, int nlists, int k, int lane, bool cg)
{
    constexpr int C = SELECT_C;
    const int totalk = nlists * k;
    uint64_t c[C + 1], nx[C];
    auto load = [&](int base, uint64_t *d) {
#pragma unroll
        for (int i = 0; i < C; ++i) {
            const int idx = base + i * 32 + lane;
            const int bb = idx / k, e = idx - bb * k;
            const uint64_t *src = lists + (size_t)bb * 32 * M + e;
            d[i] = idx < totalk ? (cg ? __ldcg(src) : __ldcv(src)) : 0ull;
        }
    };
    top.init();
    load(0, c);
    for (int base = 0; base < totalk; base += 32 * C) {
        const bool more = base + 32 * C < totalk;      // warp-uniform
#pragma unroll
        for (int i = 0; i < C; ++i) nx[i] = 0ull;
        if (more) load(base + 32 * C, nx);
        c[C] = top.v[0];                               // lane r < k: the r-th best so far
        warp_select<M, C + 1>(top, c, k, lane);
#pragma unroll
        for (int i = 0; i < C; ++i) c[i] = nx[i];
    }
}

// nlists whole sorted lists (32*M keys each) -> one warp's sorted list, by bitonic merges, U lists loaded ahead
template <int M>
__device__ __forceinline__ void stream_merge_lists(WarpTopK<M> &top, const uint64_t *lists, int nlists, int k, int lane, bool cg)
{
    constexpr int U = M == 4 ? 2 : 4;
    top.init();
    for (int b0 = 0; b0 < nlists; b0 += U) {
        uint64_t brev[U][M];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int bb = b0 + u;
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const uint64_t *src = lists + (size_t)bb * 32 * M + (M - 1 - j) * 32 + (31 - lane);
                brev[u][j] = bb < nlists ? (cg ? __ldcg(src) : __ldcv(src)) : 0ull;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) bitonic_merge_in<M>(top, brev[u], lane);
    }
    set_threshold<M>(top, k);
}

template <int NV, int M, int METRIC>
__global__ void __launch_bounds__(STREAM_THREADS, 1)
scan_stream_kernel(const __grid_constant__ StreamParams p)
{
    using namespace ptx;
    constexpr int CW = STREAM_CONSUMER_WARPS;
    constexpr int TILE_ROWS = stream_tile_rows<NV>();
    constexpr int R = TILE_ROWS / CW;
    constexpr int STAGE = stream_stage_bytes<NV>();
    constexpr int LIST = 32 * M;                         // keys per parked list
    extern __shared__ __align__(128) unsigned char tsm[];
    unsigned char *ctl = tsm + STREAM_STAGES * STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(ctl);              // [STAGES] tile landed
    uint64_t *empty = full + STREAM_STAGES;                          // [STAGES] all consumer warps done with the stage
    uint64_t *parked = empty + STREAM_STAGES;                        // [2] the eight lists of a query are in their slot
    uint64_t *slot_free = parked + 2;                                // [2] the finisher has read the slot
    uint32_t *tile_of = reinterpret_cast<uint32_t *>(ctl + 128);     // [STAGES]
    uint32_t *tile_q = tile_of + STREAM_STAGES;                      // [STAGES]
    uint64_t *sm_lists = reinterpret_cast<uint64_t *>(ctl + STREAM_CTL_BYTES);   // [2][CW][LIST]

    const int lane = threadIdx.x & 31, warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);
    const uint32_t n = p.n, nq = p.nq;
    const int k = (int)p.k;
    const uint32_t n_tiles = (n + TILE_ROWS - 1) / TILE_ROWS;

    // the producer's first claim goes out before anything else: its latency overlaps the set-up
    unsigned long long w0 = blockIdx.x, w1 = 0;
    if (warp == CW && lane == 0) w1 = atomicAdd(p.work_ctr, 1ull) + gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STREAM_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
        for (int b = 0; b < 2; ++b) { mbar_init(&parked[b], CW); mbar_init(&slot_free[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == CW) {
        // ===== producer: work item w = q * n_tiles + t =====
        const unsigned long long total = (unsigned long long)nq * n_tiles;
        unsigned long long qbase = 0;     // cq * n_tiles
        uint32_t cq = 0, stage = 0, phase = 0;
        auto issue = [&](unsigned long long w) {
            while (w - qbase >= n_tiles) { qbase += n_tiles; ++cq; }
            const uint32_t t = (uint32_t)(w - qbase);
            const uint32_t rows = min((uint32_t)TILE_ROWS, n - t * TILE_ROWS);
            const uint32_t bytes = rows * NV * 32 * 16;
            mbar_wait(&empty[stage], phase ^ 1);
            if (lane == 0) { tile_of[stage] = t; tile_q[stage] = cq; }
            __syncwarp();
            mbar_expect_tx(&full[stage], bytes);
            bulk_g2s(tsm + stage * STAGE, p.X + (size_t)t * TILE_ROWS * p.ld4, bytes, &full[stage]);
            if (++stage == STREAM_STAGES) { stage = 0; phase ^= 1; }
        };
        for (;;) {
            unsigned long long w = __shfl_sync(FULL, w0, 0);
            if (w >= total) break;
            if (lane == 0) w0 = atomicAdd(p.work_ctr, 1ull) + gridDim.x;
            issue(w);
            w = __shfl_sync(FULL, w1, 0);
            if (w >= total) break;
            if (lane == 0) w1 = atomicAdd(p.work_ctr, 1ull) + gridDim.x;
            issue(w);
        }
        mbar_wait(&empty[stage], phase ^ 1);
        if (lane == 0) {
            tile_of[stage] = TMA_NO_TILE;
            tile_q[stage] = nq;            // the consumers park every query up to nq - 1
            mbar_arrive(&full[stage]);     // completes the phase: no bytes expected
        }
        __syncwarp();
    } else if (warp < CW) {
        // ===== consumers: warp w owns rows w*R .. w*R+R-1 of every tile =====
        WarpTopK<M> top;
        top.init();
        float4 qv[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) qv[v] = reinterpret_cast<const float4 *>(p.Q)[v * 32 + lane];
        const int my_r = row_of_lane<R>(lane);
        const bool rep = (lane & (32 / R - 1)) == 0;
        uint32_t stage = 0, phase = 0, cur_q = 0;
        for (;;) {
            mbar_wait(&full[stage], phase);
            const uint32_t t = *reinterpret_cast<volatile uint32_t *>(&tile_of[stage]);
            const uint32_t tq = *reinterpret_cast<volatile uint32_t *>(&tile_q[stage]);
            if (tq != cur_q) {
                // query boundary: park the finished lists (empty ones for queries this block saw no tile of)
                while (cur_q < tq) {
                    const uint32_t b = cur_q & 1, u = cur_q >> 1;
                    mbar_wait(&slot_free[b], (u & 1) ^ 1);
                    top.store(sm_lists + ((size_t)b * CW + warp) * LIST, lane);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&parked[b]);
                    top.init();
                    ++cur_q;
                }
                if (tq < nq) {
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        qv[v] = reinterpret_cast<const float4 *>(p.Q + (size_t)tq * p.ld4 * 4)[v * 32 + lane];
                }
            }
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
            const uint32_t row = t * TILE_ROWS + warp * R + my_r;
            const float rank = (METRIC == METRIC_L2) ? -s : s;
            const uint64_t key = make_key(rank, p.row_base + row);
            top.offer(key, rep && row < n && s == s, lane, k);   // rows past n read stale smem: masked
            if (++stage == STREAM_STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===== finisher: block merge, post, last-block merge, exchange, result — one query after the other =====
        WarpTopK<M> top;
        const bool by_rounds = M == 1 && k <= SELECT_MAX_K;
        const uint32_t grid = gridDim.x;
        bool dead = false;                                          // warp-uniform: a wait gave up somewhere
        for (uint32_t q = 0; q < nq; ++q) {
            const uint32_t b = q & 1, u = q >> 1;
            top.init();
            mbar_wait(&parked[b], u & 1);
            if (dead) {
                if (lane == 0) { mbar_arrive(&slot_free[b]); p.res_nfound[q] = 0xffffffffu; }
                continue;
            }
            const uint64_t *lists = sm_lists + (size_t)b * CW * LIST;
            if (by_rounds) {
                static_assert(CW * SELECT_MAX_K <= 4 * 32, "four candidates per lane");
                uint64_t c[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int f = i * 32 + lane, w = f / k, e = f - w * k;
                    c[i] = f < CW * k ? lists[w * LIST + e] : 0ull;
                }
                warp_select<M, 4>(top, c, k, lane);
            } else {
#pragma unroll
                for (int j = 0; j < M; ++j) top.v[j] = lists[j * 32 + lane];
                for (int w = 1; w < CW; ++w) {
                    const uint64_t *l = lists + w * LIST;
                    if (l[0] == 0) continue;                       // empty list (warp-uniform)
                    uint64_t brev[M];
#pragma unroll
                    for (int j = 0; j < M; ++j) brev[j] = l[(M - 1 - j) * 32 + (31 - lane)];
                    bitonic_merge_in<M>(top, brev, lane);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_free[b]);             // the consumers may park query q + 2

            // the block lists of parity b are free once query q - 2 is complete
            if (q >= 2) {
                unsigned bad = 0;
                if (lane == 0) {
                    const long long t0 = clock64();
                    while (ld_acquire_gpu_u32(p.done) < q - 1) {
                        if (ld_acquire_gpu_u32(p.fault)) { bad = 1; break; }
                        if (clock64() - t0 > 16000000000ll) { atomicExch(p.fault, 1u); bad = 1; break; }   // ~8 s
                        __nanosleep(32);
                    }
                }
                dead = __shfl_sync(FULL, bad, 0) != 0;
                __syncwarp();                                      // lane 0's acquire orders every lane's stores below
                if (dead) {
                    if (lane == 0) p.res_nfound[q] = 0xffffffffu;
                    continue;
                }
            }
            uint64_t *part = p.partials + (size_t)b * grid * LIST;
            top.store(part + (size_t)blockIdx.x * LIST, lane);
            __threadfence();
            __syncwarp();
            unsigned last = 0;
            if (lane == 0) last = atomicAdd(p.ticket + b, 1u) == grid - 1 ? 1u : 0u;
            last = __shfl_sync(FULL, last, 0);
            if (!last) continue;
            __threadfence();

            // ---- last block to post query q: merge all block lists ----
            if (by_rounds) stream_select_lists<M>(top, part, (int)grid, k, lane, true);
            else stream_merge_lists<M>(top, part, (int)grid, k, lane, true);

            bool timed_out = false;
            if (p.x.world >= 1) {
                // ---- fused exchange (see finish_topk): publish, wait for every rank, merge world x k keys ----
                const uint32_t world = p.x.world;
                const uint64_t seq = p.x.seq + q;
                const uint32_t slot = (uint32_t)(seq & 1);
                uint64_t *mine = p.x.peer[p.x.rank];
                for (uint32_t g = 0; g < world; ++g) {
                    uint64_t *dst = xchg_keys(p.x.peer[g], world, slot, p.x.rank);
#pragma unroll
                    for (int j = 0; j < M; ++j) dst[j * 32 + lane] = (j * 32 + lane < k) ? top.v[j] : 0ull;
                }
                __threadfence_system();
                __syncwarp();
                if ((uint32_t)lane < world) st_release_sys(xchg_flag(p.x.peer[lane], world, slot, p.x.rank), seq);
                bool ok = true;
                if ((uint32_t)lane < world) {
                    const uint64_t *f = xchg_flag(mine, world, slot, (uint32_t)lane);
                    const long long t0 = clock64();
                    while (ld_acquire_sys(f) < seq) {
                        if (clock64() - t0 > 4000000000ll) { ok = false; break; }   // ~2 s: a rank is missing
                        __nanosleep(64);
                    }
                }
                timed_out = !__all_sync(FULL, ok);
                const uint64_t *theirs = xchg_keys(mine, world, slot, 0);           // [world][XCHG_KEYS], XCHG_KEYS == 32 * 4
                if (by_rounds) {
                    static_assert(XCHG_MAX_WORLD * SELECT_MAX_K <= 8 * 32, "eight candidates per lane");
                    uint64_t c[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int f = i * 32 + lane, g = f / k, e = f - g * k;
                        c[i] = f < (int)world * k ? __ldcv(theirs + (size_t)g * XCHG_KEYS + e) : 0ull;
                    }
                    warp_select<M, 8>(top, c, k, lane);
                } else {
                    top.init();
                    for (uint32_t g = 0; g < world; ++g) {
                        const uint64_t *l = theirs + (size_t)g * XCHG_KEYS;
                        uint64_t brev[M];
#pragma unroll
                        for (int j = 0; j < M; ++j) brev[j] = __ldcv(l + (M - 1 - j) * 32 + (31 - lane));
                        bitonic_merge_in<M>(top, brev, lane);
                    }
                    set_threshold<M>(top, k);
                }
            }
            emit_results<M, METRIC>(top, k, nullptr, p.res_ids + (size_t)q * k, p.res_scores + (size_t)q * k, p.res_nfound + q, lane);
            if (timed_out && lane == 0) p.res_nfound[q] = 0xffffffffu;   // host reports the failure
            if (lane == 0) {
                p.ticket[b] = 0;
                __threadfence();
                st_release_gpu_u32(p.done, q + 1);
            }
            __syncwarp();
        }
    }
}

}  // namespace sema

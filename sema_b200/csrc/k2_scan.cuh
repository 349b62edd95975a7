// k2_scan.cuh — kernel K2: single-query exact scan with fused top-k (shared pieces + the
// register-fed variant; the default TMA-ring variant is in k2_scan_tma.cuh).
//
// Replaces LanceDB's flat KNN behind table.query().nearest_to(q)?.limit(k).execute()
// (reference: src/storage/lance_indexer.rs:121-126).  HBM-bandwidth bound: the
// N x ld fp32 matrix is streamed exactly once (algorithmic bytes N*d*4), the query
// lives in registers, and selection never leaves the SM until the per-block lists
// are merged by the last block to finish (single launch, graph-replayable).
//
// Work split (register-fed variant): a warp owns batches of R consecutive rows.  A row is NV*32 float4
// (NV = 3 for d=384, 6 for d=768); lane l loads float4 l, l+32, ... of each row,
// i.e. every warp-level LDG.128 covers 512 contiguous bytes.  The R partial sums
// per lane are reduced with a transposed butterfly (R-1 + log2(32/R) shuffles per R
// rows instead of 5R), after which lane L holds the finished score of one row and
// offers it to the warp's top-k list.
#pragma once
#include "common.cuh"

namespace sema {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_WARPS = SCAN_THREADS / 32;

// Peer exchange of the corpus-sharded mode, fused into K2's last block: every rank stores its
// local top-k keys straight into every rank's exchange buffer over NVLink (P2P stores on
// cudaIpc-mapped memory), raises a per-(slot, source) flag with release.sys, then waits for all
// ranks' flags and merges the world*k keys itself — scan + exchange + merge in ONE launch per rank.
constexpr int XCHG_MAX_WORLD = 16;
constexpr int XCHG_KEYS = 128;   // keys per (slot, source rank)
struct Exchange {
    uint64_t *peer[XCHG_MAX_WORLD];  // peer[g]: rank g's buffer as mapped in this process (peer[rank] = own)
    uint32_t world;                  // 0 = exchange disabled (plain single-index search)
    uint32_t rank;
    uint64_t seq;                    // search sequence number, identical on all ranks, starts at 1
    // buffer layout (uint64 units): keys[2][world][XCHG_KEYS], then flags[2][world]
};
__device__ __forceinline__ uint64_t *xchg_keys(uint64_t *buf, uint32_t world, uint32_t slot, uint32_t src)
{
    return buf + ((size_t)slot * world + src) * XCHG_KEYS;
}
__device__ __forceinline__ uint64_t *xchg_flag(uint64_t *buf, uint32_t world, uint32_t slot, uint32_t src)
{
    return buf + (size_t)2 * world * XCHG_KEYS + (size_t)slot * world + src;
}
__device__ __forceinline__ void st_release_sys(uint64_t *p, uint64_t v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct ScanParams {
    const float4 *X;        // row-major, row stride ld4 float4
    const float *q;         // ld floats, zero padded, device
    uint64_t *partials;     // gridDim.x * 32*M keys
    unsigned int *ticket;   // self-resetting arrival counter
    const uint64_t *bound;  // nullable: only keys < *bound compete (multi-pass k > 128)
    uint64_t *out_keys;     // k keys, best first, 0 = empty
    // optional fused decode (single pass), all nullable together
    uint64_t *res_ids;      // k global row ids
    float *res_scores;      // k scores (cosine) or distances (L2)
    uint32_t *res_nfound;
    uint32_t n;             // rows to scan (snapshot)
    uint32_t ld4;           // row stride in float4
    uint32_t k;
    uint32_t row_base;      // global id of local row 0
    // TMA variant only: dynamic tile scheduler.  Tiles gridDim.x, gridDim.x + 1, ... are claimed
    // with atomicAdd on *work_ctr (each block's first tile is static: blockIdx.x); the last block
    // resets the counter.  Consecutive launches alternate between two counters because a launch
    // chained with programmatic dependent launch starts scanning while its predecessor still merges.
    unsigned int *work_ctr;
    // Host-visible completion (sema_index_search with host buffers): res_* then point into mapped
    // pinned host memory and the last block stores host_seq to *host_flag (release.sys) after the
    // results, so the host can poll instead of issuing a D2H copy and a stream synchronise.
    uint64_t *host_flag;
    uint64_t host_seq;
    // Query passed by kernel parameter only: apply the mean_pool normalise tail
    // (src/semantic/embeddings.rs:83-88) to it in registers, with kernel K1's arithmetic and
    // summation order, so the result is the one K1 would have written.
    uint32_t normalize_query;
    Exchange x;             // sharded mode (world > 1): fused peer exchange + global merge
};

__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

template <int METRIC>
__device__ __forceinline__ float accum4(float a, const float4 x, const float4 q)
{
    if (METRIC == METRIC_L2) {
        float t;
        t = q.x - x.x; a = fmaf(t, t, a);
        t = q.y - x.y; a = fmaf(t, t, a);
        t = q.z - x.z; a = fmaf(t, t, a);
        t = q.w - x.w; a = fmaf(t, t, a);
    } else {
        a = fmaf(x.x, q.x, a);
        a = fmaf(x.y, q.y, a);
        a = fmaf(x.z, q.z, a);
        a = fmaf(x.w, q.w, a);
    }
    return a;
}

// Transposed butterfly: in: acc[r] = this lane's partial of row r; out: the full sum
// of row rows_of_lane<R>(lane), replicated over the 32/R lanes that share it.
template <int R>
__device__ __forceinline__ float reduce_rows(float (&acc)[R], int lane)
{
#pragma unroll
    for (int half = R / 2, d = 16; half >= 1; half >>= 1, d >>= 1) {
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int r = 0; r < half; ++r) {
            const float keep = up ? acc[r + half] : acc[r];
            const float send = up ? acc[r] : acc[r + half];
            acc[r] = keep + __shfl_xor_sync(FULL, send, d);
        }
    }
#pragma unroll
    for (int d = 16 / R; d >= 1; d >>= 1) acc[0] += __shfl_xor_sync(FULL, acc[0], d);
    return acc[0];
}

template <int R>
__device__ __forceinline__ int row_of_lane(int lane)
{
    int r = 0;
#pragma unroll
    for (int half = R / 2, d = 16; half >= 1; half >>= 1, d >>= 1)
        if (lane & d) r += half;
    return r;
}

// Decode one key into the boundary's (row id, score) pair.
template <int METRIC>
__device__ __forceinline__ void decode_key(uint64_t key, uint64_t &id, float &score)
{
    const float rank = key_rank(key);
    id = key ? (uint64_t)key_gid(key) : 0ull;
    score = key ? (METRIC == METRIC_L2 ? -rank + 0.0f : rank) : 0.0f;
}

// Write the sorted key list held by warp 0 (and, optionally, its decoded form).
template <int M, int METRIC>
__device__ __forceinline__ void emit_results(const WarpTopK<M> &top, int k, uint64_t *out_keys,
                                             uint64_t *res_ids, float *res_scores,
                                             uint32_t *res_nfound, int lane)
{
    int found = 0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const int e = j * 32 + lane;
        const bool in = e < k;
        if (in && out_keys) out_keys[e] = top.v[j];
        found += __popc(__ballot_sync(FULL, in && top.v[j] != 0));
        if (res_ids && in) decode_key<METRIC>(top.v[j], res_ids[e], res_scores[e]);
    }
    if (res_nfound && lane == 0) *res_nfound = (uint32_t)found;
}

// ---- selection by rounds, for small k ------------------------------------------------------
// The merges after the scan sit on the latency path of every search (they are what is left of a
// query on a small corpus).  Inserting candidates one at a time into the sorted warp list costs a
// ~200-cycle shuffle chain per insertion, and merging P lists of k needs ~k(1 + ln P) of them.  For
// k <= SELECT_MAX_K it is cheaper to hold the candidates unsorted, C per lane, and extract the
// maximum k times: one round = C compares + a 5-step butterfly.  Keys are unique (the row id is
// part of the key), so clearing "the key equal to the maximum" removes exactly one candidate.
constexpr int SELECT_MAX_K = 16;
constexpr int SELECT_C = 10;     // candidates per lane the last-block merge can hold (148 blocks x k = 16 -> 2368 keys over 8 warps)

template <int M, int C>
__device__ __forceinline__ void warp_select(WarpTopK<M> &top, uint64_t (&c)[C], int k, int lane)
{
    uint64_t mine = 0;           // lane r ends up with the r-th best
    for (int r = 0; r < k; ++r) {
        uint64_t m = c[0];
#pragma unroll
        for (int i = 1; i < C; ++i) m = c[i] > m ? c[i] : m;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const uint64_t o = __shfl_xor_sync(FULL, m, d);
            m = o > m ? o : m;
        }
#pragma unroll
        for (int i = 0; i < C; ++i)
            if (c[i] == m) c[i] = 0;
        if (lane == r) mine = m;
    }
    top.init();
    top.v[0] = mine;
    top.thr = __shfl_sync(FULL, mine, k - 1);
}

// per-warp lists (first k entries each) -> warp 0's list, by selection; NACTIVE <= 8, k <= 16
template <int M, int NACTIVE>
__device__ __forceinline__ void block_select(WarpTopK<M> &top, uint64_t *sm, int warp, int lane, int k)
{
    static_assert(NACTIVE * SELECT_MAX_K <= 4 * 32, "four candidates per lane");
    if (warp < NACTIVE) top.store(sm + warp * 32 * M, lane);
    __syncthreads();
    if (warp == 0) {
        uint64_t c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int f = i * 32 + lane, w = f / k, e = f - w * k;
            c[i] = f < NACTIVE * k ? sm[w * 32 * M + e] : 0ull;
        }
        warp_select<M, 4>(top, c, k, lane);
    }
}

// ---- merging sorted lists, for larger k -----------------------------------------------------
// Every list in flight after the scan is sorted (descending over e = j*32 + lane, 32*M entries, zeros
// last).  Two such lists merge without any insertion: C[e] = max(A[e], B[32M-1-e]) holds the 32M
// largest keys of the union as a bitonic sequence (the half-cleaner property), which log2(32M)
// compare-exchange stages sort — the first log2(M) between registers, the last five by shuffle.
// brev[j] must hold the other list reversed: other[(M-1-j)*32 + (31-lane)].
template <int M>
__device__ __forceinline__ void bitonic_merge_in(WarpTopK<M> &top, const uint64_t (&brev)[M], int lane)
{
#pragma unroll
    for (int j = 0; j < M; ++j) top.v[j] = top.v[j] > brev[j] ? top.v[j] : brev[j];
#pragma unroll
    for (int jd = M / 2; jd >= 1; jd >>= 1)
#pragma unroll
        for (int j = 0; j < M; ++j)
            if ((j & jd) == 0) {
                const uint64_t a = top.v[j], b = top.v[j | jd];
                top.v[j] = a > b ? a : b;
                top.v[j | jd] = a > b ? b : a;
            }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const bool upper = (lane & d) != 0;
#pragma unroll
        for (int j = 0; j < M; ++j) {
            const uint64_t o = __shfl_xor_sync(FULL, top.v[j], d);
            const uint64_t hi = top.v[j] > o ? top.v[j] : o, lo = top.v[j] > o ? o : top.v[j];
            top.v[j] = upper ? lo : hi;
        }
    }
}

template <int M>
__device__ __forceinline__ void set_threshold(WarpTopK<M> &top, int k)
{
    const int kj = (k - 1) >> 5, kl = (k - 1) & 31;
    uint64_t t = 0;
#pragma unroll
    for (int j = 0; j < M; ++j)
        if (j == kj) t = top.v[j];
    top.thr = __shfl_sync(FULL, t, kl);
}

// per-warp sorted lists -> warp 0's list by bitonic merges
template <int M, int NACTIVE>
__device__ __forceinline__ void block_bitonic(WarpTopK<M> &top, uint64_t *sm, int warp, int lane, int k)
{
    if (warp < NACTIVE) top.store(sm + warp * 32 * M, lane);
    __syncthreads();
    if (warp == 0) {
        for (int w = 1; w < NACTIVE; ++w) {
            const uint64_t *l = sm + w * 32 * M;
            if (l[0] == 0) continue;                       // empty list (warp-uniform)
            uint64_t brev[M];
#pragma unroll
            for (int j = 0; j < M; ++j) brev[j] = l[(M - 1 - j) * 32 + (31 - lane)];
            bitonic_merge_in<M>(top, brev, lane);
        }
        set_threshold<M>(top, k);
    }
}

// Everything after a block's warps have scanned their rows, shared by the LDG and the TMA variant:
// per-warp lists -> block list -> global partials; the last block to arrive (atomic ticket) merges
// all block lists, optionally exchanges with the other shards (Exchange), and emits the result.
// NW = warps in the block, NACTIVE = warps (0..NACTIVE-1) that hold lists / take part in the merges.
template <int M, int METRIC, int NW, int NACTIVE>
__device__ __forceinline__ void finish_topk(WarpTopK<M> &top, const ScanParams &p, uint64_t *sm_keys, bool *is_last,
                                            int warp, int lane)
{
    const int k = (int)p.k;
    const bool by_rounds = M == 1 && k <= SELECT_MAX_K;     // small k: selection by rounds (see warp_select)
    if (by_rounds) block_select<M, NACTIVE>(top, sm_keys, warp, lane, k);
    else block_bitonic<M, NACTIVE>(top, sm_keys, warp, lane, k);
    if (warp == 0) top.store(p.partials + (size_t)blockIdx.x * 32 * M, lane);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) *is_last = (atomicAdd(p.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!*is_last) return;
    __threadfence();

    // Only the first k entries of each block list matter.  They are read as one flat array of
    // gridDim.x * k keys with all of a lane's loads in flight at once: this merge sits on the
    // critical path after the last block arrives, so its load latency must overlap, not add up.
    const int totalk = (int)gridDim.x * k;
    if (by_rounds && totalk <= NACTIVE * 32 * SELECT_C) {
        if (warp < NACTIVE) {
            uint64_t c[SELECT_C];
#pragma unroll
            for (int i = 0; i < SELECT_C; ++i) {
                const int idx = (i * NACTIVE + warp) * 32 + lane;
                const int bb = idx / k, e = idx - bb * k;
                c[i] = idx < totalk ? __ldcg(p.partials + (size_t)bb * 32 * M + e) : 0ull;
            }
            warp_select<M, SELECT_C>(top, c, k, lane);
        }
        block_select<M, NACTIVE>(top, sm_keys, warp, lane, k);
    } else {
        // whole block lists (32*M sorted keys each), U of them loaded ahead, merged without insertions
        top.init();
        constexpr int U = M == 4 ? 2 : 4;
        for (int b0 = warp; warp < NACTIVE && b0 < (int)gridDim.x; b0 += NACTIVE * U) {
            uint64_t brev[U][M];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int bb = b0 + u * NACTIVE;
#pragma unroll
                for (int j = 0; j < M; ++j)
                    brev[u][j] = bb < (int)gridDim.x ? __ldcg(p.partials + (size_t)bb * 32 * M + (M - 1 - j) * 32 + (31 - lane)) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) bitonic_merge_in<M>(top, brev[u], lane);
        }
        block_bitonic<M, NACTIVE>(top, sm_keys, warp, lane, k);
    }

    bool timed_out = false;
    if (p.x.world >= 1) {
        // ---- fused exchange: publish the shard's top-k to every rank, wait for theirs, merge ----
        const uint32_t world = p.x.world, slot = (uint32_t)(p.x.seq & 1);
        uint64_t *mine = p.x.peer[p.x.rank];
        if (warp == 0) {
            for (uint32_t g = 0; g < world; ++g) {
                uint64_t *dst = xchg_keys(p.x.peer[g], world, slot, p.x.rank);
#pragma unroll
                for (int j = 0; j < M; ++j) dst[j * 32 + lane] = (j * 32 + lane < k) ? top.v[j] : 0ull;
            }
            __threadfence_system();
            __syncwarp();
            if ((uint32_t)lane < world) st_release_sys(xchg_flag(p.x.peer[lane], world, slot, p.x.rank), p.x.seq);
            bool ok = true;
            if ((uint32_t)lane < world) {
                const uint64_t *f = xchg_flag(mine, world, slot, (uint32_t)lane);
                const long long t0 = clock64();
                while (ld_acquire_sys(f) < p.x.seq) {
                    if (clock64() - t0 > 4000000000ll) { ok = false; break; }   // ~2 s: a rank is missing
                    __nanosleep(64);
                }
            }
            ok = __all_sync(FULL, ok);
            if (lane == 0) *is_last = ok;   // reuse the shared flag to broadcast the outcome
        }
        __syncthreads();
        timed_out = !*is_last;
        if (by_rounds) {
            // world * k <= 16 * 16 keys: warp 0 selects alone, 8 candidates per lane
            static_assert(XCHG_MAX_WORLD * SELECT_MAX_K <= 8 * 32, "eight candidates per lane");
            if (warp == 0) {
                uint64_t c[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int f = i * 32 + lane, g = f / k, e = f - g * k;
                    c[i] = f < (int)world * k ? __ldcv(xchg_keys(mine, world, slot, (uint32_t)g) + e) : 0ull;
                }
                warp_select<M, 8>(top, c, k, lane);
            }
        } else {
            // every rank published a sorted list of 32*M keys (zeros past k): merge them like block lists
            top.init();
            for (int g = warp; warp < NACTIVE && g < (int)world; g += NACTIVE) {
                const uint64_t *l = xchg_keys(mine, world, slot, (uint32_t)g);
                uint64_t brev[M];
#pragma unroll
                for (int j = 0; j < M; ++j) brev[j] = __ldcv(l + (M - 1 - j) * 32 + (31 - lane));
                bitonic_merge_in<M>(top, brev, lane);
            }
            block_bitonic<M, NACTIVE>(top, sm_keys, warp, lane, k);
        }
    }
    if (warp == 0) {
        emit_results<M, METRIC>(top, k, p.out_keys, p.res_ids, p.res_scores, p.res_nfound, lane);
        if (timed_out && p.res_nfound && lane == 0) *p.res_nfound = 0xffffffffu;   // host reports the failure
        if (lane == 0) {
            *p.ticket = 0;
            if (p.work_ctr) *p.work_ctr = 0;   // every block has stopped claiming tiles (all passed the ticket)
        }
        if (p.host_flag) {
            __threadfence_system();            // each lane's result stores reach the host before the flag
            __syncwarp();
            if (lane == 0) st_release_sys(p.host_flag, p.host_seq);
        }
    }
}

// NV > 0: row is exactly NV*32 float4 (unrolled, query in registers).
// NV == 0: generic row length (query in shared memory, lane-strided loop).
template <int NV, int R, int M, int METRIC>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_topk_kernel(const ScanParams p)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ uint64_t sm_keys[SCAN_WARPS * 32 * M];
    __shared__ bool is_last;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t n = p.n;
    const uint32_t ld4 = p.ld4;
    const int k = (int)p.k;
    const uint64_t bound = p.bound ? *p.bound : ~0ull;

    constexpr int NVR = NV > 0 ? NV : 1;
    float4 qv[NVR];
    float4 *qs = reinterpret_cast<float4 *>(dyn_smem);
    if (NV > 0) {
#pragma unroll
        for (int v = 0; v < NVR; ++v) qv[v] = reinterpret_cast<const float4 *>(p.q)[v * 32 + lane];
    } else {
        for (uint32_t i = threadIdx.x; i < ld4; i += SCAN_THREADS)
            qs[i] = reinterpret_cast<const float4 *>(p.q)[i];
        __syncthreads();
    }

    WarpTopK<M> top;
    top.init();

    const uint32_t nb = (n + R - 1) / R;
    const uint32_t gw = blockIdx.x * SCAN_WARPS + warp;
    const uint32_t nw = gridDim.x * SCAN_WARPS;
    const int my_r = row_of_lane<R>(lane);
    const bool rep = (lane & (32 / R - 1)) == 0;

    // static stride over batches (a dynamic chunk scheduler was measured and gave nothing: the
    // kernel is limited by the memory system, not by SM-to-SM imbalance)
    for (uint32_t b = gw; b < nb; b += nw) {
        const uint32_t row0 = b * R;
        float acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = 0.0f;

        if (NV > 0) {
            float4 x[R][NVR];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t row = min(row0 + r, n - 1);
                const float4 *xp = p.X + (size_t)row * ld4 + lane;
#pragma unroll
                for (int v = 0; v < NVR; ++v) x[r][v] = ldg_stream(xp + v * 32);
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int v = 0; v < NVR; ++v) acc[r] = accum4<METRIC>(acc[r], x[r][v], qv[v]);
        } else {
            for (uint32_t i = lane; i < ld4; i += 32) {
                const float4 qq = qs[i];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t row = min(row0 + r, n - 1);
                    acc[r] = accum4<METRIC>(acc[r], ldg_stream(p.X + (size_t)row * ld4 + i), qq);
                }
            }
        }

        const float s = reduce_rows<R>(acc, lane);
        const uint32_t row = row0 + my_r;
        const float rank = (METRIC == METRIC_L2) ? -s : s;
        const uint64_t key = make_key(rank, p.row_base + row);
        top.offer(key, rep && row < n && s == s && key < bound, lane, k);
    }

    // ---- block merge, last-block merge, optional shard exchange, result emission ----
    finish_topk<M, METRIC, SCAN_WARPS, SCAN_WARPS>(top, p, sm_keys, &is_last, warp, lane);
}

}  // namespace sema

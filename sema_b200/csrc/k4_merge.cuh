// k4_merge.cuh — kernel K4: merge ranked key lists into the global top-k and decode.
//
// No reference analogue (the reference is single-process): this is the exchange step
// of the corpus-sharded mode: global top-k = top-k of the union of
// the per-shard top-k lists gathered over NVLink.  Latency bound: one block.
#pragma once
#include "k2_scan.cuh"

namespace sema {

constexpr int MERGE_THREADS = 256;
constexpr int MERGE_WARPS = MERGE_THREADS / 32;

// keys: total candidate keys (0 = empty).  Only keys < *bound compete (multi-pass k > 128).
template <int M, int METRIC>
__global__ void __launch_bounds__(MERGE_THREADS)
merge_topk_kernel(const uint64_t *keys, uint32_t total, int k, const uint64_t *bound_p,
                  uint64_t *out_keys, uint64_t *res_ids, float *res_scores, uint32_t *res_nfound)
{
    __shared__ uint64_t sm_keys[MERGE_WARPS * 32 * M];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t bound = bound_p ? *bound_p : ~0ull;
    WarpTopK<M> top;
    top.init();
    for (uint32_t base = warp * 32; base < total; base += MERGE_THREADS) {
        const uint32_t i = base + lane;
        const uint64_t key = i < total ? keys[i] : 0ull;
        top.offer(key, key != 0 && key < bound, lane, k);
    }
    block_merge<M, MERGE_WARPS>(top, sm_keys, warp, lane, k);
    if (warp == 0) emit_results<M, METRIC>(top, k, out_keys, res_ids, res_scores, res_nfound, lane);
}

// Batched K4 (sharded batch search): keys = [n_lists][nq][k] as all-gathered from the shards; block q
// merges the n_lists * k candidates of query q into res_*[q][k].
template <int M, int METRIC>
__global__ void __launch_bounds__(MERGE_THREADS)
merge_topk_batch_kernel(const uint64_t *keys, uint32_t n_lists, uint32_t nq, int k, uint64_t *res_ids, float *res_scores,
                        uint32_t *res_nfound)
{
    __shared__ uint64_t sm_keys[MERGE_WARPS * 32 * M];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x;
    const uint32_t total = n_lists * (uint32_t)k;
    WarpTopK<M> top;
    top.init();
    for (uint32_t base = warp * 32; base < total; base += MERGE_THREADS) {
        const uint32_t i = base + lane;
        uint64_t key = 0ull;
        if (i < total) {
            const uint32_t g = i / (uint32_t)k, e = i - g * (uint32_t)k;
            key = keys[((size_t)g * nq + q) * k + e];
        }
        top.offer(key, key != 0, lane, k);
    }
    block_merge<M, MERGE_WARPS>(top, sm_keys, warp, lane, k);
    if (warp == 0)
        emit_results<M, METRIC>(top, k, nullptr, res_ids + (size_t)q * k, res_scores + (size_t)q * k, res_nfound + q, lane);
}

// (ids, scores, n_found)[nq][k] -> packed ranking keys [nq][k] (0 past n_found): what a shard
// contributes to the all-gather of a batched search
template <int METRIC>
__global__ void pack_keys_kernel(const uint64_t *ids, const float *scores, const uint32_t *nf, uint32_t nq, uint32_t k,
                                 uint64_t *keys)
{
    const size_t total = (size_t)nq * k;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t q = (uint32_t)(i / k), e = (uint32_t)(i - (size_t)q * k);
        const float sc = scores[i];
        keys[i] = e < nf[q] ? make_key(METRIC == METRIC_L2 ? -sc : sc, (uint32_t)ids[i]) : 0ull;
    }
}

// keys[k] -> ids/scores/n_found (used after multi-pass selection)
template <int METRIC>
__global__ void decode_kernel(const uint64_t *keys, int k, uint64_t *res_ids, float *res_scores,
                              uint32_t *res_nfound)
{
    __shared__ int found;
    if (threadIdx.x == 0) found = 0;
    __syncthreads();
    int mine = 0;
    for (int e = threadIdx.x; e < k; e += blockDim.x) {
        const uint64_t key = keys[e];
        decode_key<METRIC>(key, res_ids[e], res_scores[e]);
        mine += key != 0;
    }
    if (mine) atomicAdd(&found, mine);
    __syncthreads();
    if (threadIdx.x == 0) *res_nfound = (uint32_t)found;
}

}  // namespace sema

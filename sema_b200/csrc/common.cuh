// common.cuh — ranking keys and the warp-resident top-k list shared by the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sema {

constexpr unsigned FULL = 0xffffffffu;
constexpr int METRIC_COSINE = 0;
constexpr int METRIC_L2 = 1;

// ---- packed ranking key ------------------------------------------------------
// key = ordered(rank value) << 32 | (0xFFFFFFFF - global row id); larger = better,
// exact ties resolve to the lower row id; 0 = empty slot (no finite or infinite
// float maps to ordered() == 0).
__device__ __forceinline__ uint32_t f32_ordered(float s)
{
    uint32_t u = __float_as_uint(s + 0.0f);  // -0 -> +0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_f32(uint32_t o)
{
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ uint64_t make_key(float rank, uint32_t gid)
{
    return ((uint64_t)f32_ordered(rank) << 32) | (uint64_t)(0xffffffffu - gid);
}
__device__ __forceinline__ uint32_t key_gid(uint64_t k) { return 0xffffffffu - (uint32_t)k; }
__device__ __forceinline__ float key_rank(uint64_t k) { return ordered_f32((uint32_t)(k >> 32)); }

// ---- warp-resident sorted list -------------------------------------------------
// 32*M keys, sorted descending over e = j*32 + lane.  Only the best k matter: thr is
// the key at position k-1 (0 while the list is not full), offers at or below it are
// dropped with one ballot per 32 candidates; insertions are rare (~k ln(n/k) per warp).
template <int M>
struct WarpTopK {
    uint64_t v[M];
    uint64_t thr;

    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int j = 0; j < M; ++j) v[j] = 0;
        thr = 0;
    }

    __device__ __forceinline__ void insert(uint64_t ck, int lane, int k)
    {
        int pos = 0;
#pragma unroll
        for (int j = 0; j < M; ++j) pos += __popc(__ballot_sync(FULL, v[j] > ck));
#pragma unroll
        for (int j = M - 1; j >= 0; --j) {
            uint64_t up = __shfl_up_sync(FULL, v[j], 1);
            if (j > 0) {
                uint64_t carry = __shfl_sync(FULL, v[j - 1], 31);
                if (lane == 0) up = carry;
            }
            const int e = j * 32 + lane;
            if (e > pos) v[j] = up;
            else if (e == pos) v[j] = ck;
        }
        const int kj = (k - 1) >> 5, kl = (k - 1) & 31;
        uint64_t t = 0;
#pragma unroll
        for (int j = 0; j < M; ++j)
            if (j == kj) t = v[j];
        thr = __shfl_sync(FULL, t, kl);
    }

    // every lane offers one candidate (or none); warp-convergent call
    __device__ __forceinline__ void offer(uint64_t key, bool valid, int lane, int k)
    {
        unsigned m = __ballot_sync(FULL, valid && key > thr);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t ck = __shfl_sync(FULL, key, src);
            if (ck > thr) insert(ck, lane, k);
        }
    }

    __device__ __forceinline__ void store(uint64_t *dst, int lane) const
    {
#pragma unroll
        for (int j = 0; j < M; ++j) dst[j * 32 + lane] = v[j];
    }
};

// Reduce the per-warp lists of a block into warp 0's list.  sm: WARPS*32*M keys.
template <int M, int WARPS>
__device__ __forceinline__ void block_merge(WarpTopK<M> &top, uint64_t *sm, int warp, int lane, int k)
{
    top.store(sm + warp * 32 * M, lane);
    __syncthreads();
    if (warp == 0) {
        const int chunks = (k + 31) >> 5;  // entries past k never matter
        for (int w = 1; w < WARPS; ++w)
            for (int j = 0; j < chunks; ++j) {
                const uint64_t key = sm[w * 32 * M + j * 32 + lane];
                top.offer(key, key != 0, lane, k);
            }
    }
}

}  // namespace sema

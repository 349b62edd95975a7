// k0_pool.cuh — kernel K0: masked mean pooling + L2 normalisation of token embeddings, the step
// that produces every vector the scan path stores or receives as a query.
//
// Restates mean_pool (reference: src/semantic/embeddings.rs:61-91):
//   pooled[j] = sum_i tokens[i][j] * mask[i]   (i ascending, f32, multiply then add — no FMA)
//   mask_sum  = sum_i mask[i];  pooled /= mask_sum  iff mask_sum > 0
//   norm = sqrt(sum_j pooled[j]^2) (j ascending);  pooled /= norm  iff norm > 0
// Every sum keeps the reference's order, so the result is bit-identical to it: one thread owns one
// column j and walks the tokens in order (rows are contiguous over j: coalesced), and one thread
// folds the squared pooled values in order.  HBM-bound: the tokens (n x seq x hidden fp32) are read
// once; with skip_masked the rows of padding tokens (mask == 0) are not read at all — identical
// results for finite inputs (x * 0 adds +-0 to a sum that is never -0).
#pragma once
#include "common.cuh"

namespace sema {

constexpr int POOL_MAX_THREADS = 512;   // keeps the register budget above what 16 loads in flight per thread need

// One block per text; texts are claimed from *next_text (atomicAdd, zeroed before the launch) by a
// grid of exactly the resident blocks, so texts of different attended length balance out and no
// second, partly filled wave exists.  Dynamic shared memory: (3 * seq + hidden) floats.
// out: row stride out_ld floats; columns [hidden, out_ld) are written as zeros.
// The tokens a text actually reads are first compacted into a list (index, mask value) — all of
// them, or with skip_masked those whose mask is not 0 — so that the accumulation loop is branch-free
// and unrolled with 16 independent loads in flight per thread (the adds stay in token order).
__global__ void __launch_bounds__(POOL_MAX_THREADS)
pool_kernel(const float *tokens, const float *mask, uint64_t n, uint32_t seq, uint32_t hidden, float *out,
            uint64_t out_ld, int skip_masked, unsigned long long *next_text)
{
    extern __shared__ float psm[];
    float *sm_mask = psm;                                          // [seq]   all mask values (mask_sum)
    float *sm_mv = psm + seq;                                      // [seq]   mask values of the listed tokens
    uint32_t *sm_idx = reinterpret_cast<uint32_t *>(psm + 2 * seq); // [seq]   their token indices, ascending
    float *sm_pool = psm + 3 * seq;                                // [hidden]
    __shared__ float sm_norm;
    __shared__ uint32_t sm_cnt;
    __shared__ unsigned long long sm_text;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (;;) {
        if (threadIdx.x == 0) sm_text = atomicAdd(next_text, 1ull);
        __syncthreads();
        const uint64_t t = sm_text;
        if (t >= n) break;
        const float *tok = tokens + t * (uint64_t)seq * hidden;
        for (uint32_t i = threadIdx.x; i < seq; i += blockDim.x) sm_mask[i] = mask[t * seq + i];
        __syncthreads();
        if (warp == 0) {                                           // ordered compaction: ballot + prefix count
            uint32_t cnt = 0;
            for (uint32_t i0 = 0; i0 < seq; i0 += 32) {
                const uint32_t i = i0 + lane;
                const float m = i < seq ? sm_mask[i] : 0.0f;
                const bool keep = i < seq && !(skip_masked && m == 0.0f);
                const unsigned b = __ballot_sync(FULL, keep);
                if (keep) {
                    const uint32_t pos = cnt + __popc(b & ((1u << lane) - 1u));
                    sm_idx[pos] = i;
                    sm_mv[pos] = m;
                }
                cnt += __popc(b);
            }
            if (lane == 0) sm_cnt = cnt;
        }
        __syncthreads();
        const uint32_t cnt = sm_cnt;
        float mask_sum = 0.0f;
        for (uint32_t i = 0; i < seq; ++i) mask_sum = __fadd_rn(mask_sum, sm_mask[i]);   // same value in every thread
        for (uint32_t j = threadIdx.x; j < hidden; j += blockDim.x) {
            const float *col = tok + j;
            float acc = 0.0f;
            uint32_t c = 0;
            for (; c + 16 <= cnt; c += 16) {
                float e[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) e[u] = __ldg(col + (uint64_t)sm_idx[c + u] * hidden);
#pragma unroll
                for (int u = 0; u < 16; ++u) acc = __fadd_rn(acc, __fmul_rn(e[u], sm_mv[c + u]));
            }
            for (; c < cnt; ++c) acc = __fadd_rn(acc, __fmul_rn(__ldg(col + (uint64_t)sm_idx[c] * hidden), sm_mv[c]));
            if (mask_sum > 0.0f) acc = acc / mask_sum;
            sm_pool[j] = acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float ss = 0.0f;
#pragma unroll 8
            for (uint32_t j = 0; j < hidden; ++j) ss = __fadd_rn(ss, __fmul_rn(sm_pool[j], sm_pool[j]));
            sm_norm = sqrtf(ss);
        }
        __syncthreads();
        const float norm = sm_norm;
        float *o = out + t * out_ld;
        for (uint32_t j = threadIdx.x; j < out_ld; j += blockDim.x) {
            float v = 0.0f;
            if (j < hidden) {
                v = sm_pool[j];
                if (norm > 0.0f) v = v / norm;
            }
            o[j] = v;
        }
        __syncthreads();   // the shared arrays are reused by the next text
    }
}

// float4 form for hidden % 4 == 0 (and 16-byte aligned rows): one thread owns four adjacent columns,
// so a text needs hidden / 4 threads (96 for dim 384) and an SM holds ~4x more texts at once — their
// serial pieces (mask compaction, the in-order norm) overlap with other texts' loads.  The four
// columns of a thread are independent sums in token order: the same bits as pool_kernel.
__global__ void __launch_bounds__(256)
pool_kernel_v4(const float *tokens, const float *mask, uint64_t n, uint32_t seq, uint32_t hidden, float *out,
               uint64_t out_ld, int skip_masked, unsigned long long *next_text)
{
    extern __shared__ float psm[];
    float *sm_mask = psm;
    float *sm_mv = psm + seq;
    uint32_t *sm_idx = reinterpret_cast<uint32_t *>(psm + 2 * seq);
    float *sm_pool = psm + 3 * seq;                                // [hidden], 16-byte aligned (seq % 4 == 0 is checked by the host)
    __shared__ float sm_norm;
    __shared__ uint32_t sm_cnt;
    __shared__ unsigned long long sm_text;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t h4 = hidden / 4;

    for (;;) {
        if (threadIdx.x == 0) sm_text = atomicAdd(next_text, 1ull);
        __syncthreads();
        const uint64_t t = sm_text;
        if (t >= n) break;
        const float4 *tok = reinterpret_cast<const float4 *>(tokens + t * (uint64_t)seq * hidden);
        for (uint32_t i = threadIdx.x; i < seq; i += blockDim.x) sm_mask[i] = mask[t * seq + i];
        __syncthreads();
        if (warp == 0) {
            uint32_t cnt = 0;
            for (uint32_t i0 = 0; i0 < seq; i0 += 32) {
                const uint32_t i = i0 + lane;
                const float m = i < seq ? sm_mask[i] : 0.0f;
                const bool keep = i < seq && !(skip_masked && m == 0.0f);
                const unsigned b = __ballot_sync(FULL, keep);
                if (keep) {
                    const uint32_t pos = cnt + __popc(b & ((1u << lane) - 1u));
                    sm_idx[pos] = i;
                    sm_mv[pos] = m;
                }
                cnt += __popc(b);
            }
            if (lane == 0) sm_cnt = cnt;
        }
        __syncthreads();
        const uint32_t cnt = sm_cnt;
        float mask_sum = 0.0f;
        for (uint32_t i = 0; i < seq; ++i) mask_sum = __fadd_rn(mask_sum, sm_mask[i]);
        for (uint32_t j = threadIdx.x; j < h4; j += blockDim.x) {
            const float4 *col = tok + j;
            float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            uint32_t c = 0;
            for (; c + 8 <= cnt; c += 8) {
                float4 e[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) e[u] = __ldg(col + (uint64_t)sm_idx[c + u] * h4);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float m = sm_mv[c + u];
                    acc.x = __fadd_rn(acc.x, __fmul_rn(e[u].x, m)); acc.y = __fadd_rn(acc.y, __fmul_rn(e[u].y, m));
                    acc.z = __fadd_rn(acc.z, __fmul_rn(e[u].z, m)); acc.w = __fadd_rn(acc.w, __fmul_rn(e[u].w, m));
                }
            }
            for (; c < cnt; ++c) {
                const float4 e = __ldg(col + (uint64_t)sm_idx[c] * h4);
                const float m = sm_mv[c];
                acc.x = __fadd_rn(acc.x, __fmul_rn(e.x, m)); acc.y = __fadd_rn(acc.y, __fmul_rn(e.y, m));
                acc.z = __fadd_rn(acc.z, __fmul_rn(e.z, m)); acc.w = __fadd_rn(acc.w, __fmul_rn(e.w, m));
            }
            if (mask_sum > 0.0f) { acc.x = acc.x / mask_sum; acc.y = acc.y / mask_sum; acc.z = acc.z / mask_sum; acc.w = acc.w / mask_sum; }
            reinterpret_cast<float4 *>(sm_pool)[j] = acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float ss = 0.0f;
#pragma unroll 8
            for (uint32_t j = 0; j < hidden; ++j) ss = __fadd_rn(ss, __fmul_rn(sm_pool[j], sm_pool[j]));
            sm_norm = sqrtf(ss);
        }
        __syncthreads();
        const float norm = sm_norm;
        float *o = out + t * out_ld;
        for (uint32_t j = threadIdx.x; j < out_ld; j += blockDim.x) {
            float v = 0.0f;
            if (j < hidden) {
                v = sm_pool[j];
                if (norm > 0.0f) v = v / norm;
            }
            o[j] = v;
        }
        __syncthreads();
    }
}

}  // namespace sema

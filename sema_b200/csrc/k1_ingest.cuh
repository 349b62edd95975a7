// k1_ingest.cuh — kernel K1: L2-normalise + lay out rows in the HBM matrix.
//
// Restates the normalise tail of mean_pool (reference:
// src/semantic/embeddings.rs:83-88: norm = sqrt(sum x^2) in f32; x /= norm iff
// norm > 0; a zero row stays zero) and the nullable FixedSizeList<Float32,dim>
// column of src/storage/lance_indexer.rs:41-45, 66-76: a null row (valid == 0), or
// a row holding a non-finite value, is stored as quiet NaNs so that K2's
// `score == score` test drops it without reading a validity bitmap.
// One warp per row; IEEE sqrt and divide (no fast-math), so the only difference
// from the reference's sequential sum is the summation order (<= a few ulp).
#pragma once
#include <float.h>
#include "common.cuh"

namespace sema {

constexpr int INGEST_THREADS = 256;

// value(seed,row,col) of SURVEY.md §8(d); bit-identical to oracle/cpu_scan.c synth_value.
__device__ __forceinline__ float synth_value(uint64_t seed, uint64_t row, uint64_t col)
{
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + row * 0xBF58476D1CE4E5B9ull +
                 col * 0x94D049BB133111EBull + 1ull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    const int s = (int)(x & 0xffff) + (int)((x >> 16) & 0xffff) + (int)((x >> 32) & 0xffff) +
                  (int)(x >> 48);
    return (float)(s - 131070);
}

// Fill n rows (stride ld, first d columns) of dst with the synthetic corpus.
__global__ void __launch_bounds__(INGEST_THREADS)
synth_kernel(float *dst, uint32_t ld, uint32_t d, uint64_t seed, uint64_t row0, uint64_t n)
{
    const uint64_t total = n * ld;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / ld;
        const uint32_t c = (uint32_t)(i - r * ld);
        dst[i] = c < d ? synth_value(seed, row0 + r, c) : 0.0f;
    }
}

// src: n rows, stride src_ld floats (may alias dst).  dst: stride ld floats, columns
// [d, ld) are zero padding.  valid_in nullable.  valid_out: one byte per row.
template <bool VEC4>
__global__ void __launch_bounds__(INGEST_THREADS)
ingest_kernel(const float *src, uint64_t src_ld, float *dst, uint32_t ld, uint32_t d, uint64_t n,
              const uint8_t *valid_in, uint8_t *valid_out, int normalize, float *max_norm2)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const float qnan = __uint_as_float(0x7fc00000u);

    for (uint64_t r = warp0; r < n; r += nwarps) {
        const float *s = src + r * src_ld;
        float *o = dst + r * (uint64_t)ld;
        float ss = 0.0f;
        if (VEC4) {
            for (uint32_t j = lane; j < d / 4; j += 32) {
                const float4 v = reinterpret_cast<const float4 *>(s)[j];
                ss = __fadd_rn(ss, __fmul_rn(v.x, v.x)); ss = __fadd_rn(ss, __fmul_rn(v.y, v.y));
                ss = __fadd_rn(ss, __fmul_rn(v.z, v.z)); ss = __fadd_rn(ss, __fmul_rn(v.w, v.w));
            }
        } else {
            for (uint32_t j = lane; j < d; j += 32) {
                const float v = s[j];
                ss = __fadd_rn(ss, __fmul_rn(v, v));  // x*x then add, as the reference (no FMA)
            }
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) ss += __shfl_xor_sync(FULL, ss, m);

        bool ok = ss <= FLT_MAX;  // false for NaN / inf anywhere in the row
        if (valid_in && valid_in[r] == 0) ok = false;  // may alias valid_out[r]: read by all lanes first
        const float norm = sqrtf(ss);
        const bool scale = normalize && norm > 0.0f;

        if (VEC4) {
            for (uint32_t j = lane; j < ld / 4; j += 32) {
                float4 v = reinterpret_cast<const float4 *>(s)[j];
                if (!ok) v = make_float4(qnan, qnan, qnan, qnan);
                else if (scale) { v.x = v.x / norm; v.y = v.y / norm; v.z = v.z / norm; v.w = v.w / norm; }
                reinterpret_cast<float4 *>(o)[j] = v;
            }
        } else {
            // dst may alias src with a different stride only when ld == src_ld; all
            // reads of this row happened above or happen before the write of the same j
            for (uint32_t j = lane; j < ld; j += 32) {
                float v = j < d ? s[j] : 0.0f;
                if (j < d) {
                    if (!ok) v = qnan;
                    else if (scale) v = v / norm;
                }
                o[j] = v;
            }
        }
        __syncwarp();
        if (lane == 0) {
            valid_out[r] = ok ? 1 : 0;
            // non-negative floats order like unsigned ints; K3 uses max_norm2[0] = max |x|^2 to bound its
            // error and max_norm2[1] = min |x|^2 (L2 metric: how far the dot product is from the distance)
            if (ok) {
                atomicMax(reinterpret_cast<unsigned int *>(max_norm2), __float_as_uint(scale ? 1.0000005f : ss));
                atomicMin(reinterpret_cast<unsigned int *>(max_norm2) + 1, __float_as_uint(scale ? 0.9999995f : ss));
            }
        }
    }
}

// tombstone: overwrite the listed rows with NaN (reference: table.delete(predicate),
// src/storage/lance_indexer.rs:234-250).
__global__ void __launch_bounds__(INGEST_THREADS)
tombstone_kernel(float *X, uint32_t ld, const uint64_t *rows, uint64_t n, uint64_t n_rows,
                 uint8_t *valid)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const float qnan = __uint_as_float(0x7fc00000u);
    for (uint64_t i = warp0; i < n; i += nwarps) {
        const uint64_t r = rows[i];
        if (r >= n_rows) continue;
        for (uint32_t j = lane; j < ld; j += 32) X[r * ld + j] = qnan;
        if (lane == 0) valid[r] = 0;
    }
}

// compaction: dst[i] = X[src[i]] for i in [0, m) — one warp per row, float4 granularity
__global__ void __launch_bounds__(INGEST_THREADS)
gather_rows_kernel(const float *X, uint32_t ld, const uint32_t *src, uint64_t m, float *dst)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = warp0; i < m; i += nwarps) {
        const float4 *s = reinterpret_cast<const float4 *>(X + (uint64_t)src[i] * ld);
        float4 *d = reinterpret_cast<float4 *>(dst + i * ld);
        for (uint32_t j = lane; j < ld / 4; j += 32) d[j] = s[j];
    }
}

}  // namespace sema

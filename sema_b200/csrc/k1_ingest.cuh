// k1_ingest.cuh — kernel K1: L2-normalise + lay out rows in the HBM matrix.
//
// Restates the normalise tail of mean_pool (reference:
// src/semantic/embeddings.rs:83-88: norm = sqrt(sum x^2) in f32; x /= norm iff
// norm > 0; a zero row stays zero) and the nullable FixedSizeList<Float32,dim>
// column of src/storage/lance_indexer.rs:41-45, 66-76: a null row (valid == 0), or
// a row holding a non-finite value, is stored as quiet NaNs so that K2's
// `score == score` test drops it without reading a validity bitmap.
// IEEE multiply, add, sqrt and divide in the reference's own order (no FMA, no
// fast-math, no re-association): the stored rows are bit-identical to the reference's.
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

// src: n rows, stride src_ld floats (may alias dst when the strides agree).  dst: stride ld floats,
// columns [d, ld) are zero padding.  valid_in nullable (may alias valid_out).  valid_out: one byte per row.
//
// The squared norm is summed in the REFERENCE'S ORDER — `pooled.iter().map(|x| x * x).sum::<f32>()`
// (src/semantic/embeddings.rs:84) is a sequential f32 fold over j = 0, 1, ..., d-1, multiply then add —
// so the stored rows are bit-identical to the reference's (and to oracle/cpu_scan.c:sema_oracle_normalize),
// not merely within a few ulp.  A sequential sum cannot be split over lanes, so the work is transposed:
// a warp owns 32 consecutive rows; it stages a [32 rows] x [CH columns] tile of them in shared memory
// with coalesced loads (every warp-level load covers one row's contiguous 128 / 512 bytes), then lane r
// folds row r's CH values in order (conflict-free: row stride CH + VW words).  The second pass re-reads
// the rows (L2 hits: the warp has just streamed them), scales and writes them.  HBM traffic stays one
// read + one write per row; ingest from the host is PCIe-bound long before this kernel matters.
// VW = 4: float4 accesses (d % 4 == 0, src_ld % 4 == 0, 16-byte aligned src); VW = 1: any shape.
constexpr int INGEST_SEQ_WARPS = 2;          // 2 x 16.9 KB tiles: static shared memory, up to 6 blocks per SM
template <int VW> struct IngestTile {
    static constexpr int CH = 32 * VW;          // columns per staged chunk
    static constexpr int STRIDE = CH + VW;      // words per tile row: LDS.128 / LDS.32 by row are conflict-free
    static constexpr int WORDS = 32 * STRIDE;
};

template <int VW>
__global__ void __launch_bounds__(INGEST_SEQ_WARPS * 32)
ingest_kernel(const float *src, uint64_t src_ld, float *dst, uint32_t ld, uint32_t d, uint64_t n,
              const uint8_t *valid_in, uint8_t *valid_out, int normalize, float *max_norm2)
{
    using T = IngestTile<VW>;
    __shared__ __align__(16) float tiles[INGEST_SEQ_WARPS][T::WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *tile = tiles[warp];
    const uint64_t gw = (uint64_t)blockIdx.x * INGEST_SEQ_WARPS + warp;
    const uint64_t nw = (uint64_t)gridDim.x * INGEST_SEQ_WARPS;
    const uint64_t groups = (n + 31) / 32;
    const float qnan = __uint_as_float(0x7fc00000u);

    for (uint64_t g = gw; g < groups; g += nw) {
        const uint64_t r0 = g * 32;
        const uint32_t rows = (uint32_t)((n - r0) < 32 ? (n - r0) : 32);
        float ss = 0.0f;                                   // lane r: running sum of row r0 + r
        for (uint32_t c0 = 0; c0 < d; c0 += T::CH) {
            // ---- stage rows [r0, r0+rows) x columns [c0, c0+CH): 16 row loads in flight per lane
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if (VW == 4) {
                    float4 v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t rr = half * 16 + i, c = c0 + lane * 4;
                        v[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                        if (rr < rows && c < d) v[i] = *reinterpret_cast<const float4 *>(src + (r0 + rr) * src_ld + c);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        *reinterpret_cast<float4 *>(tile + (half * 16 + i) * T::STRIDE + lane * 4) = v[i];
                } else {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t rr = half * 16 + i, c = c0 + lane;
                        v[i] = (rr < rows && c < d) ? src[(r0 + rr) * src_ld + c] : 0.0f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) tile[(half * 16 + i) * T::STRIDE + lane] = v[i];
                }
            }
            __syncwarp();
            // ---- lane r folds its row's values in column order: ss = ss + x*x (no FMA), as the reference
            const uint32_t jn = (d - c0) < (uint32_t)T::CH ? (d - c0) : (uint32_t)T::CH;
            const float *mine = tile + lane * T::STRIDE;
            if (VW == 4) {
                for (uint32_t j = 0; j < jn; j += 4) {     // d % 4 == 0
                    const float4 x = *reinterpret_cast<const float4 *>(mine + j);
                    ss = __fadd_rn(ss, __fmul_rn(x.x, x.x)); ss = __fadd_rn(ss, __fmul_rn(x.y, x.y));
                    ss = __fadd_rn(ss, __fmul_rn(x.z, x.z)); ss = __fadd_rn(ss, __fmul_rn(x.w, x.w));
                }
            } else {
                for (uint32_t j = 0; j < jn; ++j) ss = __fadd_rn(ss, __fmul_rn(mine[j], mine[j]));
            }
            __syncwarp();
        }
        // ---- per-row outcome (lane r <-> row r0 + r)
        bool ok = ss <= FLT_MAX;                           // false for NaN / inf anywhere in the row
        if ((uint32_t)lane < rows && valid_in && valid_in[r0 + lane] == 0) ok = false;
        const float norm = sqrtf(ss);
        const bool scale = normalize && norm > 0.0f;
        // ---- second pass: scale and write, one row at a time, coalesced
        for (uint32_t rr = 0; rr < rows; ++rr) {
            const float nrm = __shfl_sync(FULL, norm, rr);
            const bool okr = __shfl_sync(FULL, (int)ok, rr) != 0, scr = __shfl_sync(FULL, (int)scale, rr) != 0;
            const float *s = src + (r0 + rr) * src_ld;
            float *o = dst + (r0 + rr) * (uint64_t)ld;
            if (VW == 4) {
                for (uint32_t j = lane; j < ld / 4; j += 32) {
                    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (j * 4 < d) v = reinterpret_cast<const float4 *>(s)[j];
                    if (j * 4 < d) {
                        if (!okr) v = make_float4(qnan, qnan, qnan, qnan);
                        else if (scr) { v.x = v.x / nrm; v.y = v.y / nrm; v.z = v.z / nrm; v.w = v.w / nrm; }
                    }
                    reinterpret_cast<float4 *>(o)[j] = v;
                }
            } else {
                for (uint32_t j = lane; j < ld; j += 32) {
                    float v = j < d ? s[j] : 0.0f;
                    if (j < d) {
                        if (!okr) v = qnan;
                        else if (scr) v = v / nrm;
                    }
                    o[j] = v;
                }
            }
        }
        if ((uint32_t)lane < rows) {
            valid_out[r0 + lane] = ok ? 1 : 0;
            // non-negative floats order like unsigned ints; K3 uses max_norm2[0] = max |x|^2 to bound its
            // error and max_norm2[1] = min |x|^2 (L2 metric: how far the dot product is from the distance)
            if (ok) {
                atomicMax(reinterpret_cast<unsigned int *>(max_norm2), __float_as_uint(scale ? 1.0000005f : ss));
                atomicMin(reinterpret_cast<unsigned int *>(max_norm2) + 1, __float_as_uint(scale ? 0.9999995f : ss));
            }
        }
        __syncwarp();
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
    const uint32_t ld4 = ld / 4;
    if (ld4 <= 6 * 32) {
        // rows of up to 768 floats (the path's dims 384 / 768): four rows per warp step, every load issued before the
        // first store, so a warp keeps up to 24 x 512 B in flight instead of 3 - 6
        constexpr int U = 4, V = 6;
        for (uint64_t i0 = warp0 * U; i0 < m; i0 += nwarps * U) {
            float4 v[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t i = i0 + u < m ? i0 + u : m - 1;
                const float4 *s = reinterpret_cast<const float4 *>(X + (uint64_t)src[i] * ld);
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if (lane + j * 32 < ld4) v[u][j] = s[lane + j * 32];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (i0 + u >= m) break;
                float4 *d = reinterpret_cast<float4 *>(dst + (i0 + u) * ld);
#pragma unroll
                for (int j = 0; j < V; ++j)
                    if (lane + j * 32 < ld4) d[lane + j * 32] = v[u][j];
            }
        }
        return;
    }
    for (uint64_t i = warp0; i < m; i += nwarps) {
        const float4 *s = reinterpret_cast<const float4 *>(X + (uint64_t)src[i] * ld);
        float4 *d = reinterpret_cast<float4 *>(dst + i * ld);
        for (uint32_t j = lane; j < ld4; j += 32) d[j] = s[j];
    }
}

// ---- compaction plan on the device: flags -> prefix sums -> gather list, old->new map, compacted validity bytes ----
// Three kernels over tiles of COMPACT_TILE rows (256 threads x 16 flags): per-tile live counts, one block scanning
// the tile counts, and an emit pass that re-derives each thread's offset inside its tile.  Replaces a host loop over
// every row (the reference deletes by predicate and lets LanceDB rewrite fragments: src/storage/lance_indexer.rs:234-250).
constexpr int COMPACT_TILE = 4096;

__device__ __forceinline__ uint32_t compact_load16(const uint8_t *flags, uint64_t base, uint64_t n, uint8_t (&f)[16])
{
    uint32_t c = 0;
    if (base + 16 <= n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(flags + base);    // base is a multiple of 16, flags 256-byte aligned
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i) { f[i] = (uint8_t)(w[i >> 2] >> ((i & 3) * 8)); c += f[i] != 0; }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) { f[i] = (base + i < n) ? flags[base + i] : 0; c += f[i] != 0; }
    }
    return c;
}

// inclusive scan of one value per thread over a block of 256 threads (8 warps); returns the inclusive sum, *total = block sum
__device__ __forceinline__ uint32_t compact_block_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    uint32_t before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < INGEST_THREADS / 32; ++w) {
        const uint32_t ws = warp_sums[w];
        if (w < warp) before += ws;
        all += ws;
    }
    __syncthreads();
    *total = all;
    return x + before;
}

// tile_count[b] = flags set in tile b; *first_dropped = lowest row whose flag is clear (n if none; initialised by the host)
__global__ void __launch_bounds__(INGEST_THREADS)
compact_count_kernel(const uint8_t *flags, uint64_t n, uint32_t *tile_count, unsigned long long *first_dropped)
{
    __shared__ uint32_t warp_sums[INGEST_THREADS / 32];
    __shared__ unsigned long long first;
    if (threadIdx.x == 0) first = ~0ull;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * COMPACT_TILE + (uint64_t)threadIdx.x * 16;
    uint8_t f[16];
    const uint32_t c = compact_load16(flags, base, n, f);
    if (c != 16) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (f[i] == 0 && base + i < n) { atomicMin(&first, (unsigned long long)(base + i)); break; }
    }
    uint32_t total;
    compact_block_scan(c, warp_sums, &total);
    if (threadIdx.x == 0) {
        tile_count[blockIdx.x] = total;
        if (first != ~0ull) atomicMin(first_dropped, first);
    }
}

// exclusive scan of the tile counts by ONE block (tiles = n / 4096: 2442 for 10 M rows); *live = number of flags set
__global__ void __launch_bounds__(INGEST_THREADS)
compact_scan_tiles_kernel(const uint32_t *tile_count, uint32_t tiles, uint32_t *tile_off, unsigned long long *live)
{
    __shared__ uint32_t warp_sums[INGEST_THREADS / 32];
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < tiles; b0 += INGEST_THREADS) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < tiles ? tile_count[i] : 0u;
        uint32_t total;
        const uint32_t incl = compact_block_scan(v, warp_sums, &total);
        if (i < tiles) tile_off[i] = carry + incl - v;
        carry += total;
    }
    if (threadIdx.x == 0) *live = carry;
}

// src[new] = old (ascending), new_valid[new] = valid[old], map[old] = new or ~0 (map nullable)
__global__ void __launch_bounds__(INGEST_THREADS)
compact_emit_kernel(const uint8_t *flags, const uint8_t *valid, uint64_t n, const uint32_t *tile_off, uint32_t *src,
                    uint8_t *new_valid, unsigned long long *map)
{
    __shared__ uint32_t warp_sums[INGEST_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * COMPACT_TILE + (uint64_t)threadIdx.x * 16;
    uint8_t f[16];
    const uint32_t c = compact_load16(flags, base, n, f);
    uint32_t total;
    uint32_t o = tile_off[blockIdx.x] + compact_block_scan(c, warp_sums, &total) - c;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint64_t r = base + i;
        if (r >= n) break;
        if (f[i]) {
            src[o] = (uint32_t)r;
            new_valid[o] = valid[r];
            if (map) map[r] = o;
            ++o;
        } else if (map) {
            map[r] = ~0ull;
        }
    }
}

}  // namespace sema

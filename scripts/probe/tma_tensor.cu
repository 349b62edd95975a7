// Does the K3 pair kernel's fetch pattern cap the per-SM ingest?  One CTA per SM streams the "hi" 8 KB of every 16 KB
// group of a large buffer (HBM) through a 12 x 16 KB ring, (a) as 1-D bulk copies (2 x 8 KB per stage), (b) as one 3-D
// tensor-map copy per stage (box {256 x u64, 4, 2} over {256, 4, groups}, stride 16 KB) — the pair kernel's form —
// and (c) as one contiguous 16 KB 1-D copy (no gaps), with `share` CTAs streaming the same slice.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t *b, uint32_t par)
{
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(b)), "r"(par) : "memory");
}
constexpr uint32_t STAGE = 16384, NST = 12;
__global__ void __launch_bounds__(64, 1) k(const unsigned char *src, size_t groups_per_cta, uint64_t iters, int mode, uint32_t share,
                                           const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)NST * STAGE), *empty = full + 32;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < NST; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t slices = gridDim.x / share;
    const size_t g0 = (size_t)(blockIdx.x % slices) * groups_per_cta;      // first 16 KB group of this CTA's slice
    if (threadIdx.x == 0) {
        uint32_t st = 0, ph = 0; size_t g = 0;
        for (uint64_t i = 0; i < iters; ++i) {
            mwait(&empty[st], ph ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(STAGE) : "memory");
            unsigned char *dst = sm + (size_t)st * STAGE;
            if (mode == 0) {          // two 8 KB hi blocks, 16 KB apart
                for (int j = 0; j < 2; ++j)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(s32(dst + j * 8192)), "l"(src + (g0 + g + j) * 16384), "r"(8192u), "r"(s32(&full[st])) : "memory");
            } else if (mode == 1) {   // the same bytes as one 3-D tensor copy
                asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                             ::"r"(s32(dst)), "l"(&tmap), "r"(0), "r"(0), "r"((uint32_t)(g0 + g)), "r"(s32(&full[st])) : "memory");
            } else {                  // contiguous 16 KB
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(dst)), "l"(src + (g0 + g) * 16384), "r"(STAGE), "r"(s32(&full[st])) : "memory");
            }
            g += (mode == 2) ? 1 : 2; if (g + 2 > groups_per_cta) g = 0;
            if (++st == NST) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        uint32_t st = 0, ph = 0;
        for (uint64_t i = 0; i < iters; ++i) {
            mwait(&full[st], ph);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
            if (++st == NST) { st = 0; ph ^= 1; }
        }
    }
}
int main()
{
    unsigned char *buf; const size_t big = (size_t)12 << 30;
    cudaMalloc(&buf, big); cudaMemset(buf, 1, big);
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
    CUtensorMap tm;
    const cuuint64_t gdim[3] = {256, 4, big / 16384}; const cuuint64_t gstr[2] = {2048, 16384};
    const cuuint32_t box[3] = {256, 4, 2}, es[3] = {1, 1, 1};
    CUresult r = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"1-D 2 x 8 KB (hi blocks, 16 KB apart)", "3-D tensor copy of the same bytes", "1-D 16 KB contiguous"};
    for (uint32_t share : {1u, 4u})
        for (int mode = 0; mode < 3; ++mode) {
            const int g = 144;
            const size_t groups_per_cta = big / 16384 / (g / share);
            const uint64_t iters = ((size_t)6 << 30) / 148 / STAGE;
            k<<<g, 64, NST * STAGE + 1024>>>(buf, groups_per_cta, 64, mode, share, tm);
            cudaEventRecord(e0);
            k<<<g, 64, NST * STAGE + 1024>>>(buf, groups_per_cta, iters, mode, share, tm);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)g * iters * STAGE;
            printf("share %u %-42s: %7.1f GB/s into SMs, %6.1f GB/s per SM (%s)\n", share, names[mode], bytes / ms / 1e6, bytes / ms / 1e6 / g,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}

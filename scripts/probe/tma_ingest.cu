// Per-SM ingest rate of 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) on B200: one CTA per SM streams
// `stage`-byte pieces through a ring of `depth` bytes in shared memory from a region of `region` bytes (small region
// = L2-resident, large = HBM).  A consumer warp only waits for each stage and hands it back.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t *b, uint32_t par)
{
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(b)), "r"(par) : "memory");
}
__global__ void __launch_bounds__(64, 1) k(const unsigned char *src, size_t region, uint32_t stage, uint32_t nst, uint32_t piece, uint64_t iters, uint32_t share = 1, uint32_t lag = 0)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)nst * stage), *empty = full + 32;
    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nst; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // `share` CTAs (strided across the grid, so they sit on different SMs / GPCs) stream the SAME slice; CTA j of a
    // sharing group starts `lag` * j stages late (it spins first), so its requests find the lines already in L2
    const uint32_t slices = gridDim.x / share;
    const size_t per = region / slices / stage * stage;        // this group's slice of the region
    const unsigned char *base = src + (size_t)(blockIdx.x % slices) * per;
    if (lag && threadIdx.x == 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)lag * (blockIdx.x / slices) * 800) { }     // ~800 cycles per 48 KB stage at ~120 GB/s
    }
    if (threadIdx.x == 0) {
        uint32_t st = 0, ph = 0; size_t off = 0;
        for (uint64_t i = 0; i < iters; ++i) {
            mwait(&empty[st], ph ^ 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(stage) : "memory");
            for (uint32_t o = 0; o < stage; o += piece)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(sm + (size_t)st * stage + o)), "l"(base + off + o), "r"(piece), "r"(s32(&full[st])) : "memory");
            off += stage; if (off + stage > per) off = 0;
            if (++st == nst) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        uint32_t st = 0, ph = 0;
        for (uint64_t i = 0; i < iters; ++i) {
            mwait(&full[st], ph);
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
            if (++st == nst) { st = 0; ph ^= 1; }
        }
    }
}
int main()
{
    unsigned char *buf; size_t big = (size_t)12 << 30;
    cudaMalloc(&buf, big); cudaMemset(buf, 1, big);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t regions[] = {(size_t)32 << 20, (size_t)96 << 20, (size_t)12 << 30};
    const uint32_t stages[] = {8192, 16384, 49152};
    for (int g : {148})
    for (size_t region : regions)
        for (uint32_t stage : stages)
            for (uint32_t piece : {8192u, stage}) {
                if (piece == 8192u && stage == 8192u) continue;
                const uint32_t nst = 196608 / stage;
                const uint64_t iters = ((size_t)6 << 30) / 148 / stage;          // ~6 GB in total at 148 CTAs
                k<<<g, 64, (size_t)nst * stage + 1024>>>(buf, region, stage, nst, piece, 64);   // warm
                cudaEventRecord(e0);
                k<<<g, 64, (size_t)nst * stage + 1024>>>(buf, region, stage, nst, piece, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                const double bytes = (double)g * iters * stage;
                printf("ctas %3d region %6zu MB stage %5u piece %5u ring %u: %7.1f GB/s total, %6.1f GB/s per SM (%s)\n", g, region >> 20, stage, piece,
                       nst, bytes / ms / 1e6, bytes / ms / 1e6 / g, cudaGetErrorString(cudaGetLastError()));
            }
    // sharing: groups of 4 CTAs stream the same 12 GB-region slices (the K3 pattern: 4 query groups, one corpus range)
    for (uint32_t share : {1u, 2u, 4u})
        for (uint32_t lag : {0u, 4u, 16u}) {
            if (share == 1 && lag) continue;
            const uint32_t stage = 49152, nst = 4; const int g = 144;
            const uint64_t iters = ((size_t)6 << 30) / 148 / stage;
            k<<<g, 64, (size_t)nst * stage + 1024>>>(buf, (size_t)12 << 30, stage, nst, stage, 64, share, lag);
            cudaEventRecord(e0);
            k<<<g, 64, (size_t)nst * stage + 1024>>>(buf, (size_t)12 << 30, stage, nst, stage, iters, share, lag);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)g * iters * stage;
            printf("share %u lag %2u stages: %7.1f GB/s into SMs, %6.1f GB/s per SM (%s)\n", share, lag, bytes / ms / 1e6, bytes / ms / 1e6 / g,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}

// ptx.cuh — thin inline-PTX wrappers shared by the TMA-fed kernels (K2's bulk-copy variant, K3):
// mbarrier init / expect_tx / arrive / wait and the 1-D bulk async copy (cp.async.bulk, SASS UBLKCP).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sema {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    // one asm block with scoped labels: no C-level loop, so the surrounding code stays
    // warp-uniform for the compiler (uniform registers, no re-convergence scaffolding)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SEMA_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra SEMA_DONE;\n\t"
        "bra SEMA_WAIT;\n\t"
        "SEMA_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep in hardware, do not spin
        : "memory");
}
// Wait on a barrier whose completing arrival may come from ANOTHER CTA of the cluster (remote mbarrier.arrive,
// multicast tcgen05.commit).  No suspend-time hint: with the long hint above a waiter that is already asleep when the
// remote arrival lands is not woken by it (measured: the CTA-pair K3 kernel ran 3x slower, ~0.7 us per stage, with
// every remote-completed wait on the critical path), so this form re-polls at the hardware's own short limit.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SEMA_WAITC:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SEMA_DONEC;\n\t"
        "bra SEMA_WAITC;\n\t"
        "SEMA_DONEC:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// Pure polling wait (mbarrier.test_wait never suspends the thread): for barriers completed from outside the CTA when
// every microsecond of wake-up latency is on the critical path.  backoff_ns > 0 sleeps between polls.
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity, uint32_t backoff_ns = 0)
{
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
        if (ok) break;
        if (backoff_ns) asm volatile("nanosleep.u32 %0;" ::"r"(backoff_ns));
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t}"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace ptx
}  // namespace sema

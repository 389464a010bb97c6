// tma_utils.cuh -- mbarrier + cp.async.bulk (TMA, UBLKCP in SASS) wrappers for the statically compiled kernels.
// (The NVRTC-compiled filter/project skeleton carries its own copy: csrc/jit_tma_skeleton.inc.)
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t nqe_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void nqe_mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nqe_smem_u32(bar)), "r"(count));
}
// make freshly initialised barriers visible to the async proxy (the bulk copies complete on them)
__device__ __forceinline__ void nqe_mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void nqe_mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nqe_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void nqe_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(nqe_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void nqe_mbar_wait(unsigned long long *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "NQE_WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra NQE_WAIT_DONE;\n\t"
        "bra NQE_WAIT_LOOP;\n\t"
        "NQE_WAIT_DONE:\n\t}" ::"r"(nqe_smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy; bytes and both addresses are multiples of 16; completion is signalled on `bar`
__device__ __forceinline__ void nqe_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar,
                                             unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     nqe_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(nqe_smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ unsigned long long nqe_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

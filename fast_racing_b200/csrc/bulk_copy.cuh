// 1-D bulk (TMA-unit) copies global -> shared memory with mbarrier completion, and bulk L2 prefetch (sm_90+ / sm_100a).
// Shared by the evaluation kernels (staging of a trajectory's coefficients / factors) and the solver's direction update
// (streaming the L-BFGS history through a shared-memory ring).
#pragma once
#include <cuda_runtime.h>

namespace frx {
// ---- 1-D bulk copies global -> shared memory through the TMA unit, completion on an mbarrier (sm_90+/sm_100a) ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy before the first copy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
// bytes: a multiple of 16; both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst), a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes),
                 "r"(a)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}

// One bulk (TMA-unit) prefetch of a whole byte range into L2: the 16-byte aligned interior of [p, p + bytes).
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, size_t bytes) {
    const unsigned long long a = ((unsigned long long)p + 15ull) & ~15ull, e = ((unsigned long long)p + bytes) & ~15ull;
    if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((unsigned)(e - a)) : "memory");
}

}  // namespace frx

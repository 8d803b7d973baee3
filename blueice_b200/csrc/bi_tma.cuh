// blueice_b200 -- mbarrier and 1-D TMA bulk-copy helpers shared by the kernels that stage tiles in shared memory
// (K2: event tiles of the anchor tensor; K4: anchor rows of a bin tile).
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t bi_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bi_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bi_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bi_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bi_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool bi_mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bi_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a byte-count bug would otherwise hang the GPU; after ~2 s of polling the kernel traps
// (a reported CUDA error) instead of spinning forever.
__device__ __forceinline__ void bi_mbar_wait(uint64_t* bar, unsigned parity) {
    if (bi_mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!bi_mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bi_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     bi_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(bi_smem_u32(bar))
                 : "memory");
}

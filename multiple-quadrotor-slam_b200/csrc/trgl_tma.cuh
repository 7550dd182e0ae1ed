// Bulk-async (TMA engine) input pipeline for the streaming solvers.
//
// The per-point solvers are HBM-latency bound when every thread issues its own 16-byte loads: the bytes in flight
// are capped by registers x occupancy (ncu: 84 regs -> 16 warps/SM -> 55 % of DRAM peak).  Here one elected thread
// per CTA streams whole tiles of u1/u2 into a ring of shared-memory stages with `cp.async.bulk` (SASS UBLKCP),
// completion signalled on an mbarrier per stage, so STAGES-1 tiles per CTA are always in flight and cost no
// registers.  CTAs are persistent (grid = SMs x resident CTAs) and walk the tiles round-robin.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace trgl {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

}  // namespace trgl

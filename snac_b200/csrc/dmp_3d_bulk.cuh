// dmp_3d_bulk.cuh -- mbarrier / bulk-async-copy (TMA) helpers and the small 3D definitions shared by the 3D kernels.
#pragma once
#include "dmp_common.cuh"

namespace d3 {

constexpr uint32_t COLVALID = 0x7FFFF8u;     // padded columns 3..22 are inside the plan area
constexpr unsigned FULL = 0xFFFFFFFFu;

struct EnvT {
    int pr, pc, plan_idx, cb, cs;
    float ret;
    int cross;      // running sum(min(height, plan)): +1 per brick laid at or below the plan height
};

// direction table: 0 left (c-1), 1 right (c+1), 2 "up" (r+1), 3 "down" (r-1)   (check_sur :88-102)
__device__ __forceinline__ int dir_dr(int d) { return d == 2 ? 1 : (d == 3 ? -1 : 0); }
__device__ __forceinline__ int dir_dc(int d) { return d == 0 ? -1 : (d == 1 ? 1 : 0); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    }
}

}  // namespace d3

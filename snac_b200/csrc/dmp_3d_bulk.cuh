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

// ---- an env turns tall (a height has just reached 15: every nibble, the new one included, is still exact): the warp
// writes its wide map out, 16 cells per lane.  Rare per env, but among 262 144 random-policy envs it happens in every few
// launches, and a serial loop over 400 cells held the whole grid up for ~70 us: hence cooperative, one round trip.
__device__ __forceinline__ void warp_widen_words(uint16_t* wide, uint32_t w0, uint32_t w1, int lane) {
    const uint32_t w[2] = {w0, w1};                       // this lane's 16 nibbles (cells 16 lane .. 16 lane + 15)
    uint32_t o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint32_t by = (w[q >> 2] >> (8 * (q & 3))) & 0xFFu;
        o[q] = (by & 0xFu) | ((by >> 4) << 16);
    }
    uint4* dst = reinterpret_cast<uint4*>(wide) + 2 * lane;
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// ---- tall envs (dmp_common.cuh): a nibble that reads 15 stands for "15 or more".  The window of a tall env is formatted
// from its nibbles like any other; only when a saturated nibble lies inside the window are those cells re-read from the
// env's exact wide map.  u0 / u1: the seven rows of biased bytes (nibble + 1, 0 = frame): a saturated cell reads 16.
__device__ __forceinline__ bool window_saturated(const uint32_t (&u0)[7], const uint32_t (&u1)[7]) {
    uint32_t m = 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) m |= u0[k] | u1[k];
    return (m & 0x10101010u) != 0u;
}
// out of line, practically never: observation values of the saturated cells from the wide map
template <typename ObsT>
__device__ __noinline__ void fix_saturated_row(const uint16_t* ge, int pr, int pc, const uint32_t* u0, const uint32_t* u1, ObsT* row) {
    for (int k = 0; k < 7; ++k) {
        const uint64_t c = (uint64_t)u0[k] | ((uint64_t)u1[k] << 32);
        for (int j = 0; j < 7; ++j)
            if (((c >> (8 * j)) & 0xFFu) == 16u)
                row[k * 7 + j] = obs_from_int<ObsT>((int)__ldcg(ge + (pr - 6 + k) * 20 + (pc - 6 + j)));
    }
}
// the same for record rows: window bytes = min(height + 1, 255), patched in the packed row codes
static __device__ __noinline__ void fix_saturated_codes(const uint16_t* ge, int pr, int pc, uint64_t* c) {
    for (int k = 0; k < 7; ++k)
        for (int j = 0; j < 7; ++j)
            if (((c[k] >> (8 * j)) & 0xFFu) == 16u) {
                const uint64_t h = (uint64_t)min((int)__ldcg(ge + (pr - 6 + k) * 20 + (pc - 6 + j)) + 1, 255);
                c[k] = (c[k] & ~(0xFFull << (8 * j))) | (h << (8 * j));
            }
}


// ---- DMP_OBS_BITS (include/dmp.h): the seven window rows of biased bytes (height + 1 <= 16, 0 = frame) -> 49 four-bit
// codes min(height + 1, 15) + the trailer word: one 32 B record, two 128-bit streaming stores from registers (no tile).
// A code of 15 stands for "height >= 14": such a record carries the saturation flag.
__device__ __forceinline__ void bits32_store(void* obs, int64_t idx, const uint32_t (&u0)[7], const uint32_t (&u1)[7], int cb, int cs,
                                             float reward, bool done) {
    uint32_t r[7], any15 = 0u;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const uint32_t a = u0[k] - ((u0[k] >> 4) & 0x01010101u), b = u1[k] - ((u1[k] >> 4) & 0x01010101u);   // 16 -> 15
        r[k] = __byte_perm(a | (a >> 4), 0u, 0x4420) | (__byte_perm(b | (b >> 4), 0u, 0x4420) << 16);       // 7 nibbles, 28 bits
        uint32_t t = r[k] & (r[k] >> 1);
        t &= t >> 2;
        any15 |= t;
    }
    const bool sat = (any15 & 0x01111111u) != 0u;
    uint4* d = reinterpret_cast<uint4*>(obs) + 2 * idx;
    __stcs(d, make_uint4(r[0] | (r[1] << 28), (r[1] >> 4) | (r[2] << 24), (r[2] >> 8) | (r[3] << 20), (r[3] >> 12) | (r[4] << 16)));
    __stcs(d + 1, make_uint4((r[4] >> 16) | (r[5] << 12), (r[5] >> 20) | (r[6] << 8), r[6] >> 24,
                             bits_trailer(cb, cs, reward, done, sat)));
}

}  // namespace d3

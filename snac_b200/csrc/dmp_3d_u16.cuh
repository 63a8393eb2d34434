// dmp_3d_u16.cuh -- pieces shared by the 3D kernels that work on u16 height maps staged in shared memory by
// bulk async copies (dmp_3d_tile.cu: whole maps; dmp_3d_step.cu: only the rows a single step looks at).
#pragma once
#include "dmp_common.cuh"

namespace u16map {

constexpr uint32_t COLVALID = 0x7FFFF8u;     // padded columns 3..22 are inside the plan area

struct EnvT {
    int pr, pc, plan_idx, cb, cs;
    float ret;
    int cross;      // running sum(min(height, plan)): +1 per brick laid at or below the plan height
};

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

__device__ __forceinline__ int dir_dr(int d) { return d == 2 ? 1 : (d == 3 ? -1 : 0); }
__device__ __forceinline__ int dir_dc(int d) { return d == 0 ? -1 : (d == 1 ? 1 : 0); }

// environment_memory[r][c] (padded coordinates) of this lane's env; -1 on the frame
__device__ __forceinline__ int cell_s(const uint16_t* g, int r, int c) {
    const unsigned ir = (unsigned)(r - 3), ic = (unsigned)(c - 3);
    return (ir < 20u && ic < 20u) ? (int)g[ir * 20u + ic] : -1;
}

// two packed halfwords (biased: height+1, 0 = frame) -> two observation values
template <typename ObsT>
__device__ __forceinline__ void emit_pair(uint32_t u, ObsT& lo, ObsT& hi) {
    if constexpr (sizeof(ObsT) == 4) {
        lo = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7610u)) - 8388609.0f;
        hi = __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7632u)) - 8388609.0f;
    } else {
        lo = obs_from_int<ObsT>((int)(u & 0xFFFFu) - 1);
        hi = obs_from_int<ObsT>((int)(u >> 16) - 1);
    }
}

// stage (c): 7x7 window of this lane's env -> its row of the warp tile
template <typename ObsT>
__device__ __forceinline__ void observe_tile(const uint16_t* g, const EnvT& e, ObsT* row, bool normalise,
                                             int total_brick, int total_step) {
    const int ic0 = e.pc - 6;                           // interior column of window column 0 (may be negative)
    const int w0 = ic0 >> 1;                            // first word of the row to fetch (floor)
    const int sh16 = (ic0 & 1) * 16;
    const int sh = e.pc - 3;
    const uint32_t colvalid = (COLVALID >> sh) & 0x7Fu;
    // per-pair validity masks (0xFFFF per valid halfword)
    uint32_t m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t lo = (colvalid >> (2 * q)) & 1u, hi = (q < 3) ? ((colvalid >> (2 * q + 1)) & 1u) : 0u;
        m[q] = (lo ? 0x0000FFFFu : 0u) | (hi ? 0xFFFF0000u : 0u);
    }
    // bias (+1 per valid halfword), applied AFTER masking: a garbage 0xFFFF from the guard bytes must not
    // carry into its valid neighbour
    const uint32_t b0 = m[0] & 0x00010001u, b1 = m[1] & 0x00010001u, b2 = m[2] & 0x00010001u, b3 = m[3] & 0x00010001u;
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(g);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int ir = e.pr - 6 + k;                    // interior row of window row k
        const bool rowvalid = (unsigned)ir < 20u;
        ObsT* o = row + k * 7;
        // branch-free: rows outside the plan area read a clamped (valid) row and are masked to "frame"
        const int irc = min(max(ir, 0), 19);
        const uint32_t rm = rowvalid ? 0xFFFFFFFFu : 0u;
        const uint32_t* rw = gw + irc * 10 + w0;        // over-reads stay inside the 16 B guards
        const uint32_t x0 = rw[0], x1 = rw[1], x2 = rw[2], x3 = rw[3], x4 = rw[4];
        const uint32_t u0 = ((__funnelshift_r(x0, x1, sh16) & m[0]) + b0) & rm;
        const uint32_t u1 = ((__funnelshift_r(x1, x2, sh16) & m[1]) + b1) & rm;
        const uint32_t u2 = ((__funnelshift_r(x2, x3, sh16) & m[2]) + b2) & rm;
        const uint32_t u3 = ((__funnelshift_r(x3, x4, sh16) & m[3]) + b3) & rm;
        ObsT dummy;
        emit_pair<ObsT>(u0, o[0], o[1]);
        emit_pair<ObsT>(u1, o[2], o[3]);
        emit_pair<ObsT>(u2, o[4], o[5]);
        emit_pair<ObsT>(u3, o[6], dummy);
    }
    obs_counters<ObsT>(normalise, e.cb, e.cs, total_brick, total_step, row[49], row[50]);
}

// stage (c), first half: the 7x7 window of this lane's env as 7 x 4 words of biased halfwords (height + 1, 0 = frame;
// the upper half of word 3 is unused).  emit_pair() turns them into observation values.
__device__ __forceinline__ void window_regs(const uint16_t* g, const EnvT& e, uint32_t (&u)[7][4]) {
    const int ic0 = e.pc - 6;                           // interior column of window column 0 (may be negative)
    const int w0 = ic0 >> 1;                            // first word of the row to fetch (floor)
    const int sh16 = (ic0 & 1) * 16;
    const uint32_t colvalid = (COLVALID >> (e.pc - 3)) & 0x7Fu;
    uint32_t m[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t lo = (colvalid >> (2 * q)) & 1u, hi = (q < 3) ? ((colvalid >> (2 * q + 1)) & 1u) : 0u;
        m[q] = (lo ? 0x0000FFFFu : 0u) | (hi ? 0xFFFF0000u : 0u);
    }
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(g);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const int ir = e.pr - 6 + k;                    // interior row of window row k
        const uint32_t rm = ((unsigned)ir < 20u) ? 0xFFFFFFFFu : 0u;
        const uint32_t* rw = gw + min(max(ir, 0), 19) * 10 + w0;       // over-reads stay inside the 16 B guards
        const uint32_t x0 = rw[0], x1 = rw[1], x2 = rw[2], x3 = rw[3], x4 = rw[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t lo = q == 0 ? x0 : q == 1 ? x1 : q == 2 ? x2 : x3, hi = q == 0 ? x1 : q == 1 ? x2 : q == 2 ? x3 : x4;
            u[k][q] = ((__funnelshift_r(lo, hi, sh16) & m[q]) + (m[q] & 0x00010001u)) & rm;
        }
    }
}

}  // namespace u16map

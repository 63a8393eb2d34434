// dmp_common.cuh -- shared device helpers for libdmp.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "dmp.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdmp is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------------
// geometry (reference __init__ constants; see include/dmp.h for the citations)
// ---------------------------------------------------------------------------------------------
constexpr int D1_W = 30, D1_HW = 2, D1_OBS = 7, D1_ACT = 3, D1_LO = 2, D1_HI = 31;
constexpr int D2_W = 20, D2_HW = 3, D2_OBS = 51, D2_ACT = 5, D2_LO = 3, D2_HI = 22;
constexpr int D3_OBS = 51, D3_ACT = 8;
constexpr int PLAN2D_WORDS = 16;     // 13 used
constexpr int GRID2D_WORDS = 13;
constexpr int PLAN1D_BYTES = 32;     // 30 used
constexpr int CELLS3D = 400;

constexpr uint64_t DMP_T_INIT = 0xFFFFFFFFFFFFFFFFull;

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (stream definition: oracle/philox.py)
// ---------------------------------------------------------------------------------------------
struct Draw { uint32_t x0, x1, x2, x3; };

__device__ __forceinline__ Draw philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Draw{c0, c1, c2, c3};
}

__device__ __forceinline__ Draw env_draw(uint64_t seed, uint64_t env_id, uint64_t t) {
    return philox4x32_10((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)t, (uint32_t)(t >> 32),
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// In-step draws (stream definition: oracle/philox.py).  One Philox block serves FOUR consecutive steps:
//   block = philox4x32_10(counter = (env lo, env hi, (t >> 2) lo, (t >> 2) hi), key = seed),  w = block.x[t & 3]
//   step_size = 1 + (((w & 0xffff) * 3) >> 16)          action = ((w >> 16) * A) >> 16
// so a rollout pays the ten Philox rounds once per four steps (they were a quarter of the 1D kernel's instructions).
struct StepDraws {
    Draw blk;
    uint64_t tb;                                        // block index held in blk (~0 = none)
    __device__ __forceinline__ StepDraws() : blk{0u, 0u, 0u, 0u}, tb(~0ull) {}
    __device__ __forceinline__ uint32_t word(uint64_t seed, uint64_t env_id, uint64_t t) {
        const uint64_t b = t >> 2;
        if (b != tb) {                                  // warp-uniform: t is the same for every env of a launch
            blk = env_draw(seed, env_id, b);
            tb = b;
        }
        const uint32_t j = (uint32_t)t & 3u;
        return j == 0u ? blk.x0 : (j == 1u ? blk.x1 : (j == 2u ? blk.x2 : blk.x3));
    }
};

__device__ __forceinline__ int draw_step_size(uint32_t w) { return 1 + (int)(((w & 0xFFFFu) * 3u) >> 16); }
__device__ __forceinline__ int draw_action(uint32_t w, int n_actions, int dist) {
    const uint32_t h = w >> 16;
    if (dist == DMP_ACT_REF3D) {
        const int v = (int)((h * 20u) >> 16);
        return v < 16 ? (v >> 2) : (v - 12);
    }
    return (int)((h * (uint32_t)n_actions) >> 16);
}
// plan index of an env that auto-resets in step t: its own block, keyed apart from the step draws
constexpr uint64_t DMP_PLAN_KEY = 0x504C414E5F4B4559ull;     // "PLAN_KEY"
__device__ __forceinline__ uint32_t plan_word(uint64_t seed, uint64_t env_id, uint64_t t) {
    return env_draw(seed ^ DMP_PLAN_KEY, env_id, t).x2;
}
__device__ __forceinline__ int draw_plan(uint32_t x, int n_plans) { return (int)__umulhi(x, (uint32_t)n_plans); }

// ---------------------------------------------------------------------------------------------
// observation element conversion.  Every raw observation value of the reference is a small
// integer held in float64; the normalised counters are an fp64 quotient.
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T obs_from_int(int v);
template <> __device__ __forceinline__ float   obs_from_int<float>(int v)   { return (float)v; }
template <> __device__ __forceinline__ double  obs_from_int<double>(int v)  { return (double)v; }
template <> __device__ __forceinline__ int16_t obs_from_int<int16_t>(int v) { return (int16_t)v; }

template <typename T> __device__ __forceinline__ T obs_from_ratio(int num, int den);
template <> __device__ __forceinline__ float   obs_from_ratio<float>(int n, int d)   { return (float)__ddiv_rn((double)n, (double)d); }
template <> __device__ __forceinline__ double  obs_from_ratio<double>(int n, int d)  { return __ddiv_rn((double)n, (double)d); }
template <> __device__ __forceinline__ int16_t obs_from_ratio<int16_t>(int n, int)   { return (int16_t)n; }

// spread the 7 bits of x into the low bits of 7 bytes (byte j = bit j): the partial products of
// x * 0x0002040810204081 never overlap, so one wide IMAD (FMA pipe) + one mask does it.
__device__ __forceinline__ uint64_t spread7(uint32_t x) {
    return ((uint64_t)x * 0x0002040810204081ull) & 0x0101010101010101ull;
}

// biased byte (value + 1: 2D codes 0 frame / 1 empty / 2 occupied, 3D height + 1 with 0 = frame) -> observation value
template <typename ObsT, int BYTE>
__device__ __forceinline__ ObsT obs_from_biased(uint32_t packed) {
    if constexpr (sizeof(ObsT) == 4) {
        // float: drop the byte into the mantissa of 2^23 (one PRMT), subtract 2^23 + 1 (one FADD): exact,
        // and keeps the per-element work off the conversion (XU) pipe
        return __uint_as_float(__byte_perm(packed, 0x4B000000u, 0x7650u + BYTE)) - 8388609.0f;
    } else {
        return obs_from_int<ObsT>((int)((packed >> (8 * BYTE)) & 0xFFu) - 1);
    }
}

// a plan byte as a bare 32-bit load result: nothing touches the register until the reward is computed, so the
// load's latency hides behind the observation stage (1D and 3D rollouts)
__device__ __forceinline__ int ldg_u8(const uint8_t* p) {
    int v;
    asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// counters -> the two trailing obs columns
template <typename T>
__device__ __forceinline__ void obs_counters(bool normalise, int cb, int cs, int total_brick, int total_step, T& o_cb, T& o_cs) {
    if (normalise) {
        o_cb = obs_from_ratio<T>(cb, total_brick);
        o_cs = obs_from_ratio<T>(cs, total_step);
    } else {
        o_cb = obs_from_int<T>(cb);
        o_cs = obs_from_int<T>(cs);
    }
}

// ---------------------------------------------------------------------------------------------
// DMP_OBS_REC: packed step records (include/dmp.h).  A record type stands in for ObsT in the kernel templates; a
// "row" of the warp tile is then ONE record instead of D observation values.
//   Rec56 (2D / 3D): bytes 0..48 window value + 1, 49 flags, 50-51 count_brick, 52-53 count_step, 54 reward (i8), 55 done
//   Rec16 (1D)     : 5 x i16 raw window, u16 count_brick, u16 count_step, i8 reward, u8 done
// ---------------------------------------------------------------------------------------------
struct Rec56 { uint32_t w[14]; };
struct Rec16 { uint32_t w[4]; };
// DMP_OBS_BITS: bit-packed step records (include/dmp.h), one or two 128-bit words per env that leave the thread's registers
// through 128-bit streaming stores -- no shared-memory tile.  Bits16 (2D): 49 x 2-bit window codes + 12-bit counters + reward
// code + done; Bits32 (3D): 49 x 4-bit window codes (height + 1, saturating at 15) + the same trailer.
struct Bits16 { uint32_t w[4]; };
struct Bits32 { uint32_t w[8]; };
template <typename T> struct is_rec { static constexpr bool value = false; };
template <> struct is_rec<Rec56> { static constexpr bool value = true; };
template <> struct is_rec<Rec16> { static constexpr bool value = true; };
template <> struct is_rec<Bits16> { static constexpr bool value = true; };
template <> struct is_rec<Bits32> { static constexpr bool value = true; };
template <typename T> struct is_bits { static constexpr bool value = false; };
template <> struct is_bits<Bits16> { static constexpr bool value = true; };
template <> struct is_bits<Bits32> { static constexpr bool value = true; };
// elements of ObsT per env in an observation buffer / warp tile
template <typename ObsT, int D> __host__ __device__ constexpr int row_elems() { return is_rec<ObsT>::value ? 1 : D; }

// seven window rows of seven biased bytes each (bits 0..55 of c[k]) -> the 49-byte stream in 13 words (byte 48 is the low
// byte of w[12]).  Every shift is a compile-time constant: funnel shifts / PRMTs.
__device__ __forceinline__ void pack49(const uint64_t (&c)[7], uint32_t (&w)[13]) {
#pragma unroll
    for (int j = 0; j < 13; ++j) {
        const int k = (4 * j) / 7, o = (4 * j) % 7;        // row and byte offset of the word's first stream byte
        uint64_t v = c[k] >> (8 * o);
        if (7 - o < 4 && k + 1 < 7) v |= c[k + 1] << (8 * (7 - o));
        w[j] = (uint32_t)v;
    }
}
__device__ __forceinline__ void rec56_store(Rec56* row, const uint32_t (&w)[13], int cb, int cs, float reward, bool done,
                                            bool saturated) {
    const uint32_t flags = (done ? (uint32_t)DMP_REC_DONE : 0u) | (saturated ? (uint32_t)DMP_REC_SATURATED : 0u);
    uint2* d = reinterpret_cast<uint2*>(row);              // records are 8 B aligned (56 = 7 x 8)
#pragma unroll
    for (int j = 0; j < 6; ++j) d[j] = make_uint2(w[2 * j], w[2 * j + 1]);
    d[6] = make_uint2((w[12] & 0xFFu) | (flags << 8) | ((uint32_t)(cb & 0xFFFF) << 16),
                      (uint32_t)(cs & 0xFFFF) | (((uint32_t)(int)reward & 0xFFu) << 16) | (done ? 1u << 24 : 0u));
}

// DMP_OBS_BITS trailer: 12-bit counters (saturating, flagged), the reward as an index into {0, 1, 5, 10, -1, -100} (every reward
// of the six classes), done, and the saturation flag: 30 bits starting at bit `sh` of the record's last payload word.
__device__ __forceinline__ uint32_t bits_reward_code(float r) {
    const int v = (int)r;
    return v == 0 ? 0u : (v == 1 ? 1u : (v == 5 ? 2u : (v == 10 ? 3u : (v == -1 ? 4u : 5u))));
}
__device__ __forceinline__ uint32_t bits_trailer_code(int cb, int cs, uint32_t reward_code, bool done, bool saturated) {
    const bool over = (cb | cs) > 0xFFF;
    return (uint32_t)min(cb, 0xFFF) | ((uint32_t)min(cs, 0xFFF) << 12) | (reward_code << 24) |
           (done ? 1u << 27 : 0u) | ((saturated || over) ? 1u << 28 : 0u);
}
__device__ __forceinline__ uint32_t bits_trailer(int cb, int cs, float reward, bool done, bool saturated) {
    return bits_trailer_code(cb, cs, bits_reward_code(reward), done, saturated);
}
// two 7-bit values held in the two halves of x (bits 0..6 and 16..22): bit j of each half moves to bit 2j of its half
__device__ __forceinline__ uint32_t spread2x7(uint32_t x) {
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// ---------------------------------------------------------------------------------------------
// warp-tile copy-out: a warp has staged `n_elems` contiguous T elements in shared memory at
// `tile`; stream them to `dst` (global) with 128-bit stores where the alignment allows.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void warp_tile_store(T* __restrict__ dst, const T* tile, int n_elems, int lane) {
    constexpr int PER16 = 16 / (int)sizeof(T);
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        const int n16 = n_elems / PER16;
        const uint4* s4 = reinterpret_cast<const uint4*>(tile);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (int i = lane; i < n16; i += 32) __stcs(d4 + i, s4[i]);       // streaming: obs is write-once
        for (int i = n16 * PER16 + lane; i < n_elems; i += 32) dst[i] = tile[i];
    } else {
        for (int i = lane; i < n_elems; i += 32) dst[i] = tile[i];
    }
}

// full-warp fast path: the element count is a compile-time constant (32 envs x D values), the span is
// 16 B aligned by construction when `dst` is; fully unrolled 128-bit streaming stores.
template <typename T, int N_ELEMS>
__device__ __forceinline__ void warp_tile_store_full(T* __restrict__ dst, const T* tile, int lane) {
    constexpr int PER16 = 16 / (int)sizeof(T);
    static_assert(N_ELEMS % PER16 == 0, "tile must be a whole number of 16 B vectors");
    constexpr int N16 = N_ELEMS / PER16;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) != 0) { warp_tile_store<T>(dst, tile, N_ELEMS, lane); return; }
    const uint4* s4 = reinterpret_cast<const uint4*>(tile);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    uint4 v[(N16 + 31) / 32];
#pragma unroll
    for (int i = 0; i < (N16 + 31) / 32; ++i)
        if (i * 32 + 31 < N16 || i * 32 + lane < N16) v[i] = s4[i * 32 + lane];
#pragma unroll
    for (int i = 0; i < (N16 + 31) / 32; ++i)
        if (i * 32 + 31 < N16 || i * 32 + lane < N16) __stcs(d4 + i * 32 + lane, v[i]);
}

// the same when the caller has already established that `dst` is 16 B aligned (no fallback path in the loop)
template <typename T, int N_ELEMS>
__device__ __forceinline__ void warp_tile_store_aligned(T* __restrict__ dst, const T* tile, int lane) {
    constexpr int PER16 = 16 / (int)sizeof(T);
    static_assert(N_ELEMS % PER16 == 0, "tile must be a whole number of 16 B vectors");
    constexpr int N16 = N_ELEMS / PER16;
    const uint4* s4 = reinterpret_cast<const uint4*>(tile);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    uint4 v[(N16 + 31) / 32];
#pragma unroll
    for (int i = 0; i < (N16 + 31) / 32; ++i)
        if (i * 32 + 31 < N16 || i * 32 + lane < N16) v[i] = s4[i * 32 + lane];
#pragma unroll
    for (int i = 0; i < (N16 + 31) / 32; ++i)
        if (i * 32 + 31 < N16 || i * 32 + lane < N16) __stcs(d4 + i * 32 + lane, v[i]);
}

// tile rows (D observation values, or one record) of `nrows` envs / of a full warp through the load/store path
template <typename ObsT, int D>
__device__ __forceinline__ void tile_rows_store(ObsT* __restrict__ dst, const ObsT* tile, int nrows, int lane) {
    if constexpr (is_rec<ObsT>::value)
        warp_tile_store<uint32_t>(reinterpret_cast<uint32_t*>(dst), reinterpret_cast<const uint32_t*>(tile),
                                  nrows * (int)(sizeof(ObsT) / 4), lane);
    else
        warp_tile_store<ObsT>(dst, tile, nrows * D, lane);
}
template <typename ObsT, int D>
__device__ __forceinline__ void tile_rows_store_full(ObsT* __restrict__ dst, const ObsT* tile, int lane) {
    if constexpr (is_rec<ObsT>::value)
        warp_tile_store_full<uint32_t, 32 * (int)(sizeof(ObsT) / 4)>(reinterpret_cast<uint32_t*>(dst),
                                                                     reinterpret_cast<const uint32_t*>(tile), lane);
    else
        warp_tile_store_full<ObsT, 32 * D>(dst, tile, lane);
}

// ---------------------------------------------------------------------------------------------
// Bulk (TMA) copy-out of a warp tile: ONE cp.async.bulk shared -> global instead of N16/32 x (LDS.128 + STG.128) per
// lane.  The copy engine reads the tile itself, so the 2 x 4 wavefronts per 512 B that the load/store path spends in
// the L1 data pipe disappear -- that pipe, not HBM, was the busiest unit of the single-step kernels (ncu: 69 %).
// Protocol: every lane orders its own tile writes before the async proxy (fence.proxy.async), the warp converges, one
// lane issues the copy and commits it; `warp_tile_bulk_wait` must be called before the tile is written again or the
// block exits.  `dst` and `tile` must be 16 B aligned and `bytes` a multiple of 16.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_tile_bulk_store(void* dst, const void* tile, uint32_t bytes, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));      // write-once stream
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                     ::"l"(dst), "r"((uint32_t)__cvta_generic_to_shared(tile)), "r"(bytes), "l"(pol) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
}
__device__ __forceinline__ void warp_tile_bulk_wait(int lane) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// L2 residency hints.  Env state is re-read every step while observations are written once and never
// read back by these kernels: state accesses carry an evict_last policy, observation stores are
// streaming (__stcs), so that on B200's 126 MB L2 the state of ~10^6 2D envs (67 MB) stays on chip and
// only the observation stream goes to HBM.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_keep(const uint4* p, uint64_t pol) {
    uint4 v;
    asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_keep(uint4* p, const uint4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A vector step is one short launch (1D: 4 MB of traffic, ~1 us of work) replayed
// back to back from a CUDA graph; with stream-serialised launches every step pays the full launch + block-scheduling
// latency after the previous grid has drained.  The hot kernels therefore (1) signal `launch_dependents` as soon as
// they start, so the next step's grid is scheduled while this one runs, and (2) touch no global memory before
// `griddepcontrol.wait`, which returns once the previous grid has completed and its writes are visible.  Both
// instructions are no-ops when a kernel is launched without the attribute (DmpIO.flags & DMP_F_NO_PDL).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// host-side helpers ---------------------------------------------------------------------------
template <typename... KArgs, typename... Args>
inline cudaError_t dmp_launch_pdl(bool pdl, void (*kern)(KArgs...), unsigned blocks, unsigned threads, size_t smem,
                                  cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks, 1, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int dmp_set_error(cudaError_t e);   // records e, returns DMP_OK / DMP_ECUDA

// 3D height maps (include/dmp.h): the NIBBLE maps u8[n][208] behind the u16 area are what the hot kernels
// (dmp_3d_roll.cu, dmp_3d_step.cu) read and write: cell i of an env is nibble i (byte i >> 1, low nibble first) =
// min(height, 15); bytes 200..207 pad the env to 13 granules of 16 B (bulk copies need 16 B alignment).  An env is "tall"
// once a height reaches TALL3 = 15: its flag (bit 7 of the position-row byte of aux.x) is set, its u16 "wide" map holds
// the exact heights and its nibbles saturate at 15; the wide map of every other env is stale scratch that nothing reads.
// Why nibbles: every reachable state of the reference's plans is tiny (plan height 6), and a single step has to fetch the
// 7..10 map rows under its window from DRAM, which is read in 64 B pieces -- 20 B byte rows cost 256 B per env-step for
// 49 useful cells (the single-step kernel's bound), 10 B nibble rows 160 B; rollouts stage half the bytes per launch and
// clear half per episode.  Not writing the wide map is the other half of the layout: a 2-byte store into a line that is
// not in L2 costs a 32 B read and a 32 B write of DRAM traffic (profiles/README.md).
//   dmp3d_widen(st, clear_flags) : wide map := nibbles for every env that is not tall; optionally drops all flags
//   dmp3d_sync_bytes(st)         : nibbles := min(wide, 15) and flag := any(wide >= TALL3), for every env
// The stage kernels and import only know the wide maps: widen(clear) runs before them and sync_bytes after;
// iou reads the wide maps after widen(keep); export reads whichever map is the env's exact one.
constexpr int TALL3 = 15;                    // nibbles of a non-tall env are <= 14: exact heights
constexpr int NIB3_STRIDE = 208;             // bytes per env in the nibble area (200 used)
constexpr uint32_t AUX3_TALL = 0x80u;        // aux.x bit 7 (pos_row is 3..22)
__host__ __device__ inline uint8_t* nmap3(const DmpState& st) {
    return reinterpret_cast<uint8_t*>(st.cells) + (size_t)st.n_envs * (CELLS3D * 2);
}
__device__ __forceinline__ int sat_nib(int h) { return min(h, 15); }
// eight nibbles (cells 8j .. 8j+7 of a map word) -> two words of four bytes each, cell order kept
__device__ __forceinline__ void nib8_to_bytes(uint32_t x, uint32_t& b0, uint32_t& b1) {
    const uint32_t lo = x & 0x0F0F0F0Fu, hi = (x >> 4) & 0x0F0F0F0Fu;    // even / odd cells as bytes
    b0 = __byte_perm(lo, hi, 0x5140);                                    // cells 0 1 2 3
    b1 = __byte_perm(lo, hi, 0x7362);                                    // cells 4 5 6 7
}
// the byte of a nibble map that holds cell i, rewritten with that cell := v (this thread is the env's only writer)
__device__ __forceinline__ void nib_store_global(uint8_t* map, int i, int v) {
    uint8_t* p = map + (i >> 1);
    const int sh = (i & 1) * 4;
    const uint32_t b = __ldcg(p);
    *p = (uint8_t)((b & ~(0xFu << sh)) | ((uint32_t)v << sh));
}
int dmp3d_sync_bytes(const DmpState& st, cudaStream_t s);
int dmp3d_widen(const DmpState& st, bool clear_flags, cudaStream_t s);
inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// per-dimension entry points (defined in dmp_1d.cu / dmp_2d.cu / dmp_3d.cu)
int dmp1d_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s);
int dmp2d_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s);
int dmp3d_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s);
int dmp3d_cache_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s);      // dmp_3d_roll.cu
int dmp3d_step_bytes(const DmpState& st, const DmpIO& io, cudaStream_t s);                // dmp_3d_step.cu
int dmp1d_reset(const DmpState& st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs, int obs_kind, cudaStream_t s);
int dmp2d_reset(const DmpState& st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs, int obs_kind, cudaStream_t s);
int dmp3d_reset(const DmpState& st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs, int obs_kind, cudaStream_t s);
int dmp1d_iou(const DmpState& st, double* out, cudaStream_t s);
int dmp2d_iou(const DmpState& st, double* out, cudaStream_t s);
int dmp3d_iou(const DmpState& st, double* out, cudaStream_t s);
int dmp1d_export(const DmpState& st, int32_t* grid, int32_t* scalars, float* ret, cudaStream_t s);
int dmp2d_export(const DmpState& st, int32_t* grid, int32_t* scalars, float* ret, cudaStream_t s);
int dmp3d_export(const DmpState& st, int32_t* grid, int32_t* scalars, float* ret, cudaStream_t s);
int dmp1d_import(const DmpState& st, const int32_t* grid, const int32_t* scalars, const float* ret, cudaStream_t s);
int dmp2d_import(const DmpState& st, const int32_t* grid, const int32_t* scalars, const float* ret, cudaStream_t s);
int dmp3d_import(const DmpState& st, const int32_t* grid, const int32_t* scalars, const float* ret, cudaStream_t s);

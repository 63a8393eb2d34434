// dmp_3d.cu -- 3D (2.5-D height map) mobile-construction envs: 20x20 heights, 7x7 window,
// 8 actions (4 collision-checked moves, 4 adjacent builds).  This file holds the dispatch of dmp_rollout for 3D and the
// warp-per-env utility kernels (reset, IoU, export, import, wide <-> nibble map passes); the hot kernels are
// dmp_3d_roll.cu (K steps per launch) and dmp_3d_step.cu (single steps).
//
// Reference semantics: Env/3D/DMP_simulator_3d_static_circle.py:67-276 (static, T=1300) and
// Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-277 (dynamic, T=1000, re-check after
// placement and the -100 boxed-in penalty).
#include "dmp_common.cuh"

namespace {

constexpr int WPB3 = 8;                  // warps (= envs) per block
constexpr unsigned FULL = 0xFFFFFFFFu;

struct Env3 {
    int pr, pc, plan_idx, cb, cs;
    float ret;
    int cross;      // sum(min(height, plan)) maintained incrementally: +1 for every brick laid at or below the plan
};

__device__ __forceinline__ void unpack3(const uint4& a, Env3& e) {
    e.pr = a.x & 0x7F; e.pc = (a.x >> 8) & 0xFF; e.plan_idx = a.x >> 16;      // bit 7: AUX3_TALL, not this struct's business
    e.cb = a.y & 0xFFFF; e.cs = a.y >> 16;
    e.ret = __uint_as_float(a.z);
    e.cross = (int)a.w;
}
__device__ __forceinline__ uint4 pack3(const Env3& e) {
    return make_uint4((uint32_t)e.pr | ((uint32_t)e.pc << 8) | ((uint32_t)e.plan_idx << 16),
                      (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16), __float_as_uint(e.ret), (uint32_t)e.cross);
}

// (d) IoU of Env/3D/DMP_simulator_3d_static_circle.py:257-276, one warp per env:
// cross = sum(min(h, plan)); iou = cross / (total_brick + count_brick - cross).  Optionally clears the map.
__device__ __forceinline__ double warp_iou3(uint16_t* grid, const uint8_t* __restrict__ plan, int total_brick,
                                            int count_brick, int lane, bool clear) {
    int cross = 0;
    if (lane < 25) {                     // 16 cells per lane: two 128-bit height loads, one 128-bit plan load
        const uint4 h0 = reinterpret_cast<const uint4*>(grid)[2 * lane];
        const uint4 h1 = reinterpret_cast<const uint4*>(grid)[2 * lane + 1];
        const uint4 pp = __ldg(reinterpret_cast<const uint4*>(plan) + lane);
        const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const uint32_t pw[4] = {pp.x, pp.y, pp.z, pp.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int ha = hw[i] & 0xFFFF, hb = hw[i] >> 16;
            const int pa = (pw[i >> 1] >> ((i & 1) * 16)) & 0xFF, pb = (pw[i >> 1] >> ((i & 1) * 16 + 8)) & 0xFF;
            cross += min(ha, pa) + min(hb, pb);
        }
        if (clear) {
            const uint4 z = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4*>(grid)[2 * lane] = z;
            reinterpret_cast<uint4*>(grid)[2 * lane + 1] = z;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cross += __shfl_xor_sync(FULL, cross, o);
    return __ddiv_rn((double)cross, (double)(total_brick + count_brick - cross));
}

// reset: one warp per env (coalesced clear of the 800 B map)
template <typename ObsT>
__global__ void k3d_reset(const DmpState st, const uint8_t* __restrict__ mask, const int32_t* __restrict__ plan_idx,
                          const uint64_t t_draw, ObsT* __restrict__ obs) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs || (mask && !mask[env])) return;
    uint4* aux = reinterpret_cast<uint4*>(st.aux);
    int p;
    if (plan_idx) {
        p = plan_idx[env];
        if ((unsigned)p >= (unsigned)st.n_plans) { if (lane == 0) atomicOr(st.err, DMP_ERR_PLANIDX); p = 0; }
    } else if (st.plan_mode == DMP_PLAN_PHILOX) {
        p = draw_plan(env_draw(st.seed, (uint64_t)(st.env_base + env), t_draw).x3, st.n_plans);
    } else {
        const uint32_t ax = aux[env].x;
        p = (int)(ax >> 16);
        // sequential order starts at plan 0 on an env that has never been reset (zeroed state: row 0 is not a position),
        // index_for_non_random = 0 of Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:51-56
        if (st.plan_mode == DMP_PLAN_SEQUENTIAL) p = ((ax & 0x7Fu) == 0u) ? 0 : ((p + 1 >= st.n_plans) ? 0 : p + 1);
        if ((unsigned)p >= (unsigned)st.n_plans) p = 0;
    }
    __syncwarp();
    uint4* g4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(st.cells) + env * CELLS3D);
    const uint4 z = make_uint4(0, 0, 0, 0);
    g4[lane] = z;
    if (lane < 18) g4[lane + 32] = z;
    if (lane < NIB3_STRIDE / 16) reinterpret_cast<uint4*>(nmap3(st) + env * NIB3_STRIDE)[lane] = z;      // nibble map
    if (lane == 0) aux[env] = pack3(Env3{D2_LO, D2_LO, p, 0, 0, 0.f, 0});
    if (obs) {
        if constexpr (is_bits<ObsT>::value) {               // 49 four-bit codes: 0 frame (rows / columns 0..2), 1 empty
            if (lane < 8) {
                uint32_t w = 0u;
                for (int i = lane * 8; i < lane * 8 + 8 && i < 49; ++i)
                    if (i / 7 >= 3 && i % 7 >= 3) w |= 1u << ((i & 7) * 4);
                reinterpret_cast<uint32_t*>(obs + env)[lane] = w;
            }
        } else if constexpr (is_rec<ObsT>::value) {
            uint8_t* o = reinterpret_cast<uint8_t*>(obs + env);
            for (int i = lane; i < 56; i += 32) o[i] = (i < 49 && i / 7 >= 3 && i % 7 >= 3) ? 1 : 0;
        } else {
            ObsT* o = obs + env * D3_OBS;
            for (int j = lane; j < D3_OBS; j += 32)
                o[j] = obs_from_int<ObsT>((j < 49 && (j / 7 < 3 || j % 7 < 3)) ? -1 : 0);
        }
    }
}

__global__ void k3d_iou(const DmpState st, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs) return;
    Env3 e;
    unpack3(reinterpret_cast<const uint4*>(st.aux)[env], e);
    uint16_t* grid = reinterpret_cast<uint16_t*>(st.cells) + env * CELLS3D;
    const double v = warp_iou3(grid, reinterpret_cast<const uint8_t*>(st.plans) + e.plan_idx * CELLS3D,
                               st.plan_total[e.plan_idx], e.cb, lane, false);
    if (lane == 0) out[env] = v;
}

__global__ void k3d_export(const DmpState st, int32_t* __restrict__ grid, int32_t* __restrict__ scalars, float* __restrict__ ret) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs) return;
    // heights of an env that is not tall are its nibbles (its wide map is scratch); a tall env's are in its wide map
    const uint4 ax = reinterpret_cast<const uint4*>(st.aux)[env];
    const bool tall = (ax.x & AUX3_TALL) != 0u;
    const uint16_t* g = reinterpret_cast<const uint16_t*>(st.cells) + env * CELLS3D;
    const uint8_t* nib = nmap3(st) + env * NIB3_STRIDE;
    if (grid)
        for (int i = lane; i < 676; i += 32) {
            const int r = i / 26, c = i % 26;
            int v = -1;
            if (r >= 3 && r < 23 && c >= 3 && c < 23) {
                const int ci = (r - 3) * 20 + (c - 3);
                v = tall ? (int)g[ci] : (int)((nib[ci >> 1] >> ((ci & 1) * 4)) & 0xFu);
            }
            grid[env * 676 + i] = v;
        }
    if (lane == 0) {
        Env3 e;
        unpack3(ax, e);
        if (scalars) {
            int32_t* s = scalars + env * 8;
            s[0] = e.pr; s[1] = e.pc; s[2] = e.cb; s[3] = e.cs; s[4] = e.plan_idx;
            s[5] = st.plan_total[e.plan_idx]; s[6] = 0; s[7] = 0;
        }
        if (ret) ret[env] = e.ret;
    }
}

__global__ void k3d_import(const DmpState st, const int32_t* __restrict__ grid, const int32_t* __restrict__ scalars,
                           const float* __restrict__ ret) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs) return;
    uint16_t* g = reinterpret_cast<uint16_t*>(st.cells) + env * CELLS3D;
    if (grid)
        for (int i = lane; i < CELLS3D; i += 32) g[i] = (uint16_t)grid[env * 676 + (i / 20 + 3) * 26 + (i % 20 + 3)];
    __syncwarp();
    uint4* aux = reinterpret_cast<uint4*>(st.aux);
    Env3 e;
    unpack3(aux[env], e);
    if (scalars) {
        const int32_t* s = scalars + env * 8;
        e.pr = s[0]; e.pc = s[1]; e.cb = s[2]; e.cs = s[3]; e.plan_idx = s[4];
    }
    if (ret) e.ret = ret[env];
    // the running sum(min(height, plan)) must match the (possibly new) map and plan
    const uint8_t* plan = reinterpret_cast<const uint8_t*>(st.plans) + e.plan_idx * CELLS3D;
    int cross = 0;
    for (int i = lane; i < CELLS3D; i += 32) cross += min((int)g[i], (int)plan[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cross += __shfl_xor_sync(FULL, cross, o);
    e.cross = cross;
    if (lane == 0) aux[env] = pack3(e);
}

// nibbles := min(wide, 15) and flag := any(wide >= TALL3) for every env; one warp per env
__global__ void k3d_sync_bytes(const DmpState st) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs) return;
    bool tall = false;
    if (lane < 25) {                                     // 16 cells per lane: 32 B of heights -> 8 B of nibbles
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(st.cells) + env * CELLS3D) + 2 * lane;
        const uint4 a = src[0], b = src[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[2] = {0u, 0u};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t h0 = w[q] & 0xFFFFu, h1 = w[q] >> 16;
            tall |= max(h0, h1) >= (uint32_t)TALL3;
            o[q >> 2] |= (min(h0, 15u) | (min(h1, 15u) << 4)) << (8 * (q & 3));
        }
        reinterpret_cast<uint2*>(nmap3(st) + env * NIB3_STRIDE)[lane] = make_uint2(o[0], o[1]);
    } else if (lane == 25) {
        reinterpret_cast<uint2*>(nmap3(st) + env * NIB3_STRIDE)[25] = make_uint2(0u, 0u);      // pad bytes 200..207
    }
    const bool any = __any_sync(FULL, tall);
    if (lane == 0) {
        uint32_t* ax = reinterpret_cast<uint32_t*>(reinterpret_cast<uint4*>(st.aux) + env);
        *ax = (*ax & ~AUX3_TALL) | (any ? AUX3_TALL : 0u);
    }
}

// wide := nibbles for every env that is not tall (whose nibbles are exact); optionally drops all flags afterwards
__global__ void k3d_widen(const DmpState st, const int clear_flags) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs) return;
    uint32_t* ax = reinterpret_cast<uint32_t*>(reinterpret_cast<uint4*>(st.aux) + env);
    const uint32_t x = *ax;
    if (!(x & AUX3_TALL) && lane < 25) {
        const uint2 b = reinterpret_cast<const uint2*>(nmap3(st) + env * NIB3_STRIDE)[lane];
        const uint32_t w[2] = {b.x, b.y};
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t by = (w[q >> 2] >> (8 * (q & 3))) & 0xFFu;       // cells 2q, 2q+1 of this lane's 16
            o[q] = (by & 0xFu) | ((by >> 4) << 16);
        }
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(st.cells) + env * CELLS3D) + 2 * lane;
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    __syncwarp();
    if (clear_flags && lane == 0 && (x & AUX3_TALL)) *ax = x & ~AUX3_TALL;
}

inline unsigned blocks3(int64_t n) { return (unsigned)((n + WPB3 - 1) / WPB3); }

}  // namespace

int dmp3d_sync_bytes(const DmpState& st, cudaStream_t s) {
    k3d_sync_bytes<<<blocks3(st.n_envs), WPB3 * 32, 0, s>>>(st);
    return dmp_set_error(cudaGetLastError());
}
int dmp3d_widen(const DmpState& st, bool clear_flags, cudaStream_t s) {
    k3d_widen<<<blocks3(st.n_envs), WPB3 * 32, 0, s>>>(st, clear_flags ? 1 : 0);
    return dmp_set_error(cudaGetLastError());
}

int dmp3d_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    // rollouts (K > 1) stage whole maps in shared memory; a single step fetches only the rows it can look at
    if (K > 1 || (io.flags & DMP_F_ROLLOUT_K1)) return dmp3d_cache_rollout(st, io, K, s);
    return dmp3d_step_bytes(st, io, s);
}

int dmp3d_reset(const DmpState& st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs,
                int obs_kind, cudaStream_t s) {
    const unsigned b = blocks3(st.n_envs);
    switch (obs_kind) {
        case DMP_OBS_F32: k3d_reset<float><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (float*)obs); break;
        case DMP_OBS_F64: k3d_reset<double><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (double*)obs); break;
        case DMP_OBS_I16: k3d_reset<int16_t><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (int16_t*)obs); break;
        case DMP_OBS_REC: k3d_reset<Rec56><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (Rec56*)obs); break;
        case DMP_OBS_BITS: k3d_reset<Bits32><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (Bits32*)obs); break;
        default: return DMP_EINVAL;
    }
    return dmp_set_error(cudaGetLastError());
}

int dmp3d_iou(const DmpState& st, double* out, cudaStream_t s) {
    const int rc = dmp3d_widen(st, false, s);            // the kernel reads the wide maps
    if (rc != DMP_OK) return rc;
    k3d_iou<<<blocks3(st.n_envs), WPB3 * 32, 0, s>>>(st, out);
    return dmp_set_error(cudaGetLastError());
}
int dmp3d_export(const DmpState& st, int32_t* grid, int32_t* scalars, float* ret, cudaStream_t s) {
    k3d_export<<<blocks3(st.n_envs), WPB3 * 32, 0, s>>>(st, grid, scalars, ret);
    return dmp_set_error(cudaGetLastError());
}
int dmp3d_import(const DmpState& st, const int32_t* grid, const int32_t* scalars, const float* ret, cudaStream_t s) {
    int rc = dmp3d_widen(st, true, s);                   // the kernel rewrites aux.x and sums over the wide map
    if (rc != DMP_OK) return rc;
    k3d_import<<<blocks3(st.n_envs), WPB3 * 32, 0, s>>>(st, grid, scalars, ret);
    rc = dmp_set_error(cudaGetLastError());
    return rc != DMP_OK ? rc : dmp3d_sync_bytes(st, s);
}

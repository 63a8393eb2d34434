// dmp_3d.cu -- 3D (2.5-D height map) mobile-construction envs: 20x20 heights, 7x7 window,
// 8 actions (4 collision-checked moves, 4 adjacent builds).
//
// Reference semantics: Env/3D/DMP_simulator_3d_static_circle.py:67-276 (static, T=1300) and
// Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-277 (dynamic, T=1000, re-check after
// placement and the -100 boxed-in penalty).
//
// Mapping: one env per WARP.  The 800 B height map of an env is contiguous in HBM; the 12 cells of
// the movement cross (3 per direction) are fetched by 12 lanes in one round and exchanged with
// shuffles, the (warp-uniform) step logic runs redundantly in every lane, then lanes 0..50 gather
// the 7x7 window at the new position and write the 51 observation values as one coalesced row.
// IoU of a finished episode is a warp reduction (128-bit loads, min, shuffle tree) fused with the
// reset of the height map.
#include <stdlib.h>
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

// value of environment_memory[r][c] in padded coordinates: -1 on the 3-cell frame (:72-75)
__device__ __forceinline__ int cell3(const uint16_t* grid, int r, int c) {
    const unsigned ir = (unsigned)(r - 3), ic = (unsigned)(c - 3);
    return (ir < 20u && ic < 20u) ? (int)grid[ir * 20u + ic] : -1;
}

// direction table: 0 left (c-1), 1 right (c+1), 2 "up" (r+1), 3 "down" (r-1)   (check_sur :88-102)
__device__ __forceinline__ int dir_dr(int d) { return d == 2 ? 1 : (d == 3 ? -1 : 0); }
__device__ __forceinline__ int dir_dc(int d) { return d == 0 ? -1 : (d == 1 ? 1 : 0); }

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

template <typename ObsT>
__global__ void __launch_bounds__(WPB3 * 32) k3d_rollout(const DmpState st, const DmpIO io, const int K) {
    const int lane = threadIdx.x & 31;
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= n) return;                                            // whole warp leaves together
    uint16_t* grid = reinterpret_cast<uint16_t*>(st.cells) + env * CELLS3D;
    uint4* aux = reinterpret_cast<uint4*>(st.aux);
    const uint8_t* __restrict__ plans = reinterpret_cast<const uint8_t*>(st.plans);
    Env3 e;
    unpack3(aux[env], e);
    int total_brick = __ldg(st.plan_total + e.plan_idx);
    int errbits = 0;
    const bool dynamic = st.dynamic != 0;
    const bool autoreset = io.flags & DMP_F_AUTORESET;
    const bool normalise = io.flags & DMP_F_NORMALISE;
    const bool need_draw = (io.actions == nullptr) || (io.step_sizes == nullptr);
    const int tslot = (io.flags & DMP_F_TSLOT1) ? 1 : 0;
    const uint64_t t0 = st.t_dev ? st.t_dev[tslot] : st.t;

    // lane role in the movement cross: direction lane/3, distance lane%3 + 1
    const int xdir = lane / 3, xdist = lane % 3 + 1;
    const int xdr = dir_dr(xdir) * xdist, xdc = dir_dc(xdir) * xdist;

    StepDraws draws;

    for (int k = 0; k < K; ++k) {
        const uint64_t t = t0 + (uint64_t)k;
        const int64_t idx = (int64_t)k * n + env;
        uint32_t dw = 0;
        if (need_draw) dw = draws.word(st.seed, (uint64_t)(st.env_base + env), t);
        const int a = io.actions ? (int)io.actions[idx] : draw_action(dw, D3_ACT, st.action_dist);
        const int s = io.step_sizes ? (int)io.step_sizes[idx] : draw_step_size(dw);
        if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;
        if (a > 7) errbits |= DMP_ERR_ACTION;          // reference: treated as an unbuilt brick (:187-208)

        e.cs += 1;
        // ---- check_sur (:88-102) + move_step operands (:104-134): one load round --------------
        int cv = 0;
        if (lane < 12) cv = cell3(grid, e.pr + xdr, e.pc + xdc);
        int n1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) n1[q] = __shfl_sync(FULL, cv, q * 3);
        const bool boxed = (n1[0] != 0) && (n1[1] != 0) && (n1[2] != 0) && (n1[3] != 0);

        float reward = 0.f;
        bool done;
        bool tail = true;                                   // falls through to the common ending
        if (a <= 3) {
            // (a) move: legal iff the adjacent cell is exactly 0; walk over consecutive empty cells
            const int na = (a == 0) ? n1[0] : (a == 1) ? n1[1] : (a == 2) ? n1[2] : n1[3];
            const int v2 = __shfl_sync(FULL, cv, a * 3 + 1), v3 = __shfl_sync(FULL, cv, a * 3 + 2);
            if (na == 0) {
                int nstep = 1;
                if (s >= 2 && v2 == 0) { nstep = 2; if (s >= 3 && v3 == 0) nstep = 3; }
                e.pr = min(max(e.pr + dir_dr(a) * nstep, D2_LO), D2_HI);     // clip_position :142-151
                e.pc = min(max(e.pc + dir_dc(a) * nstep, D2_LO), D2_HI);
            }
        } else {
            // (b) build on the adjacent cell a-4 unless it is frame (:187-208)
            const int q = a - 4;
            bool built = false;
            int newh = 0, tr = 0, tc = 0, pplan = 0;
            int m1[4] = {n1[0], n1[1], n1[2], n1[3]};
            if (a <= 7) {
                const int nq = (q == 0) ? n1[0] : (q == 1) ? n1[1] : (q == 2) ? n1[2] : n1[3];
                if (nq != -1) {
                    built = true;
                    newh = nq + 1;
                    tr = e.pr + dir_dr(q); tc = e.pc + dir_dc(q);
                    e.cb += 1;
                    pplan = plans[e.plan_idx * CELLS3D + (tr - 3) * 20 + (tc - 3)];
                    if (newh <= pplan) e.cross += 1;
                    if (lane == 0) grid[(tr - 3) * 20 + (tc - 3)] = (uint16_t)newh;
#pragma unroll
                    for (int z = 0; z < 4; ++z) if (z == q) m1[z] = newh;
                }
            }
            if (dynamic) {
                // re-check AFTER placement (Env/3D/...usedata.py:199-221)
                const bool boxed2 = (m1[0] != 0) && (m1[1] != 0) && (m1[2] != 0) && (m1[3] != 0);
                if (boxed2) { reward = -100.f; done = true; tail = false; }
                else if (e.cb >= total_brick) { reward = 0.f; done = true; tail = false; }
                else if (built) { tail = false; done = false; }
            } else {
                // pre-placement check (static :210-221); a successful build never tests the step limit
                if (e.cb >= total_brick || boxed) { reward = 0.f; done = true; tail = false; }
                else if (built) { tail = false; done = false; }
            }
            if (!tail && !done)                             // reward_check :232-239
                reward = (newh > pplan) ? -1.f : (newh == pplan ? 10.f : 1.f);
        }
        if (tail) {
            done = (e.cs >= st.total_step) || (!dynamic && boxed);       // static :226 / dynamic :226
            reward = 0.f;
        }
        e.ret += reward;
        __syncwarp();                                       // lane 0's brick is visible to the gather below

        // ---- (c) observation: 7x7 window at the new position + counters, one coalesced row ----
        if (io.obs) {
            ObsT* o = reinterpret_cast<ObsT*>(io.obs) + idx * D3_OBS;
            ObsT ocb, ocs;
            obs_counters<ObsT>(normalise, e.cb, e.cs, total_brick, st.total_step, ocb, ocs);
            {
                const int j = lane;                         // 0..31
                o[j] = obs_from_int<ObsT>(cell3(grid, e.pr - 3 + j / 7, e.pc - 3 + j % 7));
            }
            if (lane < 19) {
                const int j = lane + 32;                    // 32..50
                ObsT v;
                if (j < 49) v = obs_from_int<ObsT>(cell3(grid, e.pr - 3 + j / 7, e.pc - 3 + j % 7));
                else v = (j == 49) ? ocb : ocs;
                o[j] = v;
            }
        }
        if (lane == 0) {
            if (io.reward) io.reward[idx] = reward;
            if (io.done) io.done[idx] = done ? 1 : 0;
        }

        // ---- (e) done / auto-reset ------------------------------------------------------------
        if (done && autoreset) {
            __syncwarp();
            const double iou = warp_iou3(grid, plans + e.plan_idx * CELLS3D, total_brick, e.cb, lane, true);
            if (lane == 0) {
                atomicAdd(st.ep_cnt + env, 1u);                 // fire-and-forget REDs: no read-modify-write stall
                atomicAdd(st.ep_len + env, (uint32_t)e.cs);
                atomicAdd(st.ep_ret + env, (double)e.ret);
                atomicAdd(st.ep_iou + env, iou);
            }
            if (io.next_plan) {
                const int p = io.next_plan[idx];
                if ((unsigned)p >= (unsigned)st.n_plans) errbits |= DMP_ERR_PLANIDX; else e.plan_idx = p;
            } else if (st.plan_mode == DMP_PLAN_PHILOX) {
                e.plan_idx = draw_plan(plan_word(st.seed, (uint64_t)(st.env_base + env), t), st.n_plans);
            } else if (st.plan_mode == DMP_PLAN_SEQUENTIAL) {
                e.plan_idx = (e.plan_idx + 1 == st.n_plans) ? 0 : e.plan_idx + 1;
            }
            total_brick = __ldg(st.plan_total + e.plan_idx);
            e.pr = e.pc = D2_LO; e.cb = e.cs = 0; e.ret = 0.f; e.cross = 0;
            __syncwarp();
        }
    }
    if (lane == 0) {
        aux[env] = pack3(e);
        if (errbits) atomicOr(st.err, errbits);
        if (st.t_dev && env == 0) st.t_dev[tslot ^ 1] = t0 + (uint64_t)K;
    }
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
        p = (int)(aux[env].x >> 16);
        if (st.plan_mode == DMP_PLAN_SEQUENTIAL) p = (p + 1 >= st.n_plans) ? 0 : p + 1;
        if ((unsigned)p >= (unsigned)st.n_plans) p = 0;
    }
    __syncwarp();
    uint4* g4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(st.cells) + env * CELLS3D);
    const uint4 z = make_uint4(0, 0, 0, 0);
    g4[lane] = z;
    if (lane < 18) g4[lane + 32] = z;
    if (lane < 25) reinterpret_cast<uint4*>(bmap3(st) + env * CELLS3D)[lane] = z;       // byte shadow
    if (lane == 0) aux[env] = pack3(Env3{D2_LO, D2_LO, p, 0, 0, 0.f, 0});
    if (obs) {
        ObsT* o = obs + env * D3_OBS;
        for (int j = lane; j < D3_OBS; j += 32)
            o[j] = obs_from_int<ObsT>((j < 49 && (j / 7 < 3 || j % 7 < 3)) ? -1 : 0);
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
    const uint16_t* g = reinterpret_cast<const uint16_t*>(st.cells) + env * CELLS3D;
    if (grid)
        for (int i = lane; i < 676; i += 32) grid[env * 676 + i] = cell3(g, i / 26, i % 26);
    if (lane == 0) {
        Env3 e;
        unpack3(reinterpret_cast<const uint4*>(st.aux)[env], e);
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

// bytes := min(wide, 255) and flag := any(wide >= TALL3) for every env; one warp per env
__global__ void k3d_sync_bytes(const DmpState st) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs) return;
    bool tall = false;
    if (lane < 25) {                                     // 16 cells per lane
        const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(st.cells) + env * CELLS3D) + 2 * lane;
        const uint4 a = src[0], b = src[1];
        const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t h0 = w[2 * q] & 0xFFFFu, h1 = w[2 * q] >> 16, h2 = w[2 * q + 1] & 0xFFFFu, h3 = w[2 * q + 1] >> 16;
            tall |= max(max(h0, h1), max(h2, h3)) >= (uint32_t)TALL3;
            o[q] = min(h0, 255u) | (min(h1, 255u) << 8) | (min(h2, 255u) << 16) | (min(h3, 255u) << 24);
        }
        reinterpret_cast<uint4*>(bmap3(st) + env * CELLS3D)[lane] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    const bool any = __any_sync(FULL, tall);
    if (lane == 0) {
        uint32_t* ax = reinterpret_cast<uint32_t*>(reinterpret_cast<uint4*>(st.aux) + env);
        *ax = (*ax & ~AUX3_TALL) | (any ? AUX3_TALL : 0u);
    }
}

// wide := bytes for every env that is not tall (whose bytes are exact); optionally drops all flags afterwards
__global__ void k3d_widen(const DmpState st, const int clear_flags) {
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * WPB3 + (threadIdx.x >> 5);
    if (env >= st.n_envs) return;
    uint32_t* ax = reinterpret_cast<uint32_t*>(reinterpret_cast<uint4*>(st.aux) + env);
    const uint32_t x = *ax;
    if (!(x & AUX3_TALL) && lane < 25) {
        const uint4 b = reinterpret_cast<const uint4*>(bmap3(st) + env * CELLS3D)[lane];
        const uint32_t w[4] = {b.x, b.y, b.z, b.w};
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            o[2 * q] = (w[q] & 0xFFu) | ((w[q] & 0xFF00u) << 8);
            o[2 * q + 1] = ((w[q] >> 16) & 0xFFu) | ((w[q] >> 24) << 16);
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

// warp-per-env kernel (round-1 first version; kept as a cross-check and for DMP_3D_KERNEL=wpe)
static int dmp3d_wpe_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    const unsigned b = blocks3(st.n_envs);
    switch (io.obs_kind) {
        case DMP_OBS_F32: k3d_rollout<float><<<b, WPB3 * 32, 0, s>>>(st, io, K); break;
        case DMP_OBS_F64: k3d_rollout<double><<<b, WPB3 * 32, 0, s>>>(st, io, K); break;
        case DMP_OBS_I16: k3d_rollout<int16_t><<<b, WPB3 * 32, 0, s>>>(st, io, K); break;
        default: return DMP_EINVAL;
    }
    return dmp_set_error(cudaGetLastError());
}

int dmp3d_sync_bytes(const DmpState& st, cudaStream_t s) {
    k3d_sync_bytes<<<blocks3(st.n_envs), WPB3 * 32, 0, s>>>(st);
    return dmp_set_error(cudaGetLastError());
}
int dmp3d_widen(const DmpState& st, bool clear_flags, cudaStream_t s) {
    k3d_widen<<<blocks3(st.n_envs), WPB3 * 32, 0, s>>>(st, clear_flags ? 1 : 0);
    return dmp_set_error(cudaGetLastError());
}

int dmp3d_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    // DMP_3D_KERNEL = w (warp per env) | t (u16 tile, whole maps by bulk async copies) | c (byte cache) |
    // r / s (single step, first / second generation over the u16 maps; K == 1) forces one kernel.  Default: the
    // byte-cache kernel for rollouts (K > 1), the byte-row kernel (dmp_3d_step3.cu) for single steps.
    const char* v = getenv("DMP_3D_KERNEL");             // read per call: tests switch kernels in-process
    const int forced = v ? (int)v[0] : 0;
    const bool legacy = forced == 'w' || forced == 't' || ((forced == 'r' || forced == 's') && K == 1);
    if (legacy) {                                        // these only know the wide (u16) maps
        int rc = dmp3d_widen(st, true, s);
        if (rc != DMP_OK) return rc;
        if (forced == 'w') rc = dmp3d_wpe_rollout(st, io, K, s);
        else if (forced == 't') rc = dmp3d_tile_rollout(st, io, K, s);
        else if (forced == 'r') rc = dmp3d_step_rows(st, io, s);
        else rc = dmp3d_step_span(st, io, s);
        return rc != DMP_OK ? rc : dmp3d_sync_bytes(st, s);
    }
    if (forced == 'c') return dmp3d_cache_rollout(st, io, K, s);
    return K > 1 ? dmp3d_cache_rollout(st, io, K, s) : dmp3d_step_bytes(st, io, s);
}

int dmp3d_reset(const DmpState& st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs,
                int obs_kind, cudaStream_t s) {
    const unsigned b = blocks3(st.n_envs);
    switch (obs_kind) {
        case DMP_OBS_F32: k3d_reset<float><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (float*)obs); break;
        case DMP_OBS_F64: k3d_reset<double><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (double*)obs); break;
        case DMP_OBS_I16: k3d_reset<int16_t><<<b, WPB3 * 32, 0, s>>>(st, mask, plan_idx, t_draw, (int16_t*)obs); break;
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
    if (grid) {
        const int rc = dmp3d_widen(st, false, s);
        if (rc != DMP_OK) return rc;
    }
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

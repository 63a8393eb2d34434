// dmp_1d.cu -- 1D mobile-construction envs (30-column height field, 5-cell window, 3 actions).
//
// Reference semantics: Env/1D/DMP_Env_1D_static.py:57-151 and
// Env/1D/DMP_Env_1D_dynamic_usedata_plan.py:32-133 (same step logic; dataset plan).
//
// Mapping: one env per thread, state = 4 coalesced 128-bit loads (30 x u16 heights + the two
// counters) + one 64-bit load (position, plan index, running return).  Heights are parked in shared
// memory in the reference's padded form (a -1 halfword pair either side) so the 5-cell window is
// three conflict-free word reads + funnel shifts; observations are staged per warp and streamed out
// contiguously.
#include "dmp_common.cuh"

namespace {

// B1 = threads (= envs) per block is a template parameter: 128 for large batches, 32 when the batch is too small
// to give every SM several blocks (BASELINE config 2: 65 536 envs = 2 048 warps for 592 schedulers)
constexpr int S1_WORDS = 17;          // padded 34 halfwords: word 0 and 16 are the -1 walls

struct Env1 {
    int pos, plan_idx, cb, cs;
    float ret;
};

// stage (a): clip_position, Env/1D/DMP_Env_1D_static.py:57-64
__device__ __forceinline__ void stage_move1(Env1& e, int a, int s) {
    const int p = (a == 0) ? e.pos - s : e.pos + s;
    e.pos = min(max(p, D1_LO), D1_HI);
}

// stage (b): environment_memory[0, position] += 1 (:104); returns the new height
template <int B1>
__device__ __forceinline__ int stage_deposit1(uint32_t* g, const Env1& e) {
    const int w = e.pos >> 1, sh = (e.pos & 1) * 16;
    const uint32_t x = g[w * B1];
    const uint32_t h = ((x >> sh) & 0xFFFFu) + 1u;
    g[w * B1] = (x & ~(0xFFFFu << sh)) | ((h & 0xFFFFu) << sh);
    return (int)h;
}

// stage (c): window [pos-2, pos+2] + counters (:131-133)
template <int B1>
__device__ __forceinline__ void window1(const uint32_t* g, const Env1& e, uint32_t& q0, uint32_t& q1, uint32_t& q2) {
    const int p0 = e.pos - D1_HW;
    const int w = p0 >> 1, sh = (p0 & 1) * 16;
    const uint32_t x0 = g[w * B1], x1 = g[(w + 1) * B1], x2 = g[(w + 2) * B1];
    q0 = __funnelshift_r(x0, x1, sh); q1 = __funnelshift_r(x1, x2, sh); q2 = x2 >> sh;     // five halfwords: q0, q1, low q2
}
template <typename ObsT, int B1>
__device__ __forceinline__ void stage_observe1(const uint32_t* g, const Env1& e, ObsT* row, bool normalise,
                                               int total_brick, int total_step) {
    uint32_t q0, q1, q2;
    window1<B1>(g, e, q0, q1, q2);
    row[0] = obs_from_int<ObsT>((int)(int16_t)(q0 & 0xFFFF));
    row[1] = obs_from_int<ObsT>((int)(int16_t)(q0 >> 16));
    row[2] = obs_from_int<ObsT>((int)(int16_t)(q1 & 0xFFFF));
    row[3] = obs_from_int<ObsT>((int)(int16_t)(q1 >> 16));
    row[4] = obs_from_int<ObsT>((int)(int16_t)(q2 & 0xFFFF));
    obs_counters<ObsT>(normalise, e.cb, e.cs, total_brick, total_step, row[5], row[6]);
}
// DMP_OBS_REC: the 16 B step record of one env (include/dmp.h); written straight to global memory, one coalesced
// 128-bit store per lane (no tile)
__device__ __forceinline__ uint4 make_rec16(uint32_t q0, uint32_t q1, uint32_t q2, int cb, int cs, float reward, bool done) {
    return make_uint4(q0, q1, (q2 & 0xFFFFu) | ((uint32_t)(cb & 0xFFFF) << 16),
                      (uint32_t)(cs & 0xFFFF) | (((uint32_t)(int)reward & 0xFFu) << 16) | (done ? (uint32_t)DMP_REC_DONE << 24 : 0u));
}

// IoU, Env/1D/DMP_Env_1D_static.py:138-151: cross = sum(min(g, p)); iou = cross / (sum p + sum g - cross)
__device__ __forceinline__ double iou1_words(const uint32_t (&hw)[15], const uint8_t* __restrict__ plan) {
    int a1 = 0, a2 = 0, cross = 0;
#pragma unroll
    for (int j = 0; j < 15; ++j) {
        const int h0 = hw[j] & 0xFFFF, h1 = hw[j] >> 16;
        const int p0 = plan[2 * j], p1 = plan[2 * j + 1];
        a1 += p0 + p1;
        a2 += h0 + h1;
        cross += min(h0, p0) + min(h1, p1);
    }
    return __ddiv_rn((double)cross, (double)(a1 + a2 - cross));
}

// The warp's [32][7] observation tile leaves through 2 x (LDS.128 + STG.128) per lane.  (A bulk async copy out of a ring of
// tiles was measured 20 % slower -- the 896 B tile is too small to pay for fence + elected-lane issue -- and is gone.)
// FAST = true: the throughput configuration with everything that is uniform over a launch folded at compile time --
// in-kernel Philox draws, observations / rewards / done flags all materialised, auto-reset, raw counters, full warps
// (n a multiple of the block size), 16 B aligned observation buffer.  The launcher checks those conditions.
// RF = true (DMP_F_RESET_OBS): a finished env is reset BEFORE its observation is cut (the row it writes is the next episode's
// first policy input); a template parameter so that each instantiation carries one copy of the reset block.
template <typename ObsT, int B1, bool FAST, bool RF>
__global__ void __launch_bounds__(B1) k1d_rollout(const DmpState st, const DmpIO io, const int K) {
    constexpr bool REC = is_rec<ObsT>::value;
    constexpr int ROW = row_elems<ObsT, D1_OBS>();
    extern __shared__ uint4 smem_raw[];
    uint32_t* G = reinterpret_cast<uint32_t*>(smem_raw);                 // [S1_WORDS][B1]
    ObsT* tiles = reinterpret_cast<ObsT*>(G + S1_WORDS * B1);            // [B1/32][32*7] (no tile for records)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * B1 + tid;
    const int64_t env0 = env - lane;
    const int nvalid = FAST ? 32 : (int)min((int64_t)32, n - env0);
    const bool live = FAST ? true : (env < n);
    ObsT* tile = tiles + warp * (32 * ROW);
    uint32_t* g = G + tid;

    uint4* cells = reinterpret_cast<uint4*>(st.cells);
    uint2* aux = reinterpret_cast<uint2*>(st.aux);
    const uint8_t* __restrict__ plans = reinterpret_cast<const uint8_t*>(st.plans);
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0, v3 = v0;
    uint2 ax = make_uint2(D1_LO, 0);
    if (live) {
        v0 = cells[env]; v1 = cells[n + env]; v2 = cells[2 * n + env]; v3 = cells[3 * n + env];
        ax = aux[env];
    }
    g[0] = 0xFFFFFFFFu;
    g[1 * B1] = v0.x;  g[2 * B1] = v0.y;  g[3 * B1] = v0.z;  g[4 * B1] = v0.w;
    g[5 * B1] = v1.x;  g[6 * B1] = v1.y;  g[7 * B1] = v1.z;  g[8 * B1] = v1.w;
    g[9 * B1] = v2.x;  g[10 * B1] = v2.y; g[11 * B1] = v2.z; g[12 * B1] = v2.w;
    g[13 * B1] = v3.x; g[14 * B1] = v3.y; g[15 * B1] = v3.z;
    g[16 * B1] = 0xFFFFFFFFu;
    Env1 e;
    e.pos = ax.x & 0xFFFF; e.plan_idx = ax.x >> 16; e.ret = __uint_as_float(ax.y);
    e.cb = v3.w & 0xFFFF; e.cs = v3.w >> 16;
    int total_brick = __ldg(st.plan_total + e.plan_idx);
    unsigned dirty = 0;
    int errbits = 0;

    const bool autoreset = FAST ? true : (bool)(io.flags & DMP_F_AUTORESET);
    const bool normalise = FAST ? false : (bool)(io.flags & DMP_F_NORMALISE);
    const uint8_t* act_in = FAST ? nullptr : io.actions;
    const uint8_t* size_in = FAST ? nullptr : io.step_sizes;
    const bool need_draw = FAST ? true : ((act_in == nullptr) || (size_in == nullptr));
    const bool want_obs = FAST ? true : (io.obs != nullptr);
    const bool want_rew = FAST ? true : (io.reward != nullptr);
    const bool want_done = FAST ? true : (io.done != nullptr);
    const int tslot = (io.flags & DMP_F_TSLOT1) ? 1 : 0;
    const uint64_t t0 = st.t_dev ? st.t_dev[tslot] : st.t;

    StepDraws draws;
    const uint64_t gid = (uint64_t)(st.env_base + env);
    int64_t idx = env;                                      // flat [k][env] index of this step's inputs / outputs
    ObsT* dst = reinterpret_cast<ObsT*>(io.obs) + env0 * ROW;             // this step's 32 observation rows / records
    const int64_t dst_step = n * ROW;
    for (int k = 0; k < K; ++k, idx += n, dst += dst_step) {
        const uint64_t t = t0 + (uint64_t)k;
        uint32_t dw = 0;
        if (need_draw) dw = draws.word(st.seed, gid, t);
        int a, s;
        if (act_in) a = live ? act_in[idx] : 0; else a = draw_action(dw, D1_ACT, DMP_ACT_UNIFORM);
        if (size_in) {
            s = live ? size_in[idx] : 1;
            if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;     // a drawn step size is always 1..3
        } else {
            s = draw_step_size(dw);
        }

        // ---- step(): Env/1D/DMP_Env_1D_static.py:85-136 ---------------------------------------
        e.cs += 1;
        bool done, rewarded = false;
        int h = 0, p = 0;
        if (a < 2) {                                        // (a) move left / right
            stage_move1(e, a, s);
            done = e.cs >= st.total_step;
        } else if (a == 2) {                                // (b) drop
            e.cb += 1;
            h = stage_deposit1<B1>(g, e);
            dirty |= 1u << ((e.pos - D1_HW) >> 3);          // 8 heights per 128-bit plane
            if (e.cb >= total_brick) {                      // :107-114
                done = true;
            } else {                                        // :116-123; the plan byte is requested here and consumed after
                done = e.cs >= st.total_step;               // the observation, which hides the load's latency (it was the
                p = ldg_u8(plans + e.plan_idx * PLAN1D_BYTES + (e.pos - D1_HW));      // kernel's largest single stall)
                rewarded = true;
            }
        } else {
            errbits |= DMP_ERR_ACTION;
            done = e.cs >= st.total_step;
        }

        // (e) done / auto-reset: fold the episode into the statistics, clear the state, next plan
        const bool fin = done && autoreset && live;
        auto finish = [&]() {
            uint32_t hw[15];
#pragma unroll
            for (int j = 0; j < 15; ++j) hw[j] = g[(j + 1) * B1];
            const double iou = iou1_words(hw, plans + e.plan_idx * PLAN1D_BYTES);
            atomicAdd(st.ep_cnt + env, 1u);                 // fire-and-forget REDs: no read-modify-write stall
            atomicAdd(st.ep_len + env, (uint32_t)e.cs);
            atomicAdd(st.ep_ret + env, (double)e.ret);
            atomicAdd(st.ep_iou + env, iou);
            if (io.next_plan) {
                const int p = io.next_plan[idx];
                if ((unsigned)p >= (unsigned)st.n_plans) errbits |= DMP_ERR_PLANIDX; else e.plan_idx = p;
            } else if (st.plan_mode == DMP_PLAN_PHILOX) {
                e.plan_idx = draw_plan(plan_word(st.seed, gid, t), st.n_plans);
            } else if (st.plan_mode == DMP_PLAN_SEQUENTIAL) {
                e.plan_idx = (e.plan_idx + 1 == st.n_plans) ? 0 : e.plan_idx + 1;
            }
            total_brick = __ldg(st.plan_total + e.plan_idx);
#pragma unroll
            for (int j = 1; j <= 15; ++j) g[j * B1] = 0;
            e.pos = D1_LO; e.cb = e.cs = 0; e.ret = 0.f;
            dirty = 0xFu;
        };
        // (d) reward: -1 / 10 / 1 for a column above / at / below its plan height (:117-123).  Normally evaluated after the
        // observation (the plan byte's latency hides behind it); DMP_F_RESET_OBS needs it before the reset
        float reward = 0.f;
        if constexpr (RF) {
            reward = rewarded ? ((h > p) ? -1.f : (h == p ? 10.f : 1.f)) : 0.f;
            e.ret += reward;
            if (fin) finish();                              // the observation below is the reset env's
        }

        uint32_t q0 = 0, q1 = 0, q2 = 0;
        if (want_obs) {
            if constexpr (REC) {
                window1<B1>(g, e, q0, q1, q2);                  // the record is assembled once the reward is known
            } else {
                stage_observe1<ObsT, B1>(g, e, tile + lane * D1_OBS, normalise, total_brick, st.total_step);
                __syncwarp();
                if constexpr (FAST) warp_tile_store_aligned<ObsT, 32 * D1_OBS>(dst, tile, lane);
                else if (nvalid == 32) warp_tile_store_full<ObsT, 32 * D1_OBS>(dst, tile, lane);
                else if (nvalid > 0) warp_tile_store<ObsT>(dst, tile, nvalid * D1_OBS, lane);
                __syncwarp();
            }
        }
        if constexpr (!RF) {
            reward = rewarded ? ((h > p) ? -1.f : (h == p ? 10.f : 1.f)) : 0.f;
            e.ret += reward;
        }
        if (live) {
            if constexpr (REC) {
                if (want_obs) __stcs(reinterpret_cast<uint4*>(dst) + lane, make_rec16(q0, q1, q2, e.cb, e.cs, reward, done));
            }
            if (want_rew) io.reward[idx] = reward;
            if (want_done) io.done[idx] = done ? 1 : 0;
        }

        if constexpr (!RF) { if (fin) finish(); }
    }

    if (live) {
        if ((e.cb | e.cs) > 0xFFFF) {                                    // 16-bit packed counters (include/dmp.h)
            errbits |= DMP_ERR_OVERFLOW;
            e.cb = min(e.cb, 0xFFFF); e.cs = min(e.cs, 0xFFFF);
        }
        if (dirty & 1u) cells[env] = make_uint4(g[1 * B1], g[2 * B1], g[3 * B1], g[4 * B1]);
        if (dirty & 2u) cells[n + env] = make_uint4(g[5 * B1], g[6 * B1], g[7 * B1], g[8 * B1]);
        if (dirty & 4u) cells[2 * n + env] = make_uint4(g[9 * B1], g[10 * B1], g[11 * B1], g[12 * B1]);
        cells[3 * n + env] = make_uint4(g[13 * B1], g[14 * B1], g[15 * B1],
                                        (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16));
        aux[env] = make_uint2((uint32_t)e.pos | ((uint32_t)e.plan_idx << 16), __float_as_uint(e.ret));
        if (errbits) atomicOr(st.err, errbits);
    }
    if (st.t_dev && blockIdx.x == 0 && tid == 0) st.t_dev[tslot ^ 1] = t0 + (uint64_t)K;
}

template <typename ObsT>
__global__ void k1d_reset(const DmpState st, const uint8_t* __restrict__ mask, const int32_t* __restrict__ plan_idx,
                          const uint64_t t_draw, ObsT* __restrict__ obs) {
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n || (mask && !mask[env])) return;
    uint4* cells = reinterpret_cast<uint4*>(st.cells);
    uint2* aux = reinterpret_cast<uint2*>(st.aux);
    int p;
    if (plan_idx) {
        p = plan_idx[env];
        if ((unsigned)p >= (unsigned)st.n_plans) { atomicOr(st.err, DMP_ERR_PLANIDX); p = 0; }
    } else if (st.plan_mode == DMP_PLAN_PHILOX) {
        p = draw_plan(env_draw(st.seed, (uint64_t)(st.env_base + env), t_draw).x3, st.n_plans);
    } else {
        const uint32_t ax = aux[env].x;
        p = (int)(ax >> 16);
        // sequential order starts at plan 0 on an env that has never been reset (zeroed state: 0 is not a position),
        // index_for_non_random = 0 of Env/1D/DMP_Env_1D_dynamic_usedata_plan.py:46-50
        if (st.plan_mode == DMP_PLAN_SEQUENTIAL) p = ((ax & 0xFFFFu) == 0u) ? 0 : ((p + 1 >= st.n_plans) ? 0 : p + 1);
        if ((unsigned)p >= (unsigned)st.n_plans) p = 0;
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
    cells[env] = z; cells[n + env] = z; cells[2 * n + env] = z; cells[3 * n + env] = z;
    aux[env] = make_uint2((uint32_t)D1_LO | ((uint32_t)p << 16), 0u);
    if (obs) {                       // window at pos 2: [-1, -1, 0, 0, 0], counters 0 (:81-83)
        if constexpr (is_rec<ObsT>::value) {
            reinterpret_cast<uint4*>(obs)[env] = make_rec16(0xFFFFFFFFu, 0u, 0u, 0, 0, 0.f, false);
        } else {
            ObsT* o = obs + env * D1_OBS;
            o[0] = obs_from_int<ObsT>(-1); o[1] = obs_from_int<ObsT>(-1);
            for (int j = 2; j < 7; ++j) o[j] = obs_from_int<ObsT>(0);
        }
    }
}

__device__ __forceinline__ uint32_t cell_word1(const DmpState& st, int64_t env, int w) {
    return reinterpret_cast<const uint32_t*>(st.cells)[((int64_t)(w >> 2) * st.n_envs + env) * 4 + (w & 3)];
}

__global__ void k1d_iou(const DmpState st, double* __restrict__ out) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= st.n_envs) return;
    uint32_t hw[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) hw[j] = cell_word1(st, env, j);
    const int p = (int)(reinterpret_cast<const uint2*>(st.aux)[env].x >> 16);
    out[env] = iou1_words(hw, reinterpret_cast<const uint8_t*>(st.plans) + p * PLAN1D_BYTES);
}

__global__ void k1d_export(const DmpState st, int32_t* __restrict__ grid, int32_t* __restrict__ scalars, float* __restrict__ ret) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= st.n_envs) return;
    if (grid) {
        int32_t* g = grid + env * 34;
        g[0] = g[1] = g[32] = g[33] = -1;
        for (int j = 0; j < 15; ++j) {
            const uint32_t x = cell_word1(st, env, j);
            g[2 + 2 * j] = x & 0xFFFF;
            g[3 + 2 * j] = x >> 16;
        }
    }
    const uint2 ax = reinterpret_cast<const uint2*>(st.aux)[env];
    if (scalars) {
        int32_t* s = scalars + env * 8;
        const uint32_t c = cell_word1(st, env, 15);
        const int p = ax.x >> 16;
        s[0] = ax.x & 0xFFFF; s[1] = 0; s[2] = c & 0xFFFF; s[3] = c >> 16; s[4] = p;
        s[5] = st.plan_total[p]; s[6] = 0; s[7] = 0;
    }
    if (ret) ret[env] = __uint_as_float(ax.y);
}

__global__ void k1d_import(const DmpState st, const int32_t* __restrict__ grid, const int32_t* __restrict__ scalars,
                           const float* __restrict__ ret) {
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    uint32_t* cw = reinterpret_cast<uint32_t*>(st.cells);
    uint2* aux = reinterpret_cast<uint2*>(st.aux);
    auto wref = [&](int w) -> uint32_t& { return cw[((int64_t)(w >> 2) * n + env) * 4 + (w & 3)]; };
    if (grid) {
        const int32_t* g = grid + env * 34;
        for (int j = 0; j < 15; ++j) wref(j) = (uint32_t)(g[2 + 2 * j] & 0xFFFF) | ((uint32_t)(g[3 + 2 * j] & 0xFFFF) << 16);
    }
    uint2 ax = aux[env];
    if (scalars) {
        const int32_t* s = scalars + env * 8;
        wref(15) = (uint32_t)(s[2] & 0xFFFF) | ((uint32_t)s[3] << 16);
        ax.x = (uint32_t)s[0] | ((uint32_t)s[4] << 16);
    }
    if (ret) ax.y = __float_as_uint(ret[env]);
    aux[env] = ax;
}

template <typename ObsT, int B1, bool FAST, bool RF>
int launch_rollout1_r(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    const size_t smem = (size_t)S1_WORDS * B1 * 4 + (is_rec<ObsT>::value ? 0 : (size_t)(B1 / 32) * 32 * D1_OBS * sizeof(ObsT));
    const unsigned blocks = (unsigned)((st.n_envs + B1 - 1) / B1);
    static bool attr_done = false;           // per instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k1d_rollout<ObsT, B1, FAST, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return dmp_set_error(e);
        attr_done = true;
    }
    // plain stream-ordered launch: with programmatic dependent launch the next grid's single-warp blocks become
    // resident early and unbalance the SMs (measured: 47 vs 61 G env-steps/s at 65 536 envs, K = 16)
    k1d_rollout<ObsT, B1, FAST, RF><<<blocks, B1, smem, s>>>(st, io, K);
    return dmp_set_error(cudaGetLastError());
}

template <typename ObsT, int B1, bool FAST>
int launch_rollout1_k(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    if (io.flags & DMP_F_RESET_OBS) return launch_rollout1_r<ObsT, B1, FAST, true>(st, io, K, s);
    return launch_rollout1_r<ObsT, B1, FAST, false>(st, io, K, s);
}

template <typename ObsT, int B1>
int launch_rollout1_b(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    // the throughput configuration (see k1d_rollout, FAST) gets the kernel with its launch-uniform branches folded
    const bool fast = !(io.flags & DMP_F_GENERIC) && !io.actions && !io.step_sizes && io.obs && io.reward && io.done &&
                      (io.flags & DMP_F_AUTORESET) && !(io.flags & DMP_F_NORMALISE) && st.n_envs % B1 == 0 &&
                      (reinterpret_cast<uintptr_t>(io.obs) & 15) == 0 && (st.n_envs * (int64_t)D1_OBS * (int64_t)sizeof(ObsT)) % 16 == 0;
    if (fast) return launch_rollout1_k<ObsT, B1, true>(st, io, K, s);
    return launch_rollout1_k<ObsT, B1, false>(st, io, K, s);
}

template <typename ObsT>
int launch_rollout1(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    // fewer than ~4 blocks of 128 per SM: single-warp blocks spread the warps evenly over the 148 SMs
    const bool small = st.n_envs < (int64_t)128 * 148 * 4;
    return small ? launch_rollout1_b<ObsT, 32>(st, io, K, s) : launch_rollout1_b<ObsT, 128>(st, io, K, s);
}

}  // namespace

int dmp1d_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    switch (io.obs_kind) {
        case DMP_OBS_F32: return launch_rollout1<float>(st, io, K, s);
        case DMP_OBS_F64: return launch_rollout1<double>(st, io, K, s);
        case DMP_OBS_I16: return launch_rollout1<int16_t>(st, io, K, s);
        case DMP_OBS_REC: return launch_rollout1<Rec16>(st, io, K, s);
    }
    return DMP_EINVAL;
}

int dmp1d_reset(const DmpState& st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs,
                int obs_kind, cudaStream_t s) {
    const unsigned blocks = (unsigned)((st.n_envs + 255) / 256);
    switch (obs_kind) {
        case DMP_OBS_F32: k1d_reset<float><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (float*)obs); break;
        case DMP_OBS_F64: k1d_reset<double><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (double*)obs); break;
        case DMP_OBS_I16: k1d_reset<int16_t><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (int16_t*)obs); break;
        case DMP_OBS_REC: k1d_reset<Rec16><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (Rec16*)obs); break;
        default: return DMP_EINVAL;
    }
    return dmp_set_error(cudaGetLastError());
}

int dmp1d_iou(const DmpState& st, double* out, cudaStream_t s) {
    k1d_iou<<<(unsigned)((st.n_envs + 255) / 256), 256, 0, s>>>(st, out);
    return dmp_set_error(cudaGetLastError());
}
int dmp1d_export(const DmpState& st, int32_t* grid, int32_t* scalars, float* ret, cudaStream_t s) {
    k1d_export<<<(unsigned)((st.n_envs + 127) / 128), 128, 0, s>>>(st, grid, scalars, ret);
    return dmp_set_error(cudaGetLastError());
}
int dmp1d_import(const DmpState& st, const int32_t* grid, const int32_t* scalars, const float* ret, cudaStream_t s) {
    k1d_import<<<(unsigned)((st.n_envs + 127) / 128), 128, 0, s>>>(st, grid, scalars, ret);
    return dmp_set_error(cudaGetLastError());
}

// dmp_stages.cu -- the five stages of a step as standalone kernels (include/dmp.h, "stage kernels").
//
// The fused rollout kernels (dmp_1d.cu / dmp_2d.cu / dmp_3d.cu) keep an env on chip for the whole
// step; here every stage is its own launch that reads and writes the SoA state in HBM, so each of
// (a) move, (b) deposit, (c) observe, (d) reward, (e) done/reset can be checked against the oracle
// and timed in isolation.  Running a -> b -> c -> d -> e equals one fused step.  One env per thread;
// (c) stages the observation tile of a warp in shared memory and writes it out with 128-bit stores.
//
// scratch: i32 [n][4] = { flags, new cell value at the target, target row, target col }
#include "dmp_common.cuh"

namespace {

constexpr int SB = 128;
enum : int { F_PLACED = 1, F_BOXED_PRE = 2, F_BOXED_POST = 4, F_BAD_ACTION = 8, F_WAS_OCC = 16 };

struct Sc { int pr, pc, plan_idx, cb, cs; float ret; int cross; };   // cross: 3D running sum(min(h, plan))

__device__ __forceinline__ uint32_t& word_ref(const DmpState& st, int64_t env, int w) {
    return reinterpret_cast<uint32_t*>(st.cells)[((int64_t)(w >> 2) * st.n_envs + env) * 4 + (w & 3)];
}

__device__ void load_sc(const DmpState& st, int64_t env, Sc& s) {
    if (st.dim == 1) {
        const uint2 ax = reinterpret_cast<const uint2*>(st.aux)[env];
        const uint32_t c = word_ref(st, env, 15);
        s = Sc{(int)(ax.x & 0xFFFF), 0, (int)(ax.x >> 16), (int)(c & 0xFFFF), (int)(c >> 16), __uint_as_float(ax.y), 0};
    } else if (st.dim == 2) {
        const uint32_t w13 = word_ref(st, env, 13), w14 = word_ref(st, env, 14);
        s = Sc{(int)(w13 & 0xFF), (int)((w13 >> 8) & 0xFF), (int)(w13 >> 16), (int)(w14 & 0xFFFF), (int)(w14 >> 16),
               __uint_as_float(word_ref(st, env, 15)), 0};
    } else {
        const uint4 a = reinterpret_cast<const uint4*>(st.aux)[env];
        s = Sc{(int)(a.x & 0xFF), (int)((a.x >> 8) & 0xFF), (int)(a.x >> 16), (int)(a.y & 0xFFFF), (int)(a.y >> 16),
               __uint_as_float(a.z), (int)a.w};
    }
}

__device__ void store_sc(const DmpState& st, int64_t env, const Sc& s) {
    if (st.dim == 1) {
        reinterpret_cast<uint2*>(st.aux)[env] = make_uint2((uint32_t)s.pr | ((uint32_t)s.plan_idx << 16), __float_as_uint(s.ret));
        word_ref(st, env, 15) = (uint32_t)(s.cb & 0xFFFF) | ((uint32_t)s.cs << 16);
    } else if (st.dim == 2) {
        word_ref(st, env, 13) = (uint32_t)s.pr | ((uint32_t)s.pc << 8) | ((uint32_t)s.plan_idx << 16);
        word_ref(st, env, 14) = (uint32_t)(s.cb & 0xFFFF) | ((uint32_t)s.cs << 16);
        word_ref(st, env, 15) = __float_as_uint(s.ret);
    } else {
        reinterpret_cast<uint4*>(st.aux)[env] = make_uint4((uint32_t)s.pr | ((uint32_t)s.pc << 8) | ((uint32_t)s.plan_idx << 16),
                                                          (uint32_t)(s.cb & 0xFFFF) | ((uint32_t)s.cs << 16), __float_as_uint(s.ret), (uint32_t)s.cross);
    }
}

// environment_memory[r][c] in the reference's padded coordinates (1D: r ignored)
__device__ int get_cell(const DmpState& st, int64_t env, int r, int c) {
    if (st.dim == 1) {
        const unsigned i = (unsigned)(c - 2);
        if (i >= 30u) return -1;
        return (word_ref(st, env, i >> 1) >> ((i & 1) * 16)) & 0xFFFF;
    }
    const unsigned ir = (unsigned)(r - 3), ic = (unsigned)(c - 3);
    if (ir >= 20u || ic >= 20u) return -1;
    if (st.dim == 2) {
        const unsigned b = ir * 20 + ic;
        return (word_ref(st, env, b >> 5) >> (b & 31)) & 1;
    }
    return reinterpret_cast<const uint16_t*>(st.cells)[env * CELLS3D + ir * 20 + ic];
}

__device__ void set_cell(const DmpState& st, int64_t env, int r, int c, int v) {
    if (st.dim == 1) {
        const unsigned i = (unsigned)(c - 2);
        uint32_t& w = word_ref(st, env, i >> 1);
        const int sh = (i & 1) * 16;
        w = (w & ~(0xFFFFu << sh)) | ((uint32_t)(v & 0xFFFF) << sh);
    } else if (st.dim == 2) {
        const unsigned b = (unsigned)(r - 3) * 20 + (unsigned)(c - 3);
        uint32_t& w = word_ref(st, env, b >> 5);
        w = v ? (w | (1u << (b & 31))) : (w & ~(1u << (b & 31)));
    } else {
        reinterpret_cast<uint16_t*>(st.cells)[env * CELLS3D + (r - 3) * 20 + (c - 3)] = (uint16_t)v;
    }
}

__device__ int plan_cell(const DmpState& st, int plan_idx, int r, int c) {
    if (st.dim == 1) return reinterpret_cast<const uint8_t*>(st.plans)[plan_idx * PLAN1D_BYTES + (c - 2)];
    const int b = (r - 3) * 20 + (c - 3);
    if (st.dim == 2) return (reinterpret_cast<const uint32_t*>(st.plans)[plan_idx * PLAN2D_WORDS + (b >> 5)] >> (b & 31)) & 1;
    return reinterpret_cast<const uint8_t*>(st.plans)[plan_idx * CELLS3D + b];
}

__device__ __forceinline__ void step_inputs(const DmpState& st, const DmpIO& io, int64_t env, int& a, int& s) {
    const int A = st.dim == 1 ? D1_ACT : (st.dim == 2 ? D2_ACT : D3_ACT);
    StepDraws draws;
    uint32_t dw = 0;
    if (!io.actions || !io.step_sizes) dw = draws.word(st.seed, (uint64_t)(st.env_base + env), st.t);
    a = io.actions ? (int)io.actions[env] : draw_action(dw, A, st.dim == 3 ? st.action_dist : DMP_ACT_UNIFORM);
    s = io.step_sizes ? (int)io.step_sizes[env] : draw_step_size(dw);
}

__device__ __forceinline__ int dr_of(int d) { return d == 2 ? 1 : (d == 3 ? -1 : 0); }
__device__ __forceinline__ int dc_of(int d) { return d == 0 ? -1 : (d == 1 ? 1 : 0); }

// ---- (a) move: count_step += 1, position update with clamp (3D: collision walk) ------------------
__global__ void k_stage_move(const DmpState st, const DmpIO io, int32_t* __restrict__ scratch) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= st.n_envs) return;
    int a, s;
    step_inputs(st, io, env, a, s);
    Sc e;
    load_sc(st, env, e);
    e.cs += 1;
    int flags = 0;
    if (st.dim == 1) {
        if (a < 2) e.pr = min(max(a == 0 ? e.pr - s : e.pr + s, D1_LO), D1_HI);
        else if (a != 2) flags |= F_BAD_ACTION;
    } else if (st.dim == 2) {
        if (a < 4) {
            int r = e.pr, c = e.pc;
            if (a == 0) c -= s; else if (a == 1) c += s; else if (a == 2) r += s; else r -= s;
            e.pr = min(max(r, D2_LO), D2_HI);
            e.pc = min(max(c, D2_LO), D2_HI);
        } else if (a != 4) flags |= F_BAD_ACTION;
    } else {
        bool boxed = true;
        for (int q = 0; q < 4; ++q)
            if (get_cell(st, env, e.pr + dr_of(q), e.pc + dc_of(q)) == 0) boxed = false;
        if (boxed) flags |= F_BOXED_PRE;
        if (a <= 3 && get_cell(st, env, e.pr + dr_of(a), e.pc + dc_of(a)) == 0) {
            int nstep = 0;
            for (int i = 1; i <= s && i <= 3; ++i) {
                if (get_cell(st, env, e.pr + dr_of(a) * i, e.pc + dc_of(a) * i) == 0) nstep++; else break;
            }
            e.pr = min(max(e.pr + dr_of(a) * nstep, D2_LO), D2_HI);
            e.pc = min(max(e.pc + dc_of(a) * nstep, D2_LO), D2_HI);
        }
        if (a > 7) flags |= F_BAD_ACTION;
    }
    store_sc(st, env, e);
    int32_t* sc = scratch + env * 4;
    sc[0] = flags; sc[1] = 0; sc[2] = 0; sc[3] = 0;
    if (flags & F_BAD_ACTION) atomicOr(st.err, DMP_ERR_ACTION);
}

// ---- (b) deposit: brick into the grid, count_brick += 1 --------------------------------------------
__global__ void k_stage_deposit(const DmpState st, const DmpIO io, int32_t* __restrict__ scratch) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= st.n_envs) return;
    int a, s;
    step_inputs(st, io, env, a, s);
    Sc e;
    load_sc(st, env, e);
    int32_t* sc = scratch + env * 4;
    int flags = sc[0];
    if (st.dim == 1) {
        if (a == 2) {
            const int h = get_cell(st, env, 0, e.pr) + 1;
            set_cell(st, env, 0, e.pr, h);
            e.cb += 1;
            flags |= F_PLACED; sc[1] = h; sc[2] = 0; sc[3] = e.pr;
        }
    } else if (st.dim == 2) {
        if (a == 4) {
            const int was = get_cell(st, env, e.pr, e.pc);
            set_cell(st, env, e.pr, e.pc, 1);               // += 1 then clip to 1
            e.cb += 1;
            flags |= F_PLACED | (was ? F_WAS_OCC : 0); sc[1] = was + 1; sc[2] = e.pr; sc[3] = e.pc;
        }
    } else {
        if (a >= 4 && a <= 7) {
            const int q = a - 4;
            const int tr = e.pr + dr_of(q), tc = e.pc + dc_of(q);
            const int v = get_cell(st, env, tr, tc);
            if (v != -1) {
                set_cell(st, env, tr, tc, v + 1);
                e.cb += 1;
                if (v + 1 <= plan_cell(st, e.plan_idx, tr, tc)) e.cross += 1;
                flags |= F_PLACED; sc[1] = v + 1; sc[2] = tr; sc[3] = tc;
            }
        }
        if (a > 3 && st.dynamic) {                           // re-check after the (attempted) placement
            bool boxed = true;
            for (int q = 0; q < 4; ++q)
                if (get_cell(st, env, e.pr + dr_of(q), e.pc + dc_of(q)) == 0) boxed = false;
            if (boxed) flags |= F_BOXED_POST;
        }
    }
    sc[0] = flags;
    store_sc(st, env, e);
}

// ---- (c) observe: window gather -> warp tile in shared memory -> 128-bit stores ---------------------
template <typename ObsT>
__global__ void __launch_bounds__(SB) k_stage_observe(const DmpState st, const DmpIO io) {
    extern __shared__ uint4 smem_raw[];
    ObsT* tiles = reinterpret_cast<ObsT*>(smem_raw);
    const int D = st.dim == 1 ? D1_OBS : D2_OBS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t env = (int64_t)blockIdx.x * SB + threadIdx.x;
    const int64_t env0 = env - lane;
    const int nvalid = (int)min((int64_t)32, st.n_envs - env0);
    ObsT* row = tiles + (warp * 32 + lane) * D;
    if (env < st.n_envs) {
        Sc e;
        load_sc(st, env, e);
        const int tb = st.plan_total[e.plan_idx];
        if (st.dim == 1) {
            for (int j = 0; j < 5; ++j) row[j] = obs_from_int<ObsT>(get_cell(st, env, 0, e.pr - 2 + j));
        } else {
            for (int k = 0; k < 7; ++k)
                for (int j = 0; j < 7; ++j) row[k * 7 + j] = obs_from_int<ObsT>(get_cell(st, env, e.pr - 3 + k, e.pc - 3 + j));
        }
        obs_counters<ObsT>(io.flags & DMP_F_NORMALISE, e.cb, e.cs, tb, st.total_step, row[D - 2], row[D - 1]);
    }
    __syncwarp();
    if (nvalid > 0)
        warp_tile_store<ObsT>(reinterpret_cast<ObsT*>(io.obs) + env0 * D, tiles + warp * 32 * D, nvalid * D, lane);
}

// ---- (d) reward + done ----------------------------------------------------------------------------------
__global__ void k_stage_reward(const DmpState st, const DmpIO io, const int32_t* __restrict__ scratch) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= st.n_envs) return;
    int a, s;
    step_inputs(st, io, env, a, s);
    Sc e;
    load_sc(st, env, e);
    const int32_t* sc = scratch + env * 4;
    const int flags = sc[0];
    const int tb = st.plan_total[e.plan_idx];
    float reward = 0.f;
    bool done = e.cs >= st.total_step;
    if (st.dim == 1) {
        if (flags & F_PLACED) {
            if (e.cb >= tb) done = true;
            else { const int p = plan_cell(st, e.plan_idx, 0, sc[3]); reward = sc[1] > p ? -1.f : (sc[1] == p ? 10.f : 1.f); }
        }
    } else if (st.dim == 2) {
        if (flags & F_PLACED) {
            if (e.cb >= tb) done = true;
            else reward = (!(flags & F_WAS_OCC) && plan_cell(st, e.plan_idx, sc[2], sc[3])) ? 5.f : 0.f;
        }
    } else {
        const bool boxed = flags & F_BOXED_PRE;
        bool tail = true;
        if (a > 3) {
            if (st.dynamic) {
                if (flags & F_BOXED_POST) { reward = -100.f; done = true; tail = false; }
                else if (e.cb >= tb) { done = true; tail = false; }
                else if (flags & F_PLACED) { done = false; tail = false; }
            } else {
                if (e.cb >= tb || boxed) { done = true; tail = false; }
                else if (flags & F_PLACED) { done = false; tail = false; }
            }
            if (!tail && !done) { const int p = plan_cell(st, e.plan_idx, sc[2], sc[3]); reward = sc[1] > p ? -1.f : (sc[1] == p ? 10.f : 1.f); }
        }
        if (tail) done = (e.cs >= st.total_step) || (!st.dynamic && boxed);
    }
    e.ret += reward;
    store_sc(st, env, e);
    if (io.reward) io.reward[env] = reward;
    if (io.done) io.done[env] = done ? 1 : 0;
}

// ---- (e) done / reset: IoU of the finished episode -> ep_*, then reset ------------------------------
__device__ double iou_generic(const DmpState& st, int64_t env, const Sc& e) {
    if (st.dim == 1) {
        int a1 = 0, a2 = 0, cross = 0;
        for (int c = 2; c < 32; ++c) {
            const int g = get_cell(st, env, 0, c), p = plan_cell(st, e.plan_idx, 0, c);
            a1 += p; a2 += g; cross += min(g, p);
        }
        return __ddiv_rn((double)cross, (double)(a1 + a2 - cross));
    }
    int x = 0, y = 0;
    for (int r = 3; r < 23; ++r)
        for (int c = 3; c < 23; ++c) {
            const int g = get_cell(st, env, r, c), p = plan_cell(st, e.plan_idx, r, c);
            if (st.dim == 2) { x += (g && p); y += (g || p); }
            else x += min(g, p);
        }
    if (st.dim == 2) return __ddiv_rn((double)x, (double)y);
    return __ddiv_rn((double)x, (double)(st.plan_total[e.plan_idx] + e.cb - x));
}

__global__ void k_stage_done_reset(const DmpState st, const DmpIO io) {
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= st.n_envs) return;
    if (!io.done[env]) return;
    Sc e;
    load_sc(st, env, e);
    st.ep_cnt[env] += 1;
    st.ep_len[env] += (uint32_t)e.cs;
    st.ep_ret[env] += (double)e.ret;
    st.ep_iou[env] += iou_generic(st, env, e);
    if (io.next_plan) e.plan_idx = io.next_plan[env];
    else if (st.plan_mode == DMP_PLAN_PHILOX) e.plan_idx = draw_plan(plan_word(st.seed, (uint64_t)(st.env_base + env), st.t), st.n_plans);
    else if (st.plan_mode == DMP_PLAN_SEQUENTIAL) e.plan_idx = (e.plan_idx + 1 == st.n_plans) ? 0 : e.plan_idx + 1;
    if (st.dim == 3) {
        uint4* g4 = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(st.cells) + env * CELLS3D);
        for (int i = 0; i < 50; ++i) g4[i] = make_uint4(0, 0, 0, 0);
    } else {
        for (int w = 0; w < (st.dim == 1 ? 15 : 13); ++w) word_ref(st, env, w) = 0;
    }
    e.pr = st.dim == 1 ? D1_LO : D2_LO;
    e.pc = st.dim == 1 ? 0 : D2_LO;
    e.cb = e.cs = 0;
    e.ret = 0.f;
    e.cross = 0;
    store_sc(st, env, e);
}

bool stage_args_ok(const DmpState* st, const DmpIO* io, const int32_t* scratch) {
    if (!st || !io || !scratch) return false;
    if (st->dim < 1 || st->dim > 3 || st->n_envs < 1 || !st->cells || !st->plans || !st->plan_total || !st->err) return false;
    if (st->dim != 2 && !st->aux) return false;
    return true;
}
inline unsigned sblocks(int64_t n) { return (unsigned)((n + SB - 1) / SB); }

// 3D: the stage kernels read and write the wide (u16) height maps only (dmp_common.cuh): make them current before a
// stage and fold them back into the nibble maps / tall flags after it
int stage_enter(const DmpState* st, void* stream) {
    return st->dim == 3 ? dmp3d_widen(*st, true, as_stream(stream)) : DMP_OK;
}
int stage_leave(const DmpState* st, void* stream) {
    const int rc = dmp_set_error(cudaGetLastError());
    return (rc != DMP_OK || st->dim != 3) ? rc : dmp3d_sync_bytes(*st, as_stream(stream));
}

}  // namespace

extern "C" {

int dmp_stage_move(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream) {
    if (!stage_args_ok(st, io, scratch)) return DMP_EINVAL;
    if (const int rc = stage_enter(st, stream)) return rc;
    k_stage_move<<<sblocks(st->n_envs), SB, 0, as_stream(stream)>>>(*st, *io, scratch);
    return stage_leave(st, stream);
}

int dmp_stage_deposit(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream) {
    if (!stage_args_ok(st, io, scratch)) return DMP_EINVAL;
    if (const int rc = stage_enter(st, stream)) return rc;
    k_stage_deposit<<<sblocks(st->n_envs), SB, 0, as_stream(stream)>>>(*st, *io, scratch);
    return stage_leave(st, stream);
}

int dmp_stage_observe(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream) {
    if (!stage_args_ok(st, io, scratch) || !io->obs) return DMP_EINVAL;
    if (io->obs_kind != DMP_OBS_F32 && io->obs_kind != DMP_OBS_F64 && io->obs_kind != DMP_OBS_I16) return DMP_EINVAL;
    if (io->obs_kind == DMP_OBS_I16 && (io->flags & DMP_F_NORMALISE)) return DMP_EINVAL;
    if (const int rc = stage_enter(st, stream)) return rc;
    const int D = st->dim == 1 ? D1_OBS : D2_OBS;
    const unsigned b = sblocks(st->n_envs);
    cudaError_t e = cudaSuccess;
    switch (io->obs_kind) {
        case DMP_OBS_F32:
            k_stage_observe<float><<<b, SB, SB * D * sizeof(float), as_stream(stream)>>>(*st, *io);
            break;
        case DMP_OBS_F64: {
            const int smem = SB * D * (int)sizeof(double);
            e = cudaFuncSetAttribute(k_stage_observe<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return dmp_set_error(e);
            k_stage_observe<double><<<b, SB, smem, as_stream(stream)>>>(*st, *io);
            break;
        }
        case DMP_OBS_I16:
            if (io->flags & DMP_F_NORMALISE) return DMP_EINVAL;
            k_stage_observe<int16_t><<<b, SB, SB * D * sizeof(int16_t), as_stream(stream)>>>(*st, *io);
            break;
        default: return DMP_EINVAL;
    }
    return stage_leave(st, stream);
}

int dmp_stage_reward(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream) {
    if (!stage_args_ok(st, io, scratch)) return DMP_EINVAL;
    if (const int rc = stage_enter(st, stream)) return rc;
    k_stage_reward<<<sblocks(st->n_envs), SB, 0, as_stream(stream)>>>(*st, *io, scratch);
    return stage_leave(st, stream);
}

int dmp_stage_done_reset(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream) {
    if (!stage_args_ok(st, io, scratch) || !io->done) return DMP_EINVAL;
    if (!st->ep_cnt || !st->ep_len || !st->ep_ret || !st->ep_iou) return DMP_EINVAL;
    if (const int rc = stage_enter(st, stream)) return rc;
    k_stage_done_reset<<<sblocks(st->n_envs), SB, 0, as_stream(stream)>>>(*st, *io);
    return stage_leave(st, stream);
}

}  // extern "C"

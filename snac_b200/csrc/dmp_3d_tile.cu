// dmp_3d_tile.cu -- 3D envs, tile kernel: 32 envs per warp, one env per lane, the warp's 32 height
// maps staged in shared memory by bulk async copies (cp.async.bulk -> UBLKCP, completion on an mbarrier).
//
// Why: the step logic of a 3D env is scalar (Philox, check_sur, move_step, the termination rules); with
// one env per warp (dmp_3d.cu) it is replicated 32x and the kernel is instruction bound.  Here every lane
// runs its own env (SIMT-efficient like the 2D kernel) and only the data movement is cooperative:
//   * state in : one 800 B bulk copy per env (issued by its lane, all 32 in flight at once), 16 B aux per lane;
//   * step     : neighbours / walk cells / 7x7 window are shared-memory reads; a brick is written to shared
//                memory and written through to HBM with a single 2-byte store;
//   * obs out  : [32][51] tile in shared memory, streamed out as one contiguous span of 128-bit stores;
//   * done     : ballot over the warp, then the whole warp computes that env's IoU (128-bit shared loads,
//                min, shuffle tree) and clears its height map (coalesced 128-bit stores)  -- stage (d)/(e).
// Semantics: identical to dmp_3d.cu (reference citations there); tests cross-check the two kernels.
#include "dmp_3d_u16.cuh"

namespace {

using namespace u16map;

constexpr int ENV_STRIDE = 816;              // bytes per env in smem: 800 B map + 16 B guard (window over-read)
constexpr int GUARD = 16;                    // leading guard before env 0 of a warp
constexpr unsigned FULL = 0xFFFFFFFFu;

// WPB = warps per block.  Rollout launches use the largest block that fits the 227 KB of shared memory
// (7 warps for f32/i16 observations: 7 x (26 128 B maps + 6 528 B tile) = 228.6 KB, one block per SM);
// single-step launches use 2-warp blocks (3 per SM), which schedule and drain better.
template <typename ObsT, int WPB>
__global__ void __launch_bounds__(WPB * 32) k3d_tile_rollout(const DmpState st, const DmpIO io, const int K) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = st.n_envs;
    const int64_t env0 = ((int64_t)blockIdx.x * WPB + warp) * 32;
    if (env0 >= n) return;                                            // whole warp leaves together
    const int64_t env = env0 + lane;
    const bool live = env < n;
    const int nvalid = (int)min((int64_t)32, n - env0);

    // shared memory carve-up: per warp [GUARD | 32 x ENV_STRIDE] ... then tiles, then barriers
    uint8_t* base = reinterpret_cast<uint8_t*>(smem_raw);
    constexpr int WARP_GRID_BYTES = GUARD + 32 * ENV_STRIDE;
    uint8_t* wgrid = base + warp * WARP_GRID_BYTES + GUARD;
    ObsT* tile = reinterpret_cast<ObsT*>(base + WPB * WARP_GRID_BYTES) + warp * (32 * D3_OBS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + WPB * WARP_GRID_BYTES + WPB * 32 * D3_OBS * sizeof(ObsT)) + warp;
    uint16_t* g = reinterpret_cast<uint16_t*>(wgrid + lane * ENV_STRIDE);         // this lane's height map

    uint16_t* cells = reinterpret_cast<uint16_t*>(st.cells);
    uint4* aux = reinterpret_cast<uint4*>(st.aux);
    const uint8_t* __restrict__ plans = reinterpret_cast<const uint8_t*>(st.plans);

    // ---- state in: 32 bulk copies in flight, one per lane ------------------------------------------
    if (lane == 0) mbar_init(bar, 32);
    __syncwarp();
    if (live) {
        mbar_arrive_expect_tx(bar, CELLS3D * 2);
        bulk_g2s(g, cells + env * CELLS3D, CELLS3D * 2, bar);
    } else {
        mbar_arrive(bar);
    }
    EnvT e{D2_LO, D2_LO, 0, 0, 0, 0.f, 0};
    if (live) {
        const uint4 a = aux[env];
        e.pr = a.x & 0xFF; e.pc = (a.x >> 8) & 0xFF; e.plan_idx = a.x >> 16;
        e.cb = a.y & 0xFFFF; e.cs = a.y >> 16;
        e.ret = __uint_as_float(a.z);
        e.cross = (int)a.w;
    }
    int total_brick = __ldg(st.plan_total + e.plan_idx);
    int errbits = 0;
    const bool dynamic = st.dynamic != 0;
    const bool autoreset = io.flags & DMP_F_AUTORESET;
    const bool normalise = io.flags & DMP_F_NORMALISE;
    const bool need_draw = (io.actions == nullptr) || (io.step_sizes == nullptr);
    const int tslot = (io.flags & DMP_F_TSLOT1) ? 1 : 0;
    const uint64_t t0 = st.t_dev ? st.t_dev[tslot] : st.t;
    mbar_wait(bar, 0);

    StepDraws draws;
    for (int k = 0; k < K; ++k) {
        const uint64_t t = t0 + (uint64_t)k;
        const int64_t idx = (int64_t)k * n + env;
        uint32_t dw = 0;
        if (need_draw) dw = draws.word(st.seed, (uint64_t)(st.env_base + env), t);
        int a, s;
        if (io.actions) a = live ? (int)io.actions[idx] : 0; else a = draw_action(dw, D3_ACT, st.action_dist);
        if (io.step_sizes) s = live ? (int)io.step_sizes[idx] : 1; else s = draw_step_size(dw);
        if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;
        if (a > 7) errbits |= DMP_ERR_ACTION;              // reference: an unbuilt brick (:187-208)

        e.cs += 1;
        // ---- check_sur (:88-102) --------------------------------------------------------------
        int n1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) n1[q] = cell_s(g, e.pr + dir_dr(q), e.pc + dir_dc(q));
        const bool boxed = (n1[0] != 0) && (n1[1] != 0) && (n1[2] != 0) && (n1[3] != 0);

        float reward = 0.f;
        bool done = false, tail = true;
        if (a <= 3) {
            // (a) move_step (:104-134): consecutive empty cells, at most s
            const int na = (a == 0) ? n1[0] : (a == 1) ? n1[1] : (a == 2) ? n1[2] : n1[3];
            if (na == 0) {
                const int dr = dir_dr(a), dc = dir_dc(a);
                int nstep = 1;
                if (s >= 2 && cell_s(g, e.pr + 2 * dr, e.pc + 2 * dc) == 0) {
                    nstep = 2;
                    if (s >= 3 && cell_s(g, e.pr + 3 * dr, e.pc + 3 * dc) == 0) nstep = 3;
                }
                e.pr = min(max(e.pr + dr * nstep, D2_LO), D2_HI);
                e.pc = min(max(e.pc + dc * nstep, D2_LO), D2_HI);
            }
        } else {
            // (b) build on neighbour a-4 unless it is frame
            const int q = a - 4;
            bool built = false;
            int newh = 0, ti = 0, pplan = 0;
            int m1[4] = {n1[0], n1[1], n1[2], n1[3]};
            if (a <= 7) {
                const int nq = (q == 0) ? n1[0] : (q == 1) ? n1[1] : (q == 2) ? n1[2] : n1[3];
                if (nq != -1) {
                    built = true;
                    newh = nq + 1;
                    ti = (e.pr + dir_dr(q) - 3) * 20 + (e.pc + dir_dc(q) - 3);
                    e.cb += 1;
                    pplan = plans[e.plan_idx * CELLS3D + ti];
                    if (newh <= pplan) e.cross += 1;
                    g[ti] = (uint16_t)newh;
                    if (live) cells[env * CELLS3D + ti] = (uint16_t)newh;      // write-through
#pragma unroll
                    for (int z = 0; z < 4; ++z) if (z == q) m1[z] = newh;
                }
            }
            if (dynamic) {
                const bool boxed2 = (m1[0] != 0) && (m1[1] != 0) && (m1[2] != 0) && (m1[3] != 0);
                if (boxed2) { reward = -100.f; done = true; tail = false; }
                else if (e.cb >= total_brick) { done = true; tail = false; }
                else if (built) { tail = false; }
            } else {
                if (e.cb >= total_brick || boxed) { done = true; tail = false; }
                else if (built) { tail = false; }
            }
            if (!tail && !done) reward = (newh > pplan) ? -1.f : (newh == pplan ? 10.f : 1.f);
        }
        if (tail) done = (e.cs >= st.total_step) || (!dynamic && boxed);
        e.ret += reward;

        // ---- (c) observation --------------------------------------------------------------------
        if (io.obs) {
            observe_tile<ObsT>(g, e, tile + lane * D3_OBS, normalise, total_brick, st.total_step);
            __syncwarp();
            ObsT* dst = reinterpret_cast<ObsT*>(io.obs) + ((int64_t)k * n + env0) * D3_OBS;
            if (nvalid == 32) warp_tile_store_full<ObsT, 32 * D3_OBS>(dst, tile, lane);
            else warp_tile_store<ObsT>(dst, tile, nvalid * D3_OBS, lane);
            __syncwarp();
        }
        if (live) {
            if (io.reward) io.reward[idx] = reward;
            if (io.done) io.done[idx] = done ? 1 : 0;
        }

        // ---- (d)/(e) finished episodes ---------------------------------------------------------------
        // IoU = cross / (total_brick + count_brick - cross) (:257-276) needs no scan: `cross` is kept up to date
        // by every build.  Each finished lane folds its episode into the statistics and resets its scalars; the
        // warp then clears the finished envs' height maps cooperatively (ballot, 128-bit coalesced stores).
        const bool fin = done && autoreset && live;
        if (fin) {
            const double iou = __ddiv_rn((double)e.cross, (double)(total_brick + e.cb - e.cross));
            atomicAdd(st.ep_cnt + env, 1u);                 // fire-and-forget REDs: no read-modify-write stall
            atomicAdd(st.ep_len + env, (uint32_t)e.cs);
            atomicAdd(st.ep_ret + env, (double)e.ret);
            atomicAdd(st.ep_iou + env, iou);
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
        }
        unsigned dm = __ballot_sync(FULL, fin);
        while (dm) {
            const int src = __ffs(dm) - 1;
            dm &= dm - 1;
            if (lane < 25) {
                uint4* sg = reinterpret_cast<uint4*>(wgrid + src * ENV_STRIDE);
                uint4* gg = reinterpret_cast<uint4*>(cells + (env0 + src) * CELLS3D);
                const uint4 z = make_uint4(0, 0, 0, 0);
                sg[2 * lane] = z; sg[2 * lane + 1] = z;
                gg[2 * lane] = z; gg[2 * lane + 1] = z;
            }
        }
        __syncwarp();
    }
    if (live) {
        aux[env] = make_uint4((uint32_t)e.pr | ((uint32_t)e.pc << 8) | ((uint32_t)e.plan_idx << 16),
                              (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16), __float_as_uint(e.ret), (uint32_t)e.cross);
        if (errbits) atomicOr(st.err, errbits);
        if (st.t_dev && env == 0) st.t_dev[tslot ^ 1] = t0 + (uint64_t)K;
    }
}

template <typename ObsT, int WPB>
int launch_tile_wpb(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    const size_t smem = (size_t)WPB * (GUARD + 32 * ENV_STRIDE) + (size_t)WPB * 32 * D3_OBS * sizeof(ObsT) + WPB * 8 + 16;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k3d_tile_rollout<ObsT, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return dmp_set_error(e);
        attr_done = true;
    }
    const int64_t warps = (st.n_envs + 31) / 32;
    const unsigned blocks = (unsigned)((warps + WPB - 1) / WPB);
    k3d_tile_rollout<ObsT, WPB><<<blocks, WPB * 32, smem, s>>>(st, io, K);
    return dmp_set_error(cudaGetLastError());
}

template <typename ObsT, int WPB_ROLLOUT>
int launch_tile(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    if (K == 1 || st.n_envs < 32 * WPB_ROLLOUT * 148) return launch_tile_wpb<ObsT, 2>(st, io, K, s);
    return launch_tile_wpb<ObsT, WPB_ROLLOUT>(st, io, K, s);
}

}  // namespace

int dmp3d_tile_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    switch (io.obs_kind) {
        case DMP_OBS_F32: return launch_tile<float, 7>(st, io, K, s);
        case DMP_OBS_F64: return launch_tile<double, 5>(st, io, K, s);
        case DMP_OBS_I16: return launch_tile<int16_t, 7>(st, io, K, s);
    }
    return DMP_EINVAL;
}

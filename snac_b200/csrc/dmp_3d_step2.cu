// dmp_3d_step2.cu -- 3D envs, single-step kernel (K = 1, dmp_step), second generation: TWO dependent HBM round
// trips per step instead of three, and 40 % less shared memory per env in flight.
//
// dmp_3d_step.cu fetched (1) the scalar state, (2) the six cells the move/build decision reads, (3) the map rows
// under the window at the NEW position -- three dependent DRAM latencies with ~11 warps per SM to hide them
// (profiles/README.md: 0.34 of the HBM roofline, long-scoreboard bound).  Here the action and the step size are
// known before anything is loaded (inputs or Philox), so the rows that can matter are known as soon as the position
// is: the 7 rows under the old window, plus `step_size` more rows in the direction of a vertical move.  They come in
// as ONE bulk async copy per env (cp.async.bulk -> UBLKCP, <= 416 B) right after the scalar state; the decision
// cells, the new window and the brick patch are all served from that staged span.  The observation tile of the warp
// is built in the same shared memory once every lane holds its window in registers, so a warp needs 14.5 KB instead
// of 17 KB (+ f32 tile) and 14 single-warp blocks fit an SM.
// Semantics: Env/3D/DMP_simulator_3d_static_circle.py:67-276 and
// Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-277 (see dmp_3d.cu for the line-by-line citations).
#include "dmp_3d_u16.cuh"

namespace {

using namespace u16map;

constexpr int SLOT2_B = 464;                 // per-lane staging: 16 B guard | <= 416 B of rows | 16 B guard, padded to an
                                             // ODD number of 16 B granules (29): same-offset words of the 32 lanes then fall
                                             // into 8 distinct 4-bank groups (4-way conflicts; 448 B = 28 granules gave 16-way)
constexpr unsigned FULL = 0xFFFFFFFFu;

// `tune` (DMP_3D_STEP_TUNE, default 0): bit 0 = at the end of the step ask L2 for the rows under the NEW window, so that
// the next step's round trip 2 is an L2 hit; bits 1-2 = L2 policy of the row copy (1: evict_last for half of the lines,
// 2: for all of them) -- the rows an agent looks at change by at most three per step.
template <typename ObsT>
__global__ void __launch_bounds__(32) k3d_step_span(const DmpState st, const DmpIO io, const int tune) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x;
    const int64_t n = st.n_envs;
    const int64_t env0 = (int64_t)blockIdx.x * 32;
    const int nvalid = (int)min((int64_t)32, n - env0);
    const bool live = lane < nvalid;
    const int64_t env = env0 + (live ? lane : 0);                     // idle lanes shadow env0 but never store

    uint8_t* base = reinterpret_cast<uint8_t*>(smem_raw);
    uint8_t* slot = base + (size_t)lane * SLOT2_B;
    ObsT* tile = reinterpret_cast<ObsT*>(base);                       // aliases the slots (used after they are drained)
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + (size_t)32 * SLOT2_B);

    uint16_t* cells = reinterpret_cast<uint16_t*>(st.cells);
    uint16_t* gwarp = cells + env0 * CELLS3D;
    uint16_t* ge = cells + env * CELLS3D;                             // this lane's map in HBM
    uint4* aux = reinterpret_cast<uint4*>(st.aux);
    const uint8_t* __restrict__ plans = reinterpret_cast<const uint8_t*>(st.plans);

    if (lane == 0) mbar_init(bar, 32);
    pdl_launch_dependents();
    pdl_wait();                                                       // the previous step's state is visible from here
    // ---- round trip 1: scalar state (the draws do not depend on it and overlap its latency).  16 B per env: the whole
    // array (4 MB at 262 144 envs) is kept in L2 (evict_last), so this round trip is an L2 hit, not a DRAM access
    const uint64_t keep = l2_policy_keep();
    const uint4 ax = ldg_keep(aux + env, keep);
    const bool dynamic = st.dynamic != 0;
    const bool autoreset = io.flags & DMP_F_AUTORESET;
    const bool normalise = io.flags & DMP_F_NORMALISE;
    const bool need_draw = (io.actions == nullptr) || (io.step_sizes == nullptr);
    const int tslot = (io.flags & DMP_F_TSLOT1) ? 1 : 0;
    const uint64_t t = st.t_dev ? st.t_dev[tslot] : st.t;
    const uint64_t gid = (uint64_t)(st.env_base + env0) + (uint64_t)lane;
    const int64_t idx = env0 + lane;
    int errbits = 0;

    StepDraws draws;
    uint32_t dw = 0;
    if (need_draw) dw = draws.word(st.seed, gid, t);
    int a, s;
    if (io.actions) a = live ? (int)io.actions[idx] : 0; else a = draw_action(dw, D3_ACT, st.action_dist);
    if (io.step_sizes) s = live ? (int)io.step_sizes[idx] : 1; else s = draw_step_size(dw);
    if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;
    if (a > 7) errbits |= DMP_ERR_ACTION;                  // reference: an unbuilt brick (:187-208)
    const int dir = a & 3, dr = dir_dr(dir), dc = dir_dc(dir);

    EnvT e;
    e.pr = ax.x & 0xFF; e.pc = (ax.x >> 8) & 0xFF; e.plan_idx = ax.x >> 16;
    e.cb = ax.y & 0xFFFF; e.cs = (ax.y >> 16) + 1;
    e.ret = __uint_as_float(ax.z);
    e.cross = (int)ax.w;

    // ---- round trip 2: every map row this step can look at, one bulk copy per env ---------------------------
    const int ext = min(max(s, 1), 3);                     // a move covers at most min(s, 3) cells (move_step :104-134)
    const int row_lo = max(e.pr - 6 - (a == 3 ? ext : 0), 0);
    const int row_hi = min(e.pr + (a == 2 ? ext : 0), 19);
    const int b_lo = (row_lo * 40) & ~15, b_hi = ((row_hi + 1) * 40 + 15) & ~15;      // 16 B granules, <= 416 B
    __syncwarp();                                                                       // mbarrier init visible
    if (live) {
        mbar_arrive_expect_tx(bar, (uint32_t)(b_hi - b_lo));
        if (tune & 6) {
            uint64_t pol;
            if (tune & 4) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
            else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.5;" : "=l"(pol));
            bulk_g2s_hint(slot + 16, reinterpret_cast<const uint8_t*>(ge) + b_lo, (uint32_t)(b_hi - b_lo), bar, pol);
        } else {
            bulk_g2s(slot + 16, reinterpret_cast<const uint8_t*>(ge) + b_lo, (uint32_t)(b_hi - b_lo), bar);
        }
    } else {
        mbar_arrive(bar);
    }
    const int total_brick = __ldg(st.plan_total + e.plan_idx);
    const int o = (e.pr - 3) * 20 + (e.pc - 3);
    const int ti = min(max(o + dr * 20 + dc, 0), CELLS3D - 1);        // build target (valid whenever a brick is laid)
    int pplan = 0;
    if (a >= 4) pplan = __ldg(plans + e.plan_idx * CELLS3D + ti);     // consumed after the observation
    // virtual map base: cell (r, c) of the staged rows lives at g[r * 20 + c]
    uint16_t* g = reinterpret_cast<uint16_t*>(slot + 16 - b_lo);
    mbar_wait(bar, 0);

    // ---- the six cells the decision reads: four neighbours (check_sur :88-102), second and third cell in the
    // action's direction (move_step).  Unconditional reads at an index clamped into the staged span; whether a cell
    // is frame follows from one coordinate.
    int c6[6];
    {
        const int lo_cell = row_lo * 20, hi_cell = row_hi * 20 + 19;
        const int dstep = dr * 20 + dc, sgn = dr + dc;
        const int coord = (dir < 2 ? e.pc : e.pr) - 3;
        auto at = [&](int i) { return (int)g[min(max(i, lo_cell), hi_cell)]; };
        const int vl = at(o - 1), vr = at(o + 1), vu = at(o + 20), vd = at(o - 20);
        const int v2 = at(o + 2 * dstep), v3 = at(o + 3 * dstep);
        c6[0] = (e.pc > D2_LO) ? vl : -1;
        c6[1] = (e.pc < D2_HI) ? vr : -1;
        c6[2] = (e.pr < D2_HI) ? vu : -1;
        c6[3] = (e.pr > D2_LO) ? vd : -1;
        c6[4] = ((unsigned)(coord + 2 * sgn) < 20u) ? v2 : -1;
        c6[5] = ((unsigned)(coord + 3 * sgn) < 20u) ? v3 : -1;
    }
    const bool boxed = (c6[0] != 0) && (c6[1] != 0) && (c6[2] != 0) && (c6[3] != 0);     // check_sur
    const int nsel = (dir == 0) ? c6[0] : (dir == 1) ? c6[1] : (dir == 2) ? c6[2] : c6[3];

    bool done = false, tail = true;
    bool built = false, boxed_penalty = false;
    int newh = 0;
    if (a <= 3) {
        // (a) move_step (:104-134): consecutive empty cells, at most s
        int nstep = 0;
        if (nsel == 0) nstep = (s >= 2 && c6[4] == 0) ? ((s >= 3 && c6[5] == 0) ? 3 : 2) : 1;
        e.pr = min(max(e.pr + dr * nstep, D2_LO), D2_HI);
        e.pc = min(max(e.pc + dc * nstep, D2_LO), D2_HI);
    } else {
        // (b) build on neighbour a-4 unless it is frame
        bool open_after = (c6[0] == 0) || (c6[1] == 0) || (c6[2] == 0) || (c6[3] == 0);
        if (a <= 7 && nsel != -1) {
            built = true;
            newh = nsel + 1;
            e.cb += 1;
            open_after = ((dir != 0) && c6[0] == 0) || ((dir != 1) && c6[1] == 0) ||
                         ((dir != 2) && c6[2] == 0) || ((dir != 3) && c6[3] == 0);
        }
        if (dynamic) {                                   // re-check after placement (:199-231)
            if (!open_after) { boxed_penalty = true; done = true; tail = false; }
            else if (e.cb >= total_brick) { done = true; tail = false; }
            else if (built) { tail = false; }
        } else {                                         // static (:210-230)
            if (e.cb >= total_brick || boxed) { done = true; tail = false; }
            else if (built) { tail = false; }
        }
    }
    if (tail) done = (e.cs >= st.total_step) || (!dynamic && boxed);
    if (built) {
        g[ti] = (uint16_t)newh;                          // patch the staged rows ...
        if (live) ge[ti] = (uint16_t)newh;               // ... and write the brick through to HBM
    }

    // ---- (c) observation: window -> registers, then the warp's [32][51] tile over the drained slots ----------
    bool bulk_pending = false;
    if (io.obs) {
        uint32_t u[7][4];
        window_regs(g, e, u);
        __syncwarp();                                    // every lane has read its slot
        ObsT* row = tile + lane * D3_OBS;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            ObsT dummy;
            emit_pair<ObsT>(u[k][0], row[k * 7 + 0], row[k * 7 + 1]);
            emit_pair<ObsT>(u[k][1], row[k * 7 + 2], row[k * 7 + 3]);
            emit_pair<ObsT>(u[k][2], row[k * 7 + 4], row[k * 7 + 5]);
            emit_pair<ObsT>(u[k][3], row[k * 7 + 6], dummy);
        }
        obs_counters<ObsT>(normalise, e.cb, e.cs, total_brick, st.total_step, row[49], row[50]);
        ObsT* dst = reinterpret_cast<ObsT*>(io.obs) + env0 * D3_OBS;
        if (nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            warp_tile_bulk_store(dst, tile, 32 * D3_OBS * sizeof(ObsT), lane);          // one bulk async copy per warp
            bulk_pending = true;
        } else {
            __syncwarp();
            warp_tile_store<ObsT>(dst, tile, nvalid * D3_OBS, lane);
        }
    }

    // ---- (d) reward (reward_check :232-239) -----------------------------------------------------
    float reward = 0.f;
    if (built) {
        if (newh <= pplan) e.cross += 1;
        if (!tail && !done) reward = (newh > pplan) ? -1.f : (newh == pplan ? 10.f : 1.f);
    }
    if (boxed_penalty) reward = -100.f;
    e.ret += reward;
    if (live) {
        if (io.reward) io.reward[idx] = reward;
        if (io.done) io.done[idx] = done ? 1 : 0;
    }

    // ---- (e) finished episodes: IoU = cross / (total_brick + count_brick - cross) (:257-276) ---------------
    const bool fin = done && autoreset && live;
    if (fin) {
        const int den = total_brick + e.cb - e.cross;
        const double iou = (e.cross == 0 && den != 0) ? 0.0 : __ddiv_rn((double)e.cross, (double)den);
        atomicAdd(st.ep_cnt + env, 1u);                 // fire-and-forget REDs
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
        e.pr = e.pc = D2_LO; e.cb = e.cs = 0; e.ret = 0.f; e.cross = 0;
    }
    unsigned dm = __ballot_sync(FULL, fin);
    while (dm) {                                            // the warp clears each finished env's map in HBM
        const int src = __ffs(dm) - 1;
        dm &= dm - 1;
        if (lane < 25) {
            uint4* gg = reinterpret_cast<uint4*>(gwarp + src * CELLS3D) + 2 * lane;
            const uint4 z = make_uint4(0, 0, 0, 0);
            gg[0] = z; gg[1] = z;
        }
    }
    if ((tune & 1) && live) {                               // rows under the window the next step starts from
        const int p_lo = (max(e.pr - 6, 0) * 40) & ~15, p_hi = ((min(e.pr, 19) + 1) * 40 + 15) & ~15;
        bulk_prefetch_l2(reinterpret_cast<const uint8_t*>(ge) + p_lo, (uint32_t)(p_hi - p_lo));
    }
    if (live) {
        stg_keep(aux + env, make_uint4((uint32_t)e.pr | ((uint32_t)e.pc << 8) | ((uint32_t)e.plan_idx << 16),
                                       (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16), __float_as_uint(e.ret), (uint32_t)e.cross), keep);
        if (errbits) atomicOr(st.err, errbits);
        if (st.t_dev && env == 0) st.t_dev[tslot ^ 1] = t + 1;
    }
    if (bulk_pending) warp_tile_bulk_wait(lane);                        // the tile must outlive the copy that reads it
}

template <typename ObsT>
int launch_span(const DmpState& st, const DmpIO& io, cudaStream_t s) {
    static_assert(32 * D3_OBS * sizeof(ObsT) <= 32 * SLOT2_B, "the observation tile must fit the drained slots");
    const size_t smem = (size_t)32 * SLOT2_B + 16;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k3d_step_span<ObsT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return dmp_set_error(e);
        e = cudaFuncSetAttribute(k3d_step_span<ObsT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return dmp_set_error(e);
        attr_done = true;
    }
    const unsigned blocks = (unsigned)((st.n_envs + 31) / 32);
    const char* tv = getenv("DMP_3D_STEP_TUNE");
    const int tune = tv ? atoi(tv) : 0;
    return dmp_set_error(dmp_launch_pdl(k3d_step_span<ObsT>, blocks, 32u, smem, s, st, io, tune));
}

}  // namespace

int dmp3d_step_span(const DmpState& st, const DmpIO& io, cudaStream_t s) {
    switch (io.obs_kind) {
        case DMP_OBS_F32: return launch_span<float>(st, io, s);
        case DMP_OBS_F64: return launch_span<double>(st, io, s);
        case DMP_OBS_I16: return launch_span<int16_t>(st, io, s);
    }
    return DMP_EINVAL;
}

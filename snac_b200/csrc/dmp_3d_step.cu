// dmp_3d_step.cu -- 3D envs, single-step kernel (K = 1, dmp_step): one env per lane; only the bytes one step
// can look at are fetched.
//
// A step reads the agent's movement cross (<= 6 cells, all within 3 cells of the agent) and the 7x7 window at
// the NEW position.  Staging the whole 800 B map for that (dmp_3d_tile.cu) triples the HBM traffic of a step
// (B_alg = 330 B); here
//   * the six cells are plain 2-byte loads (the other coordinate is the agent's, so one test decides "frame");
//   * the <= 7 map rows under the new window (<= 280 contiguous bytes) come in as ONE bulk async copy per env
//     (cp.async.bulk -> UBLKCP, 16 B granules, completion on the warp's mbarrier) into a 336 B slot per lane;
//   * a brick is patched into the staged rows after the copy has landed and written through to HBM with one
//     2-byte store; observation tile / copy-out / finished-episode handling are those of the tile kernel.
// Semantics: Env/3D/DMP_simulator_3d_static_circle.py:67-276 and
// Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-277 (see dmp_3d.cu for the line-by-line citations).
#include "dmp_3d_u16.cuh"

namespace {

using namespace u16map;

constexpr int SLOT_B = 336;                  // per-lane staging: 16 B guard | <= 304 B of rows | 16 B guard
constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int WPB_STEP = 4;

template <typename ObsT>
__global__ void __launch_bounds__(WPB_STEP * 32) k3d_step_rows(const DmpState st, const DmpIO io) {
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = st.n_envs;
    const int64_t env0 = ((int64_t)blockIdx.x * WPB_STEP + warp) * 32;
    if (env0 >= n) return;                                            // whole warp leaves together
    const int nvalid = (int)min((int64_t)32, n - env0);
    const bool live = lane < nvalid;
    const int64_t env = env0 + (live ? lane : 0);                     // idle lanes shadow env0 but never store

    uint8_t* base = reinterpret_cast<uint8_t*>(smem_raw);
    uint8_t* slot = base + (size_t)(warp * 32 + lane) * SLOT_B;
    ObsT* tile = reinterpret_cast<ObsT*>(base + (size_t)WPB_STEP * 32 * SLOT_B) + warp * (32 * D3_OBS);
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + (size_t)WPB_STEP * 32 * (SLOT_B + D3_OBS * sizeof(ObsT))) + warp;

    uint16_t* cells = reinterpret_cast<uint16_t*>(st.cells);
    uint16_t* gwarp = cells + env0 * CELLS3D;
    uint16_t* ge = cells + env * CELLS3D;                             // this lane's map in HBM
    uint4* aux = reinterpret_cast<uint4*>(st.aux);
    const uint8_t* __restrict__ plans = reinterpret_cast<const uint8_t*>(st.plans);

    if (lane == 0) mbar_init(bar, 32);
    EnvT e{D2_LO, D2_LO, 0, 0, 0, 0.f, 0};
    {
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
    const uint64_t t = st.t_dev ? st.t_dev[tslot] : st.t;
    const uint64_t gid = (uint64_t)(st.env_base + env0) + (uint64_t)lane;
    const int64_t idx = env0 + lane;

    StepDraws draws;
    uint32_t dw = 0;
    if (need_draw) dw = draws.word(st.seed, gid, t);
    int a, s;
    if (io.actions) a = live ? (int)io.actions[idx] : 0; else a = draw_action(dw, D3_ACT, st.action_dist);
    if (io.step_sizes) s = live ? (int)io.step_sizes[idx] : 1; else s = draw_step_size(dw);
    if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;
    if (a > 7) errbits |= DMP_ERR_ACTION;                  // reference: an unbuilt brick (:187-208)

    e.cs += 1;
    // ---- the six cells this step can depend on: the four neighbours (check_sur :88-102) and the second and third
    // cell in the action's direction (move_step :104-134).  Loads are unconditional at a clamped index; whether a
    // cell is frame follows from one coordinate.
    const int dir = a & 3, dr = dir_dr(dir), dc = dir_dc(dir);
    const int o = (e.pr - 3) * 20 + (e.pc - 3);
    int c6[6];
    {
        const int dstep = dr * 20 + dc, sgn = dr + dc;
        const int coord = (dir < 2 ? e.pc : e.pr) - 3;
        auto at = [&](int i) { return (int)ge[min(max(i, 0), CELLS3D - 1)]; };
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
    int newh = 0, pplan = 0, ti = 0;
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
            ti = o + dr * 20 + dc;
            e.cb += 1;
            pplan = __ldg(plans + e.plan_idx * CELLS3D + ti);
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

    // ---- stage the rows under the window at the new position: interior rows [pr-6, pr] clipped to the map ----
    const int row_lo = max(e.pr - 6, 0), row_hi = min(e.pr, 19);
    const int b_lo = (row_lo * 40) & ~15, b_hi = ((row_hi + 1) * 40 + 15) & ~15;      // 16 B granules, <= 304 B
    __syncwarp();                                                                       // mbarrier init visible
    if (live) {
        mbar_arrive_expect_tx(bar, (uint32_t)(b_hi - b_lo));
        bulk_g2s(slot + 16, reinterpret_cast<const uint8_t*>(ge) + b_lo, (uint32_t)(b_hi - b_lo), bar);
    } else {
        mbar_arrive(bar);
    }
    // virtual map base: cell (r, c) of the staged rows lives at g[r * 20 + c]
    uint16_t* g = reinterpret_cast<uint16_t*>(slot + 16 - b_lo);
    mbar_wait(bar, 0);
    if (built) {
        g[ti] = (uint16_t)newh;                          // the copy has landed: patch the staged rows ...
        if (live) ge[ti] = (uint16_t)newh;               // ... and write the brick through to HBM
    }

    // ---- (c) observation --------------------------------------------------------------------
    if (io.obs) {
        observe_tile<ObsT>(g, e, tile + lane * D3_OBS, normalise, total_brick, st.total_step);
        __syncwarp();
        ObsT* dst = reinterpret_cast<ObsT*>(io.obs) + env0 * D3_OBS;
        if (nvalid == 32) warp_tile_store_full<ObsT, 32 * D3_OBS>(dst, tile, lane);
        else warp_tile_store<ObsT>(dst, tile, nvalid * D3_OBS, lane);
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
    if (live) {
        aux[env] = make_uint4((uint32_t)e.pr | ((uint32_t)e.pc << 8) | ((uint32_t)e.plan_idx << 16),
                              (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16), __float_as_uint(e.ret), (uint32_t)e.cross);
        if (errbits) atomicOr(st.err, errbits);
        if (st.t_dev && env == 0) st.t_dev[tslot ^ 1] = t + 1;
    }
}

template <typename ObsT>
int launch_step(const DmpState& st, const DmpIO& io, cudaStream_t s) {
    const size_t smem = (size_t)WPB_STEP * 32 * (SLOT_B + D3_OBS * sizeof(ObsT)) + WPB_STEP * 8 + 16;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k3d_step_rows<ObsT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return dmp_set_error(e);
        e = cudaFuncSetAttribute(k3d_step_rows<ObsT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return dmp_set_error(e);
        attr_done = true;
    }
    const int64_t warps = (st.n_envs + 31) / 32;
    const unsigned blocks = (unsigned)((warps + WPB_STEP - 1) / WPB_STEP);
    k3d_step_rows<ObsT><<<blocks, WPB_STEP * 32, smem, s>>>(st, io);
    return dmp_set_error(cudaGetLastError());
}

}  // namespace

int dmp3d_step_rows(const DmpState& st, const DmpIO& io, cudaStream_t s) {
    switch (io.obs_kind) {
        case DMP_OBS_F32: return launch_step<float>(st, io, s);
        case DMP_OBS_F64: return launch_step<double>(st, io, s);
        case DMP_OBS_I16: return launch_step<int16_t>(st, io, s);
    }
    return DMP_EINVAL;
}

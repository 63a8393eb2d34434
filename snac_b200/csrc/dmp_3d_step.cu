// dmp_3d_step.cu -- 3D envs, single-step kernel (K = 1, dmp_step): the rows a step looks at are
// staged from the NIBBLE maps (include/dmp.h: u8[n][208] behind the wide u16 maps, nibble = min(h, 15)).
//
// A step is bound by what it must fetch from DRAM, and DRAM is read in 64 B pieces.  The 7..10 map rows under the window
// were 140..200 B of byte rows in the previous generation (256 B of DRAM reads per env-step, 0.58 of the HBM roofline);
// they are 70..100 B of nibble rows now (<= 128 B staged, ~160 B read).  (The generation before staged u16 rows and
// wrote every brick into a u16 map: 0.55 -- a 2-byte store into a line that is not in L2 costs a 32 B read and a 32 B
// write of DRAM traffic.)  <= 128 B of nibble rows per env land in a 144 B slot (9 granules of 16 B: an odd count spreads
// same-offset words of the 32 lanes over 8 bank groups); the window is cut out with 2 word reads + 1 funnel shift per row
// and spread to bytes with 2 LOP + 2 PRMT; the brick goes to the nibble map, whose line the step has just read.  The
// action and the step size are known before anything is loaded, so ONE bulk async copy per env (cp.async.bulk -> UBLKCP,
// completion on the warp's mbarrier) right after the scalar state fetches the 7 rows under the old window plus
// `step_size` more rows in the direction of a vertical move; the six decision cells, the window at the new position and
// the brick patch are served from that span; every lane pulls its window into registers, then the warp's [32][51]
// observation tile is built over the drained slots and leaves through one bulk async copy.
// Exactness (dmp_common.cuh): the nibbles of an env ARE its heights until one reaches 15; from then on the env is
// flagged tall and its wide u16 map is kept exact as well: a build on a saturated nibble reads the true height there, and
// so do the window cells that read 15 (out of line; plan height is 6, but among 262 144 random-policy envs a handful
// are tall at any time, so the tall path must not be slow: everything else still comes from the nibbles).
// Semantics: Env/3D/DMP_simulator_3d_static_circle.py:67-276 and
// Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-277 (see dmp_3d_roll.cu for the line-by-line citations).
#include "dmp_3d_bulk.cuh"

namespace {

using namespace d3;

constexpr int SLOT3_B = 144;                 // per-lane staging: 16 B guard | <= 128 B of nibble rows.  Word reads of the window
                                             // run up to 6 B past the rows: into the next lane's guard (masked columns)

// cell i of a nibble map addressed through its (virtual) word base
__device__ __forceinline__ int nib_at(const uint32_t* gw, int i) { return (int)((gw[i >> 3] >> ((i & 7) * 4)) & 0xFu); }

// per-warp shared memory: 32 slots, or the observation tile built over them when that is larger (f64)
template <typename ObsT>
__host__ __device__ constexpr size_t warp_area_bytes() {
    return (size_t)32 * row_elems<ObsT, D3_OBS>() * sizeof(ObsT) > (size_t)32 * SLOT3_B
               ? (size_t)32 * row_elems<ObsT, D3_OBS>() * sizeof(ObsT) : (size_t)32 * SLOT3_B;
}

// RF = true (DMP_F_RESET_OBS): a finished env's observation is the one its reset returns (gym-style auto-reset).
template <typename ObsT, bool RF>
__global__ void __launch_bounds__(128, 7) k3d_step_bytes(const DmpState st, const DmpIO io) {
    constexpr bool REC = is_rec<ObsT>::value;
    constexpr int ROW = row_elems<ObsT, D3_OBS>();
    extern __shared__ uint4 smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t n = st.n_envs;
    const int64_t env0 = ((int64_t)blockIdx.x * wpb + warp) * 32;
    if (env0 >= n) return;                                            // whole warp leaves together (no block-wide sync below)
    const int nvalid = (int)min((int64_t)32, n - env0);
    const bool live = lane < nvalid;
    const int64_t env = env0 + (live ? lane : 0);                     // idle lanes shadow env0 but never store

    // shared memory: [wpb] warp areas (32 slots, re-used as the warp's observation tile) | [wpb] mbarriers | 16 B pad
    constexpr size_t AREA_B = warp_area_bytes<ObsT>();
    uint8_t* base = reinterpret_cast<uint8_t*>(smem_raw) + (size_t)warp * AREA_B;
    uint8_t* slot = base + (size_t)lane * SLOT3_B;
    ObsT* tile = reinterpret_cast<ObsT*>(base);                       // aliases the slots (used after they are drained)
    uint64_t* bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_raw) + (size_t)wpb * AREA_B) + warp;

    uint16_t* cells = reinterpret_cast<uint16_t*>(st.cells);
    uint16_t* gwarp = cells + env0 * CELLS3D;
    uint16_t* ge = cells + env * CELLS3D;                             // this lane's u16 map in HBM (canonical)
    uint8_t* nwarp = nmap3(st) + env0 * NIB3_STRIDE;
    uint8_t* ne = nmap3(st) + env * NIB3_STRIDE;                      // ... and its nibble map
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
    const uint64_t t = st.t_dev ? st.t_dev[tslot] : st.t;             // requested here, consumed by the draws below
    const uint64_t gid = (uint64_t)(st.env_base + env0) + (uint64_t)lane;
    const int64_t idx = env0 + lane;
    int errbits = 0;

    EnvT e;
    e.pr = ax.x & 0x7F; e.pc = (ax.x >> 8) & 0xFF; e.plan_idx = ax.x >> 16;
    bool tall = live && (ax.x & AUX3_TALL);                // a height >= TALL3 somewhere: the wide map is the exact one
    e.cb = ax.y & 0xFFFF; e.cs = (ax.y >> 16) + 1;
    e.ret = __uint_as_float(ax.z);
    e.cross = (int)ax.w;

    StepDraws draws;
    uint32_t dw = 0;
    if (need_draw) dw = draws.word(st.seed, gid, t);
    int a, s;
    if (io.actions) a = live ? (int)io.actions[idx] : 0; else a = draw_action(dw, D3_ACT, st.action_dist);
    if (io.step_sizes) s = live ? (int)io.step_sizes[idx] : 1; else s = draw_step_size(dw);
    if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;
    if (a > 7) errbits |= DMP_ERR_ACTION;                  // reference: an unbuilt brick (:187-208)
    const int dir = a & 3, dr = dir_dr(dir), dc = dir_dc(dir);

    // ---- round trip 2: every map row this step can look at, one bulk copy of nibble rows (10 B each) per env.  (Requesting
    // the maximal 13-row range before the draws are known, so that the Philox rounds overlap the copy, measured 6 % slower
    // in a same-box A/B: 12.33 vs 13.06 G env-steps/s -- the extra 40 B per env cost more than the overlap saves.)
    const int ext = min(max(s, 1), 3);                     // a move covers at most min(s, 3) cells (move_step :104-134)
    const int row_lo = max(e.pr - 6 - (a == 3 ? ext : 0), 0);
    const int row_hi = min(e.pr + (a == 2 ? ext : 0), 19);
    const int b_lo = (row_lo * 10) & ~15, b_hi = ((row_hi + 1) * 10 + 15) & ~15;      // 16 B granules, <= 128 B, <= 208
    __syncwarp();                                                                       // mbarrier init visible
    if (live) {
        mbar_arrive_expect_tx(bar, (uint32_t)(b_hi - b_lo));
        bulk_g2s(slot + 16, ne + b_lo, (uint32_t)(b_hi - b_lo), bar);
    } else {
        mbar_arrive(bar);
    }
    const int total_brick = __ldg(st.plan_total + e.plan_idx);
    const int o = (e.pr - 3) * 20 + (e.pc - 3);
    const int ti = min(max(o + dr * 20 + dc, 0), CELLS3D - 1);        // build target (valid whenever a brick is laid)
    int pplan = 0;
    if (a >= 4) pplan = __ldg(plans + e.plan_idx * CELLS3D + ti);     // consumed after the observation
    // virtual map base: cell i of the staged rows is nibble i & 7 of g[i >> 3]; 16 B aligned like the slot
    uint32_t* g = reinterpret_cast<uint32_t*>(slot + 16 - b_lo);
    mbar_wait(bar, 0);

    // ---- the six cells the decision reads: four neighbours (check_sur :88-102), second and third cell in the
    // action's direction (move_step).  Unconditional reads at an index clamped into the staged span; whether a cell
    // is frame follows from one coordinate.  The decision only asks "== 0 / > 0 / frame": saturated nibbles answer it.
    int c6[6];
    {
        // unconditional reads at an index clamped into the staged span (measured 3-7 % faster than predicated reads)
        const int lo_cell = row_lo * 20, hi_cell = row_hi * 20 + 19;
        const int dstep = dr * 20 + dc, sgn = dr + dc;
        const int coord = (dir < 2 ? e.pc : e.pr) - 3;
        auto at = [&](int i) { return nib_at(g, min(max(i, lo_cell), hi_cell)); };
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
    int nsel = (dir == 0) ? c6[0] : (dir == 1) ? c6[1] : (dir == 2) ? c6[2] : c6[3];

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
            if (nsel == 15) nsel = (int)__ldcg(ge + ti);              // saturated nibble (tall envs only): the true height
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
    bool turns_tall = false;
    uint32_t brick_byte = 0;                             // the nibble-map byte that holds the new brick
    if (built) {
        uint32_t* wp = g + (ti >> 3);                    // patch the staged rows ...
        const int sh = (ti & 7) * 4;
        const uint32_t x = (*wp & ~(0xFu << sh)) | ((uint32_t)sat_nib(newh) << sh);
        *wp = x;
        brick_byte = (x >> (((ti >> 1) & 3) * 8)) & 0xFFu;
        if (live) {
            ne[ti >> 1] = (uint8_t)brick_byte;           // ... and write the brick's byte through to the nibble map
            if (tall) {
                ge[ti] = (uint16_t)newh;                 // a tall env keeps its wide map exact
            } else if (newh >= TALL3) {                  // the nibbles stop being exact after this brick: the env turns
                turns_tall = tall = true;                // tall and its wide map is made current (below, by the warp)
            }
        }
    }
    for (unsigned tw = __ballot_sync(FULL, turns_tall); tw; tw &= tw - 1) {
        const int src = __ffs(tw) - 1;
        const int tib = __shfl_sync(FULL, ti >> 1, src);
        const uint32_t nb = __shfl_sync(FULL, brick_byte, src);
        if (lane < 25) {                                 // 16 cells per lane; the new brick's byte may not have landed yet
            uint2 v = __ldcg(reinterpret_cast<const uint2*>(nwarp + src * NIB3_STRIDE) + lane);
            if ((tib >> 3) == lane) {
                const int bs = (tib & 3) * 8;
                if (tib & 4) v.y = (v.y & ~(0xFFu << bs)) | (nb << bs); else v.x = (v.x & ~(0xFFu << bs)) | (nb << bs);
            }
            warp_widen_words(gwarp + src * CELLS3D, v.x, v.y, lane);
        }
        __syncwarp();                                    // the wide map is visible to its env's lane
    }

    // ---- (d) reward (reward_check :232-239): the plan byte is consumed after the window has been formatted; a record
    // row carries the reward, so it is closed after (d), observation rows leave first
    float reward = 0.f;
    bool rewarded = false;
    auto reward_now = [&]() {
        if constexpr (RF) {                              // called twice in that mode: before the reset and by the record path
            if constexpr (RF) {                          // called twice in that mode: before the reset and by the record path
                if (rewarded) return;
                rewarded = true;
            }
        }
        if (built) {
            if (newh <= pplan) e.cross += 1;
            if (!tail && !done) reward = (newh > pplan) ? -1.f : (newh == pplan ? 10.f : 1.f);
        }
        if (boxed_penalty) reward = -100.f;
    };

    // ---- (e) finished episodes: IoU = cross / (total_brick + count_brick - cross) (:257-276) ---------------
    const bool fin = done && autoreset && live;
    const bool fin_wide = fin && tall;                      // a tall env's wide map is cleared with it
    auto finish_lane = [&]() {
        if (fin) {
            tall = false;
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
    };
    if constexpr (RF) {                                     // reward and reset first: the observation below is the reset env's
        reward_now();
        e.ret += reward;
        finish_lane();
    }

    // ---- (c) observation: window -> registers, then the warp's [32][51] tile over the drained slots ----------
    bool bulk_pending = false;
    if (io.obs) {
        // the seven window rows as 2 words of biased bytes each (height + 1, 0 = frame)
        uint32_t u0[7], u1[7];
        {
            const uint32_t cv = (COLVALID >> (e.pc - 3)) & 0x7Fu;     // window column j lies inside the plan area
            const uint64_t one = spread7(cv);                          // 0x01 per valid byte
            const uint32_t b0 = (uint32_t)one, b1 = (uint32_t)(one >> 32);
            const uint32_t m0 = b0 * 0xFFu, m1 = b1 * 0xFFu;           // 0xFF per valid byte (no carries)
            const int c0 = e.pc - 6;                                    // interior column of window column 0 (may be < 0)
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int ir = e.pr - 6 + k;                            // interior row of window row k
                const bool rowvalid = (unsigned)ir < 20u;
                const int irc = min(max(ir, row_lo), row_hi);           // rows outside the map re-read a staged row, masked
                const int ok = irc * 20 + c0;                           // nibble offset of the row's first window cell
                const uint32_t* rw = g + (ok >> 3);                     // over-reads stay inside the 16 B guards
                const uint32_t q = __funnelshift_r(rw[0], rw[1], (ok & 7) * 4);
                uint32_t q0, q1;
                nib8_to_bytes(q, q0, q1);                               // seven nibbles -> seven bytes (the eighth is masked)
                u0[k] = rowvalid ? ((q0 & m0) + b0) : 0u;
                u1[k] = rowvalid ? ((q1 & m1) + b1) : 0u;
            }
        }
        if constexpr (RF) {
            if (fin) {                                   // window at [3, 3] of an empty map: rows / columns 0..2 are frame
#pragma unroll
                for (int k = 0; k < 7; ++k) { u0[k] = k < 3 ? 0u : 0x01000000u; u1[k] = k < 3 ? 0u : 0x00010101u; }
            }
        }
        if constexpr (is_bits<ObsT>::value) {            // one 32 B record per env, straight from registers
            reward_now();
            if (live) bits32_store(io.obs, idx, u0, u1, e.cb, e.cs, reward, done);
        } else {
        __syncwarp();                                    // every lane has read its slot
        ObsT* row = tile + lane * ROW;
        const bool fix = tall && window_saturated(u0, u1);      // a cell of 15 or more in a tall env's window: practically never
        if constexpr (REC) {                             // one packed record per env; it carries reward and done
            uint64_t c[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) c[k] = (uint64_t)u0[k] | ((uint64_t)u1[k] << 32);
            if (fix) fix_saturated_codes(ge, e.pr, e.pc, c);     // exact heights from the wide map, this step's brick included
            uint32_t w[13];
            pack49(c, w);
            reward_now();
            rec56_store(row, w, e.cb, e.cs, reward, done, tall);
        } else {
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                ObsT* o7 = row + k * 7;
                o7[0] = obs_from_biased<ObsT, 0>(u0[k]);
                o7[1] = obs_from_biased<ObsT, 1>(u0[k]);
                o7[2] = obs_from_biased<ObsT, 2>(u0[k]);
                o7[3] = obs_from_biased<ObsT, 3>(u0[k]);
                o7[4] = obs_from_biased<ObsT, 0>(u1[k]);
                o7[5] = obs_from_biased<ObsT, 1>(u1[k]);
                o7[6] = obs_from_biased<ObsT, 2>(u1[k]);
            }
            if (fix) fix_saturated_row<ObsT>(ge, e.pr, e.pc, u0, u1, row);      // exact heights from the wide map
            obs_counters<ObsT>(normalise, e.cb, e.cs, total_brick, st.total_step, row[49], row[50]);
        }
        ObsT* dst = reinterpret_cast<ObsT*>(io.obs) + env0 * ROW;
        if (!(io.flags & DMP_F_TILE_LDST) && nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            warp_tile_bulk_store(dst, tile, 32 * ROW * sizeof(ObsT), lane);             // one bulk async copy per warp
            bulk_pending = true;
        } else {
            __syncwarp();
            tile_rows_store<ObsT, D3_OBS>(dst, tile, nvalid, lane);
        }
        if constexpr (!REC) reward_now();
        }
    } else {
        reward_now();
    }
    if constexpr (!RF) e.ret += reward;
    if (live) {
        if (io.reward) io.reward[idx] = reward;
        if (io.done) io.done[idx] = done ? 1 : 0;
    }

    if constexpr (!RF) finish_lane();
    unsigned dm = __ballot_sync(FULL, fin);
    const unsigned dmw = __ballot_sync(FULL, fin_wide);
    while (dm) {                                            // the warp clears each finished env's map in HBM
        const int src = __ffs(dm) - 1;
        dm &= dm - 1;
        const uint4 z = make_uint4(0, 0, 0, 0);
        if (lane < NIB3_STRIDE / 16) reinterpret_cast<uint4*>(nwarp + src * NIB3_STRIDE)[lane] = z;
        if (lane < 25 && ((dmw >> src) & 1u)) {
            uint4* gg = reinterpret_cast<uint4*>(gwarp + src * CELLS3D) + 2 * lane;
            gg[0] = z; gg[1] = z;
        }
    }
    if (live) {
        if ((e.cb | e.cs) > 0xFFFF) {                                 // 16-bit packed counters (include/dmp.h)
            errbits |= DMP_ERR_OVERFLOW;
            e.cb = min(e.cb, 0xFFFF); e.cs = min(e.cs, 0xFFFF);
        }
        stg_keep(aux + env, make_uint4((uint32_t)e.pr | (tall ? AUX3_TALL : 0u) | ((uint32_t)e.pc << 8) | ((uint32_t)e.plan_idx << 16),
                                       (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16), __float_as_uint(e.ret), (uint32_t)e.cross), keep);
        if (errbits) atomicOr(st.err, errbits);
        if (st.t_dev && env == 0) st.t_dev[tslot ^ 1] = t + 1;
    }
    if (bulk_pending) warp_tile_bulk_wait(lane);                        // the tile must outlive the copy that reads it
}

// Launch shape: two-warp blocks, 28 warps per SM by registers (72): 8 192 warps (BASELINE's 262 144 envs) are 1.98 waves.  Measured 12.2 G env-steps/s
// against 12.0 G with single-warp blocks.  (Per-lane cp.async copies of the rows instead of the per-lane bulk copy:
// 10.9 G, MIO-throttled -- removed.)
template <typename ObsT, bool RF>
int launch_bytes_r(const DmpState& st, const DmpIO& io, cudaStream_t s) {
    static_assert(warp_area_bytes<ObsT>() % 16 == 0, "warp areas and the mbarriers behind them must stay 16 B aligned");
    constexpr int wpb = 2;
    const size_t smem = (size_t)wpb * warp_area_bytes<ObsT>() + (size_t)wpb * 8 + 16;
    static bool attr_done = false;                           // per instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k3d_step_bytes<ObsT, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return dmp_set_error(e);
        e = cudaFuncSetAttribute(k3d_step_bytes<ObsT, RF>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return dmp_set_error(e);
        attr_done = true;
    }
    const int64_t warps = (st.n_envs + 31) / 32;
    const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
    return dmp_set_error(dmp_launch_pdl(!(io.flags & DMP_F_NO_PDL), k3d_step_bytes<ObsT, RF>, blocks, (unsigned)(wpb * 32), smem,
                                        s, st, io));
}

template <typename ObsT>
int launch_bytes(const DmpState& st, const DmpIO& io, cudaStream_t s) {
    if (io.flags & DMP_F_RESET_OBS) return launch_bytes_r<ObsT, true>(st, io, s);
    return launch_bytes_r<ObsT, false>(st, io, s);
}

}  // namespace

int dmp3d_step_bytes(const DmpState& st, const DmpIO& io, cudaStream_t s) {
    switch (io.obs_kind) {
        case DMP_OBS_F32: return launch_bytes<float>(st, io, s);
        case DMP_OBS_F64: return launch_bytes<double>(st, io, s);
        case DMP_OBS_I16: return launch_bytes<int16_t>(st, io, s);
        case DMP_OBS_REC: return launch_bytes<Rec56>(st, io, s);
        case DMP_OBS_BITS: return launch_bytes<Bits32>(st, io, s);
    }
    return DMP_EINVAL;
}

// dmp_3d_roll.cu -- 3D envs, rollout kernel (K steps per launch): one env per lane, the warp's 32 height
// maps cached in shared memory as NIBBLES.
//
// Why a nibble cache: the step logic only asks "is this cell -1 / 0 / > 0" and the observation wants the raw
// height, which for every reachable state of the reference's plans is tiny (plan height 6; an env would have
// to stack 15 bricks on one cell to leave the nibble range).  212 B per env (instead of 816 B of u16, or 404 B of bytes)
// lets 16 warps share an SM instead of 7 (11 with bytes), which is what this latency-bound kernel needs, and an odd word
// stride (53) makes same-offset accesses of the 32 lanes bank-conflict free.  HBM keeps the same nibbles
// (dmp_common.cuh); every brick is written through.
//   * state in : the warp's 32 nibble maps are one contiguous 6 656 B span (include/dmp.h): 13 coalesced 128-bit
//                loads per lane in one round trip.  An env whose tall flag is set (a height >= 15 somewhere) keeps its
//                wide u16 map in HBM exact as well and reads it only where a nibble says "15 or more" (a build on such
//                a cell, window cells that read 15); an env that reaches that height inside the launch writes its
//                wide map out and carries on like that.
//   * step     : neighbours / walk cells are nibble reads, the 7x7 window is 2 word reads + 1 funnel shift per row;
//   * obs out  : a row's seven nibbles are spread to bytes (2 LOP + 2 PRMT), biased by +1 (0 = frame), each byte
//                is dropped into the mantissa of 2^23 (PRMT) and turned into -1/0/h by one FADD; the warp's
//                [32][51] tile leaves as one contiguous span (one bulk async copy per step);
//   * done     : IoU needs no scan (`cross` is maintained by every build); the warp clears a finished env's
//                map cooperatively in shared memory and in HBM.
// Semantics: Env/3D/DMP_simulator_3d_static_circle.py:67-276 and
// Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-277 (line-by-line citations in the kernel body).
#include "dmp_3d_bulk.cuh"

namespace {

using namespace d3;

// Cache slot of one env: 1 guard word | 50 words = 400 nibbles | 2 guard words = 53 words (odd stride: same-offset accesses
// of the 32 lanes hit 32 banks).  Every read of a lane stays inside its own slot -- a window row that starts up to 3 cells
// before row 0 / ends up to 4 cells behind row 19 reads the guards, rows outside the map re-read the agent's row and are
// masked, frame cells are not read at all -- so no lane ever reads what another lane writes (racecheck-clean).
constexpr int MAP_W = 53;
constexpr int MAP_B = MAP_W * 4;              // 212 B
constexpr int MAP_G0 = 1;                     // guard words in front of the data
constexpr int WARP_MAP_B = 32 * MAP_B;        // 6 784 (multiple of 16)
constexpr size_t SMEM_MAX = 232448;           // 227 KB opt-in limit per block

struct EnvR {
    int pr, pc, plan_idx, cb, cs;
    float ret;
    int cross;      // running sum(min(height, plan)): +1 per brick laid at or below the plan height
};

// cell i of a cached nibble map (i may run into the guards: the caller masks what it reads there)
__device__ __forceinline__ int nib_at(const uint32_t* gw, int i) { return (int)((gw[i >> 3] >> ((i & 7) * 4)) & 0xFu); }

// stage (c): 7x7 window of this lane's env (nibble cache) -> seven rows of biased bytes (height + 1, 0 = frame), two
// words per row.  Row k starts 20 k nibbles after row 0: two word reads and one funnel shift cut its seven nibbles out
// (a valid row reads at most one word before / behind the map: the slot's guards), 2 LOP + 2 PRMT spread them to bytes.
__device__ __forceinline__ void window_cache(const uint32_t* gw, int pr, int pc, uint32_t (&u0)[7], uint32_t (&u1)[7]) {
    const uint32_t cv = (COLVALID >> (pc - 3)) & 0x7Fu;           // window column j lies inside the plan area
    const uint64_t one = spread7(cv);                              // 0x01 per valid byte
    const uint32_t b0 = (uint32_t)one, b1 = (uint32_t)(one >> 32);
    const uint32_t m0 = b0 * 0xFFu, m1 = b1 * 0xFFu;               // 0xFF per valid byte (no carries)
    const int o0 = (pr - 6) * 20 + (pc - 6);                       // nibble offset of window cell (0,0): -63 .. 336
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        const bool rowvalid = (unsigned)(pr - 6 + k) < 20u;        // interior row of window row k
        const int ok = o0 + (rowvalid ? 20 * k : 60);              // a row outside the map re-reads the agent's row (masked)
        const uint32_t* rw = gw + (ok >> 3);
        const uint32_t q = __funnelshift_r(rw[0], rw[1], (ok & 7) * 4);
        uint32_t q0, q1;
        nib8_to_bytes(q, q0, q1);
        u0[k] = rowvalid ? ((q0 & m0) + b0) : 0u;                  // biased bytes: height + 1, 0 = frame
        u1[k] = rowvalid ? ((q1 & m1) + b1) : 0u;
    }
}
template <typename ObsT>
__device__ __forceinline__ void observe_cache(const uint32_t* gw, const uint16_t* ge, bool tall, int pr, int pc, ObsT* row) {
    uint32_t u0[7], u1[7];
    window_cache(gw, pr, pc, u0, u1);
    const bool fix = tall && window_saturated(u0, u1);            // a cell of 15 or more in a tall env's window
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
    if (fix) fix_saturated_row<ObsT>(ge, pr, pc, u0, u1, row);    // exact heights from the wide map
}

// shared memory: [WPB][WARP_MAP_B] nibble maps | [WPB][32*51] ObsT tiles
template <typename ObsT>
__host__ __device__ constexpr size_t warp_smem_bytes() {                 // bit records leave from registers: no tile
    return (size_t)WARP_MAP_B + (is_bits<ObsT>::value ? 0 : 32 * row_elems<ObsT, D3_OBS>() * sizeof(ObsT));
}

// TMA = true: each step's warp tile leaves through one bulk async copy (dmp_common.cuh: warp_tile_bulk_store).
// RF = true (DMP_F_RESET_OBS): finished envs are reset BEFORE the observation is cut (gym-style auto-reset observation).
template <typename ObsT, bool TMA, bool RF>
__global__ void __launch_bounds__(32) k3d_cache_rollout(const DmpState st, const DmpIO io, const int K) {
    extern __shared__ uint4 smem_raw[];
    constexpr bool REC = is_rec<ObsT>::value;
    constexpr int ROW = row_elems<ObsT, D3_OBS>();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t n = st.n_envs;
    const int64_t env0 = ((int64_t)blockIdx.x * wpb + warp) * 32;
    if (env0 >= n) return;                                            // whole warp leaves together
    const int nvalid = (int)min((int64_t)32, n - env0);
    const bool live = lane < nvalid;
    const int64_t env = env0 + (live ? lane : 0);                     // idle lanes shadow env0 but never store

    uint8_t* base = reinterpret_cast<uint8_t*>(smem_raw);
    uint8_t* wmap = base + (size_t)warp * WARP_MAP_B;
    ObsT* tile = reinterpret_cast<ObsT*>(base + (size_t)wpb * WARP_MAP_B) + warp * (32 * ROW);
    uint32_t* gw = reinterpret_cast<uint32_t*>(wmap + lane * MAP_B) + MAP_G0;   // this lane's nibble map (word 0 = cells 0..7)

    uint16_t* cells = reinterpret_cast<uint16_t*>(st.cells);
    uint16_t* gwarp = cells + env0 * CELLS3D;                         // the warp's 32 WIDE maps in HBM (tall envs only)
    uint16_t* ge = cells + env * CELLS3D;                             // this lane's wide map in HBM
    uint8_t* nwarp = nmap3(st) + env0 * NIB3_STRIDE;                  // the warp's 32 nibble maps in HBM (contiguous)
    uint8_t* ne = nmap3(st) + env * NIB3_STRIDE;
    uint4* aux = reinterpret_cast<uint4*>(st.aux);
    const uint8_t* __restrict__ plans = reinterpret_cast<const uint8_t*>(st.plans);
    const bool autoreset = io.flags & DMP_F_AUTORESET;

    // ---- state in: coalesced 128-bit loads of the warp's contiguous maps, packed to bytes -------------
    pdl_launch_dependents();
    pdl_wait();                                                       // the previous launch's state is visible from here
    __syncwarp();
    EnvR e{D2_LO, D2_LO, 0, 0, 0, 0.f, 0};
    bool tall = false;
    double acc_iou = 0.0;                                             // this env's sum of episode IoUs (sequential, exact)
    {
        // the warp's 32 nibble maps are one contiguous 6 656 B span: 13 coalesced 128-bit loads per lane, all in flight at
        // once (ONE DRAM round trip per launch); env el's 13 vectors land in its 51-word slot (the last one carries the
        // 8 pad bytes of the 208 B record: only three of its words are stored)
        const uint4* src = reinterpret_cast<const uint4*>(nwarp);
        constexpr int U = NIB3_STRIDE / 16;                           // 13 vectors per env
        const int nvec = nvalid * U;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = u * 32 + lane;
            v[u] = (idx < nvec) ? src[idx] : make_uint4(0, 0, 0, 0);
        }
        if (live) {                                                   // scalar state rides along
            const uint4 a = aux[env];
            e.pr = a.x & 0x7F; e.pc = (a.x >> 8) & 0xFF; e.plan_idx = a.x >> 16;
            tall = (a.x & AUX3_TALL) != 0u;                           // this env runs from its wide map in HBM
            e.cb = a.y & 0xFFFF; e.cs = a.y >> 16;
            e.ret = __uint_as_float(a.z);
            e.cross = (int)a.w;
            if (autoreset) acc_iou = st.ep_iou[env];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = u * 32 + lane;
            const int el = idx / U, j = idx - el * U;
            uint32_t* dst = reinterpret_cast<uint32_t*>(wmap + el * MAP_B) + MAP_G0 + j * 4;
            dst[0] = v[u].x; dst[1] = v[u].y; dst[2] = v[u].z;       // lanes beyond nvec write zeros into their own slots
            if (j < U - 1) dst[3] = v[u].w;                           // (the last vector's fourth word would be the next slot's)
        }
    }
    int total_brick = __ldg(st.plan_total + e.plan_idx);
    int errbits = 0;
    const bool dynamic = st.dynamic != 0;
    const bool normalise = io.flags & DMP_F_NORMALISE;
    const bool need_draw = (io.actions == nullptr) || (io.step_sizes == nullptr);
    const int tslot = (io.flags & DMP_F_TSLOT1) ? 1 : 0;
    const uint64_t t0 = st.t_dev ? st.t_dev[tslot] : st.t;
    const uint64_t gid = (uint64_t)(st.env_base + env0) + (uint64_t)lane;
    uint32_t acc_cnt = 0, acc_len = 0;                                // episodes finished by this env in this launch
    float acc_ret = 0.f;                                              // (integer-valued: exact in any order)
    __syncwarp();
    int64_t idx = env0 + lane;                                        // flat [k][env] index of this step's outputs

    StepDraws draws;
    bool bulk_pending = false;
    for (int k = 0; k < K; ++k, idx += n) {
        const uint64_t t = t0 + (uint64_t)k;
        uint32_t dw = 0;
        if (need_draw) dw = draws.word(st.seed, gid, t);
        int a, s;
        if (io.actions) a = live ? (int)io.actions[idx] : 0; else a = draw_action(dw, D3_ACT, st.action_dist);
        if (io.step_sizes) s = live ? (int)io.step_sizes[idx] : 1; else s = draw_step_size(dw);
        if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;
        if (a > 7) errbits |= DMP_ERR_ACTION;              // reference: an unbuilt brick (:187-208)

        e.cs += 1;
        // ---- the six cells this step can depend on: the four neighbours (check_sur :88-102) and the second and
        // third cell in the action's direction (move_step :104-134); whether a cell is frame follows from ONE
        // coordinate, the other is the agent's.
        const int dir = a & 3, dr = dir_dr(dir), dc = dir_dc(dir);
        const int o = (e.pr - 3) * 20 + (e.pc - 3);
        int c6[6];
        {
            const int dstep = dr * 20 + dc, sgn = dr + dc;
            const int coord = (dir < 2 ? e.pc : e.pr) - 3;
            // (predicated reads: a frame cell is not read, so every read stays inside the env's own map)
            c6[0] = (e.pc > D2_LO) ? nib_at(gw, o - 1) : -1;
            c6[1] = (e.pc < D2_HI) ? nib_at(gw, o + 1) : -1;
            c6[2] = (e.pr < D2_HI) ? nib_at(gw, o + 20) : -1;
            c6[3] = (e.pr > D2_LO) ? nib_at(gw, o - 20) : -1;
            c6[4] = ((unsigned)(coord + 2 * sgn) < 20u) ? nib_at(gw, o + 2 * dstep) : -1;
            c6[5] = ((unsigned)(coord + 3 * sgn) < 20u) ? nib_at(gw, o + 3 * dstep) : -1;
        }
        const bool boxed = (c6[0] != 0) && (c6[1] != 0) && (c6[2] != 0) && (c6[3] != 0);     // check_sur
        int nsel = (dir == 0) ? c6[0] : (dir == 1) ? c6[1] : (dir == 2) ? c6[2] : c6[3];

        bool done = false, tail = true;
        bool built = false, boxed_penalty = false, turns_tall = false;
        int newh = 0, pplan = 0;
        if (a <= 3) {
            // (a) move_step (:104-134): consecutive empty cells, at most s
            int nstep = 0;
            if (nsel == 0) nstep = (s >= 2 && c6[4] == 0) ? ((s >= 3 && c6[5] == 0) ? 3 : 2) : 1;
            e.pr = min(max(e.pr + dr * nstep, D2_LO), D2_HI);
            e.pc = min(max(e.pc + dc * nstep, D2_LO), D2_HI);
        } else {
            // (b) build on neighbour a-4 unless it is frame
            bool open_after = (c6[0] == 0) || (c6[1] == 0) || (c6[2] == 0) || (c6[3] == 0);   // some neighbour still empty
            if (a <= 7 && nsel != -1) {
                const int ti = o + dr * 20 + dc;
                if (nsel == 15) nsel = (int)__ldcg(ge + ti);                   // saturated nibble (tall envs only): true height
                built = true;
                newh = nsel + 1;
                e.cb += 1;
                pplan = ldg_u8(plans + e.plan_idx * CELLS3D + ti);             // consumed after the observation
                uint32_t* wp = gw + (ti >> 3);
                const int sh = (ti & 7) * 4;
                const uint32_t x = (*wp & ~(0xFu << sh)) | ((uint32_t)sat_nib(newh) << sh);
                *wp = x;                                                       // the cache stays current for tall envs too
                if (live) {
                    ne[ti >> 1] = (uint8_t)(x >> (((ti >> 1) & 3) * 8));         // write-through: the byte holding the nibble
                    if (tall) {
                        ge[ti] = (uint16_t)newh;                               // a tall env keeps its wide map exact
                    } else if (newh >= TALL3) {                                // the nibbles stop being exact after this brick:
                        turns_tall = tall = true;                              // the env turns tall, its wide map is made
                    }                                                          // current below (by the warp)
                }
                // neighbours after placement: only neighbour `dir` changed, and it is now > 0
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
        if (__any_sync(FULL, turns_tall)) {                 // wide map := the (still exact) nibble cache, 16 cells per lane
            __syncwarp();
            for (unsigned tw = __ballot_sync(FULL, turns_tall); tw; tw &= tw - 1) {
                const int src = __ffs(tw) - 1;
                const uint32_t* cw = reinterpret_cast<const uint32_t*>(wmap + src * MAP_B) + MAP_G0;
                if (lane < 25) warp_widen_words(gwarp + src * CELLS3D, cw[2 * lane], cw[2 * lane + 1], lane);
            }
            __syncwarp();
        }

        // ---- (c) observation; (d) reward (reward_check :232-239) -------------------------------------------
        // The plan byte requested at build time is consumed only after the window has been formatted, which hides its
        // latency.  A record row (DMP_OBS_REC) carries the reward, so it is closed after (d); observation rows leave first.
        float reward = 0.f;
        bool rewarded = false;
        auto reward_now = [&]() {
            if constexpr (RF) {                          // called twice in that mode: before the reset and by the record path
                if (rewarded) return;
                rewarded = true;
            }
            if (built) {
                if (newh <= pplan) e.cross += 1;
                if (!tail && !done) reward = (newh > pplan) ? -1.f : (newh == pplan ? 10.f : 1.f);
            }
            if (boxed_penalty) reward = -100.f;
        };
        auto finish_all = [&]() {
            // ---- (e) finished episodes --------------------------------------------------------------------
            // IoU = cross / (total_brick + count_brick - cross) (:257-276); `cross` is kept up to date by every build.
            // Episode statistics accumulate in registers and are folded into HBM once, after the last step.
            const bool fin = done && autoreset && live;
            const bool fin_wide = fin && tall;                  // a tall env's wide map is cleared with it
            if (fin) {
                const int den = total_brick + e.cb - e.cross;
                const double iou = (e.cross == 0 && den != 0) ? 0.0 : __ddiv_rn((double)e.cross, (double)den);
                acc_cnt += 1u; acc_len += (uint32_t)e.cs; acc_ret += e.ret; acc_iou += iou;
                if (io.next_plan) {
                    const int p = io.next_plan[idx];
                    if ((unsigned)p >= (unsigned)st.n_plans) errbits |= DMP_ERR_PLANIDX; else e.plan_idx = p;
                } else if (st.plan_mode == DMP_PLAN_PHILOX) {
                    e.plan_idx = draw_plan(plan_word(st.seed, gid, t), st.n_plans);
                } else if (st.plan_mode == DMP_PLAN_SEQUENTIAL) {
                    e.plan_idx = (e.plan_idx + 1 == st.n_plans) ? 0 : e.plan_idx + 1;
                }
                total_brick = __ldg(st.plan_total + e.plan_idx);
                e.pr = e.pc = D2_LO; e.cb = e.cs = 0; e.ret = 0.f; e.cross = 0;
                tall = false;
            }
            __syncwarp();                                       // slots are cleared by OTHER lanes: this lane's brick write and
            unsigned dm = __ballot_sync(FULL, fin);             // window reads are ordered before them
            const unsigned dmw = __ballot_sync(FULL, fin_wide);
            while (dm) {                                        // the warp clears each finished env's map
                const int src = __ffs(dm) - 1;
                dm &= dm - 1;
                uint32_t* sg = reinterpret_cast<uint32_t*>(wmap + src * MAP_B) + lane;      // the whole slot, guards included
                sg[0] = 0u;
                if (lane < MAP_W - 32) sg[32] = 0u;
                const uint4 z = make_uint4(0, 0, 0, 0);
                if (lane < NIB3_STRIDE / 16) reinterpret_cast<uint4*>(nwarp + src * NIB3_STRIDE)[lane] = z;
                if (lane < 25 && ((dmw >> src) & 1u)) {
                    uint4* gg = reinterpret_cast<uint4*>(gwarp + src * CELLS3D) + 2 * lane;
                    gg[0] = z; gg[1] = z;
                }
            }
            __syncwarp();
        };
        if constexpr (RF) {                                 // reward and reset first: the observation below is the reset env's
            reward_now();
            e.ret += reward;
            finish_all();
        }
        if constexpr (is_bits<ObsT>::value) {
            if (io.obs) {
                uint32_t u0[7], u1[7];
                window_cache(gw, e.pr, e.pc, u0, u1);
                reward_now();
                if (live) bits32_store(io.obs, idx, u0, u1, e.cb, e.cs, reward, done);
            } else {
                reward_now();
            }
        } else if (io.obs) {
            if (TMA && bulk_pending) { warp_tile_bulk_wait(lane); bulk_pending = false; }   // previous copy has drained the tile
            ObsT* row = tile + lane * ROW;
            if constexpr (REC) {
                uint64_t c[7];
                uint32_t u0[7], u1[7];
                window_cache(gw, e.pr, e.pc, u0, u1);
#pragma unroll
                for (int k = 0; k < 7; ++k) c[k] = (uint64_t)u0[k] | ((uint64_t)u1[k] << 32);
                if (tall && window_saturated(u0, u1)) fix_saturated_codes(ge, e.pr, e.pc, c);
                uint32_t w[13];
                pack49(c, w);
                reward_now();
                rec56_store(row, w, e.cb, e.cs, reward, done, tall);
            } else {
                observe_cache<ObsT>(gw, ge, tall, e.pr, e.pc, row);
                obs_counters<ObsT>(normalise, e.cb, e.cs, total_brick, st.total_step, row[49], row[50]);
            }
            ObsT* dst = reinterpret_cast<ObsT*>(io.obs) + ((int64_t)k * n + env0) * ROW;
            if (TMA && nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                warp_tile_bulk_store(dst, tile, 32 * ROW * sizeof(ObsT), lane);
                bulk_pending = true;
            } else {
                __syncwarp();
                if (nvalid == 32) tile_rows_store_full<ObsT, D3_OBS>(dst, tile, lane);
                else tile_rows_store<ObsT, D3_OBS>(dst, tile, nvalid, lane);
                __syncwarp();
            }
            if constexpr (!REC) reward_now();
        } else {
            reward_now();
        }
        if constexpr (!RF) e.ret += reward;
        if (live) {
            if (io.reward) io.reward[idx] = reward;
            if (io.done) io.done[idx] = done ? 1 : 0;
        }

        if constexpr (!RF) finish_all();
    }
    if (live) {
        if ((e.cb | e.cs) > 0xFFFF) {                                 // 16-bit packed counters (include/dmp.h)
            errbits |= DMP_ERR_OVERFLOW;
            e.cb = min(e.cb, 0xFFFF); e.cs = min(e.cs, 0xFFFF);
        }
        aux[env] = make_uint4((uint32_t)e.pr | (tall ? AUX3_TALL : 0u) | ((uint32_t)e.pc << 8) | ((uint32_t)e.plan_idx << 16),
                              (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16), __float_as_uint(e.ret), (uint32_t)e.cross);
        if (acc_cnt) {                                      // this thread is the only writer of its env's statistics
            atomicAdd(st.ep_cnt + env, acc_cnt);            // fire-and-forget REDs
            atomicAdd(st.ep_len + env, acc_len);
            atomicAdd(st.ep_ret + env, (double)acc_ret);
            st.ep_iou[env] = acc_iou;
        }
        if (errbits) atomicOr(st.err, errbits);
        if (st.t_dev && env == 0) st.t_dev[tslot ^ 1] = t0 + (uint64_t)K;
    }
    if (TMA && bulk_pending) warp_tile_bulk_wait(lane);                 // the tile must outlive the copy that reads it
}

// Launch shape.  Every warp is an independent tile of 32 envs, so blocks are single warps: the block scheduler then refills
// an SM warp by warp instead of waiting for the slowest warp of a big block (three-warp blocks: 16.6 vs 19.3 G env-steps/s).
// Shared memory limits residency to 16 warps per SM for f32 observations (6.6 KB of maps + 6.5 KB of tile per warp).
template <typename ObsT, bool TMA, bool RF>
int launch_cache_r(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    constexpr int wpb = 1;
    const size_t smem = (size_t)wpb * warp_smem_bytes<ObsT>() + 16;
    static_assert(warp_smem_bytes<ObsT>() + 16 <= SMEM_MAX, "one warp's byte cache + tile must fit a block");
    static bool attr_done = false;                           // per instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k3d_cache_rollout<ObsT, TMA, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return dmp_set_error(e);
        e = cudaFuncSetAttribute(k3d_cache_rollout<ObsT, TMA, RF>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return dmp_set_error(e);
        attr_done = true;
    }
    const int64_t warps = (st.n_envs + 31) / 32;
    const unsigned blocks = (unsigned)((warps + wpb - 1) / wpb);
    return dmp_set_error(dmp_launch_pdl(!(io.flags & DMP_F_NO_PDL), k3d_cache_rollout<ObsT, TMA, RF>, blocks,
                                        (unsigned)(wpb * 32), smem, s, st, io, K));
}

template <typename ObsT, bool TMA>
int launch_cache_t(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    if (io.flags & DMP_F_RESET_OBS) return launch_cache_r<ObsT, TMA, true>(st, io, K, s);
    return launch_cache_r<ObsT, TMA, false>(st, io, K, s);
}

// the observation tile leaves through one bulk async copy per warp and step unless DMP_F_TILE_LDST asks for load/store pairs
template <typename ObsT>
int launch_cache(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    if (io.flags & DMP_F_TILE_LDST) return launch_cache_t<ObsT, false>(st, io, K, s);
    return launch_cache_t<ObsT, true>(st, io, K, s);
}

}  // namespace

int dmp3d_cache_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    switch (io.obs_kind) {
        case DMP_OBS_F32: return launch_cache<float>(st, io, K, s);
        case DMP_OBS_F64: return launch_cache<double>(st, io, K, s);
        case DMP_OBS_I16: return launch_cache<int16_t>(st, io, K, s);
        case DMP_OBS_REC: return launch_cache<Rec56>(st, io, K, s);
        case DMP_OBS_BITS: return launch_cache_t<Bits32, false>(st, io, K, s);
    }
    return DMP_EINVAL;
}

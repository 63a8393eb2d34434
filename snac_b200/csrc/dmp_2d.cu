// dmp_2d.cu -- 2D mobile-construction envs (20x20 occupancy grid, 7x7 window, 5 actions).
//
// Reference semantics: Env/2D/DMP_Env_2D_static.py:54-154 and
// Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:34-147 (identical step logic; the dynamic class
// only changes where the plan comes from and normalises the two counter columns).
//
// Mapping: one env per thread.  The 64 B state record is fetched with four coalesced 128-bit
// loads (SoA planes), the 400 occupancy bits are parked in shared memory ([word][thread], bank =
// lane, so the data-dependent row select is conflict free), the step is applied, the 7x7 window is
// cut out of the bit grid with funnel shifts, expanded into the warp's [32][51] observation tile in
// shared memory and streamed out as one contiguous, 16 B aligned span with 128-bit stores.
// K > 1 keeps the state on chip between steps (rollout mode).
#include "dmp_common.cuh"

namespace {

// Threads (= envs) per block is a template parameter BT.  Shared memory per env: 13 grid words + the 51-word row of the warp
// tile = 256 B, so an SM holds 6 blocks of 128 threads (80 registers) or 4 blocks of 224 threads (72 registers; 4 x (224 x 256
// + 1 KB reserved) is exactly the SM's 228 KB): 768 or 896 resident envs per SM.  The launcher picks the shape that covers
// the shard in ONE wave where the other needs two -- BASELINE's 1 048 576 envs over 8 GPUs are 131 072 per GPU: 1.15 waves
// of 128-thread blocks, 0.99 of 224-thread blocks (wide_blocks2 below has the measurements).
constexpr int S2_WORDS = GRID2D_WORDS;  // smem words per env: the 13 grid words, [word][thread]
constexpr uint32_t COLVALID = 0x7FFFF8u;  // padded columns 3..22 are inside the plan area

struct Env2 {
    int pr, pc, plan_idx, cb, cs;
    float ret;
};

__device__ __forceinline__ void unpack2(const uint4& v3, Env2& e) {
    e.pr = v3.y & 0xFF;
    e.pc = (v3.y >> 8) & 0xFF;
    e.plan_idx = v3.y >> 16;
    e.cb = v3.z & 0xFFFF;
    e.cs = v3.z >> 16;
    e.ret = __uint_as_float(v3.w);
}
__device__ __forceinline__ void pack2(const Env2& e, uint32_t word12, uint4& v3) {
    v3.x = word12;
    v3.y = (uint32_t)e.pr | ((uint32_t)e.pc << 8) | ((uint32_t)e.plan_idx << 16);
    v3.z = (uint32_t)(e.cb & 0xFFFF) | ((uint32_t)e.cs << 16);
    v3.w = __float_as_uint(e.ret);
}

// stage (a): move + clamp.  clip_position, Env/2D/DMP_Env_2D_static.py:84-93; action map :100-117
__device__ __forceinline__ void stage_move2(Env2& e, int a, int s) {
    int r = e.pr, c = e.pc;
    if (a == 0) c -= s;
    else if (a == 1) c += s;
    else if (a == 2) r += s;
    else r -= s;
    e.pr = min(max(r, D2_LO), D2_HI);
    e.pc = min(max(c, D2_LO), D2_HI);
}

// stage (b): brick deposition into the bit grid (g = this thread's smem column).  Returns the
// previous occupancy of the cell (nonzero = was already occupied); :120-125, :129-130, :143-144
// (increment then clip to 1 == OR of one bit).
template <int BT>
__device__ __forceinline__ uint32_t stage_deposit2(uint32_t* g, const Env2& e, int& word, uint32_t& bit) {
    const int b = (e.pr - D2_HW) * D2_W + (e.pc - D2_HW);
    word = b >> 5;
    bit = 1u << (b & 31);
    const uint32_t old = g[word * BT];
    g[word * BT] = old | bit;
    return old & bit;
}

// stage (c): observation window.  observation_ :78-82 + hstack.
// The seven window rows as padded rows (bits 3..22 valid, rest garbage): rows pr-3 .. pr+3 of the padded grid are interior rows
// pr-6 .. pr: a 143-bit span [B0, B0+143) of the bit grid that starts up to 63 bits before it and ends up to 124 bits behind it.
// Starting 3 bits early puts interior column c of every row at bit c+3 = its padded column.  Words outside the grid are read
// at a clamped index: whatever they hold only reaches rows that are masked as frame by the callers (a valid row's 20 bits
// always lie inside words 0..12).
template <int BT>
__device__ __forceinline__ void window_rows2(const uint32_t* g, const Env2& e, uint32_t (&R)[7]) {
    const int B0 = (e.pr - 2 * D2_HW) * D2_W - D2_HW;                  // -63 .. 317
    const int w0 = B0 >> 5, off = B0 & 31;                               // arithmetic shift: floor for negative B0
    auto gw = [&](int j) { return g[min(max(w0 + j, 0), GRID2D_WORDS - 1) * BT]; };
    const uint32_t x0 = gw(0), x1 = gw(1), x2 = gw(2), x3 = gw(3), x4 = gw(4), x5 = gw(5);
    const uint32_t q0 = __funnelshift_r(x0, x1, off), q1 = __funnelshift_r(x1, x2, off);
    const uint32_t q2 = __funnelshift_r(x2, x3, off), q3 = __funnelshift_r(x3, x4, off);
    const uint32_t q4 = __funnelshift_r(x4, x5, off);
    R[0] = q0;
    R[1] = __funnelshift_r(q0, q1, 20);
    R[2] = q1 >> 8;
    R[3] = __funnelshift_r(q1, q2, 28);
    R[4] = __funnelshift_r(q2, q3, 16);
    R[5] = q3 >> 4;
    R[6] = __funnelshift_r(q3, q4, 24);
}

// -> this thread's row of the warp tile.  For a record row (Rec56) the seven row codes ARE the window bytes; reward and done
// ride along.
template <typename ObsT, int BT>
__device__ __forceinline__ void stage_observe2(const uint32_t* g, const Env2& e, ObsT* row,
                                               bool normalise, int total_brick, int total_step,
                                               float reward = 0.f, bool done = false) {
    uint32_t R[7];
    window_rows2<BT>(g, e, R);
    const int sh = e.pc - D2_HW;                       // window column 0 = padded column pc-3
    const uint32_t colvalid = (COLVALID >> sh) & 0x7Fu;
    const uint64_t vcode = spread7(colvalid);          // 1 where the window column is inside the plan area
    // inside: 0/1 ; frame: -1  (environment_memory[...] = -1, :61-64)  ->  code = value + 1
    auto row_code = [&](int k) -> uint64_t {
        const int p = e.pr - D2_HW + k;                // padded row of window row k
        const bool rowvalid = (unsigned)(p - D2_HW) < (unsigned)D2_W;
        const uint32_t occ = (R[k] >> sh) & colvalid;
        return rowvalid ? spread7(occ) + vcode : 0ull;
    };
    if constexpr (is_rec<ObsT>::value) {
        uint64_t codes[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) codes[k] = row_code(k);
        uint32_t w[13];
        pack49(codes, w);
        rec56_store(row, w, e.cb, e.cs, reward, done, false);
    } else {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const uint64_t code = row_code(k);
            const uint32_t lo = (uint32_t)code, hi = (uint32_t)(code >> 32);
            ObsT* o = row + k * 7;
            o[0] = obs_from_biased<ObsT, 0>(lo);
            o[1] = obs_from_biased<ObsT, 1>(lo);
            o[2] = obs_from_biased<ObsT, 2>(lo);
            o[3] = obs_from_biased<ObsT, 3>(lo);
            o[4] = obs_from_biased<ObsT, 0>(hi);
            o[5] = obs_from_biased<ObsT, 1>(hi);
            o[6] = obs_from_biased<ObsT, 2>(hi);
        }
        obs_counters<ObsT>(normalise, e.cb, e.cs, total_brick, total_step, row[49], row[50]);
    }
}

// DMP_OBS_BITS: the same window as 49 two-bit codes (0 frame / 1 empty / 2 occupied, row-major from bit 0), then the 29-bit
// trailer at bit 98 -- one 128-bit word per env, built in registers.  Two rows share a register while their bits are spread
// (bit j -> bit 2j in each 16-bit half); the code of a cell is spread(valid) + spread(occupied).
template <int BT>
__device__ __forceinline__ uint4 observe2_bits(const uint32_t* g, const Env2& e, float reward, bool done) {
    uint32_t R[7];
    window_rows2<BT>(g, e, R);
    const int sh = e.pc - D2_HW;
    const uint32_t colvalid = (COLVALID >> sh) & 0x7Fu;
    const uint32_t rowvalid = (COLVALID >> (e.pr - D2_HW)) & 0x7Fu;     // bit k: window row k is inside the plan area
    const uint32_t vv = spread2x7(colvalid | (colvalid << 16));
    auto occ = [&](int k) { return (R[k] >> sh) & colvalid; };
    auto pair = [&](int k) -> uint32_t {                                // codes of rows k (low half) and k + 1 (high half)
        const uint32_t m = ((rowvalid >> k) & 1u ? 0x0000FFFFu : 0u) | ((rowvalid >> (k + 1)) & 1u ? 0xFFFF0000u : 0u);
        return (spread2x7(occ(k) | (occ(k + 1) << 16)) + vv) & m;
    };
    const uint32_t p01 = pair(0), p23 = pair(2), p45 = pair(4);
    const uint32_t c6 = ((rowvalid >> 6) & 1u) ? ((spread2x7(occ(6)) + vv) & 0x3FFFu) : 0u;
    const uint32_t c0 = p01 & 0xFFFFu, c1 = p01 >> 16, c2 = p23 & 0xFFFFu, c3 = p23 >> 16, c4 = p45 & 0xFFFFu, c5 = p45 >> 16;
    uint4 r;
    r.x = c0 | (c1 << 14) | (c2 << 28);
    r.y = (c2 >> 4) | (c3 << 10) | (c4 << 24);
    r.z = (c4 >> 8) | (c5 << 6) | (c6 << 20);
    // a 2D reward is 0 or 5 (code 0 / 2): no float-to-int conversion, no select chain
    r.w = (c6 >> 12) | (bits_trailer_code(e.cb, e.cs, reward != 0.f ? 2u : 0u, done, false) << 2);
    return r;
}

// stage (d) helper: IoU = |G & P| / |G | P| over the interior (render :169-175) with warp-free popc.
template <int BT>
__device__ __forceinline__ double iou2(const uint32_t* g, const uint32_t* __restrict__ plan) {
    int inter = 0, uni = 0;
#pragma unroll
    for (int w = 0; w < GRID2D_WORDS; ++w) {
        const uint32_t gw = g[w * BT], pw = __ldg(plan + w);
        inter += __popc(gw & pw);
        uni += __popc(gw | pw);
    }
    return __ddiv_rn((double)inter, (double)uni);
}

// TMA = true: the warp tile leaves through one bulk async copy (dmp_common.cuh: warp_tile_bulk_store) instead of 13 x
// (LDS.128 + STG.128) per lane: the L1 data pipe was the busiest unit of this kernel (ncu: 69 %).
// ObsT = float / double / int16_t: [n][51] observation rows; ObsT = Rec56: one packed 56 B step record per env.
// RF = true (DMP_F_RESET_OBS): a finished env is reset BEFORE its observation is cut, so the row it writes is the next
// episode's first policy input.  A template parameter, not a run-time branch: two inlined copies of the reset block
// cost the kernel registers it does not have (it sits at its 80-register cap).
template <typename ObsT, bool TMA, int BT, bool RF>
__global__ void __launch_bounds__(BT, BT == 128 ? 6 : 4) k2d_rollout(const DmpState st, const DmpIO io, const int K) {
    constexpr int ROW = row_elems<ObsT, D2_OBS>();                       // tile elements per env (51 values or 1 record)
    constexpr int B2 = BT;
    extern __shared__ uint4 smem_raw[];
    uint32_t* G = reinterpret_cast<uint32_t*>(smem_raw);                 // [S2_WORDS][BT]
    ObsT* tiles = reinterpret_cast<ObsT*>(G + S2_WORDS * B2);            // [BT/32][32*ROW]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * B2 + tid;
    const int64_t env0 = env - lane;                                     // first env of this warp
    const int nvalid = (int)min((int64_t)32, n - env0);                  // <= 0: idle warp
    const bool live = env < n;
    ObsT* tile = tiles + warp * (32 * ROW);
    uint32_t* g = G + tid;

    uint4* cells = reinterpret_cast<uint4*>(st.cells);
    const uint32_t* __restrict__ plans = reinterpret_cast<const uint32_t*>(st.plans);
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0, v2 = v0, v3 = v0;
    const uint64_t keep = l2_policy_keep();
    const bool hint = !(io.flags & DMP_F_NO_L2_HINT);
    pdl_launch_dependents();
    pdl_wait();                                                          // the previous step's state is visible from here
    if (live) {
        if (hint) {
            v0 = ldg_keep(cells + env, keep);
            v1 = ldg_keep(cells + n + env, keep);
            v2 = ldg_keep(cells + 2 * n + env, keep);
            v3 = ldg_keep(cells + 3 * n + env, keep);
        } else {
            v0 = cells[env]; v1 = cells[n + env]; v2 = cells[2 * n + env]; v3 = cells[3 * n + env];
        }
    }
    g[0 * B2] = v0.x;  g[1 * B2] = v0.y;  g[2 * B2] = v0.z;  g[3 * B2] = v0.w;
    g[4 * B2] = v1.x;  g[5 * B2] = v1.y;  g[6 * B2] = v1.z;  g[7 * B2] = v1.w;
    g[8 * B2] = v2.x;  g[9 * B2] = v2.y;  g[10 * B2] = v2.z; g[11 * B2] = v2.w;
    g[12 * B2] = v3.x;
    Env2 e;
    unpack2(v3, e);
    if (!live) { e.pr = e.pc = D2_LO; e.plan_idx = 0; }
    int total_brick = __ldg(st.plan_total + e.plan_idx);
    unsigned dirty = 0;                                                   // bit v: plane v must be written back
    int errbits = 0;

    const bool autoreset = io.flags & DMP_F_AUTORESET;
    const bool normalise = io.flags & DMP_F_NORMALISE;
    const bool need_draw = (io.actions == nullptr) || (io.step_sizes == nullptr);
    const int tslot = (io.flags & DMP_F_TSLOT1) ? 1 : 0;
    const uint64_t t0 = st.t_dev ? st.t_dev[tslot] : st.t;

    StepDraws draws;

    for (int k = 0; k < K; ++k) {
        const uint64_t t = t0 + (uint64_t)k;
        const int64_t idx = (int64_t)k * n + env;
        uint32_t dw = 0;
        if (need_draw) dw = draws.word(st.seed, (uint64_t)(st.env_base + env), t);
        int a, s;
        if (io.actions) a = live ? io.actions[idx] : 0; else a = draw_action(dw, D2_ACT, DMP_ACT_UNIFORM);
        if (io.step_sizes) s = live ? io.step_sizes[idx] : 1; else s = draw_step_size(dw);
        if ((unsigned)(s - 1) > 2u) errbits |= DMP_ERR_STEPSIZE;

        // ---- step(): Env/2D/DMP_Env_2D_static.py:95-154 ------------------------------------
        e.cs += 1;
        float reward = 0.f;
        bool done;
        if (a < 4) {                                        // (a) move
            stage_move2(e, a, s);
            done = e.cs >= st.total_step;
        } else if (a == 4) {                                // (b) drop + (d) reward
            e.cb += 1;
            int word; uint32_t bit;
            const uint32_t was = stage_deposit2<BT>(g, e, word, bit);
            dirty |= 1u << (word >> 2);
            if (e.cb >= total_brick) {                      // :127-135 budget exhausted: reward 0.0
                done = true;
            } else {                                        // :137-147
                done = e.cs >= st.total_step;
                const uint32_t pw = __ldg(plans + e.plan_idx * PLAN2D_WORDS + word);
                reward = (!was && (pw & bit)) ? 5.f : 0.f;  // pre-clip value == plan  <=> first brick on a plan cell
            }
        } else {                                            // reference: UnboundLocalError
            errbits |= DMP_ERR_ACTION;
            done = e.cs >= st.total_step;
        }
        e.ret += reward;

        // ---- (e) done / auto-reset: fold the episode into the statistics, clear the state, next plan -------------
        const bool fin = done && autoreset && live;
        auto finish = [&]() {
            const double iou = iou2<BT>(g, plans + e.plan_idx * PLAN2D_WORDS);
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
#pragma unroll
            for (int w = 0; w < GRID2D_WORDS; ++w) g[w * B2] = 0;
            e.pr = e.pc = D2_LO;
            e.cb = e.cs = 0;
            e.ret = 0.f;
            dirty = 0xFu;
        };
        if constexpr (RF) { if (fin) finish(); }            // DMP_F_RESET_OBS: the observation below is the reset env's

        // ---- (c) observation ------------------------------------------------------------------
        if constexpr (is_bits<ObsT>::value) {               // one 128-bit record per env, straight from registers
            if (io.obs && live) __stcs(reinterpret_cast<uint4*>(io.obs) + idx, observe2_bits<BT>(g, e, reward, done));
        } else if (io.obs) {
            ObsT* dst = reinterpret_cast<ObsT*>(io.obs) + ((int64_t)k * n + env0) * ROW;
            if constexpr (TMA) {
                if (k > 0) warp_tile_bulk_wait(lane);               // the previous step's copy has drained the tile
                stage_observe2<ObsT, BT>(g, e, tile + lane * ROW, normalise, total_brick, st.total_step, reward, done);
                if (nvalid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                    warp_tile_bulk_store(dst, tile, 32 * ROW * sizeof(ObsT), lane);
                } else {
                    __syncwarp();
                    if (nvalid > 0) tile_rows_store<ObsT, D2_OBS>(dst, tile, nvalid, lane);
                }
            } else {
                stage_observe2<ObsT, BT>(g, e, tile + lane * ROW, normalise, total_brick, st.total_step, reward, done);
                __syncwarp();
                if (nvalid == 32) tile_rows_store_full<ObsT, D2_OBS>(dst, tile, lane);
                else if (nvalid > 0) tile_rows_store<ObsT, D2_OBS>(dst, tile, nvalid, lane);
            }
            __syncwarp();
        }
        if (live) {
            if (io.reward) io.reward[idx] = reward;
            if (io.done) io.done[idx] = done ? 1 : 0;
        }

        if constexpr (!RF) { if (fin) finish(); }
    }

    if (live) {
        if ((e.cb | e.cs) > 0xFFFF) {                                    // 16-bit packed counters (include/dmp.h)
            errbits |= DMP_ERR_OVERFLOW;
            e.cb = min(e.cb, 0xFFFF); e.cs = min(e.cs, 0xFFFF);
        }
        pack2(e, g[12 * B2], v3);
        if (hint) {
            if (dirty & 1u) stg_keep(cells + env, make_uint4(g[0 * B2], g[1 * B2], g[2 * B2], g[3 * B2]), keep);
            if (dirty & 2u) stg_keep(cells + n + env, make_uint4(g[4 * B2], g[5 * B2], g[6 * B2], g[7 * B2]), keep);
            if (dirty & 4u) stg_keep(cells + 2 * n + env, make_uint4(g[8 * B2], g[9 * B2], g[10 * B2], g[11 * B2]), keep);
            stg_keep(cells + 3 * n + env, v3, keep);
        } else {
            if (dirty & 1u) cells[env] = make_uint4(g[0 * B2], g[1 * B2], g[2 * B2], g[3 * B2]);
            if (dirty & 2u) cells[n + env] = make_uint4(g[4 * B2], g[5 * B2], g[6 * B2], g[7 * B2]);
            if (dirty & 4u) cells[2 * n + env] = make_uint4(g[8 * B2], g[9 * B2], g[10 * B2], g[11 * B2]);
            cells[3 * n + env] = v3;
        }
        if (errbits) atomicOr(st.err, errbits);
    }
    if (st.t_dev && blockIdx.x == 0 && tid == 0) st.t_dev[tslot ^ 1] = t0 + (uint64_t)K;
    if constexpr (TMA && !is_bits<ObsT>::value) warp_tile_bulk_wait(lane);   // the tile must outlive the copy that reads it
}

// ---------------------------------------------------------------------------------------------
// reset / iou / export / import: thread per env, no shared memory
// ---------------------------------------------------------------------------------------------
template <typename ObsT>
__global__ void k2d_reset(const DmpState st, const uint8_t* __restrict__ mask, const int32_t* __restrict__ plan_idx,
                          const uint64_t t_draw, ObsT* __restrict__ obs) {
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n || (mask && !mask[env])) return;
    uint4* cells = reinterpret_cast<uint4*>(st.cells);
    int p;
    if (plan_idx) {
        p = plan_idx[env];
        if ((unsigned)p >= (unsigned)st.n_plans) { atomicOr(st.err, DMP_ERR_PLANIDX); p = 0; }
    } else if (st.plan_mode == DMP_PLAN_PHILOX) {
        p = draw_plan(env_draw(st.seed, (uint64_t)(st.env_base + env), t_draw).x3, st.n_plans);
    } else {
        const uint32_t w13 = cells[3 * n + env].y;
        p = (int)(w13 >> 16);
        // sequential order starts at plan 0 on an env that has never been reset (zeroed state: row 0 is not a position),
        // index_for_non_random = 0 of Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:39-44
        if (st.plan_mode == DMP_PLAN_SEQUENTIAL) p = ((w13 & 0xFFu) == 0u) ? 0 : ((p + 1 >= st.n_plans) ? 0 : p + 1);
        if ((unsigned)p >= (unsigned)st.n_plans) p = 0;
    }
    const uint4 z = make_uint4(0, 0, 0, 0);
    cells[env] = z; cells[n + env] = z; cells[2 * n + env] = z;
    Env2 e{D2_LO, D2_LO, p, 0, 0, 0.f};
    uint4 v3; pack2(e, 0u, v3);
    cells[3 * n + env] = v3;
    if (obs) {                       // window at [3,3] of an empty grid: rows/cols 0..2 are frame
        if constexpr (is_bits<ObsT>::value) {
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            for (int i = 0; i < 49; ++i)
                if (i / 7 >= 3 && i % 7 >= 3) w[(2 * i) >> 5] |= 1u << ((2 * i) & 31);
            reinterpret_cast<uint4*>(obs)[env] = make_uint4(w[0], w[1], w[2], w[3]);
        } else if constexpr (is_rec<ObsT>::value) {
            uint8_t* o = reinterpret_cast<uint8_t*>(obs + env);
            for (int i = 0; i < 56; ++i) o[i] = (i < 49 && i / 7 >= 3 && i % 7 >= 3) ? 1 : 0;
        } else {
            ObsT* o = obs + env * D2_OBS;
            for (int k = 0; k < 7; ++k)
                for (int j = 0; j < 7; ++j) o[k * 7 + j] = obs_from_int<ObsT>((k < 3 || j < 3) ? -1 : 0);
            o[49] = obs_from_int<ObsT>(0);
            o[50] = obs_from_int<ObsT>(0);
        }
    }
}

__global__ void k2d_iou(const DmpState st, double* __restrict__ out) {
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    const uint4* cells = reinterpret_cast<const uint4*>(st.cells);
    const uint4 v[4] = {cells[env], cells[n + env], cells[2 * n + env], cells[3 * n + env]};
    const uint32_t gw[13] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                             v[2].x, v[2].y, v[2].z, v[2].w, v[3].x};
    const uint32_t* plan = reinterpret_cast<const uint32_t*>(st.plans) + (v[3].y >> 16) * PLAN2D_WORDS;
    int inter = 0, uni = 0;
#pragma unroll
    for (int w = 0; w < 13; ++w) {
        inter += __popc(gw[w] & plan[w]);
        uni += __popc(gw[w] | plan[w]);
    }
    out[env] = __ddiv_rn((double)inter, (double)uni);
}

// one warp per env: the 676 cells of the padded grid leave as coalesced stores (the scalar drop-in classes export one env
// after every step, into mapped host memory: a single thread's 676 stores were a third of that step's latency)
__global__ void k2d_export(const DmpState st, int32_t* __restrict__ grid, int32_t* __restrict__ scalars, float* __restrict__ ret) {
    const int64_t n = st.n_envs;
    const int lane = threadIdx.x & 31;
    const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (env >= n) return;
    const uint32_t* cw = reinterpret_cast<const uint32_t*>(st.cells);
    auto word = [&](int w) { return cw[((int64_t)(w >> 2) * n + env) * 4 + (w & 3)]; };
    if (grid) {
        int32_t* g = grid + env * 676;
        for (int i = lane; i < 676; i += 32) {
            const int r = i / 26, c = i - r * 26;
            int v = -1;
            if (r >= 3 && r < 23 && c >= 3 && c < 23) {
                const int b = (r - 3) * 20 + (c - 3);
                v = (word(b >> 5) >> (b & 31)) & 1;
            }
            g[i] = v;
        }
    }
    if (lane == 0) {
        const uint32_t w13 = word(13), w14 = word(14);
        if (scalars) {
            int32_t* s = scalars + env * 8;
            const int p = w13 >> 16;
            s[0] = w13 & 0xFF; s[1] = (w13 >> 8) & 0xFF; s[2] = w14 & 0xFFFF; s[3] = w14 >> 16; s[4] = p;
            s[5] = st.plan_total[p]; s[6] = 0; s[7] = 0;
        }
        if (ret) ret[env] = __uint_as_float(word(15));
    }
}

__global__ void k2d_import(const DmpState st, const int32_t* __restrict__ grid, const int32_t* __restrict__ scalars,
                           const float* __restrict__ ret) {
    const int64_t n = st.n_envs;
    const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= n) return;
    uint32_t* cw = reinterpret_cast<uint32_t*>(st.cells);
    auto wref = [&](int w) -> uint32_t& { return cw[((int64_t)(w >> 2) * n + env) * 4 + (w & 3)]; };
    if (grid) {
        const int32_t* g = grid + env * 676;
        for (int w = 0; w < 13; ++w) {
            uint32_t acc = 0;
            for (int b = w * 32; b < w * 32 + 32 && b < 400; ++b)
                if (g[(b / 20 + 3) * 26 + (b % 20 + 3)] > 0) acc |= 1u << (b & 31);
            wref(w) = acc;
        }
    }
    if (scalars) {
        const int32_t* s = scalars + env * 8;
        wref(13) = (uint32_t)s[0] | ((uint32_t)s[1] << 8) | ((uint32_t)s[4] << 16);
        wref(14) = (uint32_t)(s[2] & 0xFFFF) | ((uint32_t)s[3] << 16);
    }
    if (ret) wref(15) = __float_as_uint(ret[env]);
}

template <typename ObsT, bool TMA, int BT, bool RF>
int launch_rollout2_r(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    const size_t smem = (size_t)S2_WORDS * BT * 4 +                       // bit records need no tile
                        (is_bits<ObsT>::value ? 0 : (size_t)(BT / 32) * 32 * row_elems<ObsT, D2_OBS>() * sizeof(ObsT));
    static bool attr_done = false;           // per instantiation
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(k2d_rollout<ObsT, TMA, BT, RF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return dmp_set_error(e);
        e = cudaFuncSetAttribute(k2d_rollout<ObsT, TMA, BT, RF>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) return dmp_set_error(e);
        attr_done = true;
    }
    const unsigned blocks = (unsigned)((st.n_envs + BT - 1) / BT);
    return dmp_set_error(dmp_launch_pdl(!(io.flags & DMP_F_NO_PDL), k2d_rollout<ObsT, TMA, BT, RF>, blocks, (unsigned)BT, smem, s,
                                        st, io, K));
}

template <typename ObsT, bool TMA, int BT>
int launch_rollout2_b(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    if (io.flags & DMP_F_RESET_OBS) return launch_rollout2_r<ObsT, TMA, BT, true>(st, io, K, s);
    return launch_rollout2_r<ObsT, TMA, BT, false>(st, io, K, s);
}

// Launch shape by shard size: 128-thread blocks hold 768 envs per SM, 224-thread blocks 896 (see the top of the file).
// Measured at 20 steps per launch (f32 observations, G env-steps/s, narrow / wide): 1 048 576 envs 30.4 / 29.3, 524 288
// 30.2 / 28.9, 262 144 29.6 / 27.7 -- but 131 072 envs (BASELINE's batch over 8 GPUs) 22.7 / 26.5: there the narrow shape
// needs a second, nearly empty wave whose blocks run latency-bound while the wide shape holds every env at once.  So the
// wide shape is taken exactly when it turns two waves into one.  DMP_F_GENERIC keeps 128-thread blocks.
constexpr int B2_NARROW = 128, B2_WIDE = 224, SMS = 148;
inline bool wide_blocks2(int64_t n, uint32_t flags) {
    if (flags & DMP_F_GENERIC) return false;
    const int64_t cap_n = (int64_t)SMS * 6 * B2_NARROW, cap_w = (int64_t)SMS * 4 * B2_WIDE;     // 113 664 / 132 608 envs
    return n > cap_n && n <= cap_w;
}

template <typename ObsT, bool TMA>
int launch_rollout2_t(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    // (the residency figures above hold for tile rows of at most 204 B; float64 observations keep the narrow shape)
    if (row_elems<ObsT, D2_OBS>() * sizeof(ObsT) <= 204 && wide_blocks2(st.n_envs, io.flags))
        return launch_rollout2_b<ObsT, TMA, B2_WIDE>(st, io, K, s);
    return launch_rollout2_b<ObsT, TMA, B2_NARROW>(st, io, K, s);
}

// the observation tile leaves through one bulk async copy per warp and step unless DMP_F_TILE_LDST asks for load/store pairs
template <typename ObsT>
int launch_rollout2(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    if (io.flags & DMP_F_TILE_LDST) return launch_rollout2_t<ObsT, false>(st, io, K, s);
    return launch_rollout2_t<ObsT, true>(st, io, K, s);
}

}  // namespace

int dmp2d_rollout(const DmpState& st, const DmpIO& io, int K, cudaStream_t s) {
    switch (io.obs_kind) {
        case DMP_OBS_F32: return launch_rollout2<float>(st, io, K, s);
        case DMP_OBS_F64: return launch_rollout2<double>(st, io, K, s);
        case DMP_OBS_I16: return launch_rollout2<int16_t>(st, io, K, s);
        case DMP_OBS_REC: return launch_rollout2<Rec56>(st, io, K, s);
        case DMP_OBS_BITS: return launch_rollout2_t<Bits16, false>(st, io, K, s);
    }
    return DMP_EINVAL;
}

int dmp2d_reset(const DmpState& st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs,
                int obs_kind, cudaStream_t s) {
    const unsigned blocks = (unsigned)((st.n_envs + 255) / 256);
    switch (obs_kind) {
        case DMP_OBS_F32: k2d_reset<float><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (float*)obs); break;
        case DMP_OBS_F64: k2d_reset<double><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (double*)obs); break;
        case DMP_OBS_I16: k2d_reset<int16_t><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (int16_t*)obs); break;
        case DMP_OBS_REC: k2d_reset<Rec56><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (Rec56*)obs); break;
        case DMP_OBS_BITS: k2d_reset<Bits16><<<blocks, 256, 0, s>>>(st, mask, plan_idx, t_draw, (Bits16*)obs); break;
        default: return DMP_EINVAL;
    }
    return dmp_set_error(cudaGetLastError());
}

int dmp2d_iou(const DmpState& st, double* out, cudaStream_t s) {
    k2d_iou<<<(unsigned)((st.n_envs + 255) / 256), 256, 0, s>>>(st, out);
    return dmp_set_error(cudaGetLastError());
}
int dmp2d_export(const DmpState& st, int32_t* grid, int32_t* scalars, float* ret, cudaStream_t s) {
    k2d_export<<<(unsigned)((st.n_envs + 3) / 4), 128, 0, s>>>(st, grid, scalars, ret);        // four warps = four envs per block
    return dmp_set_error(cudaGetLastError());
}
int dmp2d_import(const DmpState& st, const int32_t* grid, const int32_t* scalars, const float* ret, cudaStream_t s) {
    k2d_import<<<(unsigned)((st.n_envs + 127) / 128), 128, 0, s>>>(st, grid, scalars, ret);
    return dmp_set_error(cudaGetLastError());
}

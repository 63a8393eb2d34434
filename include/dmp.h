/*
 * dmp.h -- C ABI of libdmp.so, the B200 (sm_100a) batched simulator for SNAC's
 * "deep mobile printing" (DMP) mobile-construction environments.
 *
 * The reference (ai4ce/SNAC) is pure Python and has NO FFI for this path; its boundary is the
 * duck-type of the env classes.  Each entry point below names the reference interface it
 * replaces (paths relative to the reference root).  A binding a reference maintainer would add
 * is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *   - every pointer inside DmpState / DmpIO is a DEVICE pointer owned by the caller
 *     (PyTorch tensors in this repo); the library never allocates, frees or synchronises;
 *   - every launch goes on the caller's stream (`stream` is a cudaStream_t passed as void*);
 *   - functions return DMP_OK or a DMP_E* code and never throw;
 *   - envs are independent: entry points are thread-safe for distinct DmpState objects.
 *
 * State layout in HBM (structure-of-arrays, 128-bit accesses; n = n_envs)
 *   1D  cells : uint4 [4][n]   32 x u16 per env: [0..29] column heights, [30] count_brick, [31] count_step
 *       aux   : uint2 [n]      .x = pos | plan_idx << 16 ; .y = episode return so far (f32 bits)
 *   2D  cells : uint4 [4][n]   16 x u32 per env: words 0..12 = 400 occupancy bits (bit r*20+c of the
 *                              20x20 interior), word 13 = pos_row | pos_col << 8 | plan_idx << 16,
 *                              word 14 = count_brick | count_step << 16, word 15 = episode return (f32)
 *       aux   : unused (NULL)
 *   3D  cells : two areas.  At byte offset 800 n: the NIBBLE maps u8 [n][208], per env the 20x20 interior row-major with two
 *               cells per byte (cell i = nibble i & 1 of byte i >> 1, low nibble first; bytes 200..207 are zero padding
 *               that keeps every env 16 B aligned), nibble = min(height, 15) -- what single steps and rollouts read and
 *               write.  At offset 0: the WIDE maps u16 [n][400] (800 B per env), exact and maintained only for envs whose
 *               tall flag is set (a height >= 15 somewhere); for every other env the wide map is scratch (the nibbles ARE
 *               the heights: the reference's plans are 6 high).
 *               dmp_export_state / dmp_iou / dmp_import_state convert as needed; callers never see the difference.
 *       aux   : uint4 [n]      .x = pos_row | tall flag << 7 | pos_col << 8 | plan_idx << 16 ;
 *                              .y = count_brick | count_step << 16 ;
 *                              .z = episode return (f32 bits) ; .w = running sum(min(height, plan)), the IoU
 *                              numerator, +1 for every brick laid at or below the plan height
 *   positions are stored in the reference's padded coordinates (1D: 2..31, 2D/3D: 3..22).
 * Plan tables (n_plans rows; static envs have exactly one row)
 *   1D  plans : u8  [n_plans][32]    target heights (30 used)
 *   2D  plans : u32 [n_plans][16]    400 plan bits in words 0..12 (same bit order as the grid)
 *   3D  plans : u8  [n_plans][400]   target heights
 *       plan_total : i32 [n_plans]   brick budget (the reference's total_brick, incl. 2D's floor of 30)
 * Per-env episode statistics (updated on done when DMP_F_AUTORESET is set)
 *   ep_cnt u32[n] episodes finished, ep_len u32[n] sum of their lengths,
 *   ep_ret f64[n] sum of their returns, ep_iou f64[n] sum of their final IoUs
 */
#ifndef DMP_H_
#define DMP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMP_ABI_VERSION 4

/* return codes */
#define DMP_OK       0
#define DMP_EINVAL   1   /* bad argument (dim, obs kind, null pointer, K < 1, ...) */
#define DMP_ECUDA    2   /* CUDA launch error; dmp_last_error() has the cudaError_t */

/* observation element types */
#define DMP_OBS_F32  0
#define DMP_OBS_F64  1
#define DMP_OBS_I16  2   /* raw counters only */
#define DMP_OBS_REC  3   /* packed step record: window + counters + reward + done in one contiguous record per env -- the
                            host-facing kind: a step's whole result is one buffer and one device-to-host copy, 3.9x
                            smaller than f32 observations + reward + done.  Raw counters only.
                            2D / 3D: 56 B  { u8 win[49]; u8 flags; u16 count_brick; u16 count_step; i8 reward; u8 done; }
                                     win = window value + 1 (0 = the -1 frame, 1 = empty, 2D: 2 = occupied, 3D: height + 1)
                                     flags = DMP_REC_DONE | DMP_REC_SATURATED ; done = 0 / 1 (a bool the host can view in place)
                            1D     : 16 B  { i16 win[5]; u16 count_brick; u16 count_step; i8 reward; u8 done; }
                                     win = raw heights (-1 = wall)
                            Every reward of the six classes (-100, -1, 0, 1, 5, 10) fits the i8.  DmpIO.reward / DmpIO.done
                            are still written when they are not NULL. */
#define DMP_OBS_BITS 4   /* bit-packed step record, 2D and 3D only (1D: DMP_EINVAL -- its DMP_OBS_REC record is 16 B already): the
                            smallest host-facing kind, for callers that are bound by the device-to-host link (a replay buffer
                            holds these records as they are; dmp_records_unpack / unpack_records() turn any batch of them
                            back into observation rows).  Little-endian bit string, field i of width b at bits [i b, i b + b):
                            2D : 16 B  49 x 2-bit window code (value + 1: 0 = the -1 frame, 1 = empty, 2 = occupied), row-major,
                                       then the trailer at bit 98
                            3D : 32 B  49 x 4-bit window code (0 = the -1 frame, else min(height + 1, 15)), row-major; the
                                       trailer is the last 32-bit word (bit 224)
                            trailer: count_brick (12 bits, saturating), count_step (12 bits, saturating), reward code (3 bits:
                                     index into {0, 1, 5, 10, -1, -100}), done (1 bit), saturated (1 bit: a counter passed
                                     4 095 or -- 3D -- a window cell holds a height >= 14; read that env's exact observation
                                     through another kind).  Raw counters only.  DmpIO.obs must be 16 B aligned (the records
                                     leave as 128-bit stores).  DmpIO.reward / DmpIO.done are still written when they are
                                     not NULL. */
#define DMP_REC_DONE       1
#define DMP_REC_SATURATED  2   /* 3D only: the env holds a height >= 15 ("tall"); its window bytes are exact up to 253 and saturate
                                  at 255; read that env's exact observation through another obs kind / dmp_export_state */

/* DmpIO.flags */
#define DMP_F_AUTORESET  1   /* on done: fold the episode into ep_* and reset the env in the same launch */
#define DMP_F_NORMALISE  2   /* obs columns D-2, D-1 = count_brick/total_brick, count_step/total_step
                                (the dynamic envs' format, e.g. Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:64-65);
                                computed as an IEEE fp64 division, then cast to the obs type. f32/f64 only */
#define DMP_F_TSLOT1     4   /* read the step counter from t_dev[1] (see DmpState.t_dev) */
#define DMP_F_NO_L2_HINT 8   /* tuning switch: plain loads/stores for the env state instead of L2 evict_last */
#define DMP_F_NO_PDL     16  /* tuning switch: plain stream-ordered launch instead of programmatic dependent launch */
#define DMP_F_TILE_LDST  32  /* tuning switch: the observation tile leaves shared memory through 128-bit load/store pairs
                                instead of one bulk async copy per warp and step (2D, 3D) */
#define DMP_F_GENERIC    64  /* tuning switch: never pick a specialised instantiation of a kernel (1D: the throughput
                                configuration has one with its launch-uniform branches folded) */
#define DMP_F_ROLLOUT_K1 128 /* tuning switch: 3D single steps run through the rollout kernel (whole maps staged) instead
                                of the single-step kernel (only the rows a step can look at) */
#define DMP_F_RESET_OBS  256 /* with DMP_F_AUTORESET: the observation written for an env whose episode ends in this step is
                                the one its reset returns -- the next episode's first policy input, as in gym's vector
                                envs; the terminal observation is not materialised.  reward / done are the finished
                                episode's.  Acting loops (state = env.reset() after done, e.g.
                                script/DQN/2d/DQN_2d_static.py:186-206) read their next input straight from the buffer */

/* DmpState.plan_mode: which plan an env gets when it auto-resets and DmpIO.next_plan is NULL */
#define DMP_PLAN_PHILOX      0   /* random_choose_paln=True : counter-based draw                */
#define DMP_PLAN_SEQUENTIAL  1   /* random_choose_paln=False: (plan_idx + 1) % n_plans            */
#define DMP_PLAN_KEEP        2   /* keep the current plan (static envs)                           */

/* DmpState.action_dist: distribution of in-kernel synthetic actions (DmpIO.actions == NULL); 3D only, 1D/2D draw uniformly */
#define DMP_ACT_UNIFORM  0
#define DMP_ACT_REF3D    1   /* p = [.2,.2,.2,.2,.05,.05,.05,.05], Env/3D/DMP_simulator_3d_static_circle.py:361-362 */

/* bits latched into *DmpState.err (device int32) */
#define DMP_ERR_ACTION   1   /* action outside the env's action set (the reference raises UnboundLocalError,
                                Env/1D/DMP_Env_1D_static.py:130-133) */
#define DMP_ERR_STEPSIZE 2   /* injected step size outside {1,2,3} */
#define DMP_ERR_PLANIDX  4   /* injected plan index outside [0, n_plans) */
#define DMP_ERR_OVERFLOW 8   /* count_step or count_brick of an env passed 65 535 (the packed state holds 16 bits; reference
                                episodes end after <= 1 300 steps -- only envs stepped on after done without a reset get
                                there); the stored counter saturates */

typedef struct DmpState {
    int32_t  dim;          /* 1, 2 or 3 */
    int32_t  dynamic;      /* 0: static plan classes, 1: dataset-plan classes (affects 3D termination rules) */
    int32_t  n_plans;
    int32_t  total_step;   /* 750 / 600 / 1300 (3D static) / 1000 (3D dynamic) */
    int32_t  plan_mode;    /* DMP_PLAN_* */
    int32_t  action_dist;  /* DMP_ACT_*  */
    int64_t  n_envs;       /* envs resident on this device */
    int64_t  env_base;     /* global index of local env 0 (Philox counter; multi-GPU sharding) */
    uint64_t seed;         /* Philox key */
    uint64_t t;            /* global step index of the NEXT step (host-maintained Philox counter) */
    uint64_t* t_dev;       /* nullable. Device-resident step counter for CUDA-graph replay (2 slots): when set,
                              a launch reads t from t_dev[slot] (slot = DMP_F_TSLOT1 ? 1 : 0) instead of `t`
                              and writes t + K to t_dev[slot ^ 1]; consecutive launches alternate the slot */
    void*    cells;
    void*    aux;
    const void*    plans;
    const int32_t* plan_total;
    uint32_t* ep_cnt;
    uint32_t* ep_len;
    double*   ep_ret;
    double*   ep_iou;
    int32_t*  err;         /* one int32: OR of DMP_ERR_* */
} DmpState;

typedef struct DmpIO {
    const uint8_t* actions;     /* [K][n] ; NULL -> Philox synthetic actions                              */
    const uint8_t* step_sizes;  /* [K][n] in {1,2,3}; NULL -> Philox (replaces np.random.randint(1,4))     */
    const int32_t* next_plan;   /* [K][n] plan index to use if the env auto-resets at that step; NULL -> plan_mode */
    void*    obs;               /* [K][n][D] of obs_kind ([K][n] records for DMP_OBS_REC) ; NULL -> not materialised */
    float*   reward;            /* [K][n] ; NULL -> not materialised                                        */
    uint8_t* done;              /* [K][n] ; NULL -> not materialised                                        */
    int32_t  obs_kind;          /* DMP_OBS_* */
    int32_t  flags;             /* DMP_F_*   */
} DmpIO;

typedef struct DmpLayout {
    int64_t cells_bytes, aux_bytes;        /* state arrays for n envs            */
    int64_t plan_row_bytes;                /* bytes per plan-table row           */
    int32_t obs_dim, n_actions;            /* D and A of the env family          */
    int32_t grid_rows, grid_cols;          /* padded grid of dmp_export_state    */
    int32_t total_step_static, total_step_dynamic;
    int32_t rec_bytes, bits_bytes;         /* bytes of one DMP_OBS_REC record (16 / 56 / 56) and of one DMP_OBS_BITS record
                                              (0 = not available / 16 / 32) */
} DmpLayout;

/* version / diagnostics */
int dmp_abi_version(void);
int dmp_last_error(void);                  /* last cudaError_t seen by this library (per process) */

/* sizes for host-side allocation of the arrays described above */
int dmp_layout(int dim, int64_t n_envs, DmpLayout* out);

/* ---- plan sources (init path) ------------------------------------------------------------------
 * dmp_plan_static: on-device generator of the static plans
 *   replaces create_plan(): Env/1D/DMP_Env_1D_static.py:34-55 (0 sine, 1 Gaussian, 2 step),
 *   Env/2D/DMP_Env_2D_static.py:31-52 and Env/3D/DMP_simulator_3d_static_circle.py:42-65
 *   (0 dense / 1 sparse 20-gon of matplotlib.patches.CirclePolygon).
 *   Writes one plan-table row + its brick budget.  Returns DMP_EINVAL for a bad plan_choose
 *   (the reference raises ValueError at reset()).
 * dmp_plans_pack: converts a dataset in the reference's own format -- float64 arrays as stored
 *   in Env/ *D/data_*_envplan_500_*.pkl (1D [n][30], 2D/3D [n][26][26]) already copied to the
 *   device -- into plan-table rows + budgets (Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:36-46,
 *   Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:47-49). */
int dmp_plan_static(int dim, int plan_choose, void* plans_row_out, int32_t* plan_total_out, void* stream);
int dmp_plans_pack(int dim, const double* raw, int n_plans, void* plans_out, int32_t* plan_total_out, void* stream);

/* ---- on-device random plan generators and hindsight relabelling (init path) --------------------------------
 * dmp_plans_generate replaces create_plan() of the reference's generator classes:
 *   dim 1: random sinusoid, Env/1D/DMP_Env_1D_dynamic_hindsight_replay.py:29-42
 *          (k_1 = uniform(3,12), k_2 = randint(1,4), phase = uniform(-1,1)*pi; y = round(k_1 sin(2pi/30 (k_2 x + phase)) + 20))
 *   dim 2: random triangle (plan_choose 0 dense: cv2.polylines + cv2.fillPoly, 1 sparse: outline only), redrawn until
 *          area > 50 / 20, Env/2D/DMP_Env_2D_dynamic_hindsight_replay_usedata.py:37-59 (= DMP_ENV_2D_dynamic_MCTS.py:40-62);
 *          budget = max(area, 30) (:70-71)
 *   dim 3: the same masks times z = 6, budget = area * 6 (Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:47-49)
 * Writes n_plans plan-table rows + budgets, plan ids first_id .. first_id + n_plans - 1.
 *   draws : nullable injected draws of the reference's numpy stream.  dim 1: f64 [n_plans][3] = k_1, k_2, phase;
 *           dim 2/3: i32 [n_plans][max_attempts][6] = x0,x1,x2,y0,y1,y2 of every attempt in draw order.
 *           NULL -> Philox4x32-10, counter (plan id lo, plan id hi, attempt, "PLAN"), key = seed
 *           (stream definition: oracle/plangen.py).
 *   aux_out : nullable.  dim 1: f64 [n_plans][3] the parameters used (the reference's `one_hot`);
 *           dim 2/3: i32 [n_plans] attempts used.
 *   err   : device int32 (dim 2/3, required): DMP_ERR_PLANIDX is latched if a vertex is outside 0..19 or
 *           max_attempts draws were all rejected.
 * dmp_plans_from_state: plan row i := the structure env i has built so far, budget (nullable) := its brick count /
 *   height sum -- hindsight relabelling, `env_hindsight.plan = env.environment_memory[...]`
 *   (script/DRQN_hindsight/1d/DRQN_hindsight_1D_static.py:242-245).  plans_out has n_envs rows. */
int dmp_plans_generate(int dim, int plan_choose, uint64_t seed, int64_t first_id, int n_plans, const void* draws,
                       int max_attempts, void* plans_out, int32_t* plan_total_out, void* aux_out, int32_t* err,
                       void* stream);
int dmp_plans_from_state(const DmpState* st, void* plans_out, int32_t* plan_total_out, void* stream);

/* ---- reset ---------------------------------------------------------------------------------------
 * replaces reset(): Env/1D/DMP_Env_1D_static.py:66-83, Env/1D/DMP_Env_1D_dynamic_usedata_plan.py:40-70,
 * Env/2D/DMP_Env_2D_static.py:54-76, Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:34-66,
 * Env/3D/DMP_simulator_3d_static_circle.py:67-86, Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-75
 * and VectorizedEnvWrapper.reset / reset_at (multiprocess.py:20-23).
 *   mask     : u8[n] nullable (NULL = all envs); only envs with mask != 0 are touched
 *   plan_idx : i32[n] nullable; NULL -> by plan_mode (Philox draw keyed by (env, t_draw) / +1 / keep).  Sequential
 *              mode starts at plan 0 on an env that has never been reset (zero-initialised state), like the reference's
 *              index_for_non_random = 0 (Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:39-44)
 *   obs      : [n][D] nullable; rows of reset envs are written (raw counters are 0 either way)
 * Episode statistics ep_* are NOT touched (use dmp_stats_clear). */
int dmp_reset(const DmpState* st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw,
              void* obs, int obs_kind, void* stream);

/* ---- step / rollout (the hot path) --------------------------------------------------------------
 * dmp_step replaces step(action) of the six classes (Env/1D/DMP_Env_1D_static.py:85-136,
 * Env/1D/DMP_Env_1D_dynamic_usedata_plan.py:71-120, Env/2D/DMP_Env_2D_static.py:95-154,
 * Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:85-147, Env/3D/DMP_simulator_3d_static_circle.py:153-230,
 * Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:142-231), their *_hindsight_replay
 * step(action, step_size) forms, and VectorizedEnvWrapper.step (multiprocess.py:24-32).
 * dmp_rollout advances every env K steps in ONE launch (state stays on chip between steps);
 * it replaces the `for t in range(T)` loop of multiprocess.py:82-84.  dmp_step == dmp_rollout(K=1).
 * The caller advances DmpState.t by K afterwards. */
int dmp_step(const DmpState* st, const DmpIO* io, void* stream);
int dmp_rollout(const DmpState* st, const DmpIO* io, int K, void* stream);

/* ---- packed records back to observation rows ----------------------------------------------------
 * dmp_records_unpack expands n device-resident step records of kind DMP_OBS_REC or DMP_OBS_BITS (as written by dmp_step /
 * dmp_rollout for env family `dim`) into what the numeric kinds would have written for the same steps: obs [n][D] of
 * obs_kind (DMP_OBS_F32 / F64 / I16, raw counters), reward f32 [n], done u8 [n], saturated u8 [n] (the record's
 * saturation flag: its observation is not exact) -- each nullable.  The learner-side inverse of the compact kinds: a
 * minibatch sampled from a host replay buffer of records is uploaded as records and expanded here
 * (the reference stores float64 rows, e.g. script/DQN/2d/DQN_2d_static.py:196-204). */
int dmp_records_unpack(int dim, int rec_kind, const void* records, int64_t n, void* obs, int obs_kind, float* reward,
                       uint8_t* done, uint8_t* saturated, void* stream);

/* ---- IoU / statistics ----------------------------------------------------------------------------
 * dmp_iou replaces iou(): Env/1D/DMP_Env_1D_static.py:138-151, Env/3D/DMP_simulator_3d_static_circle.py:257-276
 * and the 2D formula of Env/2D/DMP_Env_2D_static.py:169-175.  iou_out: f64[n].
 * dmp_stats_reduce sums the per-env episode statistics into out[4] = {sum return, sum IoU, episodes, steps}
 * (device doubles; deterministic two-pass tree; `scratch` >= dmp_stats_scratch_bytes(n) bytes). The 4-vector
 * is what the multi-GPU driver all-reduces over NCCL. */
int dmp_iou(const DmpState* st, double* iou_out, void* stream);
int64_t dmp_stats_scratch_bytes(int64_t n_envs);
int dmp_stats_reduce(const DmpState* st, double* out4, void* scratch, void* stream);
int dmp_stats_clear(const DmpState* st, void* stream);

/* ---- state export / import (attribute views, get_state/set_state, parity tests) -----------------
 * grid    : i32 [n][rows][cols] in the reference's padded form (environment_memory, -1 frame)
 * scalars : i32 [n][8] = pos_row (1D: pos), pos_col (1D: 0), count_brick, count_step, plan_idx, total_brick, 0, 0
 * ret_acc : f32 [n] nullable */
int dmp_export_state(const DmpState* st, int32_t* grid, int32_t* scalars, float* ret_acc, void* stream);
int dmp_import_state(const DmpState* st, const int32_t* grid, const int32_t* scalars, const float* ret_acc, void* stream);

/* ---- standalone stage kernels (a)-(e) of the step, for unit parity and per-stage timing (obs kinds F32/F64/I16) ---
 * They run the same __device__ stage functions the fused kernels are built from, one stage per launch,
 * communicating through `scratch` (i32 [n][4]: action-valid/brick-placed flag, target cell row, col, reserved).
 *   (a) move      : position update + boundary clamp (+3D collision walk)       [clip_position, move_step]
 *   (b) deposit   : brick deposition into the grid + count_brick                 [step(): drop / build branch]
 *   (c) observe   : observation-window gather -> obs (smem-staged, 128-bit stores) [observation_, np.hstack]
 *   (d) reward    : reward + done decision, IoU of finished episodes              [reward rules, iou()]
 *   (e) done_reset: fold finished episodes into ep_* and reset them               [caller-side reset()] */
int dmp_stage_move(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream);
int dmp_stage_deposit(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream);
int dmp_stage_observe(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream);
int dmp_stage_reward(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream);
int dmp_stage_done_reset(const DmpState* st, const DmpIO* io, int32_t* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DMP_H_ */

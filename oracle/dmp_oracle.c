/*
 * dmp_oracle.c -- plain-C restatement of the reference algorithm (TEST INFRASTRUCTURE ONLY).
 *
 * Same role and same pinning as oracle/dmp_oracle.py (see its header): a CPU checker for the CUDA
 * path, fast enough for 10^5-env parity cases and for a strong multi-threaded CPU baseline.  It is
 * never linked into or called by the product (snac_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Arithmetic mirrors the reference: grids and plans are float64 arrays, counters are ints, the
 * brick budget is a float64.  Citations are relative to the reference root.
 *   1D  Env/1D/DMP_Env_1D_static.py:57-151, Env/1D/DMP_Env_1D_dynamic_usedata_plan.py:32-133
 *   2D  Env/2D/DMP_Env_2D_static.py:54-175, Env/2D/DMP_Env_2D_dynamic_usedata_plan.py:34-147
 *   3D  Env/3D/DMP_simulator_3d_static_circle.py:67-276, Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:45-277
 * Static 2D/3D plans are NOT generated here (matplotlib polygon, parity unpinned): the caller passes
 * plans in the reference's array format (the python oracle's circle_polygon_mask for static envs).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GRID_MAX 676

typedef struct OrcCfg {
    int32_t dim, dynamic, n_plans, total_step;
    const double* plans;        /* [n_plans][30] (1D) or [n_plans][26*26] (2D/3D), reference format */
    const double* total_brick;  /* [n_plans] brick budgets as the reference computes them (float64)  */
} OrcCfg;

typedef struct OrcEnv {
    double grid[GRID_MAX];      /* 1D: [34]; 2D/3D: [26][26] with the -1 frame */
    int32_t pos[2];             /* 1D: pos[0] */
    int32_t count_brick, count_step, plan_idx, pad;
    double ret;                 /* running episode return */
} OrcEnv;

static int obs_dim(int dim) { return dim == 1 ? 7 : 51; }
static int plan_len(int dim) { return dim == 1 ? 30 : 676; }
static const double* plan_of(const OrcCfg* c, const OrcEnv* e) { return c->plans + (size_t)e->plan_idx * plan_len(c->dim); }

int orc_sizeof_env(void) { return (int)sizeof(OrcEnv); }

/* reset(): 1D static :66-83, 2D static :54-76, 3D static :67-86 (and the dynamic twins) */
void orc_reset(OrcEnv* e, const OrcCfg* c, int plan_idx) {
    memset(e->grid, 0, sizeof(e->grid));
    if (c->dim == 1) {
        e->grid[0] = e->grid[1] = e->grid[32] = e->grid[33] = -1.0;
        e->pos[0] = 2; e->pos[1] = 0;
    } else {
        for (int r = 0; r < 26; ++r)
            for (int col = 0; col < 26; ++col)
                if (r < 3 || r > 22 || col < 3 || col > 22) e->grid[r * 26 + col] = -1.0;
        e->pos[0] = 3; e->pos[1] = 3;
    }
    e->count_brick = 0; e->count_step = 0; e->plan_idx = plan_idx; e->ret = 0.0;
}

static void write_obs(const OrcEnv* e, const OrcCfg* c, int normalise, double* obs) {
    if (!obs) return;
    const double tb = c->total_brick[e->plan_idx];
    if (c->dim == 1) {
        for (int j = 0; j < 5; ++j) obs[j] = e->grid[e->pos[0] - 2 + j];
        obs[5] = normalise ? (double)e->count_brick / tb : (double)e->count_brick;
        obs[6] = normalise ? (double)e->count_step / (double)c->total_step : (double)e->count_step;
        return;
    }
    for (int k = 0; k < 7; ++k)
        for (int j = 0; j < 7; ++j) obs[k * 7 + j] = e->grid[(e->pos[0] - 3 + k) * 26 + (e->pos[1] - 3 + j)];
    obs[49] = normalise ? (double)e->count_brick / tb : (double)e->count_brick;
    obs[50] = normalise ? (double)e->count_step / (double)c->total_step : (double)e->count_step;
}

static int clampi(int v, int lo, int hi) { return v <= lo ? lo : (v >= hi ? hi : v); }

/* returns 0 ok, 1 = action outside the action set (the reference raises UnboundLocalError in 1D/2D) */
int orc_step(OrcEnv* e, const OrcCfg* c, int a, int s, double* reward, int* done) {
    const double tb = c->total_brick[e->plan_idx];
    const double* plan = plan_of(c, e);
    e->count_step += 1;
    *reward = 0.0;
    if (c->dim == 1) {                                   /* Env/1D/DMP_Env_1D_static.py:85-136 */
        if (a == 0 || a == 1) {
            e->pos[0] = clampi(a == 0 ? e->pos[0] - s : e->pos[0] + s, 2, 31);
            *done = e->count_step >= c->total_step;
            return 0;
        }
        if (a != 2) { *done = e->count_step >= c->total_step; return 1; }
        e->count_brick += 1;
        e->grid[e->pos[0]] += 1.0;
        if ((double)e->count_brick >= tb) { *done = 1; return 0; }
        *done = e->count_step >= c->total_step;
        const double h = e->grid[e->pos[0]], p = plan[e->pos[0] - 2];
        *reward = h > p ? -1.0 : (h == p ? 10.0 : 1.0);
        return 0;
    }
    if (c->dim == 2) {                                   /* Env/2D/DMP_Env_2D_static.py:95-154 */
        int r = e->pos[0], col = e->pos[1];
        if (a >= 0 && a <= 3) {
            if (a == 0) col -= s; else if (a == 1) col += s; else if (a == 2) r += s; else r -= s;
            e->pos[0] = clampi(r, 3, 22); e->pos[1] = clampi(col, 3, 22);
            *done = e->count_step >= c->total_step;
            return 0;
        }
        if (a != 4) { *done = e->count_step >= c->total_step; return 1; }
        e->count_brick += 1;
        double* cell = &e->grid[r * 26 + col];
        *cell += 1.0;
        if ((double)e->count_brick >= tb) {
            if (*cell > 1.0) *cell = 1.0;
            *done = 1;
            return 0;
        }
        *done = e->count_step >= c->total_step;
        if (*cell == plan[r * 26 + col]) *reward = 5.0;   /* > plan: 0 */
        if (*cell > 1.0) *cell = 1.0;
        return 0;
    }
    /* 3D: static ...static_circle.py:153-230, dynamic ...triangle_usedata.py:142-231 */
    static const int DR[4] = {0, 0, 1, -1}, DC[4] = {-1, 1, 0, 0};
    const int r = e->pos[0], col = e->pos[1];
    double nb[4];
    int boxed = 1;
    for (int q = 0; q < 4; ++q) {                        /* check_sur :88-102 */
        nb[q] = e->grid[(r + DR[q]) * 26 + (col + DC[q])];
        if (nb[q] == 0.0) boxed = 0;
    }
    if (a >= 0 && a <= 3 && nb[a] == 0.0) {              /* move_step :104-134 */
        int n = 0;
        for (int i = 1; i <= s; ++i) {
            if (e->grid[(r + DR[a] * i) * 26 + (col + DC[a] * i)] == 0.0) n++; else break;
        }
        e->pos[0] = clampi(r + DR[a] * n, 3, 22);
        e->pos[1] = clampi(col + DC[a] * n, 3, 22);
    } else if (a > 3) {
        int built = 0, tr = 0, tc = 0;
        if (a <= 7 && nb[a - 4] != -1.0) {
            tr = r + DR[a - 4]; tc = col + DC[a - 4];
            e->count_brick += 1;
            e->grid[tr * 26 + tc] += 1.0;
            built = 1;
        }
        if (c->dynamic) {
            int boxed2 = 1;                              /* re-check after placement :199-206 */
            for (int q = 0; q < 4; ++q)
                if (e->grid[(r + DR[q]) * 26 + (col + DC[q])] == 0.0) boxed2 = 0;
            if (boxed2) { *reward = -100.0; *done = 1; return a > 7; }
            if ((double)e->count_brick >= tb) { *done = 1; return a > 7; }
            if (built) {
                const double v = e->grid[tr * 26 + tc], p = plan[tr * 26 + tc];
                *reward = v > p ? -1.0 : (v == p ? 10.0 : 1.0);
                *done = 0;
                return 0;
            }
        } else {
            if ((double)e->count_brick >= tb || boxed) { *done = 1; return a > 7; }
            if (built) {
                const double v = e->grid[tr * 26 + tc], p = plan[tr * 26 + tc];
                *reward = v > p ? -1.0 : (v == p ? 10.0 : 1.0);
                *done = 0;
                return 0;
            }
        }
    }
    *done = (e->count_step >= c->total_step) || (!c->dynamic && boxed);
    return a > 7;
}

double orc_iou(const OrcEnv* e, const OrcCfg* c) {
    const double* plan = plan_of(c, e);
    if (c->dim == 1) {                                   /* Env/1D/DMP_Env_1D_static.py:138-151 */
        double a1 = 0, a2 = 0, over = 0;
        for (int i = 0; i < 30; ++i) {
            a1 += plan[i]; a2 += e->grid[2 + i];
            if (e->grid[2 + i] > plan[i]) over += e->grid[2 + i] - plan[i];
        }
        const double cross = a2 - over;
        return cross / (a1 + a2 - cross);
    }
    if (c->dim == 2) {                                   /* Env/2D/DMP_Env_2D_static.py:169-175 */
        double inter = 0, uni = 0;
        for (int r = 3; r < 23; ++r)
            for (int col = 3; col < 23; ++col) {
                const int p = plan[r * 26 + col] != 0.0, g = e->grid[r * 26 + col] != 0.0;
                inter += p && g; uni += p || g;
            }
        return inter / uni;
    }
    double cross = 0;                                    /* Env/3D/DMP_simulator_3d_static_circle.py:257-276 */
    for (int r = 3; r < 23; ++r)
        for (int col = 3; col < 23; ++col) {
            const double p = plan[r * 26 + col], g = e->grid[r * 26 + col];
            cross += g > p ? p : g;
        }
    return cross / (c->total_brick[e->plan_idx] + (double)e->count_brick - cross);
}

/*
 * K steps of n envs with the vector env's auto-reset + statistics semantics (include/dmp.h).
 * actions, sizes: u8 [K][n]; next_plan: i32 [K][n] or NULL (sequential (+1) if seq != 0, else keep);
 * obs [K][n][D] f64 or NULL; reward f32 [K][n] or NULL; done u8 [K][n] or NULL;
 * ep_cnt i64[n], ep_len i64[n], ep_ret f64[n], ep_iou f64[n] accumulate when auto_reset.
 * Parallel over envs with OpenMP when compiled with -fopenmp.  Returns the OR of per-step error flags.
 */
int orc_rollout(const OrcCfg* c, OrcEnv* envs, int64_t n, int K, const uint8_t* actions, const uint8_t* sizes,
                const int32_t* next_plan, int seq, int auto_reset, int normalise, double* obs, float* reward,
                uint8_t* done, int64_t* ep_cnt, int64_t* ep_len, double* ep_ret, double* ep_iou) {
    const int D = obs_dim(c->dim);
    int err = 0;
#pragma omp parallel for schedule(static) reduction(| : err)
    for (int64_t i = 0; i < n; ++i) {
        OrcEnv* e = &envs[i];
        for (int k = 0; k < K; ++k) {
            const int64_t idx = (int64_t)k * n + i;
            double r; int d;
            err |= orc_step(e, c, actions[idx], sizes[idx], &r, &d);
            e->ret += r;
            write_obs(e, c, normalise, obs ? obs + idx * D : NULL);
            if (reward) reward[idx] = (float)r;
            if (done) done[idx] = (uint8_t)d;
            if (d && auto_reset) {
                ep_cnt[i] += 1; ep_len[i] += e->count_step; ep_ret[i] += e->ret; ep_iou[i] += orc_iou(e, c);
                int p = e->plan_idx;
                if (next_plan) p = next_plan[idx];
                else if (seq) p = (p + 1) % c->n_plans;
                orc_reset(e, c, p);
            }
        }
    }
    return err;
}

/* dense export of one env for comparisons: grid as int32 (1D: 34, else 676), scalars[6] */
void orc_export(const OrcCfg* c, const OrcEnv* envs, int64_t n, int32_t* grid, int32_t* scalars) {
    const int G = c->dim == 1 ? 34 : 676;
    for (int64_t i = 0; i < n; ++i) {
        for (int j = 0; j < G; ++j) grid[i * G + j] = (int32_t)envs[i].grid[j];
        int32_t* s = scalars + i * 8;
        s[0] = envs[i].pos[0]; s[1] = envs[i].pos[1]; s[2] = envs[i].count_brick; s[3] = envs[i].count_step;
        s[4] = envs[i].plan_idx; s[5] = (int32_t)ceil(c->total_brick[envs[i].plan_idx]); s[6] = 0; s[7] = 0;
    }
}

void orc_reset_all(const OrcCfg* c, OrcEnv* envs, int64_t n, const int32_t* plan_idx, double* obs) {
    const int D = obs_dim(c->dim);
    for (int64_t i = 0; i < n; ++i) {
        orc_reset(&envs[i], c, plan_idx ? plan_idx[i] : 0);
        write_obs(&envs[i], c, 0, obs ? obs + i * D : NULL);
    }
}

void orc_iou_all(const OrcCfg* c, const OrcEnv* envs, int64_t n, double* out) {
    for (int64_t i = 0; i < n; ++i) out[i] = orc_iou(&envs[i], c);
}

/* ---------------------------------------------------------------------------------------------------------
 * Random triangle plans (oracle/plangen.py is the commented statement; this is its fast twin used for the
 * exhaustive comparison with cv2 and for large GPU parity cases).
 * create_plan: Env/2D/DMP_Env_2D_dynamic_hindsight_replay_usedata.py:37-59.
 * rows[r] bit c = 1 where the reference image is black (row = y, column = x). */
static void orc_line_rows(int x1, int y1, int x2, int y2, uint32_t* rows) {
    int dx = x2 - x1, dy = y2 - y1, t;
    if (dx < 0) { t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; dx = -dx; dy = -dy; }
    int ystep = 1;
    if (dy < 0) { dy = -dy; ystep = -1; }
    int x = x1, y = y1;
    if (dy > dx) {
        int err = dy - 2 * dx;
        for (int i = 0; i <= dy; ++i) {
            rows[y] |= 1u << x;
            const int neg = err < 0;
            err += -2 * dx + (neg ? 2 * dy : 0);
            y += ystep;
            x += neg;
        }
    } else {
        int err = dx - 2 * dy;
        for (int i = 0; i <= dx; ++i) {
            rows[y] |= 1u << x;
            const int neg = err < 0;
            err += -2 * dy + (neg ? 2 * dx : 0);
            x += 1;
            if (neg) y += ystep;
        }
    }
}

/* out: 13 words, bit r*20+c (the 2D grid's bit order); returns the area */
int orc_triangle_mask(const int32_t* xs, const int32_t* ys, int dense, uint32_t* out13) {
    uint32_t rows[20];
    for (int r = 0; r < 20; ++r) rows[r] = 0;
    for (int i = 0; i < 3; ++i) orc_line_rows(xs[i], ys[i], xs[(i + 1) % 3], ys[(i + 1) % 3], rows);
    int area = 0;
    for (int w = 0; w < 13; ++w) out13[w] = 0;
    for (int r = 0; r < 20; ++r) {
        uint32_t m = rows[r];
        if (dense && m) {
            const int lo = __builtin_ctz(m), hi = 31 - __builtin_clz(m);
            m = (uint32_t)((2ull << hi) - (1ull << lo));
        }
        area += __builtin_popcount(m);
        const int b = r * 20;
        out13[b >> 5] |= m << (b & 31);
        if ((b & 31) > 12) out13[(b >> 5) + 1] |= m >> (32 - (b & 31));
    }
    return area;
}

void orc_triangle_masks(int64_t n, const int32_t* xs, const int32_t* ys, int dense, uint32_t* out, int32_t* area) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const int a = orc_triangle_mask(xs + 3 * i, ys + 3 * i, dense, out + 13 * i);
        if (area) area[i] = a;
    }
}

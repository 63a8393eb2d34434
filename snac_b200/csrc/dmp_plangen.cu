// dmp_plangen.cu -- on-device random plan generators (init path).
//
// Replaces create_plan() of the reference's generator classes:
//   1D random sinusoid   Env/1D/DMP_Env_1D_dynamic_hindsight_replay.py:29-42
//   2D random triangle   Env/2D/DMP_Env_2D_dynamic_hindsight_replay_usedata.py:37-59, Env/2D/DMP_ENV_2D_dynamic_MCTS.py:40-62
//   3D                   the same masks times z = 6 (Env/3D/DMP_simulator_3d_dynamic_triangle_usedata.py:47-49)
// The triangle image is cv2.polylines (1-pixel, 8-connected outline) plus, for dense plans, cv2.fillPoly; both are
// third-party arithmetic restated in oracle/plangen.py and pinned there against cv2 for every vertex triple of the
// 20x20 grid.  The stochastic draws are either injected (the reference's numpy stream) or taken from Philox with
// counter (plan id lo, plan id hi, attempt, "PLAN") -- oracle/plangen.py:philox_vertices / philox_sin_params.
#include <math.h>
#include "dmp_common.cuh"

namespace {

constexpr uint32_t PLAN_TAG = 0x504C414Eu;      // "PLAN"
constexpr int PG_BLOCK = 128;

// OpenCV's 8-connected LineIterator for integer end points: swap so that x increases, Bresenham with the
// error term dx - 2*dy (shallow) / dy - 2*dx (steep) and the "err < 0" step rule.  rows[y][tid] bit x.
__device__ __forceinline__ void line_rows(int x1, int y1, int x2, int y2, uint32_t (*rows)[PG_BLOCK], int tid) {
    int dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) { int t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; dx = -dx; dy = -dy; }
    int ystep = 1;
    if (dy < 0) { dy = -dy; ystep = -1; }
    int x = x1, y = y1;
    if (dy > dx) {
        int err = dy - 2 * dx;
        for (int i = 0; i <= dy; ++i) {
            rows[y][tid] |= 1u << x;
            const bool neg = err < 0;
            err += -2 * dx + (neg ? 2 * dy : 0);
            y += ystep;
            x += neg ? 1 : 0;
        }
    } else {
        int err = dx - 2 * dy;
        for (int i = 0; i <= dx; ++i) {
            rows[y][tid] |= 1u << x;
            const bool neg = err < 0;
            err += -2 * dy + (neg ? 2 * dx : 0);
            x += 1;
            y += neg ? ystep : 0;
        }
    }
}

// One plan per thread.  verts (nullable): i32 [n][max_attempts][6] = x0,x1,x2,y0,y1,y2 per attempt -- the
// reference's np.random.randint(0, 20, size=3) pairs in draw order; attempt a of plan p is used only if all
// earlier attempts were rejected (area <= 50 dense / 20 sparse), exactly like the reference's while loop.
__global__ void __launch_bounds__(PG_BLOCK)
k_plans_triangle(int dim, int plan_choose, uint64_t seed, int64_t first_id, int n, const int32_t* __restrict__ verts,
                 int max_attempts, void* plans_out, int32_t* total_out, int32_t* attempts_out, int32_t* err) {
    __shared__ uint32_t rows[20][PG_BLOCK];
    const int tid = threadIdx.x;
    const int p = blockIdx.x * PG_BLOCK + tid;
    if (p >= n) return;
    const int area_min = plan_choose == 0 ? 50 : 20;
    const uint64_t id = (uint64_t)(first_id + p);
    int area = 0, att = 0;
    while (att < max_attempts) {
        int xs[3], ys[3];
        if (verts) {
            const int32_t* v = verts + ((int64_t)p * max_attempts + att) * 6;
#pragma unroll
            for (int i = 0; i < 3; ++i) { xs[i] = v[i]; ys[i] = v[3 + i]; }
            bool bad = false;
#pragma unroll
            for (int i = 0; i < 3; ++i) bad |= (unsigned)xs[i] >= 20u || (unsigned)ys[i] >= 20u;
            if (bad) { atomicOr(err, DMP_ERR_PLANIDX); xs[0] = xs[1] = xs[2] = ys[0] = ys[1] = ys[2] = 0; }
        } else {
            const Draw d = philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), (uint32_t)att, PLAN_TAG,
                                         (uint32_t)seed, (uint32_t)(seed >> 32));
            const uint32_t w[3] = {d.x0, d.x1, d.x2};
#pragma unroll
            for (int i = 0; i < 3; ++i) { xs[i] = (int)(((w[i] & 0xFFFFu) * 20u) >> 16); ys[i] = (int)(((w[i] >> 16) * 20u) >> 16); }
        }
        ++att;
#pragma unroll
        for (int r = 0; r < 20; ++r) rows[r][tid] = 0;
        line_rows(xs[0], ys[0], xs[1], ys[1], rows, tid);
        line_rows(xs[1], ys[1], xs[2], ys[2], rows, tid);
        line_rows(xs[2], ys[2], xs[0], ys[0], rows, tid);
        area = 0;
#pragma unroll
        for (int r = 0; r < 20; ++r) {
            uint32_t m = rows[r][tid];
            if (plan_choose == 0 && m) {
                // fillPoly adds only pixels between the left-most and right-most outline pixel of a row
                const int lo = __ffs(m) - 1, hi = 31 - __clz(m);
                m = (uint32_t)((2ull << hi) - (1ull << lo));
                rows[r][tid] = m;
            }
            area += __popc(m);
        }
        if (area > area_min) break;
    }
    if (area <= area_min) atomicOr(err, DMP_ERR_PLANIDX);       // max_attempts exhausted: the last image is kept
    if (attempts_out) attempts_out[p] = att;
    if (dim == 2) {
        uint32_t w[PLAN2D_WORDS];
#pragma unroll
        for (int i = 0; i < PLAN2D_WORDS; ++i) w[i] = 0;
#pragma unroll
        for (int r = 0; r < 20; ++r) {
            const uint32_t m = rows[r][tid];
            const int b = r * 20;
            w[b >> 5] |= m << (b & 31);
            if ((b & 31) > 12) w[(b >> 5) + 1] |= m >> (32 - (b & 31));
        }
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(plans_out) + (int64_t)p * PLAN2D_WORDS);
#pragma unroll
        for (int q = 0; q < 4; ++q) dst[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        total_out[p] = max(area, 30);                           // reset(): `if self.total_brick < 30`, :70-71
    } else {
        uint8_t* dst = reinterpret_cast<uint8_t*>(plans_out) + (int64_t)p * CELLS3D;
        for (int r = 0; r < 20; ++r) {
            const uint32_t m = rows[r][tid];
            uint32_t* d4 = reinterpret_cast<uint32_t*>(dst + r * 20);
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const uint32_t nib = (m >> (4 * q)) & 15u;
                // bit k of nib -> byte k = 6
                const uint32_t spread = (nib * 0x00204081u) & 0x01010101u;
                d4[q] = spread * 6u;
            }
        }
        total_out[p] = area * 6;                                // sum(sum(plan / z)) * z
    }
}

// 1D random sinusoid.  params (nullable): f64 [n][3] = k_1, k_2, phase (the reference's `one_hot`).
__global__ void k_plans_sin(uint64_t seed, int64_t first_id, int n, const double* __restrict__ params,
                            uint8_t* plans_out, int32_t* total_out, double* params_out) {
    const int p = blockIdx.x, t = threadIdx.x;                  // 32 threads per plan
    if (p >= n) return;
    double k1, k2, phase;
    if (params) {
        k1 = params[3 * p]; k2 = params[3 * p + 1]; phase = params[3 * p + 2];
    } else {
        const uint64_t id = (uint64_t)(first_id + p);
        const Draw d = philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), 0u, PLAN_TAG, (uint32_t)seed, (uint32_t)(seed >> 32));
        k1 = 3.0 + 9.0 * ((double)d.x0 * 0x1p-32);
        k2 = (double)(1 + (int)__umulhi(d.x1, 3u));
        phase = ((double)d.x2 * 0x1p-31 - 1.0) * M_PI;
    }
    if (params_out && t < 3) params_out[3 * p + t] = t == 0 ? k1 : (t == 1 ? k2 : phase);
    int v = 0;
    if (t < 30) {
        // np.round(k_1 * np.sin(2 * np.pi / 30 * (k_2 * x + phase)) + 20), same association as the numpy expression
        const double arg = (2.0 * M_PI / 30.0) * (k2 * (double)t + phase);
        v = (int)rint(__dadd_rn(__dmul_rn(k1, sin(arg)), 20.0));          // no FMA contraction: numpy rounds twice
        v = min(max(v, 0), 255);
    }
    plans_out[(int64_t)p * PLAN1D_BYTES + t] = (uint8_t)v;
    int s = v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
    if (t == 0) total_out[p] = s;
}

// Hindsight relabelling: plan row i := what env i has built so far
// (script/DRQN_hindsight/1d/DRQN_hindsight_1D_static.py:243 `env_hindsight.plan = env.environment_memory[...]`).
__global__ void k_plans_from_state(const DmpState st, void* plans_out, int32_t* total_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n_envs) return;
    const int64_t n = st.n_envs;
    int total = 0;
    if (st.dim == 1) {
        const uint4* cells = reinterpret_cast<const uint4*>(st.cells);
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const uint4 v = cells[q * n + i]; w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w; }
        uint32_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t b = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * k + j;
                uint32_t h = (c & 1) ? (w[c >> 1] >> 16) : (w[c >> 1] & 0xFFFFu);
                if (c >= 30) h = 0;
                h = min(h, 255u);
                total += (int)h;
                b |= h << (8 * j);
            }
            o[k] = b;
        }
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(plans_out) + i * PLAN1D_BYTES);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    } else if (st.dim == 2) {
        const uint4* cells = reinterpret_cast<const uint4*>(st.cells);
        uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint32_t*>(plans_out) + i * PLAN2D_WORDS);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 v = cells[q * n + i];
            if (q == 3) { v.x &= 0xFFFFu; v.y = 0; v.z = 0; v.w = 0; }        // word 12 holds bits 384..399
            total += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
            dst[q] = v;
        }
    } else {
        // heights of an env that is not tall are its nibbles; a tall env's are in its wide map (plan rows are bytes)
        const bool tall = (reinterpret_cast<const uint4*>(st.aux)[i].x & AUX3_TALL) != 0u;
        const uint8_t* nib = nmap3(st) + i * NIB3_STRIDE;
        const uint16_t* wide = reinterpret_cast<const uint16_t*>(st.cells) + i * CELLS3D;
        uint8_t* dst = reinterpret_cast<uint8_t*>(plans_out) + i * CELLS3D;
        for (int c = 0; c < CELLS3D; ++c) {
            const int h = tall ? min((int)wide[c], 255) : ((nib[c >> 1] >> ((c & 1) * 4)) & 0xF);
            total += h;
            dst[c] = (uint8_t)h;
        }
    }
    if (total_out) total_out[i] = total;
}

}  // namespace

extern "C" {

int dmp_plans_generate(int dim, int plan_choose, uint64_t seed, int64_t first_id, int n_plans, const void* draws,
                       int max_attempts, void* plans_out, int32_t* plan_total_out, void* aux_out, int32_t* err,
                       void* stream) {
    if (dim < 1 || dim > 3 || n_plans < 1 || !plans_out || !plan_total_out) return DMP_EINVAL;
    if (dim == 1) {
        k_plans_sin<<<n_plans, 32, 0, as_stream(stream)>>>(seed, first_id, n_plans, reinterpret_cast<const double*>(draws),
                                                           reinterpret_cast<uint8_t*>(plans_out), plan_total_out,
                                                           reinterpret_cast<double*>(aux_out));
        return dmp_set_error(cudaGetLastError());
    }
    if (plan_choose < 0 || plan_choose > 1 || max_attempts < 1 || !err) return DMP_EINVAL;   // reference: ValueError
    k_plans_triangle<<<(n_plans + PG_BLOCK - 1) / PG_BLOCK, PG_BLOCK, 0, as_stream(stream)>>>(
        dim, plan_choose, seed, first_id, n_plans, reinterpret_cast<const int32_t*>(draws), max_attempts, plans_out,
        plan_total_out, reinterpret_cast<int32_t*>(aux_out), err);
    return dmp_set_error(cudaGetLastError());
}

int dmp_plans_from_state(const DmpState* st, void* plans_out, int32_t* plan_total_out, void* stream) {
    if (!st || st->dim < 1 || st->dim > 3 || st->n_envs < 1 || !st->cells || !plans_out) return DMP_EINVAL;
    k_plans_from_state<<<(unsigned)((st->n_envs + 127) / 128), 128, 0, as_stream(stream)>>>(*st, plans_out, plan_total_out);
    return dmp_set_error(cudaGetLastError());
}

}  // extern "C"

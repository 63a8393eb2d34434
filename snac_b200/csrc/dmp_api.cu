// dmp_api.cu -- the C ABI of libdmp.so (include/dmp.h): argument checks, per-dimension dispatch,
// plan generators / dataset packing (init path) and the episode-statistics reduction.
#include <math.h>
#include "dmp_common.cuh"

static int g_last_cuda_error = 0;

int dmp_set_error(cudaError_t e) {
    if (e == cudaSuccess) return DMP_OK;
    g_last_cuda_error = (int)e;
    return DMP_ECUDA;
}

namespace {

// ---------------------------------------------------------------------------------------------
// static plan generators (on-device restatement of create_plan)
// ---------------------------------------------------------------------------------------------
// matplotlib CirclePolygon(xy=(12.5,12.5), radius=R, resolution=20).contains_point((i, j)):
// regular 20-gon with vertices at angle pi/2 + 2*pi*k/20, crossing-number test.
__device__ bool in_polygon20(double px, double py, double radius) {
    if (radius <= 0.0) return false;
    bool inside = false;
    const double step = 2.0 * M_PI / 20.0;
    double th = step * 19 + M_PI / 2;
    double x0 = 12.5 + radius * cos(th), y0 = 12.5 + radius * sin(th);
    bool f0 = y0 >= py;
    for (int k = 0; k < 20; ++k) {
        th = step * k + M_PI / 2;
        const double x1 = 12.5 + radius * cos(th), y1 = 12.5 + radius * sin(th);
        const bool f1 = y1 >= py;
        if (f0 != f1) {
            if (((y1 - py) * (x0 - x1) >= (x1 - px) * (y0 - y1)) == f1) inside = !inside;
        }
        f0 = f1; x0 = x1; y0 = y1;
    }
    return inside;
}

// one block of 416 threads
__global__ void k_plan_static(int dim, int plan_choose, void* row_out, int32_t* total_out) {
    __shared__ int cell[416];
    const int t = threadIdx.x;
    if (dim == 1) {
        // Env/1D/DMP_Env_1D_static.py:34-55
        int v = 0;
        if (t < 30) {
            if (plan_choose == 0) {
                v = (int)rint(10.0 * sin(2.0 * M_PI / 30.0 * (double)t) + 20.0);
            } else if (plan_choose == 1) {
                const double step = 36.0 / 29.0;                       // np.linspace(-18, 18, 30)
                const double x = (t == 29) ? 18.0 : -18.0 + step * (double)t;
                const double pdf = exp(-1.0 * (x * x) / 18.0) / (sqrt(2.0 * M_PI) * 3.0);
                v = (int)rint(pdf * 100.0 + 17.0);
            } else {
                v = ((t / 5) % 2 == 0) ? 25 : 15;                       // y[0:5]=y[10:15]=y[20:25]=25 else 15
            }
        }
        cell[t] = v;
        __syncthreads();
        if (t < 32) reinterpret_cast<uint8_t*>(row_out)[t] = (uint8_t)(t < 30 ? cell[t] : 0);
        if (t == 0) {
            int s = 0;
            for (int i = 0; i < 30; ++i) s += cell[i];
            *total_out = s;
        }
        return;
    }
    // 2D / 3D: Env/2D/DMP_Env_2D_static.py:31-52, Env/3D/DMP_simulator_3d_static_circle.py:42-65
    const double r_out = plan_choose == 0 ? 7.0 : 8.0, r_in = plan_choose == 0 ? 0.0 : 7.0;
    int v = 0;
    if (t < 400) {
        const double pi = (double)(t / 20 + 3), pj = (double)(t % 20 + 3);   // point (i, j) = (row, col)
        v = (in_polygon20(pi, pj, r_out) && !in_polygon20(pi, pj, r_in)) ? 1 : 0;
    }
    cell[t] = v;
    __syncthreads();
    if (dim == 2) {
        if (t < PLAN2D_WORDS) {
            uint32_t w = 0;
            for (int b = 0; b < 32; ++b) {
                const int i = t * 32 + b;
                if (i < 400 && cell[i]) w |= 1u << b;
            }
            reinterpret_cast<uint32_t*>(row_out)[t] = w;
        }
    } else {
        if (t < 400) reinterpret_cast<uint8_t*>(row_out)[t] = (uint8_t)(cell[t] * 6);       // plan * z
    }
    if (t == 0) {
        int s = 0;
        for (int i = 0; i < 400; ++i) s += cell[i];
        // the frame of the 26x26 reference plan can never be inside a radius-8 polygon centred at 12.5
        *total_out = (dim == 2) ? max(s, 30) : s * 6;                   // 2D floor :56-57 ; 3D area*z :62-64
    }
}

// dataset rows in the reference's float64 format -> packed plan rows + brick budgets.
// One block per plan.  Sums follow the reference's evaluation order (python sum over rows of numpy
// arrays = per-column sequential adds, then a sequential sum over columns) so that the fp64 budget
// is reproduced exactly even for plans whose heights are not multiples of z.
__global__ void k_plans_pack(int dim, const double* __restrict__ raw, void* plans_out, int32_t* total_out) {
    const int p = blockIdx.x, t = threadIdx.x;
    __shared__ double colsum[26];
    if (dim == 1) {
        const double* src = raw + (int64_t)p * 30;
        uint8_t* dst = reinterpret_cast<uint8_t*>(plans_out) + (int64_t)p * PLAN1D_BYTES;
        if (t < 32) dst[t] = (t < 30) ? (uint8_t)src[t] : 0;
        if (t == 0) {
            double s = 0.0;                                             // total_brick = sum(plan), 1D dynamic :44
            for (int i = 0; i < 30; ++i) s += src[i];
            total_out[p] = (int32_t)ceil(s);
        }
        return;
    }
    const double* src = raw + (int64_t)p * 676;
    const double z = 6.0;
    if (t < 26) {
        double s = 0.0;
        for (int r = 0; r < 26; ++r) s += (dim == 3) ? src[r * 26 + t] / z : src[r * 26 + t];
        colsum[t] = s;
    }
    __syncthreads();
    if (t == 0) {
        double s = 0.0;
        for (int c = 0; c < 26; ++c) s += colsum[c];
        if (dim == 3) s = s * z;                                        // sum(sum(plan/z))*z, 3D dynamic :49
        else if (s < 30.0) s = 30.0;                                    // 2D dynamic :45-46
        total_out[p] = (int32_t)ceil(s);
    }
    if (dim == 2) {
        if (t < PLAN2D_WORDS) {
            uint32_t w = 0;
            for (int b = 0; b < 32; ++b) {
                const int i = t * 32 + b;
                if (i < 400 && src[(i / 20 + 3) * 26 + (i % 20 + 3)] > 0.0) w |= 1u << b;
            }
            reinterpret_cast<uint32_t*>(plans_out)[(int64_t)p * PLAN2D_WORDS + t] = w;
        }
    } else {
        for (int i = t; i < 400; i += blockDim.x)
            reinterpret_cast<uint8_t*>(plans_out)[(int64_t)p * CELLS3D + i] = (uint8_t)src[(i / 20 + 3) * 26 + (i % 20 + 3)];
    }
}

// ---------------------------------------------------------------------------------------------
// episode statistics: deterministic two-pass sum of {ep_ret, ep_iou, ep_cnt, ep_len}
// ---------------------------------------------------------------------------------------------
constexpr int STATS_BLOCK = 256;
constexpr int STATS_MAX_BLOCKS = 1024;

__device__ __forceinline__ void block_sum4(double (&v)[4], double* out4) {
    __shared__ double sh[STATS_BLOCK / 32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xFFFFFFFFu, v[q], o);
    if (lane == 0)
        for (int q = 0; q < 4; ++q) sh[warp][q] = v[q];
    __syncthreads();
    if (threadIdx.x < 4) {
        double s = 0.0;
        for (int w = 0; w < STATS_BLOCK / 32; ++w) s += sh[w][threadIdx.x];
        out4[threadIdx.x] = s;
    }
}

__global__ void __launch_bounds__(STATS_BLOCK) k_stats_pass1(const DmpState st, double* __restrict__ partial) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * STATS_BLOCK + threadIdx.x; i < st.n_envs; i += (int64_t)gridDim.x * STATS_BLOCK) {
        v[0] += st.ep_ret[i];
        v[1] += st.ep_iou[i];
        v[2] += (double)st.ep_cnt[i];
        v[3] += (double)st.ep_len[i];
    }
    block_sum4(v, partial + (int64_t)blockIdx.x * 4);
}

__global__ void __launch_bounds__(STATS_BLOCK) k_stats_pass2(const double* __restrict__ partial, int nblocks, double* __restrict__ out4) {
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = threadIdx.x; i < nblocks; i += STATS_BLOCK)
        for (int q = 0; q < 4; ++q) v[q] += partial[i * 4 + q];
    block_sum4(v, out4);
}

__global__ void k_stats_clear(const DmpState st) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.n_envs) return;
    st.ep_cnt[i] = 0; st.ep_len[i] = 0; st.ep_ret[i] = 0.0; st.ep_iou[i] = 0.0;
}

inline int stats_blocks(int64_t n) {
    int64_t b = (n + STATS_BLOCK - 1) / STATS_BLOCK;
    if (b > STATS_MAX_BLOCKS) b = STATS_MAX_BLOCKS;
    if (b < 1) b = 1;
    return (int)b;
}

// ---------------------------------------------------------------------------------------------
// packed step records -> observation rows (dmp_records_unpack): one thread per output value, so the rows leave as
// coalesced stores; the few record words a warp needs stay in L1
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bits_reward_value(uint32_t code) {
    return code == 0u ? 0.f : (code == 1u ? 1.f : (code == 2u ? 5.f : (code == 3u ? 10.f : (code == 4u ? -1.f : -100.f))));
}

template <typename ObsT>
__global__ void k_records_unpack(int dim, int kind, const uint8_t* __restrict__ rec, int64_t n, ObsT* __restrict__ obs,
                                 float* __restrict__ reward, uint8_t* __restrict__ done, uint8_t* __restrict__ sat) {
    const int D = dim == 1 ? D1_OBS : D2_OBS, W = D - 2;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * D) return;
    const int64_t env = i / D;
    const int col = (int)(i - env * D);
    int v, cb, cs, dn, st_;
    float rw;
    if (kind == DMP_OBS_BITS) {
        const int rb = dim == 2 ? 16 : 32, cw = dim == 2 ? 2 : 4;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(rec + env * rb);
        const uint32_t tr = dim == 2 ? (w[3] >> 2) : w[7];
        cb = tr & 0xFFF; cs = (tr >> 12) & 0xFFF; rw = bits_reward_value((tr >> 24) & 7u); dn = (tr >> 27) & 1; st_ = (tr >> 28) & 1;
        const int b = col * cw;
        v = col < W ? (int)((w[b >> 5] >> (b & 31)) & ((1u << cw) - 1u)) - 1 : 0;
    } else if (dim == 1) {
        const uint8_t* r = rec + env * 16;
        const int16_t* h = reinterpret_cast<const int16_t*>(r);
        cb = reinterpret_cast<const uint16_t*>(r)[5]; cs = reinterpret_cast<const uint16_t*>(r)[6];
        rw = (float)(int8_t)r[14]; dn = r[15]; st_ = 0;
        v = col < W ? (int)h[col] : 0;
    } else {
        const uint8_t* r = rec + env * 56;
        cb = reinterpret_cast<const uint16_t*>(r)[25]; cs = reinterpret_cast<const uint16_t*>(r)[26];
        rw = (float)(int8_t)r[54]; dn = r[55]; st_ = (r[49] & DMP_REC_SATURATED) ? 1 : 0;
        v = col < W ? (int)r[col] - 1 : 0;
    }
    if (col == W) v = cb;
    if (col == W + 1) v = cs;
    if (obs) obs[i] = obs_from_int<ObsT>(v);
    if (col == 0) {
        if (reward) reward[env] = rw;
        if (done) done[env] = (uint8_t)dn;
        if (sat) sat[env] = (uint8_t)st_;
    }
}

bool state_ok(const DmpState* st) {
    if (!st) return false;
    if (st->dim < 1 || st->dim > 3) return false;
    if (st->n_envs < 1 || st->n_plans < 1) return false;
    if (!st->cells || !st->plans || !st->plan_total || !st->err) return false;
    if (st->dim != 2 && !st->aux) return false;
    if (st->n_plans > 65535) return false;
    return true;
}

}  // namespace

// =============================================================================================
// exported C ABI
// =============================================================================================
extern "C" {

int dmp_abi_version(void) { return DMP_ABI_VERSION; }
int dmp_last_error(void) { return g_last_cuda_error; }

int dmp_layout(int dim, int64_t n, DmpLayout* out) {
    if (!out || n < 1) return DMP_EINVAL;
    switch (dim) {
        case 1:
            *out = DmpLayout{64 * n, 8 * n, PLAN1D_BYTES, D1_OBS, D1_ACT, 1, 34, 750, 750, (int32_t)sizeof(Rec16), 0};
            return DMP_OK;
        case 2:
            *out = DmpLayout{64 * n, 0, PLAN2D_WORDS * 4, D2_OBS, D2_ACT, 26, 26, 600, 600, (int32_t)sizeof(Rec56),
                             (int32_t)sizeof(Bits16)};
            return DMP_OK;
        case 3:
            *out = DmpLayout{(800 + NIB3_STRIDE) * n, 16 * n, CELLS3D, D3_OBS, D3_ACT, 26, 26, 1300, 1000,    // u16 maps + nibble maps
                             (int32_t)sizeof(Rec56), (int32_t)sizeof(Bits32)};
            return DMP_OK;
    }
    return DMP_EINVAL;
}

int dmp_plan_static(int dim, int plan_choose, void* row_out, int32_t* total_out, void* stream) {
    if (dim < 1 || dim > 3 || !row_out || !total_out) return DMP_EINVAL;
    if (plan_choose < 0 || plan_choose > (dim == 1 ? 2 : 1)) return DMP_EINVAL;     // reference: ValueError
    k_plan_static<<<1, 416, 0, as_stream(stream)>>>(dim, plan_choose, row_out, total_out);
    return dmp_set_error(cudaGetLastError());
}

int dmp_plans_pack(int dim, const double* raw, int n_plans, void* plans_out, int32_t* total_out, void* stream) {
    if (dim < 1 || dim > 3 || !raw || n_plans < 1 || !plans_out || !total_out) return DMP_EINVAL;
    k_plans_pack<<<n_plans, 128, 0, as_stream(stream)>>>(dim, raw, plans_out, total_out);
    return dmp_set_error(cudaGetLastError());
}

int dmp_reset(const DmpState* st, const uint8_t* mask, const int32_t* plan_idx, uint64_t t_draw, void* obs,
              int obs_kind, void* stream) {
    if (!state_ok(st)) return DMP_EINVAL;
    switch (st->dim) {
        case 1: return dmp1d_reset(*st, mask, plan_idx, t_draw, obs, obs_kind, as_stream(stream));
        case 2: return dmp2d_reset(*st, mask, plan_idx, t_draw, obs, obs_kind, as_stream(stream));
        default: return dmp3d_reset(*st, mask, plan_idx, t_draw, obs, obs_kind, as_stream(stream));
    }
}

int dmp_rollout(const DmpState* st, const DmpIO* io, int K, void* stream) {
    if (!state_ok(st) || !io || K < 1) return DMP_EINVAL;
    if (io->obs_kind < DMP_OBS_F32 || io->obs_kind > DMP_OBS_BITS) return DMP_EINVAL;
    if (io->obs_kind == DMP_OBS_BITS && (st->dim == 1 || (reinterpret_cast<uintptr_t>(io->obs) & 15))) return DMP_EINVAL;
    if ((io->flags & DMP_F_NORMALISE) && io->obs_kind >= DMP_OBS_I16) return DMP_EINVAL;
    if ((io->flags & DMP_F_AUTORESET) && (!st->ep_cnt || !st->ep_len || !st->ep_ret || !st->ep_iou)) return DMP_EINVAL;
    switch (st->dim) {
        case 1: return dmp1d_rollout(*st, *io, K, as_stream(stream));
        case 2: return dmp2d_rollout(*st, *io, K, as_stream(stream));
        default: return dmp3d_rollout(*st, *io, K, as_stream(stream));
    }
}

int dmp_step(const DmpState* st, const DmpIO* io, void* stream) { return dmp_rollout(st, io, 1, stream); }

int dmp_records_unpack(int dim, int rec_kind, const void* records, int64_t n, void* obs, int obs_kind, float* reward,
                       uint8_t* done, uint8_t* saturated, void* stream) {
    if (dim < 1 || dim > 3 || !records || n < 1) return DMP_EINVAL;
    if (rec_kind != DMP_OBS_REC && !(rec_kind == DMP_OBS_BITS && dim != 1)) return DMP_EINVAL;
    if (obs && (obs_kind < DMP_OBS_F32 || obs_kind > DMP_OBS_I16)) return DMP_EINVAL;
    const int64_t total = n * (dim == 1 ? D1_OBS : D2_OBS);
    const unsigned blocks = (unsigned)((total + 255) / 256);
    const uint8_t* r = reinterpret_cast<const uint8_t*>(records);
    cudaStream_t s = as_stream(stream);
    switch (obs ? obs_kind : DMP_OBS_F32) {
        case DMP_OBS_F32: k_records_unpack<float><<<blocks, 256, 0, s>>>(dim, rec_kind, r, n, (float*)obs, reward, done, saturated); break;
        case DMP_OBS_F64: k_records_unpack<double><<<blocks, 256, 0, s>>>(dim, rec_kind, r, n, (double*)obs, reward, done, saturated); break;
        default: k_records_unpack<int16_t><<<blocks, 256, 0, s>>>(dim, rec_kind, r, n, (int16_t*)obs, reward, done, saturated); break;
    }
    return dmp_set_error(cudaGetLastError());
}

int dmp_iou(const DmpState* st, double* iou_out, void* stream) {
    if (!state_ok(st) || !iou_out) return DMP_EINVAL;
    switch (st->dim) {
        case 1: return dmp1d_iou(*st, iou_out, as_stream(stream));
        case 2: return dmp2d_iou(*st, iou_out, as_stream(stream));
        default: return dmp3d_iou(*st, iou_out, as_stream(stream));
    }
}

int64_t dmp_stats_scratch_bytes(int64_t n_envs) { return (int64_t)stats_blocks(n_envs) * 4 * (int64_t)sizeof(double); }

int dmp_stats_reduce(const DmpState* st, double* out4, void* scratch, void* stream) {
    if (!state_ok(st) || !out4 || !scratch || !st->ep_cnt || !st->ep_len || !st->ep_ret || !st->ep_iou) return DMP_EINVAL;
    const int b = stats_blocks(st->n_envs);
    k_stats_pass1<<<b, STATS_BLOCK, 0, as_stream(stream)>>>(*st, reinterpret_cast<double*>(scratch));
    k_stats_pass2<<<1, STATS_BLOCK, 0, as_stream(stream)>>>(reinterpret_cast<const double*>(scratch), b, out4);
    return dmp_set_error(cudaGetLastError());
}

int dmp_stats_clear(const DmpState* st, void* stream) {
    if (!state_ok(st) || !st->ep_cnt || !st->ep_len || !st->ep_ret || !st->ep_iou) return DMP_EINVAL;
    k_stats_clear<<<(unsigned)((st->n_envs + 255) / 256), 256, 0, as_stream(stream)>>>(*st);
    return dmp_set_error(cudaGetLastError());
}

int dmp_export_state(const DmpState* st, int32_t* grid, int32_t* scalars, float* ret_acc, void* stream) {
    if (!state_ok(st)) return DMP_EINVAL;
    switch (st->dim) {
        case 1: return dmp1d_export(*st, grid, scalars, ret_acc, as_stream(stream));
        case 2: return dmp2d_export(*st, grid, scalars, ret_acc, as_stream(stream));
        default: return dmp3d_export(*st, grid, scalars, ret_acc, as_stream(stream));
    }
}

int dmp_import_state(const DmpState* st, const int32_t* grid, const int32_t* scalars, const float* ret_acc, void* stream) {
    if (!state_ok(st)) return DMP_EINVAL;
    switch (st->dim) {
        case 1: return dmp1d_import(*st, grid, scalars, ret_acc, as_stream(stream));
        case 2: return dmp2d_import(*st, grid, scalars, ret_acc, as_stream(stream));
        default: return dmp3d_import(*st, grid, scalars, ret_acc, as_stream(stream));
    }
}

}  // extern "C"

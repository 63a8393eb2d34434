// l2_affinity.cu -- probe for the 3D single-step kernel's access pattern: can 54.5 MB of per-env maps stay in B200's L2 across
// launches while 53.5 MB of observations stream out per launch?  Each warp owns tiles of 32 envs; per tile every lane reads
// 128 B of its env's 208 B map (8 x LDG.128... as 16 B pieces at a step-dependent offset) and the warp writes a 6 528 B tile.
//   grid: "wave" = one block per env pair of warps (4 096 blocks, ~2 waves, block -> SM assignment free), or
//         "persist" = 148 x 14 blocks that loop over a FIXED set of tiles (block -> SM assignment stable if the hardware
//         assigns the first wave deterministically: also probed, via %smid)
//   hint: none / evict_last on the map reads (writes are always evict_first streaming)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/l2_affinity tools/probe/l2_affinity.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint4 ld_hint(const uint4* p, uint64_t pol) {
    uint4 v; asm volatile("ld.global.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ void st_hint(uint4* p, uint4 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// the real kernel's read path: ONE bulk async copy (TMA) of 128 B per lane into a shared-memory slot, completion on the warp's
// mbarrier; MODE 0 plain, 1 evict_last, 2 evict_first hint on the copies
template <int MODE>
__global__ void __launch_bounds__(64) k(const uint8_t* __restrict__ maps, uint4* __restrict__ obs, int64_t n_tiles, int step, unsigned* smid_out, int what = 3, int passes = 1, size_t ring_stride = 0, int wmode = 0) {
    __shared__ __align__(16) uint8_t slots[2][32][144];
    __shared__ uint64_t bars[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t wg = (int64_t)blockIdx.x * 2 + warp, wtot = (int64_t)gridDim.x * 2;
    if (smid_out && threadIdx.x == 0) { unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s)); smid_out[blockIdx.x] = s; }
    const uint64_t pl = MODE == 2 ? pol_first() : pol_last(), pf = pol_first();
    uint64_t* bar = &bars[warp];
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(32) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0;
    for (int pass = 0; pass < passes; ++pass, ++step)
    for (int64_t t = wg; t < n_tiles; t += wtot) {
        const int64_t env = t * 32 + lane;
        const unsigned off = ((unsigned)(env * 2654435761u + step * 40503u) >> 7) % 6u * 16u;
        const uint8_t* src = maps + env * 208 + off;
        uint8_t* slot = slots[warp][lane];
        uint4 acc = make_uint4(lane, step, 0, 0);
        if (what & 1) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(128) : "memory");
        if (MODE)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(smem_u32(slot)), "l"(src), "r"(128), "r"(smem_u32(bar)), "l"(pl) : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(slot)), "l"(src), "r"(128), "r"(smem_u32(bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        parity ^= 1;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 v = reinterpret_cast<const uint4*>(slot)[i];
            acc.x ^= v.x; acc.y += v.y; acc.z ^= v.z; acc.w += v.w;
        }
        __syncwarp();
        }
        uint4* dst = obs + (size_t)(pass & 3) * ring_stride + t * 408 + lane;                 // 6 528 B per tile = 408 x 16 B
        if (what & 2) {
#pragma unroll
        for (int i = 0; i < 13; ++i)
            if (i * 32 + lane < 408) {
                if (wmode == 0) st_hint(dst + i * 32, acc, pf);
                else if (wmode == 1) dst[i * 32] = acc;
                else __stcs(dst + i * 32, acc);
            }
        } else if (acc.x == 0x12345678u && acc.y == 77u) {
            dst[0] = acc;
        }
    }
}

int main() {
    const int64_t n = 262144, n_tiles = n / 32;
    uint8_t* maps; uint4* obs; unsigned* smid;
    const size_t obs_bytes = (size_t)n_tiles * 6528;
    CK(cudaMalloc(&maps, n * 208 + 256));
    CK(cudaMalloc(&obs, obs_bytes * 4));
    CK(cudaMalloc(&smid, 8192 * sizeof(unsigned)));
    CK(cudaMemset(maps, 1, n * 208 + 256));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // (1) is the block -> SM assignment of a one-wave grid stable across launches?
    {
        const int G = 148 * 14;
        std::vector<unsigned> a(G), b(G);
        k<0><<<G, 64>>>(maps, obs, n_tiles, 0, smid); CK(cudaMemcpy(a.data(), smid, G * 4, cudaMemcpyDeviceToHost));
        int same = 0, tot = 0;
        for (int r = 0; r < 20; ++r) {
            k<0><<<G, 64>>>(maps, obs, n_tiles, r, smid); CK(cudaMemcpy(b.data(), smid, G * 4, cudaMemcpyDeviceToHost));
            for (int i = 0; i < G; ++i) { same += a[i] == b[i]; ++tot; }
        }
        printf("one-wave grid of %d blocks: block -> SM assignment identical to the first launch for %.1f %% of blocks over 20 launches\n", G, 100.0 * same / tot);
        k<0><<<4096, 64>>>(maps, obs, n_tiles, 0, smid); CK(cudaMemcpy(a.data(), smid, G * 4, cudaMemcpyDeviceToHost));
        same = tot = 0;
        for (int r = 0; r < 20; ++r) {
            k<0><<<4096, 64>>>(maps, obs, n_tiles, r, smid); CK(cudaMemcpy(b.data(), smid, G * 4, cudaMemcpyDeviceToHost));
            for (int i = 0; i < G; ++i) { same += a[i] == b[i]; ++tot; }
        }
        printf("two-wave grid of 4096 blocks: %.1f %% (first %d blocks)\n", 100.0 * same / tot, G);
    }
    // (2) time per launch, 300 launches back to back, obs ring of 4 buffers (214 MB > L2); with and without an L2 set-aside
    // for persisting accesses (cudaLimitPersistingL2CacheSize), at the full and at half the batch
    int dev = 0, maxp = 0, l2 = 0;
    cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
    printf("L2 %d MB, max persisting set-aside %d MB\n", l2 >> 20, maxp >> 20);
    // (3) the same persistent grid looping over its tiles `passes` times inside ONE launch (what PDL-overlapped launches
    // approach): time per pass with the launch overhead amortised.  Variants: read hint (none / evict_last), write policy
    // (evict_first hint / plain / st.cs), and the driver's own mechanism -- an access-policy window over the maps
    // (hitProp persisting, missProp streaming) with an L2 set-aside.
    cudaStream_t strm; CK(cudaStreamCreate(&strm));
    for (int window_mb : {0, 60, 79})
      for (int64_t nt : {n_tiles, n_tiles / 2})
        for (int hint = 0; hint < 2; ++hint)
          for (int wmode = 0; wmode < 3; ++wmode) {
            if (window_mb && (hint || wmode == 1)) continue;
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)window_mb << 20));
            cudaStreamAttrValue av = {};
            av.accessPolicyWindow.base_ptr = maps;
            av.accessPolicyWindow.num_bytes = window_mb ? (size_t)nt * 32 * 208 : 0;
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = window_mb ? cudaAccessPropertyPersisting : cudaAccessPropertyNormal;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            CK(cudaStreamSetAttribute(strm, cudaStreamAttributeAccessPolicyWindow, &av));
            const int G = 148 * 14, P = 32, what = 3;
            auto launch = [&]() {
                if (hint) k<1><<<G, 64, 0, strm>>>(maps, obs, nt, 0, nullptr, what, P, obs_bytes / 16, wmode);
                else k<0><<<G, 64, 0, strm>>>(maps, obs, nt, 0, nullptr, what, P, obs_bytes / 16, wmode);
            };
            launch(); launch();
            CK(cudaStreamSynchronize(strm));
            cudaEventRecord(e0, strm);
            const int R = 10;
            for (int r = 0; r < R; ++r) launch();
            cudaEventRecord(e1, strm);
            CK(cudaStreamSynchronize(strm));
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double req = (double)nt * 32 * (128 + 204);
            const double us = ms * 1e3 / R / P;
            printf("window %2d MB  envs %6lld (maps %4.1f MB)  reads %-10s writes %-11s %6.2f us per pass  %5.0f GB/s of requested bytes\n",
                   window_mb, (long long)nt * 32, nt * 32 * 208 / 1e6, hint ? "evict_last" : "no hint",
                   wmode == 0 ? "evict_first" : (wmode == 1 ? "plain" : "st.cs"), us, req / (us * 1e-6) / 1e9);
            CK(cudaCtxResetPersistingL2Cache());
          }
    return 0;
}

// Backward for channels-last tensors:  dz_b = M_b z_b  (+ [z_b > 0] * grelu_b, the fused DeepWT tail), z / grelu / dz all
// [B][P][16] -- the memory of channels-last B x 16 x H x W tensors, so a channels-last backbone needs no layout
// conversion around the loss.
//
// A pixel's 16 channels are 64 contiguous bytes: every thread loads its own pixel with four 128-bit loads (a warp
// covers 2 KB contiguous) through a 4-deep register ring (three pixels per thread in flight), multiplies by M_b held in
// shared memory (64 warp-uniform LDS.128 per pixel) and stores 64 contiguous bytes.  No shared-memory staging of the
// streams: there is no reuse and the per-thread accesses are already contiguous.  Persistent CTAs, blocks of 1024
// pixels dealt round-robin (adjacent CTAs stream adjacent memory).  Rounding as in the NCHW kernels: the
// 16-term product is accumulated first, the masked ReLU gradient added last.
#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kThreads = 256;
constexpr int kBlockSteps = 4;             // 4 x 256 pixels = 64 KB of each stream per schedule block

__device__ __forceinline__ void st_cs4(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_cs4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

template <bool kReluGrad>
__global__ void __launch_bounds__(kThreads, 1)
apply_cl_kernel(const float* __restrict__ z, const float* __restrict__ grelu, const float* __restrict__ mmat,
                float* __restrict__ dz, long long P, long long steps_per_sample, long long total_steps) {
    __shared__ __align__(16) float msh[256];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    // A step = 256 consecutive pixels of one sample (one per thread).  Blocks of kBlockSteps consecutive steps (1024
    // pixels, 64 KB per stream) are dealt round-robin to the CTAs, so the grid streams adjacent memory at any moment.
    const long long G = gridDim.x, bx = blockIdx.x;
    const long long nblk = (total_steps + kBlockSteps - 1) / kBlockSteps;
    const long long my_steps = (nblk > bx ? (nblk - bx + G - 1) / G : 0) * kBlockSteps;
    auto global_step = [&](long long k) { return ((k / kBlockSteps) * G + bx) * kBlockSteps + (k % kBlockSteps); };

    struct Px { float4 x[4]; float4 g[kReluGrad ? 4 : 1]; };
    auto load_step = [&](Px& r, long long k) {
        if (k >= my_steps) return;
        const long long s = global_step(k);
        if (s >= total_steps) return;
        const long long b = s / steps_per_sample;
        const long long p = (s - b * steps_per_sample) * kThreads + tid;
        if (p >= P) return;
        const float4* src = reinterpret_cast<const float4*>(z + (b * P + p) * kC);
#pragma unroll
        for (int q = 0; q < 4; ++q) r.x[q] = ld_cs4(src + q);
        if (kReluGrad) {
            const float4* gs = reinterpret_cast<const float4*>(grelu + (b * P + p) * kC);
#pragma unroll
            for (int q = 0; q < 4; ++q) r.g[q] = ld_cs4(gs + q);
        }
    };
    long long cur_b = -1;
    auto use_step = [&](const Px& r, long long k) {
        if (k >= my_steps) return;                            // every condition up to the barrier is uniform over the CTA
        const long long s = global_step(k);
        if (s >= total_steps) return;
        const long long b = s / steps_per_sample;
        if (b != cur_b) {
            __syncthreads();                                  // previous matrix no longer in use
            if (cur_b < 0) asm volatile("griddepcontrol.wait;" ::: "memory");   // the matrices come from the primary kernel
            msh[tid] = __ldcg(mmat + b * 256 + tid);   // coherent load: an invariant (.nc) one may be hoisted above the wait
            __syncthreads();
            cur_b = b;
        }
        const long long p = (s - b * steps_per_sample) * kThreads + tid;
        if (p >= P) return;
        const float x[kC] = {r.x[0].x, r.x[0].y, r.x[0].z, r.x[0].w, r.x[1].x, r.x[1].y, r.x[1].z, r.x[1].w,
                             r.x[2].x, r.x[2].y, r.x[2].z, r.x[2].w, r.x[3].x, r.x[3].y, r.x[3].z, r.x[3].w};
        float out[kC];
#pragma unroll
        for (int i = 0; i < kC; ++i) {
            float a = 0.f;
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                const float4 m = *reinterpret_cast<const float4*>(msh + i * kC + 4 * jq);
                a = fmaf(m.x, x[4 * jq + 0], a);
                a = fmaf(m.y, x[4 * jq + 1], a);
                a = fmaf(m.z, x[4 * jq + 2], a);
                a = fmaf(m.w, x[4 * jq + 3], a);
            }
            out[i] = a;
        }
        if (kReluGrad) {
            const float g[kC] = {r.g[0].x, r.g[0].y, r.g[0].z, r.g[0].w, r.g[1].x, r.g[1].y, r.g[1].z, r.g[1].w,
                                 r.g[2].x, r.g[2].y, r.g[2].z, r.g[2].w, r.g[3].x, r.g[3].y, r.g[3].z, r.g[3].w};
#pragma unroll
            for (int i = 0; i < kC; ++i) out[i] += x[i] <= 0.f ? 0.f : g[i];
        }
        float4* d = reinterpret_cast<float4*>(dz + (b * P + p) * kC);
#pragma unroll
        for (int q = 0; q < 4; ++q) st_cs4(d + q, make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]));
    };

    Px r0, r1, r2, r3;
    load_step(r0, 0);
    load_step(r1, 1);
    load_step(r2, 2);
    for (long long k = 0; k < my_steps; k += 4) {
        load_step(r3, k + 3); use_step(r0, k);
        load_step(r0, k + 4); use_step(r1, k + 1);
        load_step(r1, k + 5); use_step(r2, k + 2);
        load_step(r2, k + 6); use_step(r3, k + 3);
    }
}

}  // namespace

cudaError_t launch_apply_cl(const float* z, const float* grelu, const float* mmat, float* dz, int B, long long P, int sm_count,
                            cudaStream_t stream, bool programmatic_dependent) {
    const long long sps = (P + kThreads - 1) / kThreads;       // steps per sample
    const long long total = sps * B;
    const long long nblk = (total + kBlockSteps - 1) / kBlockSteps;
    const long long G = nblk < sm_count ? nblk : sm_count;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(G));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = programmatic_dependent ? 1 : 0;
    if (grelu) return cudaLaunchKernelEx(&cfg, apply_cl_kernel<true>, z, grelu, mmat, dz, P, sps, total);
    return cudaLaunchKernelEx(&cfg, apply_cl_kernel<false>, z, grelu, mmat, dz, P, sps, total);
}

}  // namespace wtpse

// Backward for channels-last tensors:  dz_b = M_b z_b  (+ [z_b > 0] * grelu_b, the fused DeepWT tail), z / grelu / dz all
// [B][P][16] -- the memory of channels-last B x 16 x H x W tensors, so a channels-last backbone needs no layout
// conversion around the loss.
//
// A pixel's 16 channels are 64 contiguous bytes: every thread loads its own pixels with four 128-bit loads each (a warp
// covers 2 KB contiguous) through a two-deep register ring, multiplies by M_b held in shared memory (warp-uniform
// LDS.128, amortised over the 2-4 pixels a thread owns) and stores 64 contiguous bytes per pixel.  No shared-memory
// staging of the streams: there is no reuse and the per-thread accesses are already contiguous.  Persistent CTAs,
// steps of 512 / 1024 pixels dealt round-robin (adjacent CTAs stream adjacent memory).  Rounding as in the NCHW
// kernels: the 16-term product is accumulated first, the masked ReLU gradient added last.
#include "common.cuh"
#include "kernels.h"
#include "whitening_matrix.cuh"

namespace wtpse {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ void st_cs4(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_cs4(const float4* p) {
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// kPx pixels per thread (tid, tid + 256, ...): one warp-uniform LDS.128 of M_b then feeds 4 * kPx FMAs.  With one pixel
// per thread the kernel was bound by the shared-memory pipe (64 LDS.128 per pixel, ~4 cycles each: 427 us for the fused
// variant at 32x16x512x512 = 0.58 of the HBM roofline).  Plain: 4 pixels per thread; fused (the ReLU gradient doubles
// the staged registers): 2.  A two-deep register ring keeps the next step's loads (64 KB per SM) in flight.
template <bool kReluGrad>
__global__ void __launch_bounds__(kThreads, 1)
apply_cl_kernel(const float* __restrict__ z, const float* __restrict__ grelu, float* __restrict__ dz, long long P,
                long long steps_per_sample, long long total_steps, SeedArgs sa) {
    constexpr int kPx = kReluGrad ? 2 : 4;
    constexpr int kStepPx = kPx * kThreads;
    __shared__ __align__(16) float msh[256];
    __shared__ IndexTables tab;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    build_index_tables(tab, tid, kThreads);
    // A step = kStepPx consecutive pixels of one sample, dealt round-robin to the CTAs: the grid streams adjacent memory.
    const long long G = gridDim.x, bx = blockIdx.x;
    const long long my_steps = total_steps > bx ? (total_steps - bx + G - 1) / G : 0;

    struct Stage { float4 x[kPx][4]; float4 g[kReluGrad ? kPx : 1][4]; };
    auto load_step = [&](Stage& r, long long k) {
        if (k >= my_steps) return;
        const long long s = k * G + bx;
        const long long b = s / steps_per_sample;
        const long long p0 = (s - b * steps_per_sample) * kStepPx + tid;
#pragma unroll
        for (int u = 0; u < kPx; ++u) {
            const long long p = p0 + u * kThreads;
            if (p < P) {
                const float4* src = reinterpret_cast<const float4*>(z + (b * P + p) * kC);
#pragma unroll
                for (int q = 0; q < 4; ++q) r.x[u][q] = ld_cs4(src + q);
                if (kReluGrad) {
                    const float4* gs = reinterpret_cast<const float4*>(grelu + (b * P + p) * kC);
#pragma unroll
                    for (int q = 0; q < 4; ++q) r.g[u][q] = ld_cs4(gs + q);
                }
            }
        }
    };
    long long cur_b = -1;
    SeedCtx sc{};
    auto use_step = [&](const Stage& r, long long k) {
        if (k >= my_steps) return;                            // every condition up to the barrier is uniform over the CTA
        const long long s = k * G + bx;
        const long long b = s / steps_per_sample;
        if (b != cur_b) {
            __syncthreads();                                  // previous matrix no longer in use
            seed_matrix(sa, sc, tab, int(b), msh, tid);
            __syncthreads();
            cur_b = b;
        }
        const long long p0 = (s - b * steps_per_sample) * kStepPx + tid;
        float4 out[kPx][4];
#pragma unroll
        for (int u = 0; u < kPx; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) out[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
        // out[u][i] = sum_j M[i][j] x[u][j]; per (i, jq) one LDS.128 of M feeds 4 FMAs per pixel
#pragma unroll
        for (int i = 0; i < kC; ++i) {
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                const float4 m = *reinterpret_cast<const float4*>(msh + i * kC + 4 * jq);
#pragma unroll
                for (int u = 0; u < kPx; ++u) {
                    float& o = (i & 3) == 0 ? out[u][i >> 2].x : (i & 3) == 1 ? out[u][i >> 2].y : (i & 3) == 2 ? out[u][i >> 2].z : out[u][i >> 2].w;
                    o = fmaf(m.x, r.x[u][jq].x, o);
                    o = fmaf(m.y, r.x[u][jq].y, o);
                    o = fmaf(m.z, r.x[u][jq].z, o);
                    o = fmaf(m.w, r.x[u][jq].w, o);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kPx; ++u) {
            const long long p = p0 + u * kThreads;
            if (p >= P) continue;
            float4* d = reinterpret_cast<float4*>(dz + (b * P + p) * kC);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 o = out[u][q];
                if (kReluGrad) {                               // added last: the rounding of autograd's dz_loss + dz_relu
                    const float4 xv = r.x[u][q], gv = r.g[u][q];
                    o.x += xv.x <= 0.f ? 0.f : gv.x; o.y += xv.y <= 0.f ? 0.f : gv.y;
                    o.z += xv.z <= 0.f ? 0.f : gv.z; o.w += xv.w <= 0.f ? 0.f : gv.w;
                }
                st_cs4(d + q, o);
            }
        }
    };

    Stage r0, r1;
    // z was written before the forward pass ran and may be prefetched at once; grelu, the saved tensors, the upstream
    // scalars and dz may belong to the kernel in front of us (programmatic dependent launch): wait for it first.
    if (kReluGrad) asm volatile("griddepcontrol.wait;" ::: "memory");
    load_step(r0, 0);
    if (!kReluGrad) asm volatile("griddepcontrol.wait;" ::: "memory");
    sc = seed_context(sa, P);
    __syncthreads();                                          // index tables
    for (long long k = 0; k < my_steps; k += 2) {
        load_step(r1, k + 1); use_step(r0, k);
        load_step(r0, k + 2); use_step(r1, k + 1);
    }
}

}  // namespace

cudaError_t launch_apply_cl(const float* z, const float* grelu, const SeedArgs& seed, float* dz, int B, long long P, int sm_count,
                            cudaStream_t stream) {
    const long long step_px = (grelu ? 2 : 4) * kThreads;      // kPx * kThreads of the instantiation launched below
    const long long sps = (P + step_px - 1) / step_px;         // steps per sample
    const long long total = sps * B;
    const long long G = total < sm_count ? total : sm_count;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(G));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (grelu) return cudaLaunchKernelEx(&cfg, apply_cl_kernel<true>, z, grelu, dz, P, sps, total, seed);
    return cudaLaunchKernelEx(&cfg, apply_cl_kernel<false>, z, grelu, dz, P, sps, total, seed);
}

}  // namespace wtpse

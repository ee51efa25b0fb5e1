// Backward stage 2: dz_b = M_b z_b, the adjoint of the Gram, fused with the gradient write.
//
// What autograd derives from the bmm of algorithms.py:1283 (two more passes over z in the reference:
// grad @ z and grad^T @ z) collapses to ONE 16x16 symmetric matrix per sample applied to every pixel:
// read z once (64 B/pixel), write dz once (64 B/pixel), 256 FMAs per pixel.
//
// Same persistent layout as the forward Gram kernel: contiguous tile ranges per CTA, a producer warp
// feeding a 3-stage shared-memory ring with 1-D TMA bulk copies, 8 consumer warps with 4 pixels per
// thread.  M_b (1 KB) sits in shared memory and is read with warp-uniform LDS.128 (broadcast).
#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kConsumers = 256;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kThreads = kConsumers + 32;
constexpr int kTilePx = kConsumers * 4;
constexpr int kStages = 3;
constexpr int kStageFloats = kC * kTilePx;
constexpr size_t kSmemBytes = size_t(kStages) * kStageFloats * sizeof(float) + 256 * sizeof(float) + 2 * kStages * sizeof(uint64_t);

__device__ __forceinline__ void st_stream(float* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
apply_tma_kernel(const float* __restrict__ z, const float* __restrict__ mmat, float* __restrict__ dz, long long P,
                 long long tiles_per_sample, long long T) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* msh = stage_buf + size_t(kStages) * kStageFloats;
    uint64_t* full = reinterpret_cast<uint64_t*>(msh + 256);
    uint64_t* empty = full + kStages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    const long long t0 = part_begin(k, T, G), t1 = part_begin(k + 1, T, G);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (long long t = t0; t < t1; ++t) {
                mbar_wait(&empty[stage], phase ^ 1);
                const long long b = t / tiles_per_sample;
                const long long px0 = (t - b * tiles_per_sample) * kTilePx;
                const long long rem = P - px0;
                const uint32_t npx = rem < kTilePx ? uint32_t(rem) : uint32_t(kTilePx);
                const uint32_t bytes = npx * 4u;
                mbar_arrive_expect_tx(&full[stage], bytes * kC);
                const float* src = z + (b * kC) * P + px0;
                float* dst = stage_buf + size_t(stage) * kStageFloats;
#pragma unroll
                for (int c = 0; c < kC; ++c) tma_load_1d(dst + c * kTilePx, src + c * P, bytes, &full[stage]);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    int stage = 0;
    uint32_t phase = 0;
    long long cur_b = -1;
    for (long long t = t0; t < t1; ++t) {
        const long long b = t / tiles_per_sample;
        const long long px0 = (t - b * tiles_per_sample) * kTilePx;
        const long long rem = P - px0;
        if (b != cur_b) {
            named_bar_sync(1, kConsumers);          // everyone is done with the previous sample's matrix
            msh[tid] = __ldg(mmat + b * 256 + tid);
            named_bar_sync(1, kConsumers);
            cur_b = b;
        }
        mbar_wait(&full[stage], phase);
        const bool active = 4LL * tid < rem;
        float4 out[kC];
        if (active) {
#pragma unroll
            for (int i = 0; i < kC; ++i) out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* src = stage_buf + size_t(stage) * kStageFloats + 4 * tid;
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                float4 x[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) x[r] = *reinterpret_cast<const float4*>(src + (4 * jq + r) * kTilePx);
#pragma unroll
                for (int i = 0; i < kC; ++i) {
                    const float4 m = *reinterpret_cast<const float4*>(msh + i * kC + 4 * jq);
                    out[i].x = fmaf(m.x, x[0].x, out[i].x); out[i].y = fmaf(m.x, x[0].y, out[i].y);
                    out[i].z = fmaf(m.x, x[0].z, out[i].z); out[i].w = fmaf(m.x, x[0].w, out[i].w);
                    out[i].x = fmaf(m.y, x[1].x, out[i].x); out[i].y = fmaf(m.y, x[1].y, out[i].y);
                    out[i].z = fmaf(m.y, x[1].z, out[i].z); out[i].w = fmaf(m.y, x[1].w, out[i].w);
                    out[i].x = fmaf(m.z, x[2].x, out[i].x); out[i].y = fmaf(m.z, x[2].y, out[i].y);
                    out[i].z = fmaf(m.z, x[2].z, out[i].z); out[i].w = fmaf(m.z, x[2].w, out[i].w);
                    out[i].x = fmaf(m.w, x[3].x, out[i].x); out[i].y = fmaf(m.w, x[3].y, out[i].y);
                    out[i].z = fmaf(m.w, x[3].z, out[i].z); out[i].w = fmaf(m.w, x[3].w, out[i].w);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (active) {
            float* dst = dz + (b * kC) * P + px0 + 4 * tid;
#pragma unroll
            for (int i = 0; i < kC; ++i) st_stream(dst + i * P, out[i]);
        }
    }
}

// Fallback (P % 4 != 0 or unaligned pointers): one pixel per thread, scalar coalesced accesses.
// grid = (ceil(P/256), B)
__global__ void __launch_bounds__(256)
apply_generic_kernel(const float* __restrict__ z, const float* __restrict__ mmat, float* __restrict__ dz, long long P) {
    __shared__ float msh[256];
    const long long b = blockIdx.y;
    msh[threadIdx.x] = mmat[b * 256 + threadIdx.x];
    __syncthreads();
    const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
    if (p >= P) return;
    const float* zb = z + b * kC * P + p;
    float x[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) x[c] = __ldg(zb + c * P);
    float* out = dz + b * kC * P + p;
#pragma unroll
    for (int i = 0; i < kC; ++i) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < kC; ++j) s = fmaf(msh[i * kC + j], x[j], s);
        out[i * P] = s;
    }
}

}  // namespace

cudaError_t launch_apply(const float* z, const float* mmat, float* dz, int B, long long P, int sm_count,
                         cudaStream_t stream) {
    const bool tma = (P % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(dz) & 15u) == 0);
    if (tma) {
        cudaError_t e = cudaFuncSetAttribute(apply_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
        if (e != cudaSuccess) return e;
        const long long tps = (P + kTilePx - 1) / kTilePx;
        const long long T = tps * B;
        const long long G = T < sm_count ? T : sm_count;
        apply_tma_kernel<<<dim3(unsigned(G)), kThreads, kSmemBytes, stream>>>(z, mmat, dz, P, tps, T);
    } else {
        apply_generic_kernel<<<dim3(unsigned((P + 255) / 256), unsigned(B)), 256, 0, stream>>>(z, mmat, dz, P);
    }
    return cudaGetLastError();
}

}  // namespace wtpse

// Backward of the fused DeepWT tail:  dz_b = M_b z_b + [z_b > 0] * grelu_b.
//
// In DeepWT.forward (algorithms.py:1099-1113) every embedding z that feeds the whitening loss is followed by
// F.relu(z).  Autograd then makes three passes for dz: the loss adjoint (whitening_apply.cu: read z, write),
// the ReLU backward (read grelu, read relu(z), write) and the sum of the two (read, read, write) -- 512 B per
// pixel.  Here it is one pass: read z, read grelu, write dz = 192 B per pixel.
//
// Same structure as apply_tma_kernel (persistent CTAs, tiles dealt round-robin, M_b derived in the kernel from the
// forward's seed: one launch), but BOTH streams are staged by the producer warp with 1-D TMA bulk copies:
// a stage holds 16 channel rows of 896 pixels of z and of grelu (112 KB), two stages fill the shared memory.
// (Reading grelu with per-thread 128-bit global loads instead -- no lead time, 8 warps per SM -- reached 0.78 of
// the HBM roofline: 315 us at 32x16x512x512 against 246 us for the bytes.)  7 consumer warps + the producer warp
// = 256 threads, so each thread may hold 255 registers: 64 outputs + 64 masked gradients.
//
// Rounding: the masked gradient is added AFTER the 16-term matrix product, so dz is bit-identical to autograd's
// `dz_loss + dz_relu` (tests/test_gpu_fusion.py).  ReLU backward is ATen's threshold_backward: the gradient
// passes unless z <= 0 (so it passes for NaN).
#include "common.cuh"
#include "kernels.h"
#include "whitening_matrix.cuh"

namespace wtpse {

namespace {

constexpr int kConsumers = 224;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kThreads = kConsumers + 32;
constexpr int kTilePx = kConsumers * 4;            // 896 pixels
constexpr int kStages = 2;
constexpr int kRowsFloats = kC * kTilePx;          // one tensor's share of a stage: 56 KB
constexpr int kStageFloats = 2 * kRowsFloats;      // z rows, then grelu rows
constexpr size_t kSmemBytes = size_t(kStages) * kStageFloats * sizeof(float) + 2 * 256 * sizeof(float) + 2 * kStages * sizeof(uint64_t);

__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
apply_relu_tma_kernel(const float* __restrict__ z, const float* __restrict__ grelu, float* __restrict__ dz, long long P,
                      long long tiles_per_sample, long long T, SeedArgs sa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* msh2 = stage_buf + size_t(kStages) * kStageFloats;          // two matrices: the current sample's and the next one's
    uint64_t* full = reinterpret_cast<uint64_t*>(msh2 + 512);
    uint64_t* empty = full + kStages;
    __shared__ IndexTables tab;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;

    build_index_tables(tab, tid, kThreads);
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
        // producer.  z was written before the forward pass ran and is streamed at once; grelu is the output of the kernel
        // in FRONT of this one (the backward of whatever consumed relu(z)), and under programmatic dependent launch that
        // kernel may still be running: the first grelu load is issued only after griddepcontrol.wait.  (The consumers'
        // own wait does not cover the producer's loads.)
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();      // both inputs are read exactly once
            auto tile_geom = [&](long long t, long long& off, uint32_t& bytes) {
                const long long b = t / tiles_per_sample;
                const long long px0 = (t - b * tiles_per_sample) * kTilePx;
                const long long rem = P - px0;
                const uint32_t npx = rem < kTilePx ? uint32_t(rem) : uint32_t(kTilePx);
                bytes = npx * 4u;
                off = (b * kC) * P + px0;
            };
            // the ring's first kStages tiles: z rows now, grelu rows after the wait
            long long t = k;
            int primed = 0;
            for (; t < T && primed < kStages; t += G, ++primed) {
                long long off; uint32_t bytes;
                tile_geom(t, off, bytes);
                mbar_arrive_expect_tx(&full[primed], bytes * 2 * kC);
                float* dst = stage_buf + size_t(primed) * kStageFloats;
#pragma unroll
                for (int c = 0; c < kC; ++c) tma_load_1d_hint(dst + c * kTilePx, z + off + c * P, bytes, &full[primed], policy);
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            {
                long long tt = k;
                for (int q = 0; q < primed; ++q, tt += G) {
                    long long off; uint32_t bytes;
                    tile_geom(tt, off, bytes);
                    float* dst = stage_buf + size_t(q) * kStageFloats;
#pragma unroll
                    for (int c = 0; c < kC; ++c)
                        tma_load_1d_hint(dst + kRowsFloats + c * kTilePx, grelu + off + c * P, bytes, &full[q], policy);
                }
            }
            int stage = primed == kStages ? 0 : primed;
            uint32_t phase = primed == kStages ? 1 : 0;
            for (; t < T; t += G) {
                mbar_wait(&empty[stage], phase ^ 1);
                long long off; uint32_t bytes;
                tile_geom(t, off, bytes);
                mbar_arrive_expect_tx(&full[stage], bytes * 2 * kC);
                float* dst = stage_buf + size_t(stage) * kStageFloats;
#pragma unroll
                for (int c = 0; c < kC; ++c) tma_load_1d_hint(dst + c * kTilePx, z + off + c * P, bytes, &full[stage], policy);
#pragma unroll
                for (int c = 0; c < kC; ++c)
                    tma_load_1d_hint(dst + kRowsFloats + c * kTilePx, grelu + off + c * P, bytes, &full[stage], policy);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    asm volatile("griddepcontrol.wait;" ::: "memory");   // saved tensors, upstream scalars and dz may belong to the kernel in front
    TileIter it;
    it.init(k, T, G, tiles_per_sample);                   // tiles dealt round-robin, as the producer walks them
    int stage = 0, cur = 0;
    uint32_t phase = 0;
    long long cur_b = it.valid() ? it.b : -1;
    SeedRegs ahead = seed_load(sa, tab, int(cur_b), tid);
    const SeedCtx sc = seed_context(sa, P);
    seed_store(ahead, sc, tab, msh2, tid);
    ahead = seed_load(sa, tab, it.valid() ? int(it.next_sample()) : -1, tid);     // one sample ahead: no latency at a sample change
    named_bar_sync(1, kConsumers);
    for (; it.valid(); it.next()) {
        const long long b = it.b;
        const long long px0 = it.tin * kTilePx;
        const long long rem = P - px0;
        if (b != cur_b) {
            seed_store(ahead, sc, tab, msh2 + (cur ^ 1) * 256, tid);      // the other buffer was last read two samples ago
            cur ^= 1;
            cur_b = b;
            ahead = seed_load(sa, tab, int(it.next_sample()), tid);
            named_bar_sync(1, kConsumers);
        }
        const float* msh = msh2 + cur * 256;
        mbar_wait(&full[stage], phase);
        const bool active = 4LL * tid < rem;
        float4 out[kC];
        if (active) {
            float4 gm[kC];
#pragma unroll
            for (int i = 0; i < kC; ++i) out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* src = stage_buf + size_t(stage) * kStageFloats + 4 * tid;
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                float4 x[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) x[r] = *reinterpret_cast<const float4*>(src + (4 * jq + r) * kTilePx);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 g = *reinterpret_cast<const float4*>(src + kRowsFloats + (4 * jq + r) * kTilePx);
                    gm[4 * jq + r] = make_float4(x[r].x <= 0.f ? 0.f : g.x, x[r].y <= 0.f ? 0.f : g.y,
                                                 x[r].z <= 0.f ? 0.f : g.z, x[r].w <= 0.f ? 0.f : g.w);
                }
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
#pragma unroll
            for (int i = 0; i < kC; ++i) {
                out[i].x += gm[i].x; out[i].y += gm[i].y; out[i].z += gm[i].z; out[i].w += gm[i].w;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
        if (active) {
            float* dst = dz + (b * kC) * P + px0 + 4 * tid;
#pragma unroll
            for (int i = 0; i < kC; ++i) st_stream4(dst + i * P, out[i]);
        }
    }
}

}  // namespace

bool apply_relu_tma_ok(const float* z, const float* grelu, const float* dz, long long P) {
    return (P % 4 == 0) && (((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(grelu) | reinterpret_cast<uintptr_t>(dz)) & 15u) == 0);
}

cudaError_t launch_apply_relu(const float* z, const float* grelu, const SeedArgs& seed, float* dz, int B, long long P,
                              int sm_count, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(apply_relu_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
    if (e != cudaSuccess) return e;
    const long long tps = (P + kTilePx - 1) / kTilePx;
    const long long T = tps * B;
    const long long G = T < sm_count ? T : sm_count;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(G));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, apply_relu_tma_kernel, z, grelu, dz, P, tps, T, seed);
}

}  // namespace wtpse

// Backward: dz_b = M_b z_b, the adjoint of the Gram, fused with the gradient write.
//
// What autograd derives from the bmm of algorithms.py:1283 (two more passes over z in the reference:
// grad @ z and grad^T @ z) collapses to ONE 16x16 symmetric matrix per sample applied to every pixel:
// read z once (64 B/pixel), write dz once (64 B/pixel), 256 FMAs per pixel.
//
// Persistent CTAs (one per SM), a producer warp feeding a 3-stage shared-memory ring with 1-D TMA bulk
// copies, 8 consumer warps with 4 pixels per thread.  M_b (1 KB) sits in shared memory and is read with
// warp-uniform LDS.128 (broadcast).  Tiles are dealt ROUND-ROBIN to the CTAs: the 148 CTAs then stream
// adjacent 4 KB pieces of each of the 16 channel rows at the same time, which is worth 10 % of DRAM
// throughput over giving every CTA its own contiguous range (2 368 scattered streams).
//
// The matrix M_b is derived in the kernel itself at every sample change from the forward's saved tensors and the three
// upstream scalars (whitening_matrix.cuh): the backward is ONE launch, launched as a programmatic dependent of whatever
// runs in front of it -- z is already streaming while that kernel finishes, and the consumers wait
// (griddepcontrol.wait) only before the first matrix.
#include "common.cuh"
#include "kernels.h"
#include "mmd_device.cuh"
#include "whitening_matrix.cuh"

namespace wtpse {

namespace {

constexpr int kConsumers = 256;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kThreads = kConsumers + 32;
constexpr int kTilePx = kConsumers * 4;
constexpr int kStages = 3;
constexpr int kStageFloats = kC * kTilePx;
constexpr size_t kSmemPipe = size_t(kStages) * kStageFloats * sizeof(float);
constexpr size_t kSmemPlain = kSmemPipe + 2 * 256 * sizeof(float) + 2 * kStages * sizeof(uint64_t);

__device__ __forceinline__ void st_stream(float* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
apply_tma_kernel(const float* __restrict__ z, float* __restrict__ dz, long long P, long long tiles_per_sample, long long T,
                 int round_robin, int l2_hint, SeedArgs sa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* msh2 = stage_buf + size_t(kStages) * kStageFloats;          // two matrices: the current sample's and the next one's
    uint64_t* full = reinterpret_cast<uint64_t*>(msh2 + 512);
    uint64_t* empty = full + kStages;
    __shared__ IndexTables tab;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // e.g. the next Gram kernel: it waits for us itself
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    // tile schedule: tiles dealt round-robin over the grid (default: at any moment the CTAs stream adjacent tiles of
    // every channel row) or one contiguous range per CTA
    TileIter it;
    if (round_robin) it.init(k, T, G, tiles_per_sample);
    else it.init(part_begin(k, T, G), part_begin(k + 1, T, G), 1, tiles_per_sample);

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
        // z was written before the forward pass ran: it is streamed without waiting for the kernel in front of us
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();
            const bool g_use_hint = l2_hint != 0;
            int stage = 0;
            uint32_t phase = 0;
            for (; it.valid(); it.next()) {
                mbar_wait(&empty[stage], phase ^ 1);
                const long long px0 = it.tin * kTilePx;
                const long long rem = P - px0;
                const uint32_t npx = rem < kTilePx ? uint32_t(rem) : uint32_t(kTilePx);
                const uint32_t bytes = npx * 4u;
                mbar_arrive_expect_tx(&full[stage], bytes * kC);
                const float* src = z + (it.b * kC) * P + px0;
                float* dst = stage_buf + size_t(stage) * kStageFloats;
#pragma unroll
                for (int c = 0; c < kC; ++c) {
                    if (g_use_hint) tma_load_1d_hint(dst + c * kTilePx, src + c * P, bytes, &full[stage], policy);
                    else tma_load_1d(dst + c * kTilePx, src + c * P, bytes, &full[stage]);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // Everything this kernel reads besides z (saved tensors, upstream scalars) and everything it writes may belong to
    // the kernel in front of it: wait for that kernel here (no-op without a programmatic dependency).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int stage = 0, cur = 0;
    uint32_t phase = 0;
    long long cur_b = it.valid() ? it.b : -1;
    SeedRegs ahead = seed_load(sa, tab, int(cur_b), tid);                 // the first matrix pays its latency once,
    const SeedCtx sc = seed_context(sa, P);                               // together with the upstream scalars
    seed_store(ahead, sc, tab, msh2, tid);
    ahead = seed_load(sa, tab, it.valid() ? int(it.next_sample()) : -1, tid);       // in flight while the first sample is processed
    named_bar_sync(1, kConsumers);
    for (; it.valid(); it.next()) {
        const long long b = it.b;
        const long long px0 = it.tin * kTilePx;
        const long long rem = P - px0;
        if (b != cur_b) {
            // the other buffer was last read two samples ago (a barrier has passed since): fill it from the registers
            seed_store(ahead, sc, tab, msh2 + (cur ^ 1) * 256, tid);
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
apply_generic_kernel(const float* __restrict__ z, const float* __restrict__ grelu, float* __restrict__ dz, long long P, SeedArgs sa) {
    __shared__ float msh[256];
    __shared__ IndexTables tab;
    const long long b = blockIdx.y;
    build_index_tables(tab, threadIdx.x, 256);
    __syncthreads();
    const SeedCtx sc = seed_context(sa, P);
    seed_matrix(sa, sc, tab, int(b), msh, threadIdx.x);
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
        if (grelu) s += x[i] <= 0.f ? 0.f : __ldg(grelu + (b * kC + i) * P + p);
        out[i * P] = s;
    }
}

bool tma_ok(const float* z, const float* dz, long long P) {
    return (P % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(dz) & 15u) == 0);
}

}  // namespace

// Tile schedule of the unfused apply kernel: blocks of this many tiles dealt round-robin to the CTAs, so that at
// any moment the 148 CTAs stream ADJACENT tiles of each channel row (0 = one contiguous range per CTA).
// Measured at 32x16x512x512: contiguous 184 us, round-robin 1: 167-171 us, 2: 168 us, 4: 180 us, 8: 181 us.
int g_apply_round_robin = 1;
// L2 evict-first policy on the TMA loads of z (read exactly once; keeps the dz write-back lines in L2 longer).
// Measured: step 281.5 -> 276 us, apply 174.9 -> 171.4 us, gram 104.2 -> 102.7 us.
int g_l2_evict_first = 1;

cudaError_t launch_apply(const float* z, const SeedArgs& seed, float* dz, int B, long long P, int sm_count, cudaStream_t stream,
                         const float* grelu) {
    if (grelu && apply_relu_tma_ok(z, grelu, dz, P)) return launch_apply_relu(z, grelu, seed, dz, B, P, sm_count, stream);
    if (!grelu && tma_ok(z, dz, P)) {
        auto kern = apply_tma_kernel;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemPlain));
        if (e != cudaSuccess) return e;
        const long long tps = (P + kTilePx - 1) / kTilePx;
        const long long T = tps * B;
        const long long G = T < sm_count ? T : sm_count;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned(G));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = kSmemPlain;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // prologue + z prefetch overlap the kernel in front
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
                return cudaLaunchKernelEx(&cfg, kern, z, dz, P, tps, T, g_apply_round_robin, g_l2_evict_first, seed);
    }
    apply_generic_kernel<<<dim3(unsigned((P + 255) / 256), unsigned(B)), 256, 0, stream>>>(z, grelu, dz, P, seed);
    return cudaGetLastError();
}

}  // namespace wtpse

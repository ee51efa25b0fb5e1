// Forward stage 1, variant 2: the per-sample 16x16 Gram with the 136 running sums SPLIT over two threads.
//
// The first kernel (whitening_gram.cu) keeps all 136 sums of a pixel quad in one thread: 255 registers, so an SM
// holds only 8 warps and the kernel ends up paced by instruction issue/latency, not by HBM (12 % of its warp
// samples wait for data).  Here every pixel quad is owned by TWO threads in different warps -- rows 0..4 of the
// upper triangle (70 sums, needs all 16 channels) and rows 5..15 (66 sums, needs channels 5..15) -- which brings
// the register count under 168 and the SM to 11 warps (10 consumers + producer), at the price of reading the staged
// tile 1.7x from shared memory (27 instead of 16 LDS.128 per quad; the pipe has the headroom).
// Tiles are 640 pixels (5 quad-warps x 32 lanes x 4 px), 40 KB per stage, 4 stages.
#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kQuadWarps = 5;                       // warps per half
constexpr int kConsumerWarps = 2 * kQuadWarps;      // 10
constexpr int kConsumers = kConsumerWarps * 32;     // 320
constexpr int kThreads = kConsumers + 32;           // + producer warp = 352
constexpr int kTilePx = kQuadWarps * 32 * 4;        // 640
constexpr int kStages = 4;
constexpr int kStageFloats = kC * kTilePx;          // 40 KB
constexpr int kHalfA = 70;                          // packed entries of rows 0..4
constexpr int kHalfB = kTri - kHalfA;               // 66, rows 5..15
constexpr int kPad = 72;                            // both halves are reduced as 72 = 8 * 9 values
constexpr size_t kSmemBytes =
    size_t(kStages) * kStageFloats * sizeof(float) + size_t(kConsumerWarps) * kPad * sizeof(float) + 2 * kStages * sizeof(uint64_t);

template <int HALF>
__device__ __forceinline__ void halve(float (&a)[kPad], int lane, int mask) {
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
        const float keep = up ? a[k + HALF] : a[k];
        const float send = up ? a[k] : a[k + HALF];
        a[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
}

// warp-level butterfly of 72 values, then cross-warp over the 5 warps of the same half; `out` points at the half's
// first packed entry (0 for rows 0..4, 70 for rows 5..15), `count` = 70 or 66
__device__ __forceinline__ void flush_half(float (&acc)[kPad], float* red, int warp, int lane, int tid, int half, int count,
                                           float* out) {
    halve<36>(acc, lane, 16);
    halve<18>(acc, lane, 8);
    halve<9>(acc, lane, 4);
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
    }
    if ((lane & 3) == 0) {
        const int base = ((lane >> 4) & 1) * 36 + ((lane >> 3) & 1) * 18 + ((lane >> 2) & 1) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) red[warp * kPad + base + k] = acc[k];
    }
    named_bar_sync(1, kConsumers);
    // threads 0..69 finish half A, threads 160..225 half B (warps 0-4 hold half A, warps 5-9 half B)
    const int e = tid - half * (kQuadWarps * 32);
    if (e >= 0 && e < count) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kQuadWarps; ++w) s += red[(half * kQuadWarps + w) * kPad + e];
        out[e] = s;
    }
    named_bar_sync(1, kConsumers);
}

template <int ROW0, int ROW1, int BASE>
__device__ __forceinline__ void accumulate_rows(float (&acc)[kPad], const float4 (&x)[kC]) {
#pragma unroll
    for (int i = ROW0; i < ROW1; ++i) {
#pragma unroll
        for (int j = i; j < kC; ++j) {
            float a = acc[tri_idx(i, j) - BASE];
            a = fmaf(x[i].x, x[j].x, a);
            a = fmaf(x[i].y, x[j].y, a);
            a = fmaf(x[i].z, x[j].z, a);
            a = fmaf(x[i].w, x[j].w, a);
            acc[tri_idx(i, j) - BASE] = a;
        }
    }
}

__global__ void __launch_bounds__(kThreads, 1)
gram_split_kernel(const float* __restrict__ z, float* __restrict__ partial, int* __restrict__ slot_count, long long P,
                  long long tiles_per_sample, long long T, int nslots, int group, int hint) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* red = stage_buf + size_t(kStages) * kStageFloats;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + kConsumerWarps * kPad);
    uint64_t* empty = full + kStages;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    const TileWalk walk(k, G, T, tiles_per_sample, group);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();
            int stage = 0;
            uint32_t phase = 0;
            for (long long b = walk.b_first; b <= walk.b_last; ++b) {
                long long t, tend;
                walk.segment(b, tiles_per_sample, t, tend);
                for (; t < tend; t += walk.g) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    const long long px0 = (t - b * tiles_per_sample) * kTilePx;
                    const long long rem = P - px0;
                    const uint32_t npx = rem < kTilePx ? uint32_t(rem) : uint32_t(kTilePx);
                    const uint32_t bytes = npx * 4u;
                    mbar_arrive_expect_tx(&full[stage], bytes * kC);
                    const float* src = z + (b * kC) * P + px0;
                    float* dst = stage_buf + size_t(stage) * kStageFloats;
#pragma unroll
                    for (int c = 0; c < kC; ++c) {
                        if (hint) tma_load_1d_hint(dst + c * kTilePx, src + c * P, bytes, &full[stage], policy);
                        else tma_load_1d(dst + c * kTilePx, src + c * P, bytes, &full[stage]);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    const int half = warp >= kQuadWarps ? 1 : 0;                  // warp-uniform
    const int quad = (warp - half * kQuadWarps) * 32 + lane;      // pixel quad inside the tile
    float acc[kPad];
#pragma unroll
    for (int e = 0; e < kPad; ++e) acc[e] = 0.f;

    int stage = 0;
    uint32_t phase = 0;
    const long long Gg = G / walk.g, grp = k / walk.g;
    for (long long b = walk.b_first; b <= walk.b_last; ++b) {
        long long t, tend;
        walk.segment(b, tiles_per_sample, t, tend);
        for (; t < tend; t += walk.g) {
            const long long px0 = (t - b * tiles_per_sample) * kTilePx;
            const long long rem = P - px0;
            mbar_wait(&full[stage], phase);
            if (4LL * quad < rem) {
                const float* src = stage_buf + size_t(stage) * kStageFloats + 4 * quad;
                float4 x[kC];
                if (half == 0) {
#pragma unroll
                    for (int c = 0; c < kC; ++c) x[c] = *reinterpret_cast<const float4*>(src + c * kTilePx);
                    accumulate_rows<0, 5, 0>(acc, x);
                } else {
#pragma unroll
                    for (int c = 5; c < kC; ++c) x[c] = *reinterpret_cast<const float4*>(src + c * kTilePx);
                    accumulate_rows<5, kC, kHalfA>(acc, x);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        const long long first_grp = part_owner(b * tiles_per_sample, T, Gg);
        const long long slot = (grp - first_grp) * walk.g + walk.r;
        float* out = partial + (b * nslots + slot) * kTri;
        flush_half(acc, red, warp, lane, tid, half, half ? kHalfB : kHalfA, out + (half ? kHalfA : 0));
        if (tid == 0 && walk.r == 0 && (b + 1) * tiles_per_sample <= walk.R1) slot_count[b] = int((grp - first_grp + 1) * walk.g);
#pragma unroll
        for (int e = 0; e < kPad; ++e) acc[e] = 0.f;
    }
}

}  // namespace

long long gram_split_tile_px() { return kTilePx; }

cudaError_t launch_gram_split(const float* z, float* partial, int* slot_count, long long P, const GramPlan& g,
                              cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(gram_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(g.G));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gram_split_kernel, z, partial, slot_count, P, g.tiles_per_sample, g.T, g.nslots, g.group,
                              g_l2_evict_first);
}

}  // namespace wtpse

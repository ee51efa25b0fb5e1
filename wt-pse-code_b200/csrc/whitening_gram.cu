// Forward pass, stage 1: per-sample 16x16 Gram  z_b z_b^T  (the one pass over z).
//
// Replaces the torch.bmm of algorithms.py:1283 / shape_networks.py:567.  HBM-bound: 64 B per pixel
// are read once; 136 FMAs per pixel (symmetric half) run on the FP32 pipes underneath.
//
// Design (B200, 148 SMs, one persistent CTA per SM):
//   * the B * ceil(P/896) pixel tiles are split into contiguous ranges, one per CTA;
//   * a producer warp streams each tile -- 16 channel rows of 896 pixels, 56 KB -- into a 3-stage
//     shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * 7 consumer warps read one float4 (4 pixels) per channel from shared memory (conflict-free
//     LDS.128) and keep the 136 running sums in registers;
//   * at a sample boundary the 136 sums are reduced across the warp with a halving butterfly
//     (153 shuffles instead of 680), across warps through shared memory, and written to a
//     per-(sample, slot) partial.  No float atomics: the slots are summed in a fixed order;
//   * the rest of the forward pass -- slot reduction, f_cor, instance terms, MMD, the backward's MMD seed -- runs
//     inside this kernel too, in whichever CTA arrives last (whitening_tail.cuh): the forward is ONE launch.
#include "common.cuh"
#include "kernels.h"
#include "whitening_tail.cuh"

namespace wtpse {

namespace {

// 7 consumer warps + 1 producer warp = 256 threads: the 136 accumulators + 64 staged inputs need 255 registers, so
// 256 threads is the most an SM holds (a 9th warp pushes ptxas to 168 registers and spills).  Folding the producer
// into a consumer warp (8 consumers, 1024-pixel tiles) was tried and is slower (110 vs 102 us): the issuing lane's
// waits on the ring stall its warp and couple the look-ahead to the consumers' pace.
constexpr int kConsumers = 224;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kThreads = kConsumers + 32;       // + producer warp
constexpr int kTilePx = kConsumers * 4;         // 896 pixels per tile
constexpr int kStages = 3;
constexpr int kStageFloats = kC * kTilePx;      // 56 KB per stage

constexpr size_t kSmemBytes =
    size_t(kStages) * kStageFloats * sizeof(float) + size_t(kConsumerWarps) * kTri * sizeof(float) + 2 * kStages * sizeof(uint64_t);

template <int HALF>
__device__ __forceinline__ void halve(float (&a)[kTri], int lane, int mask) {
    const bool up = (lane & mask) != 0;
#pragma unroll
    for (int k = 0; k < HALF; ++k) {
        const float keep = up ? a[k + HALF] : a[k];
        const float send = up ? a[k] : a[k + HALF];
        a[k] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
    }
}

// Reduce the 136 per-thread sums over the kWarps consumer warps and store them to `out[136]`.
template <int kWarps>
__device__ __forceinline__ void flush_gram_t(float (&acc)[kTri], float* red, int warp, int lane, int tid, float* out) {
    halve<68>(acc, lane, 16);
    halve<34>(acc, lane, 8);
    halve<17>(acc, lane, 4);
#pragma unroll
    for (int k = 0; k < 17; ++k) {
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2);
        acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1);
    }
    if ((lane & 3) == 0) {
        const int base = ((lane >> 4) & 1) * 68 + ((lane >> 3) & 1) * 34 + ((lane >> 2) & 1) * 17;
#pragma unroll
        for (int k = 0; k < 17; ++k) red[warp * kTri + base + k] = acc[k];
    }
    named_bar_sync(1, kWarps * 32);
    if (tid < kTri) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += red[w * kTri + tid];
        out[tid] = s;
    }
    named_bar_sync(1, kWarps * 32);
}

__device__ __forceinline__ void flush_gram(float (&acc)[kTri], float* red, int warp, int lane, int tid, float* out) {
    flush_gram_t<kConsumerWarps>(acc, red, warp, lane, tid, out);
}

__device__ __forceinline__ void gram_accumulate(float (&acc)[kTri], const float4 (&x)[kC]) {
#pragma unroll
    for (int i = 0; i < kC; ++i) {
#pragma unroll
        for (int j = i; j < kC; ++j) {
            float a = acc[tri_idx(i, j)];
            a = fmaf(x[i].x, x[j].x, a);
            a = fmaf(x[i].y, x[j].y, a);
            a = fmaf(x[i].z, x[j].z, a);
            a = fmaf(x[i].w, x[j].w, a);
            acc[tri_idx(i, j)] = a;
        }
    }
}

// ReLU of the DeepWT tail (algorithms.py:1105,1112: `F.relu(z_instance)` right after the embedding that feeds the loss).
// `x < 0 ? 0 : x` keeps NaN, like ATen's clamp_min.
__device__ __forceinline__ float4 relu4(const float4& v) {
    return make_float4(v.x < 0.f ? 0.f : v.x, v.y < 0.f ? 0.f : v.y, v.z < 0.f ? 0.f : v.z, v.w < 0.f ? 0.f : v.w);
}

// kRelu: the same pass also writes relu(z) (SURVEY 8(f).1 -- the activation that follows the embedding no longer
// re-reads z); the consumers store the 4 pixels x 16 channels they hold anyway, one coalesced 512 B row piece per
// warp and channel.  Default-policy stores: the next convolution reads relu(z) right away.
template <bool kRelu>
__global__ void __launch_bounds__(kThreads, 1)
gram_tma_kernel(const float* __restrict__ z, float* __restrict__ relu_out, float* __restrict__ partial,
                int* __restrict__ slot_count, long long P, long long tiles_per_sample, long long T, int nslots,
                int hint, int fused_tail, TailParams tp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* red = stage_buf + size_t(kStages) * kStageFloats;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + kConsumerWarps * kTri);
    uint64_t* empty = full + kStages;
    __shared__ IndexTables tab;
    __shared__ float wred[16];
    __shared__ int ticket_flag;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the dependents get resident early
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    const TileWalk walk(k, G, T, tiles_per_sample, 1);
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
    // launched with programmatic stream serialisation: everything above overlapped the previous kernel's tail; z
    // (and the workspace) may only be touched once that kernel has completed
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == kConsumerWarps) {
        // ---------------- producer: one lane issues the bulk copies ----------------
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();
            const bool g_use_hint = hint != 0;
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
                        if (g_use_hint) tma_load_1d_hint(dst + c * kTilePx, src + c * P, bytes, &full[stage], policy);
                        else tma_load_1d(dst + c * kTilePx, src + c * P, bytes, &full[stage]);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    TailClock clk;
    float acc[kTri];
#pragma unroll
    for (int e = 0; e < kTri; ++e) acc[e] = 0.f;

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
            if (4LL * tid < rem) {
                const float* src = stage_buf + size_t(stage) * kStageFloats + 4 * tid;
                float4 x[kC];
#pragma unroll
                for (int c = 0; c < kC; ++c) x[c] = *reinterpret_cast<const float4*>(src + c * kTilePx);
                if (kRelu) {
                    float* dst = relu_out + (b * kC) * P + px0 + 4 * tid;
#pragma unroll
                    for (int c = 0; c < kC; ++c) *reinterpret_cast<float4*>(dst + c * P) = relu4(x[c]);
                }
                gram_accumulate(acc, x);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        const long long first_grp = part_owner(b * tiles_per_sample, T, Gg);
        const long long slot = (grp - first_grp) * walk.g + walk.r;
        clk.mark(0);
        flush_gram(acc, red, warp, lane, tid, partial + (b * nslots + slot) * kTri);
        clk.mark(1);
        if (fused_tail) {
            // every CTA whose range touches sample b stores exactly one partial for it
            const int expected = int(part_owner((b + 1) * tiles_per_sample - 1, T, Gg) - first_grp + 1);
            tail_after_flush<kConsumers>(tp, int(b), expected, tab, wred, &ticket_flag, stage_buf, tid, 1, clk);
        } else if (tid == 0 && (b + 1) * tiles_per_sample <= walk.R1) {
            // legacy tail (separate kernels): the CTA that owns a sample's last tile knows how many slots it used
            slot_count[b] = int(grp - first_grp + 1);
        }
#pragma unroll
        for (int e = 0; e < kTri; ++e) acc[e] = 0.f;
    }
}

// Fallback for inputs the bulk copies cannot take (P % 4 != 0 or a base pointer that is not 16-byte
// aligned): same arithmetic, plain coalesced scalar loads, one pixel per thread per step.
// grid = (nslots, B); block = 256.
template <bool kRelu>
__global__ void __launch_bounds__(kConsumers)
gram_generic_kernel(const float* __restrict__ z, float* __restrict__ relu_out, float* __restrict__ partial,
                    int* __restrict__ slot_count, long long P, int nslots) {
    __shared__ float red[kConsumerWarps * kTri];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long b = blockIdx.y;
    const int slot = blockIdx.x;
    const long long chunk = (P + nslots - 1) / nslots;
    const long long p0 = slot * chunk;
    const long long p1 = (p0 + chunk < P) ? p0 + chunk : P;
    const float* zb = z + b * kC * P;

    float acc[kTri];
#pragma unroll
    for (int e = 0; e < kTri; ++e) acc[e] = 0.f;
    for (long long p = p0 + tid; p < p1; p += kConsumers) {
        float x[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) x[c] = __ldg(zb + c * P + p);
        if (kRelu) {
#pragma unroll
            for (int c = 0; c < kC; ++c) relu_out[(b * kC + c) * P + p] = x[c] < 0.f ? 0.f : x[c];
        }
#pragma unroll
        for (int i = 0; i < kC; ++i)
#pragma unroll
            for (int j = i; j < kC; ++j) acc[tri_idx(i, j)] = fmaf(x[i], x[j], acc[tri_idx(i, j)]);
    }
    flush_gram(acc, red, warp, lane, tid, partial + (b * nslots + slot) * kTri);
    if (slot == 0 && tid == 0) slot_count[b] = nslots;
}


// ------------------------------------------------------------------------------------------------
// Channels-last input: z is [B][P][16] (the memory of a channels-last B x 16 x H x W tensor, what a channels-last
// backbone produces).  A pixel's 16 channels are 64 contiguous bytes, so every thread reads its own pixel with four
// 128-bit loads (a warp covers 2 KB contiguous) and no shared-memory staging is needed; a 4-deep register ring keeps
// three pixels per thread in flight (148 SMs x 256 threads x 192 B = 7.3 MB).  Work items of `item_px` pixels of one
// sample are dealt round-robin to the persistent CTAs; each item ends with one cross-thread flush into its own slot,
// so the slot reduction that follows is the same fixed-order sum as for the NCHW kernels.
constexpr int kClThreads = 256;
constexpr int kClWarps = kClThreads / 32;

__device__ __forceinline__ void gram_accumulate1(float (&acc)[kTri], const float (&x)[kC]) {
#pragma unroll
    for (int i = 0; i < kC; ++i)
#pragma unroll
        for (int j = i; j < kC; ++j) acc[tri_idx(i, j)] = fmaf(x[i], x[j], acc[tri_idx(i, j)]);
}

template <bool kRelu>
__global__ void __launch_bounds__(kClThreads, 1)
gram_cl_kernel(const float* __restrict__ z, float* __restrict__ relu_out, float* __restrict__ partial, int* __restrict__ slot_count,
               long long P, int B, int nslots, long long item_px) {
    __shared__ float red[kClWarps * kTri];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const long long items = (long long)B * nslots;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const long long b = item / nslots;
        const int slot = int(item - b * nslots);
        const long long p0 = slot * item_px;
        const long long p1 = p0 + item_px < P ? p0 + item_px : P;
        const float* base = z + b * P * kC;
        float acc[kTri];
#pragma unroll
        for (int e = 0; e < kTri; ++e) acc[e] = 0.f;
        float4 r0[4], r1[4], r2[4], r3[4];
        auto ld = [&](float4 (&r)[4], long long p) {
            if (p < p1) {
                const float4* s = reinterpret_cast<const float4*>(base + p * kC);
#pragma unroll
                for (int q = 0; q < 4; ++q) r[q] = __ldg(s + q);
            }
        };
        auto use = [&](const float4 (&r)[4], long long p) {
            if (p < p1) {
                if (kRelu) {
                    float4* d = reinterpret_cast<float4*>(relu_out + (b * P + p) * kC);
#pragma unroll
                    for (int q = 0; q < 4; ++q) d[q] = relu4(r[q]);
                }
                const float x[kC] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w,
                                     r[2].x, r[2].y, r[2].z, r[2].w, r[3].x, r[3].y, r[3].z, r[3].w};
                gram_accumulate1(acc, x);
            }
        };
        long long p = p0 + tid;
        ld(r0, p);
        ld(r1, p + kClThreads);
        ld(r2, p + 2 * kClThreads);
        for (; p < p1; p += 4 * kClThreads) {
            ld(r3, p + 3 * kClThreads); use(r0, p);
            ld(r0, p + 4 * kClThreads); use(r1, p + kClThreads);
            ld(r1, p + 5 * kClThreads); use(r2, p + 2 * kClThreads);
            ld(r2, p + 6 * kClThreads); use(r3, p + 3 * kClThreads);
        }
        flush_gram_t<kClWarps>(acc, red, warp, lane, tid, partial + (b * nslots + slot) * kTri);
        if (tid == 0 && slot == 0) slot_count[b] = nslots;
    }
}
}  // namespace

int g_cl_tma = 1;        // channels-last kernels: tensor-map TMA pipelines (whitening_cl_tma.cu) or the per-thread kernels

GramPlan plan_gram(const float* z, int B, long long P, int sm_count, const float* relu_out) {
    GramPlan g{};
    g.tma = (P % 4 == 0) && ((reinterpret_cast<uintptr_t>(z) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(relu_out) & 15u) == 0);
    g.group = 1;
    if (g.tma) {
        g.tiles_per_sample = (P + kTilePx - 1) / kTilePx;
        g.T = g.tiles_per_sample * B;
        g.G = g.T < sm_count ? g.T : sm_count;
        int nslots = 1;
        for (int b = 0; b < B; ++b) {
            const long long first = part_owner((long long)b * g.tiles_per_sample, g.T, g.G);
            const long long last = part_owner((long long)(b + 1) * g.tiles_per_sample - 1, g.T, g.G);
            if (last - first + 1 > nslots) nslots = int(last - first + 1);
        }
        g.nslots = nslots;
    } else {
        long long want = (2LL * sm_count + B - 1) / B;           // ~2 CTAs per SM in total
        const long long max_useful = (P + 4 * kConsumers - 1) / (4 * kConsumers);
        if (want > max_useful) want = max_useful;
        if (want < 1) want = 1;
        g.nslots = int(want);
    }
    return g;
}

// The whole-batch phase of the in-kernel tail runs in the pipeline buffers of one CTA.
bool gram_tail_fits(int B, int n_per_domain, int n_domains) {
    long long m = n_domains > 1 ? (long long)n_per_domain * n_domains : 0;
    if (m > B) m = B;
    return tail_smem_bytes(B, int(m), n_domains) <= size_t(kStages) * kStageFloats * sizeof(float);
}

// Channels-last schedule: items of >= 8192 pixels, at most 256 slots per sample (gram_partial_floats covers that).
GramPlan plan_gram_cl(int B, long long P, int sm_count) {
    GramPlan g{};
    g.tma = false;
    long long item_px = 8192;
    if ((P + item_px - 1) / item_px > 256) item_px = (P + 255) / 256;
    g.item_px = item_px;
    g.nslots = int((P + item_px - 1) / item_px);
    const long long items = (long long)B * g.nslots;
    g.G = items < sm_count ? items : sm_count;
    g.group = 1;
    return g;
}

cudaError_t launch_gram_cl(const float* z, float* relu_out, float* partial, int* slot_count, int B, long long P, const GramPlan& g,
                           cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(g.G));
    cfg.blockDim = dim3(kClThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (relu_out) return cudaLaunchKernelEx(&cfg, gram_cl_kernel<true>, z, relu_out, partial, slot_count, P, B, g.nslots, g.item_px);
    return cudaLaunchKernelEx(&cfg, gram_cl_kernel<false>, z, relu_out, partial, slot_count, P, B, g.nslots, g.item_px);
}

size_t gram_partial_floats(int B, long long P, int sm_count) {
    // upper bound over all paths.  Persistent pipelines (contiguous ranges of `tile`-pixel tiles, NCHW: 896, channels-last
    // TMA: 448): a sample spans at most ceil(tiles per sample / smallest range) + 1 CTAs.  Generic kernel: <= 2 * sm_count
    // slots in total.  Per-thread channels-last kernel: items of >= 8192 pixels, <= 256 per sample.
    long long slots = 1;
    for (long long tile : {896LL, 448LL}) {
        const long long tps = (P + tile - 1) / tile;
        const long long T = tps * B;
        const long long G = T < sm_count ? T : sm_count;
        long long per_cta = T / G;
        if (per_cta < 1) per_cta = 1;
        const long long s = (tps + per_cta - 1) / per_cta + 1;
        if (s > slots) slots = s;
    }
    long long slots_gen = (2LL * sm_count + B - 1) / B;
    if (slots_gen > slots) slots = slots_gen;
    long long slots_cl = (P + 8191) / 8192;
    if (slots_cl > 256) slots_cl = 256;
    if (slots_cl > slots) slots = slots_cl;
    return size_t(B) * size_t(slots) * kTri;
}

namespace {
template <bool kRelu>
cudaError_t launch_gram_t(const float* z, float* relu_out, float* partial, int* slot_count, int B, long long P,
                          const GramPlan& g, cudaStream_t stream, const TailParams* tail) {
    if (g.tma) {
        cudaError_t e = cudaFuncSetAttribute(gram_tma_kernel<kRelu>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemBytes));
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
        TailParams tp{};
        if (tail) tp = *tail;
        return cudaLaunchKernelEx(&cfg, gram_tma_kernel<kRelu>, z, relu_out, partial, slot_count, (long long)P,
                                  g.tiles_per_sample, g.T, g.nslots, g_l2_evict_first, tail ? 1 : 0, tp);
    } else {
        gram_generic_kernel<kRelu><<<dim3(unsigned(g.nslots), unsigned(B)), kConsumers, 0, stream>>>(z, relu_out, partial,
                                                                                                   slot_count, P, g.nslots);
    }
    return cudaGetLastError();
}
}  // namespace

// relu_out != nullptr: fused ReLU write (plan_gram must have been made with the same relu_out, see its alignment rule).
// tail != nullptr (TMA plans only): the rest of the forward runs inside the kernel (whitening_tail.cuh).
cudaError_t launch_gram(const float* z, float* partial, int* slot_count, int B, long long P, const GramPlan& g,
                        cudaStream_t stream, float* relu_out, const TailParams* tail) {
    if (tail && !g.tma) return cudaErrorInvalidValue;
    if (relu_out) return launch_gram_t<true>(z, relu_out, partial, slot_count, B, P, g, stream, tail);
    return launch_gram_t<false>(z, nullptr, partial, slot_count, B, P, g, stream, tail);
}

}  // namespace wtpse

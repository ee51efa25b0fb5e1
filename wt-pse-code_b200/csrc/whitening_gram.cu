// Forward pass, stage 1: per-sample 16x16 Gram  z_b z_b^T  (the one pass over z).
//
// Replaces the torch.bmm of algorithms.py:1283 / shape_networks.py:567.  HBM-bound: 64 B per pixel
// are read once; 136 FMAs per pixel (symmetric half) run on the FP32 pipes underneath.
//
// Design (B200, 148 SMs, one persistent CTA per SM):
//   * the B * ceil(P/768) pixel tiles are split into contiguous ranges, one per CTA;
//   * a producer warp streams each tile -- 16 channel rows of 768 pixels, 48 KB -- into a 4-stage
//     shared-memory ring with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx);
//   * 6 consumer warps work as 3 PAIRS.  Both warps of a pair read the same 4 pixels per thread (conflict-free
//     LDS.128) and each keeps HALF of the 136 running sums, as 68 packed (even pixel, odd pixel) accumulators fed by
//     fma.rn.f32x2 (SASS FFMA2): 76 thread instructions per pixel instead of the 140 of one-thread-per-pixel-quad with
//     scalar FMAs.  The scalar version left the kernel issue-bound (2 warps per scheduler at 0.49 IPC each way,
//     tools/gram_lab.cu: 86.1 us alone, 0.952 of the HBM peak); this one runs at what the memory side delivers
//     (82.4 us, 0.995).  Packing over PIXELS needs no operand shuffling: a 128-bit shared load leaves
//     (px0, px1) and (px2, px3) in aligned register pairs.
//   * which warp keeps what: the upper triangle in 4x4 channel blocks A B C D.  Even warp: AA AB BB AC AD (reads all
//     16 channels), odd warp: CC CD DD BC BD (reads 12) -- 68 entries each;
//   * at a sample boundary the sums are folded (even + odd pixel), reduced across the warp with a halving butterfly,
//     across the three pairs through shared memory, and written to a per-(sample, slot) partial in packed
//     upper-triangle order.  No float atomics: the slots are summed in a fixed order;
//   * the rest of the forward pass -- slot reduction, f_cor, instance terms, MMD, the backward's MMD seed -- runs
//     inside this kernel too, in whichever CTA arrives last (whitening_tail.cuh): the forward is ONE launch.
#include "common.cuh"
#include "kernels.h"
#include "whitening_tail.cuh"

namespace wtpse {

namespace {

// 6 consumer warps + 1 producer warp: at most 2 warps per scheduler, so each thread may hold 255 registers (68 packed
// accumulators = 136, one 4-channel block of inputs per operand = 48 more).  A 9th warp would cap every thread at 168.
constexpr int kPairs = 3;
constexpr int kConsumerWarps = 2 * kPairs;
constexpr int kConsumers = kConsumerWarps * 32;
constexpr int kThreads = kConsumers + 32;       // + producer warp
constexpr int kPassPx = kPairs * 128;           // pixels one pass of the three pairs covers (4 per thread)
constexpr int kPasses = 2;
constexpr int kTilePx = kPassPx * kPasses;      // 768 pixels per tile
constexpr int kStages = 4;
constexpr int kStageFloats = kC * kTilePx;      // 48 KB per stage
constexpr int kHalf = kTri / 2;                 // 68 entries per warp of a pair

constexpr size_t kSmemBytes =
    size_t(kStages) * kStageFloats * sizeof(float) + size_t(kConsumerWarps) * kHalf * sizeof(float) + 2 * kStages * sizeof(uint64_t);

// generic / channels-last per-thread kernels: one thread per pixel (quad), all 136 sums per thread
constexpr int kGenThreads = 224;
constexpr int kGenWarps = kGenThreads / 32;

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

// ---- the pair kernel's accumulators ------------------------------------------------------------------------------
// index of entry (a, b) inside a 4x4 channel block: 10 entries (a <= b) for a diagonal block, 16 otherwise
__host__ __device__ constexpr int blk_idx(bool tri, int a, int b) { return tri ? a * 4 - (a * (a - 1)) / 2 + (b - a) : a * 4 + b; }

// packed upper-triangle index of accumulator e of warp type ty (0: AA AB BB AC AD, 1: CC CD DD BC BD)
struct PairEntries { unsigned char v[2][kHalf]; };
constexpr PairEntries make_pair_entries() {
    PairEntries t{};
    const int seq[2][5][2] = {{{0, 0}, {0, 1}, {1, 1}, {0, 2}, {0, 3}}, {{2, 2}, {2, 3}, {3, 3}, {1, 2}, {1, 3}}};
    for (int ty = 0; ty < 2; ++ty) {
        int e0 = 0;
        for (int k = 0; k < 5; ++k) {
            const int I = seq[ty][k][0], J = seq[ty][k][1];
            const bool tri = I == J;
            for (int a = 0; a < 4; ++a)
                for (int b = tri ? a : 0; b < 4; ++b) t.v[ty][e0 + blk_idx(tri, a, b)] = (unsigned char)tri_idx(4 * I + a, 4 * J + b);
            e0 += tri ? 10 : 16;
        }
    }
    return t;
}
__constant__ PairEntries kPairEntries = make_pair_entries();

__device__ __forceinline__ void load_block(float4 (&x)[4], const float* src, int c0) {
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = *reinterpret_cast<const float4*>(src + (c0 + c) * kTilePx);
}

// acc[E0 + (a, b)] += xi[a] * xj[b] over the thread's four pixels: two packed FMAs (px0, px1), (px2, px3)
template <bool TRI, int E0>
__device__ __forceinline__ void block_fma(float2 (&acc)[kHalf], const float4 (&xi)[4], const float4 (&xj)[4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = TRI ? a : 0; b < 4; ++b) {
            const int e = E0 + blk_idx(TRI, a, b);
            acc[e] = __ffma2_rn(make_float2(xi[a].x, xi[a].y), make_float2(xj[b].x, xj[b].y), acc[e]);
            acc[e] = __ffma2_rn(make_float2(xi[a].z, xi[a].w), make_float2(xj[b].z, xj[b].w), acc[e]);
        }
}

// ReLU of the DeepWT tail (algorithms.py:1105,1112: `F.relu(z_instance)` right after the embedding that feeds the loss).
// `x < 0 ? 0 : x` keeps NaN, like ATen's clamp_min.
__device__ __forceinline__ float4 relu4(const float4& v) {
    return make_float4(v.x < 0.f ? 0.f : v.x, v.y < 0.f ? 0.f : v.y, v.z < 0.f ? 0.f : v.z, v.w < 0.f ? 0.f : v.w);
}
__device__ __forceinline__ void store_relu_block(float* dst, const float4 (&x)[4], int c0, long long P) {
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<float4*>(dst + (c0 + c) * P) = relu4(x[c]);
}

// One pass of one warp over its four pixels per thread.  src: the thread's first pixel in channel row 0 of the stage;
// relu_dst: the same pixel in channel 0 of relu_out.  kRelu: the even warp also writes relu of channels 0..7, the odd
// warp of channels 8..15 (each writes what it has loaded anyway; 512 B coalesced per warp and channel).
template <int TYPE, bool kRelu>
__device__ __forceinline__ void pair_accumulate(float2 (&acc)[kHalf], const float* src, float* relu_dst, long long P) {
    float4 p[4], q[4], r[4];
    if constexpr (TYPE == 0) {
        load_block(p, src, 0);                   // A
        load_block(q, src, 4);                   // B
        if (kRelu) { store_relu_block(relu_dst, p, 0, P); store_relu_block(relu_dst, q, 4, P); }
        block_fma<true, 0>(acc, p, p);           // AA
        block_fma<false, 10>(acc, p, q);         // AB
        block_fma<true, 26>(acc, q, q);          // BB
        load_block(r, src, 8);                   // C
        block_fma<false, 36>(acc, p, r);         // AC
        load_block(q, src, 12);                  // D
        block_fma<false, 52>(acc, p, q);         // AD
    } else {
        load_block(p, src, 8);                   // C
        load_block(q, src, 12);                  // D
        if (kRelu) { store_relu_block(relu_dst, p, 8, P); store_relu_block(relu_dst, q, 12, P); }
        block_fma<true, 0>(acc, p, p);           // CC
        block_fma<false, 10>(acc, p, q);         // CD
        block_fma<true, 26>(acc, q, q);          // DD
        load_block(r, src, 4);                   // B
        block_fma<false, 36>(acc, r, p);         // BC
        block_fma<false, 52>(acc, r, q);         // BD
    }
}

// The warp's 68 sums over its 32 lanes -> red_w[68]: even + odd pixel, halving butterfly 68 -> 34 -> 17 (51 shuffles),
// then the 17 survivors over the remaining 8 lanes (51 shuffles).  Fixed order.
__device__ __forceinline__ void pair_warp_fold(const float2 (&acc)[kHalf], float* red_w, int lane) {
    float s[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) s[e] = acc[e].x + acc[e].y;
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int k = 0; k < 34; ++k) {
            const float keep = up ? s[k + 34] : s[k], send = up ? s[k] : s[k + 34];
            s[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int k = 0; k < 17; ++k) {
            const float keep = up ? s[k + 17] : s[k], send = up ? s[k] : s[k + 17];
            s[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
#pragma unroll
    for (int k = 0; k < 17; ++k) {
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 4);
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 2);
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 1);
    }
    if ((lane & 7) == 0) {
        const int b0 = ((lane >> 4) & 1) * 34 + ((lane >> 3) & 1) * 17;
#pragma unroll
        for (int k = 0; k < 17; ++k) red_w[b0 + k] = s[k];
    }
}

// kRelu: the same pass also writes relu(z) (SURVEY 8(f).1 -- the activation that follows the embedding no longer
// re-reads z).  Default-policy stores: the next convolution reads relu(z) right away.
template <bool kRelu>
__global__ void __launch_bounds__(kThreads, 1)
gram_tma_kernel(const float* __restrict__ z, float* __restrict__ relu_out, float* __restrict__ partial,
                int* __restrict__ slot_count, long long P, long long tiles_per_sample, long long T, int nslots,
                int hint, int fused_tail, TailParams tp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* red = stage_buf + size_t(kStages) * kStageFloats;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + kConsumerWarps * kHalf);
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
    const int type = warp & 1;                                  // which half of the triangle this warp keeps
    const int px_in_pass = ((warp >> 1) * 32 + lane) * 4;       // both warps of a pair: the same pixels
    // the packed upper-triangle entry thread tid < 136 sums across the pairs at every flush
    const int out_entry = tid < kTri ? kPairEntries.v[tid >= kHalf ? 1 : 0][tid >= kHalf ? tid - kHalf : tid] : 0;
    float2 acc[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) acc[e] = make_float2(0.f, 0.f);

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
            const float* src = stage_buf + size_t(stage) * kStageFloats + px_in_pass;
            float* rdst = kRelu ? relu_out + (b * kC) * P + px0 + px_in_pass : nullptr;
#pragma unroll
            for (int pass = 0; pass < kPasses; ++pass) {
                if (pass * kPassPx + px_in_pass < rem) {
                    if (type == 0) pair_accumulate<0, kRelu>(acc, src + pass * kPassPx, rdst + pass * kPassPx, P);
                    else pair_accumulate<1, kRelu>(acc, src + pass * kPassPx, rdst + pass * kPassPx, P);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        const long long first_grp = part_owner(b * tiles_per_sample, T, Gg);
        const long long slot = (grp - first_grp) * walk.g + walk.r;
        clk.mark(0);
        // flush: per-warp fold, then thread tid < 136 adds the three pairs' values of its entry in pair order
        pair_warp_fold(acc, red + warp * kHalf, lane);
        named_bar_sync(1, kConsumers);
        if (tid < kTri) {
            const int ty = tid >= kHalf ? 1 : 0, e = tid - ty * kHalf;
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < kPairs; ++g) s += red[(2 * g + ty) * kHalf + e];
            partial[(b * nslots + slot) * kTri + out_entry] = s;
        }
        named_bar_sync(1, kConsumers);
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
        for (int e = 0; e < kHalf; ++e) acc[e] = make_float2(0.f, 0.f);
    }
}

// Fallback for inputs the bulk copies cannot take (P % 4 != 0 or a base pointer that is not 16-byte
// aligned): same arithmetic, plain coalesced scalar loads, one pixel per thread per step.
// grid = (nslots, B); block = 256.
template <bool kRelu>
__global__ void __launch_bounds__(kGenThreads)
gram_generic_kernel(const float* __restrict__ z, float* __restrict__ relu_out, float* __restrict__ partial,
                    int* __restrict__ slot_count, long long P, int nslots) {
    __shared__ float red[kGenWarps * kTri];
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
    for (long long p = p0 + tid; p < p1; p += kGenThreads) {
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
    flush_gram_t<kGenWarps>(acc, red, warp, lane, tid, partial + (b * nslots + slot) * kTri);
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
        const long long max_useful = (P + 4 * kGenThreads - 1) / (4 * kGenThreads);
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
    // upper bound over all paths.  Persistent pipelines (contiguous ranges of `tile`-pixel tiles, NCHW: 768, channels-last
    // TMA: 448): a sample spans at most ceil(tiles per sample / smallest range) + 1 CTAs.  Generic kernel: <= 2 * sm_count
    // slots in total.  Per-thread channels-last kernel: items of >= 8192 pixels, <= 256 per sample.
    long long slots = 1;
    for (long long tile : {(long long)kTilePx, 448LL}) {
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
        gram_generic_kernel<kRelu><<<dim3(unsigned(g.nslots), unsigned(B)), kGenThreads, 0, stream>>>(z, relu_out, partial,
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

// Gram-kernel laboratory (GPU box): times variants of the forward streaming pass at the bench shape so that the
// shipped kernel's gap to the HBM roofline can be attributed (memory side vs issue side) before it is changed.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/gram_lab tools/gram_lab.cu
//   tools/gram_lab [B=32] [H=512] [iters=20]
//
// Variants (all: persistent CTAs, one per SM, contiguous tile ranges, 1-D TMA bulk copies into a shared-memory ring):
//   base      the shipped design: 7 consumer warps x 4 px/thread, 136 scalar-FMA accumulators, 3 stages of 896 px
//   base_mem  same pipeline, consumers only wait/arrive           -> what the memory side alone delivers
//   base_alu  same consumers, no loads / no waits                 -> what the issue side alone costs
//   pair      10 consumer warps in 5 pairs; both warps of a pair read the same 2 px/thread (LDS.64) and each keeps 68
//             of the 136 entries as packed (even px, odd px) accumulators, fma.rn.f32x2: 84 instead of 140 thread
//             instructions per pixel, 11 warps per SM instead of 8
//   pair_mem / pair_alu  as above
// Every full variant is checked against a double-precision reference Gram.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int kC = 16, kTri = 136;
__host__ __device__ constexpr int tri_idx(int i, int j) { return i * kC - (i * (i - 1)) / 2 + (j - i); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void tma_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t make_evict_first_policy() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__host__ __device__ __forceinline__ long long part_begin(long long k, long long T, long long G) { return (k * T) / G; }
__host__ __device__ __forceinline__ long long part_owner(long long t, long long T, long long G) { return ((t + 1) * G - 1) / T; }

enum Mode { FULL = 0, MEM = 1, ALU = 2 };

// ------------------------------------------------------------------------------------------------------------------
// producer shared by both designs
// ------------------------------------------------------------------------------------------------------------------
template <int kTilePx, int kStages>
__device__ __forceinline__ void producer(const float* z, float* stage_buf, uint64_t* full, uint64_t* empty, long long P, long long tps,
                                         long long t0, long long t1, long long step = 1) {
    const uint64_t policy = make_evict_first_policy();
    int stage = 0;
    uint32_t phase = 0;
    for (long long t = t0; t < t1; t += step) {
        const long long b = t / tps;
        mbar_wait(&empty[stage], phase ^ 1);
        const long long px0 = (t - b * tps) * kTilePx;
        const long long rem = P - px0;
        const uint32_t npx = rem < kTilePx ? uint32_t(rem) : uint32_t(kTilePx);
        const uint32_t bytes = npx * 4u;
        mbar_arrive_expect_tx(&full[stage], bytes * kC);
        const float* src = z + (b * kC) * P + px0;
        float* dst = stage_buf + size_t(stage) * kC * kTilePx;
#pragma unroll
        for (int c = 0; c < kC; ++c) tma_load_1d_hint(dst + c * kTilePx, src + c * P, bytes, &full[stage], policy);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// base design
// ------------------------------------------------------------------------------------------------------------------
namespace base {
constexpr int kConsumers = 224, kWarps = 7, kThreads = 256, kTilePx = 896, kStages = 3;
constexpr size_t kSmem = size_t(kStages) * kC * kTilePx * 4 + kWarps * kTri * 4 + 2 * kStages * 8;

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
__device__ __forceinline__ void flush(float (&acc)[kTri], float* red, int warp, int lane, int tid, float* out) {
    halve<68>(acc, lane, 16); halve<34>(acc, lane, 8); halve<17>(acc, lane, 4);
#pragma unroll
    for (int k = 0; k < 17; ++k) { acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 2); acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], 1); }
    if ((lane & 3) == 0) {
        const int b0 = ((lane >> 4) & 1) * 68 + ((lane >> 3) & 1) * 34 + ((lane >> 2) & 1) * 17;
#pragma unroll
        for (int k = 0; k < 17; ++k) red[warp * kTri + b0 + k] = acc[k];
    }
    named_bar_sync(1, kConsumers);
    if (tid < kTri) { float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += red[w * kTri + tid];
        out[tid] = s; }
    named_bar_sync(1, kConsumers);
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) kernel(const float* __restrict__ z, float* __restrict__ partial, long long P, long long tps,
                                                      long long T, int nslots, int rr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* red = stage_buf + size_t(kStages) * kC * kTilePx;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + kWarps * kTri);
    uint64_t* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    const long long step = (MODE == MEM && rr) ? G : 1;                       // MEM only: tiles k, k + G, ... (DRAM locality probe)
    const long long t0 = step > 1 ? k : part_begin(k, T, G), t1 = step > 1 ? T : part_begin(k + 1, T, G);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == kWarps) {
        if (lane == 0 && MODE != ALU) producer<kTilePx, kStages>(z, stage_buf, full, empty, P, tps, t0, t1, step);
        return;
    }
    float acc[kTri];
#pragma unroll
    for (int e = 0; e < kTri; ++e) acc[e] = 0.f;
    int stage = 0; uint32_t phase = 0;
    long long b_cur = t0 / tps;
    for (long long t = t0; t < t1; t += step) {
        const long long b = step > 1 ? b_cur : t / tps;
        if (b != b_cur) {
            const long long first = part_owner(b_cur * tps, T, G);
            flush(acc, red, warp, lane, tid, partial + (b_cur * nslots + (k - first)) * kTri);
#pragma unroll
            for (int e = 0; e < kTri; ++e) acc[e] = 0.f;
            b_cur = b;
        }
        const long long rem = P - (t - b * tps) * kTilePx;
        if (MODE != ALU) mbar_wait(&full[stage], phase);
        if (MODE != MEM && 4LL * tid < rem) {
            const float* src = stage_buf + size_t(stage) * kC * kTilePx + 4 * tid;
            float4 x[kC];
#pragma unroll
            for (int c = 0; c < kC; ++c) x[c] = *reinterpret_cast<const float4*>(src + c * kTilePx);
#pragma unroll
            for (int i = 0; i < kC; ++i)
#pragma unroll
                for (int j = i; j < kC; ++j) {
                    float a = acc[tri_idx(i, j)];
                    a = fmaf(x[i].x, x[j].x, a); a = fmaf(x[i].y, x[j].y, a); a = fmaf(x[i].z, x[j].z, a); a = fmaf(x[i].w, x[j].w, a);
                    acc[tri_idx(i, j)] = a;
                }
        }
        if (MODE != ALU) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[stage]); }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    const long long first = part_owner(b_cur * tps, T, G);
    flush(acc, red, warp, lane, tid, partial + (b_cur * nslots + (k - first)) * kTri);
}
}  // namespace base

// ------------------------------------------------------------------------------------------------------------------
// pair design
// ------------------------------------------------------------------------------------------------------------------
namespace pairk {
constexpr int kHalf = 68;
// The 16 x 16 upper triangle in 4 x 4 channel blocks A B C D:  AA AB AC AD / BB BC BD / CC CD / DD  (10 or 16 entries each).
// Warp type 0 keeps AA AB BB AC AD (68 entries, reads 16 channels), type 1 keeps CC CD DD BC BD (68 entries, reads 12).
// kEntry[type][e] = packed upper-triangle index of accumulator e of that type.
struct EntryTable { int v[2][kHalf]; };
__host__ __device__ constexpr int blk_idx(bool tri, int a, int b) { return tri ? a * 4 - (a * (a - 1)) / 2 + (b - a) : a * 4 + b; }
constexpr EntryTable make_entries() {
    EntryTable t{};
    const int seq[2][5][2] = {{{0, 0}, {0, 1}, {1, 1}, {0, 2}, {0, 3}}, {{2, 2}, {2, 3}, {3, 3}, {1, 2}, {1, 3}}};
    for (int ty = 0; ty < 2; ++ty) {
        int e0 = 0;
        for (int k = 0; k < 5; ++k) {
            const int I = seq[ty][k][0], J = seq[ty][k][1];
            const bool tri = I == J;
            for (int a = 0; a < 4; ++a)
                for (int b = tri ? a : 0; b < 4; ++b) t.v[ty][e0 + blk_idx(tri, a, b)] = tri_idx(4 * I + a, 4 * J + b);
            e0 += tri ? 10 : 16;
        }
    }
    return t;
}
__constant__ EntryTable kEntryDev = make_entries();

template <int kPairs, int kPasses, int kStagesT, int kPxT>
struct Cfg {
    static constexpr int kPx = kPxT;                       // pixels per thread and pass: 2 (LDS.64) or 4 (LDS.128)
    static constexpr int kCWarps = 2 * kPairs;
    static constexpr int kConsumers = kCWarps * 32;
    static constexpr int kThreads = kConsumers + 32;
    static constexpr int kPassPx = kPairs * 32 * kPx;
    static constexpr int kTilePx = kPassPx * kPasses;
    static constexpr int kStages = kStagesT;
    static constexpr size_t kSmem = size_t(kStages) * kC * kTilePx * 4 + size_t(kCWarps) * kHalf * 4 + 2 * kStages * 8;
};

template <int PX> struct Vec;
template <> struct Vec<2> { float2 v; };
template <> struct Vec<4> { float4 v; };

template <int PX>
__device__ __forceinline__ void load4(Vec<PX> (&x)[4], const float* src, int c0, int row_floats) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if constexpr (PX == 2) x[c].v = *reinterpret_cast<const float2*>(src + (c0 + c) * row_floats);
        else x[c].v = *reinterpret_cast<const float4*>(src + (c0 + c) * row_floats);
    }
}
template <int PX, bool TRI, int E0>
__device__ __forceinline__ void blk(float2 (&acc)[kHalf], const Vec<PX> (&xi)[4], const Vec<PX> (&xj)[4]) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = TRI ? a : 0; b < 4; ++b) {
            const int e = E0 + blk_idx(TRI, a, b);
            if constexpr (PX == 2) {
                acc[e] = __ffma2_rn(xi[a].v, xj[b].v, acc[e]);
            } else {
                acc[e] = __ffma2_rn(make_float2(xi[a].v.x, xi[a].v.y), make_float2(xj[b].v.x, xj[b].v.y), acc[e]);
                acc[e] = __ffma2_rn(make_float2(xi[a].v.z, xi[a].v.w), make_float2(xj[b].v.z, xj[b].v.w), acc[e]);
            }
        }
}
template <int PX, int TYPE>
__device__ __forceinline__ void accumulate(float2 (&acc)[kHalf], const float* src, int row_floats) {
    Vec<PX> p[4], q[4], r[4];
    if constexpr (TYPE == 0) {
        load4<PX>(p, src, 0, row_floats);            // A
        load4<PX>(q, src, 4, row_floats);            // B
        blk<PX, true, 0>(acc, p, p);                 // AA
        blk<PX, false, 10>(acc, p, q);               // AB
        blk<PX, true, 26>(acc, q, q);                // BB
        load4<PX>(r, src, 8, row_floats);            // C
        blk<PX, false, 36>(acc, p, r);               // AC
        load4<PX>(q, src, 12, row_floats);           // D
        blk<PX, false, 52>(acc, p, q);               // AD
    } else {
        load4<PX>(p, src, 8, row_floats);            // C
        load4<PX>(q, src, 12, row_floats);           // D
        blk<PX, true, 0>(acc, p, p);                 // CC
        blk<PX, false, 10>(acc, p, q);               // CD
        blk<PX, true, 26>(acc, q, q);                // DD
        load4<PX>(r, src, 4, row_floats);            // B
        blk<PX, false, 36>(acc, r, p);               // BC
        blk<PX, false, 52>(acc, r, q);               // BD
    }
}

// per-warp: 68 sums over the warp's 32 lanes (after folding even+odd), written to red_w[68]
__device__ __forceinline__ void warp_fold(float2 (&acc)[kHalf], float* red_w, int lane) {
    float s[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) s[e] = acc[e].x + acc[e].y;
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int k = 0; k < 34; ++k) { const float keep = up ? s[k + 34] : s[k], send = up ? s[k] : s[k + 34]; s[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int k = 0; k < 17; ++k) { const float keep = up ? s[k + 17] : s[k], send = up ? s[k] : s[k + 17]; s[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
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

template <class C, int MODE>
__global__ void __launch_bounds__(C::kThreads, 1) kernel(const float* __restrict__ z, float* __restrict__ partial, long long P, long long tps,
                                                         long long T, int nslots) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* red = stage_buf + size_t(C::kStages) * kC * C::kTilePx;
    uint64_t* full = reinterpret_cast<uint64_t*>(red + C::kCWarps * kHalf);
    uint64_t* empty = full + C::kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    const long long t0 = part_begin(k, T, G), t1 = part_begin(k + 1, T, G);
    if (tid == 0) {
        for (int s = 0; s < C::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], C::kCWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == C::kCWarps) {
        if (lane == 0 && MODE != ALU) producer<C::kTilePx, C::kStages>(z, stage_buf, full, empty, P, tps, t0, t1);
        return;
    }
    const int type = warp & 1, grp = warp >> 1;
    float2 acc[kHalf];
#pragma unroll
    for (int e = 0; e < kHalf; ++e) acc[e] = make_float2(0.f, 0.f);
    int stage = 0; uint32_t phase = 0;
    long long b_cur = t0 / tps;
    const int px_in_pass = (grp * 32 + lane) * C::kPx;

    auto flush = [&](long long b) {
        warp_fold(acc, red + warp * kHalf, lane);
        named_bar_sync(1, C::kConsumers);
        if (tid < kTri) {
            const int ty = tid >= kHalf ? 1 : 0, e = tid - ty * kHalf;
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < C::kCWarps / 2; ++g) s += red[(2 * g + ty) * kHalf + e];
            const long long first = part_owner(b * tps, T, G);
            partial[(b * nslots + (k - first)) * kTri + kEntryDev.v[ty][e]] = s;
        }
        named_bar_sync(1, C::kConsumers);
#pragma unroll
        for (int e = 0; e < kHalf; ++e) acc[e] = make_float2(0.f, 0.f);
    };

    for (long long t = t0; t < t1; ++t) {
        const long long b = t / tps;
        if (b != b_cur) { flush(b_cur); b_cur = b; }
        const long long rem = P - (t - b * tps) * C::kTilePx;
        if (MODE != ALU) mbar_wait(&full[stage], phase);
        if (MODE != MEM) {
            const float* src = stage_buf + size_t(stage) * kC * C::kTilePx + px_in_pass;
#pragma unroll
            for (int p = 0; p < C::kTilePx / C::kPassPx; ++p) {
                if (p * C::kPassPx + px_in_pass < rem) {
                    if (type == 0) accumulate<C::kPx, 0>(acc, src + p * C::kPassPx, C::kTilePx);
                    else accumulate<C::kPx, 1>(acc, src + p * C::kPassPx, C::kTilePx);
                }
            }
        }
        if (MODE != ALU) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[stage]); }
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
    }
    flush(b_cur);
}
}  // namespace pairk

// ------------------------------------------------------------------------------------------------------------------
// channels-last designs: z is [B][P][16]; a stage is 448 contiguous pixels (28 KB); warp w owns pixels 64w..64w+63 of a
// stage, lane l the pixels l and l + 32; packed accumulation over channel pairs (72 FFMA2 per pixel)
//   cl_tm   the stage arrives as 7 tensor-map boxes of 16 x 64 (cp.async.bulk.tensor.3d, 64-byte swizzle): the shipped path
//   cl_1d   the stage arrives as ONE 1-D bulk copy of 28 KB, no swizzle: the per-pixel LDS.128 are 4-way bank-conflicted
// ------------------------------------------------------------------------------------------------------------------
namespace cl {
constexpr int kConsumers = 224, kWarps = 7, kThreads = 256, kStagePx = 448, kStages = 6, kBoxPx = 64;
constexpr int kPartBytes = kStagePx * kC * 4;
constexpr size_t kSmem = 1024 + size_t(kStages) * kPartBytes + kWarps * kTri * 4 + 2 * kStages * 8;
constexpr int kPairAcc = 72;
__host__ __device__ constexpr int blk36(int I, int J) { return I * 8 - (I * (I - 1)) / 2 + (J - I); }

__device__ __forceinline__ void tma_load_box(void* dst, const CUtensorMap* tm, int px, int b, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::
                 "r"(smem_u32(dst)), "l"(tm), "r"(0), "r"(px), "r"(b), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint32_t swz(int p, int q) { return uint32_t(p) * 64u + (uint32_t(q ^ ((p >> 1) & 3)) << 4); }

__device__ __forceinline__ void accumulate(float2 (&acc2)[kPairAcc], const float4 (&r)[4]) {
    const float2 P[8] = {make_float2(r[0].x, r[0].y), make_float2(r[0].z, r[0].w), make_float2(r[1].x, r[1].y), make_float2(r[1].z, r[1].w),
                         make_float2(r[2].x, r[2].y), make_float2(r[2].z, r[2].w), make_float2(r[3].x, r[3].y), make_float2(r[3].z, r[3].w)};
    float2 S[8];
#pragma unroll
    for (int J = 0; J < 8; ++J) S[J] = make_float2(P[J].y, P[J].x);
#pragma unroll
    for (int I = 0; I < 8; ++I)
#pragma unroll
        for (int J = I; J < 8; ++J) {
            acc2[2 * blk36(I, J)] = __ffma2_rn(P[I], P[J], acc2[2 * blk36(I, J)]);
            acc2[2 * blk36(I, J) + 1] = __ffma2_rn(P[I], S[J], acc2[2 * blk36(I, J) + 1]);
        }
}
__device__ __forceinline__ void unpack(const float2 (&acc2)[kPairAcc], float (&acc)[kTri]) {
#pragma unroll
    for (int I = 0; I < 8; ++I)
#pragma unroll
        for (int J = I; J < 8; ++J) {
            const float2 D = acc2[2 * blk36(I, J)], A = acc2[2 * blk36(I, J) + 1];
            acc[tri_idx(2 * I, 2 * J)] = D.x; acc[tri_idx(2 * I + 1, 2 * J + 1)] = D.y; acc[tri_idx(2 * I, 2 * J + 1)] = A.x;
            if (I < J) acc[tri_idx(2 * I + 1, 2 * J)] = A.y;
        }
}

template <int MODE, bool kTensorMap>
__global__ void __launch_bounds__(kThreads, 1) kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ z, float* __restrict__ partial,
                                                      long long P, long long sps, long long T, int nslots) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* stage_buf = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    float* red = reinterpret_cast<float*>(stage_buf + size_t(kStages) * kPartBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + kWarps * kTri);
    uint64_t* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    const long long t0 = part_begin(k, T, G), t1 = part_begin(k + 1, T, G);
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], kWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == kWarps) {
        if (lane == 0 && MODE != ALU) {
            const uint64_t policy = make_evict_first_policy();
            int stage = 0; uint32_t phase = 0;
            for (long long t = t0; t < t1; ++t) {
                const long long b = t / sps;
                mbar_wait(&empty[stage], phase ^ 1);
                const long long px0 = (t - b * sps) * kStagePx;
                unsigned char* dst = stage_buf + size_t(stage) * kPartBytes;
                if (kTensorMap) {
                    mbar_arrive_expect_tx(&full[stage], kPartBytes);
#pragma unroll
                    for (int q = 0; q < kWarps; ++q) tma_load_box(dst + q * kBoxPx * kC * 4, &tm, int(px0) + q * kBoxPx, int(b), &full[stage], policy);
                } else {
                    const long long rem = P - px0;
                    const uint32_t bytes = uint32_t(rem < kStagePx ? rem : kStagePx) * kC * 4u;
                    mbar_arrive_expect_tx(&full[stage], bytes);
                    tma_load_1d_hint(dst, z + (b * P + px0) * kC, bytes, &full[stage], policy);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }
    float2 acc2[kPairAcc];
#pragma unroll
    for (int e = 0; e < kPairAcc; ++e) acc2[e] = make_float2(0.f, 0.f);
    int stage = 0; uint32_t phase = 0;
    long long b_cur = t0 / sps;
    auto flush = [&](long long b) {
        float acc[kTri];
        unpack(acc2, acc);
        const long long first = part_owner(b * sps, T, G);
        base::flush(acc, red, warp, lane, tid, partial + (b * nslots + (k - first)) * kTri);
#pragma unroll
        for (int e = 0; e < kPairAcc; ++e) acc2[e] = make_float2(0.f, 0.f);
    };
    for (long long t = t0; t < t1; ++t) {
        const long long b = t / sps;
        if (b != b_cur) { flush(b_cur); b_cur = b; }
        const long long rem = P - (t - b * sps) * kStagePx;
        if (MODE != ALU) mbar_wait(&full[stage], phase);
        if (MODE != MEM) {
            const unsigned char* sb = stage_buf + size_t(stage) * kPartBytes;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int p = warp * kBoxPx + j * 32 + lane;
                if (kTensorMap || p < rem) {
                    float4 r[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) r[q] = *reinterpret_cast<const float4*>(sb + (kTensorMap ? swz(p, q) : uint32_t(p) * 64u + q * 16u));
                    accumulate(acc2, r);
                }
            }
        }
        if (MODE != ALU) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[stage]); }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    flush(b_cur);
}
}  // namespace cl

__global__ void ref_gram_cl(const float* z, double* out, long long P) {   // grid (136, B), block 256; z [B][P][16]
    const int e = blockIdx.x; const long long b = blockIdx.y;
    int i = 0, rem = e; while (rem >= kC - i) { rem -= kC - i; ++i; } const int j = i + rem;
    const float* zb = z + b * P * kC;
    double s = 0; for (long long p = threadIdx.x; p < P; p += blockDim.x) s += double(zb[p * kC + i]) * double(zb[p * kC + j]);
    __shared__ double sh[256]; sh[threadIdx.x] = s; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[b * kTri + e] = sh[0];
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void ref_gram(const float* z, double* out, long long P) {   // grid (136, B), block 256
    const int e = blockIdx.x; const long long b = blockIdx.y;
    int i = 0, rem = e; while (rem >= kC - i) { rem -= kC - i; ++i; } const int j = i + rem;
    const float* zi = z + (b * kC + i) * P; const float* zj = z + (b * kC + j) * P;
    double s = 0; for (long long p = threadIdx.x; p < P; p += blockDim.x) s += double(zi[p]) * double(zj[p]);
    __shared__ double sh[256]; sh[threadIdx.x] = s; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) out[b * kTri + e] = sh[0];
}
__global__ void fill(float* z, long long n, unsigned seed) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned h = unsigned(i) * 2654435761u ^ seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        z[i] = (float(h & 0xffffff) / 8388608.f - 1.f) * 0.5f + 0.1f;
    }
}

struct Ctx { float* z[2]; float* partial; double* ref; int B; long long P; int sms; int iters; size_t partial_floats; };

template <class F>
static void run(const char* name, Ctx& c, int tile_px, bool check, F launch) {
    const long long tps = (c.P + tile_px - 1) / tile_px, T = tps * c.B;
    const long long G = T < c.sms ? T : c.sms;
    int nslots = 1;
    for (int b = 0; b < c.B; ++b) { const long long f = part_owner((long long)b * tps, T, G), l = part_owner((long long)(b + 1) * tps - 1, T, G); if (l - f + 1 > nslots) nslots = int(l - f + 1); }
    if (size_t(c.B) * nslots * kTri > c.partial_floats) { printf("%s: partial too small\n", name); return; }
    CK(cudaMemset(c.partial, 0, c.partial_floats * 4));
    for (int w = 0; w < 3; ++w) launch(c.z[w & 1], c.partial, c.P, tps, T, nslots, G);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int it = 0; it < c.iters; ++it) launch(c.z[it & 1], c.partial, c.P, tps, T, nslots, G);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / c.iters, gbs = double(c.B) * kC * c.P * 4 / (us * 1e-6) / 1e9;
    double err = -1;
    if (check) {
        CK(cudaMemset(c.partial, 0, c.partial_floats * 4));
        launch(c.z[0], c.partial, c.P, tps, T, nslots, G);
        CK(cudaDeviceSynchronize());
        std::vector<float> hp(size_t(c.B) * nslots * kTri); std::vector<double> hr(size_t(c.B) * kTri);
        CK(cudaMemcpy(hp.data(), c.partial, hp.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(hr.data(), c.ref, hr.size() * 8, cudaMemcpyDeviceToHost));
        err = 0; double mx = 0;
        for (int b = 0; b < c.B; ++b) for (int e = 0; e < kTri; ++e) {
            double s = 0; for (int sl = 0; sl < nslots; ++sl) s += hp[(size_t(b) * nslots + sl) * kTri + e];
            err = fmax(err, fabs(s - hr[size_t(b) * kTri + e])); mx = fmax(mx, fabs(hr[size_t(b) * kTri + e]));
        }
        err /= mx;
    }
    printf("%-22s tile %4d px  G %3lld  nslots %d  %8.2f us  %7.1f GB/s  frac(6548.2) %.3f  rel.err %s%.2e\n", name, tile_px, G, nslots, us, gbs,
           gbs / 6548.2, check ? "" : "(n/a) ", err);
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

template <int MODE, int RR = 0>
static void launch_base(const float* z, float* partial, long long P, long long tps, long long T, int nslots, long long G) {
    CK(cudaFuncSetAttribute(base::kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(base::kSmem)));
    base::kernel<MODE><<<unsigned(G), base::kThreads, base::kSmem>>>(z, partial, P, tps, T, nslots, RR);
}
template <class C, int MODE>
static void launch_pair(const float* z, float* partial, long long P, long long tps, long long T, int nslots, long long G) {
    CK(cudaFuncSetAttribute(pairk::kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(C::kSmem)));
    pairk::kernel<C, MODE><<<unsigned(G), C::kThreads, C::kSmem>>>(z, partial, P, tps, T, nslots);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap g_tm[2];
static const float* g_tm_base[2];
static void make_maps(Ctx& c) {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
    for (int i = 0; i < 2; ++i) {
        const cuuint64_t dims[3] = {16, cuuint64_t(c.P), cuuint64_t(c.B)};
        const cuuint64_t strides[2] = {64, cuuint64_t(c.P) * 64};
        const cuuint32_t box[3] = {16, 64, 1}, estr[3] = {1, 1, 1};
        if (fn(&g_tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, c.z[i], dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("tensor map failed\n"); exit(1); }
        g_tm_base[i] = c.z[i];
    }
}
template <int MODE, bool TM>
static void launch_cl(const float* z, float* partial, long long P, long long sps, long long T, int nslots, long long G) {
    CK(cudaFuncSetAttribute(cl::kernel<MODE, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(cl::kSmem)));
    cl::kernel<MODE, TM><<<unsigned(G), cl::kThreads, cl::kSmem>>>(g_tm[z == g_tm_base[0] ? 0 : 1], z, partial, P, sps, T, nslots);
}

int main(int argc, char** argv) {
    Ctx c{};
    c.B = argc > 1 ? atoi(argv[1]) : 32;
    const int H = argc > 2 ? atoi(argv[2]) : 512;
    c.iters = argc > 3 ? atoi(argv[3]) : 20;
    c.P = (long long)H * H;
    CK(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, 0));
    const long long n = (long long)c.B * kC * c.P;
    for (int i = 0; i < 2; ++i) { CK(cudaMalloc(&c.z[i], n * 4)); fill<<<1184, 256>>>(c.z[i], n, 17u + i); }
    c.partial_floats = size_t(c.B) * 16 * kTri;
    CK(cudaMalloc(&c.partial, c.partial_floats * 4));
    CK(cudaMalloc(&c.ref, size_t(c.B) * kTri * 8));
    ref_gram<<<dim3(kTri, c.B), 256>>>(c.z[0], c.ref, c.P);
    CK(cudaDeviceSynchronize());
    printf("B=%d H=%d P=%lld  SMs=%d  iters=%d  (%.1f MB per launch, %.1f us at 6548.2 GB/s)\n", c.B, H, c.P, c.sms, c.iters, n * 4 / 1e6, n * 4 / 6548.2e3);

    run("base", c, base::kTilePx, true, launch_base<FULL>);
    run("base_mem", c, base::kTilePx, false, launch_base<MEM>);
    run("base_mem_roundrobin", c, base::kTilePx, false, launch_base<MEM, 1>);
    run("base_alu", c, base::kTilePx, false, launch_base<ALU>);
    {
        using C = pairk::Cfg<5, 2, 4, 2>;   // 10 + 1 warps (168 regs), 2 px/thread, 640-px tiles (40 KB), 4 stages
        run("p2_5x2s4", c, C::kTilePx, true, launch_pair<C, FULL>);
        run("p2_5x2s4_mem", c, C::kTilePx, false, launch_pair<C, MEM>);
        run("p2_5x2s4_alu", c, C::kTilePx, false, launch_pair<C, ALU>);
    }
    { using C = pairk::Cfg<5, 3, 3, 2>; run("p2_5x3s3", c, C::kTilePx, true, launch_pair<C, FULL>); }   // 960-px tiles (60 KB), 3 stages
    { using C = pairk::Cfg<5, 1, 8, 2>; run("p2_5x1s8", c, C::kTilePx, true, launch_pair<C, FULL>); }   // 320-px tiles (20 KB), 8 stages
    {
        using C = pairk::Cfg<3, 2, 4, 4>;   // 6 + 1 warps (255 regs), 4 px/thread, 768-px tiles (48 KB), 4 stages
        run("p4_3x2s4", c, C::kTilePx, true, launch_pair<C, FULL>);
        run("p4_3x2s4_mem", c, C::kTilePx, false, launch_pair<C, MEM>);
        run("p4_3x2s4_alu", c, C::kTilePx, false, launch_pair<C, ALU>);
    }
    { using C = pairk::Cfg<3, 3, 3, 4>; run("p4_3x3s3", c, C::kTilePx, true, launch_pair<C, FULL>); }   // 1152-px tiles (72 KB), 3 stages
    { using C = pairk::Cfg<3, 1, 8, 4>; run("p4_3x1s8", c, C::kTilePx, true, launch_pair<C, FULL>); }   // 384-px tiles (24 KB), 8 stages
    { using C = pairk::Cfg<3, 2, 4, 2>; run("p2_3x2s4", c, C::kTilePx, true, launch_pair<C, FULL>); }   // 6 + 1 warps, 2 px/thread, 384-px tiles
    // channels-last: the same buffers read as [B][P][16]
    make_maps(c);
    ref_gram_cl<<<dim3(kTri, c.B), 256>>>(c.z[0], c.ref, c.P);
    CK(cudaDeviceSynchronize());
    run("cl_tm (shipped)", c, cl::kStagePx, true, launch_cl<FULL, true>);
    run("cl_tm_mem", c, cl::kStagePx, false, launch_cl<MEM, true>);
    run("cl_tm_alu", c, cl::kStagePx, false, launch_cl<ALU, true>);
    run("cl_1d", c, cl::kStagePx, true, launch_cl<FULL, false>);
    run("cl_1d_mem", c, cl::kStagePx, false, launch_cl<MEM, false>);
    run("cl_1d_alu", c, cl::kStagePx, false, launch_cl<ALU, false>);
    return 0;
}

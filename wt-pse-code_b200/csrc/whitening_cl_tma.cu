// Channels-last loss kernels as tensor-map TMA pipelines:  z, relu_out, grelu, dz are [B][P][16] fp32 -- the memory of
// channels-last B x 16 x H x W tensors, what the channels-last backbone of the train step hands over.
//
//   gram_cl_tma_kernel<relu>    forward:  per-sample Gram (+ relu(z) written by the same pass) + the in-kernel tail
//   apply_cl_tma_kernel<relu>   backward: dz = M_b z (+ [z > 0] * grelu)
//
// Why TMA here.  A pixel's 16 channels are 64 contiguous bytes.  The per-thread kernels these replace
// (whitening_gram.cu: gram_cl_kernel, whitening_apply_cl.cu) read them with one 128-bit load per lane at a 64-byte lane
// stride: 32 sectors per request, half of each used -- twice the useful bytes through L1, the LSU data pipe 71-75 % busy
// (profiles/r1_ncu_fusion_channels_last_summary.txt): 0.61-0.75 of the HBM roofline.  Here the tensor is described to
// the TMA unit as a 3-D tensor {16 channels, P pixels, B samples}; a producer lane moves boxes of 16 x 64 pixels
// (4 KB) global -> shared with cp.async.bulk.tensor.3d (SASS UTMALDG), bypassing the LSU, with the 64-byte swizzle so
// that the consumers' per-pixel LDS.128 are bank-conflict free; results leave the same way (in place in the stage,
// cp.async.bulk.tensor.3d shared -> global, SASS UTMASTG).  Rows past the end of a sample are zero-filled on load and
// clipped on store by the TMA unit, so there is no tail predication anywhere: a zero pixel adds nothing to a Gram.
//
// Pipeline: persistent CTAs (one per SM), one contiguous range of 448-pixel stages per CTA (a stage is one contiguous
// 28 KB piece of memory: 148 sequential streams), producer warp + 7 consumer warps (256 threads: 255 registers each),
// full/empty mbarriers, 6 stages (3 for the fused backward, whose stage carries z and grelu).  A consumer warp owns 64
// consecutive pixels of a stage -- one TMA box -- lane l the pixels l and l + 32; its results go back in place and lane
// 0 stores the warp's box.  A stage is handed back to the producer one stage late (cp.async.bulk.wait_group.read 1:
// the store before the newest has finished reading shared memory), so stores never stall the consumers.
//
// Arithmetic: per pixel exactly that of the per-thread kernels (136 / 256 FMAs in the same order, the masked ReLU
// gradient added last), so `relu_out` and `dz` are bit-identical to theirs; the Gram's summation order over pixels
// differs (results agree to fp32 rounding, tests/test_gpu_fusion.py).  Forward tail and backward matrices:
// whitening_tail.cuh / whitening_matrix.cuh, shared with the NCHW kernels.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "whitening_matrix.cuh"
#include "whitening_tail.cuh"

namespace wtpse {

namespace {

// 7 consumer warps + the producer warp = 256 threads: ptxas sizes registers for a multiple of 128 threads, so 8 + 1 warps
// would cap every thread at 168 registers (the Gram's 136 accumulators then spill)
constexpr int kConsumers = 224;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kThreads = kConsumers + 32;            // + producer warp
constexpr int kBoxPx = 64;                           // one box = 16 channels x 64 pixels = 4 KB = one warp's share of a stage
constexpr int kBoxBytes = kBoxPx * kC * 4;
constexpr int kStagePx = kConsumerWarps * kBoxPx;    // 448
constexpr int kPartBytes = kStagePx * kC * 4;        // 28 KB: one tensor's share of a stage
constexpr int kPx = kBoxPx / 32;                     // pixels per lane and stage

// ---- tensor-map TMA primitives -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_box(void* smem_dst, const CUtensorMap* tm, int px, int b, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::
            "r"(smem_u32(smem_dst)), "l"(tm), "r"(0), "r"(px), "r"(b), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tm, int px, int b, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(0), "r"(px),
                 "r"(b), "r"(smem_u32(smem_src))
                 : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// CU_TENSOR_MAP_SWIZZLE_64B: inside shared memory the 16-byte chunk index (address bits 4-5) is XORed with address bits
// 7-8.  Rows are 64 bytes and every box starts on a 1024-byte boundary, so for pixel row p of a stage those are bits
// 1-2 of p: chunk q of pixel p lives at p * 64 + ((q ^ ((p >> 1) & 3)) << 4).  Eight consecutive lanes (one LDS.128
// phase) then touch all 32 banks exactly once.
__device__ __forceinline__ uint32_t swz(int p, int q) { return uint32_t(p) * 64u + (uint32_t(q ^ ((p >> 1) & 3)) << 4); }

// First 1024-byte boundary at or after p (the swizzled boxes need it).  Plain pointer arithmetic on the shared-memory
// pointer: rounding through an integer makes the compiler lose the address space and every access to the stage buffers,
// the matrices and the tail's working set becomes a generic LD / ST (the address-divergence unit).
__device__ __forceinline__ unsigned char* align1024(unsigned char* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }

__device__ __forceinline__ float4 lds4(const unsigned char* base, uint32_t off) { return *reinterpret_cast<const float4*>(base + off); }
__device__ __forceinline__ void sts4(unsigned char* base, uint32_t off, const float4& v) { *reinterpret_cast<float4*>(base + off) = v; }

__device__ __forceinline__ float4 relu4(const float4& v) {       // `x < 0 ? 0 : x` keeps NaN, like ATen's clamp_min
    return make_float4(v.x < 0.f ? 0.f : v.x, v.y < 0.f ? 0.f : v.y, v.z < 0.f ? 0.f : v.z, v.w < 0.f ? 0.f : v.w);
}

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

// 136 per-thread sums -> out[136]: halving butterfly inside the warp (153 shuffles), fixed-order sum over the warps
__device__ __forceinline__ void flush_gram(float (&acc)[kTri], float* red, int warp, int lane, int tid, float* out) {
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
    named_bar_sync(1, kConsumers);
    if (tid < kTri) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) s += red[w * kTri + tid];
        out[tid] = s;
    }
    named_bar_sync(1, kConsumers);
}

// ---- packed accumulation (fma.rn.f32x2, SASS FFMA2) ----------------------------------------------------------------------
// A pixel's 16 channels arrive as four float4: eight natural register pairs P_I = (x_2I, x_2I+1).  For a 2x2 block of
// Gram entries (channel pairs I <= J) two packed FMAs cover all four products:
//     D[I][J] += P_I * P_J        = (x_2I x_2J   , x_2I+1 x_2J+1)      "diagonal" of the block
//     A[I][J] += P_I * swap(P_J)  = (x_2I x_2J+1 , x_2I+1 x_2J  )      "anti-diagonal"
// so a pixel costs 72 FFMA2 + 16 moves (the eight swapped pairs) instead of 136 FFMA -- the kernel was bound by
// instruction issue, not by HBM.  On the diagonal blocks (I == J) A's two lanes hold the same product; eight of the
// 144 lanes are redundant.  Every lane is an IEEE fma of the same operands in the same pixel order as the scalar
// loop it replaces: the sums are bit-identical to it.
constexpr int kPairAcc = 72;                         // 36 blocks (I <= J) x {D, A}
__host__ __device__ constexpr int blk36(int I, int J) { return I * 8 - (I * (I - 1)) / 2 + (J - I); }

__device__ __forceinline__ void gram_accumulate_pairs(float2 (&acc2)[kPairAcc], const float4 (&r)[4]) {
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

// packed accumulators -> the 136 upper-triangle sums in packed (tri_idx) order
__device__ __forceinline__ void unpack_pairs(const float2 (&acc2)[kPairAcc], float (&acc)[kTri]) {
#pragma unroll
    for (int I = 0; I < 8; ++I)
#pragma unroll
        for (int J = I; J < 8; ++J) {
            const float2 D = acc2[2 * blk36(I, J)], A = acc2[2 * blk36(I, J) + 1];
            acc[tri_idx(2 * I, 2 * J)] = D.x;
            acc[tri_idx(2 * I + 1, 2 * J + 1)] = D.y;
            acc[tri_idx(2 * I, 2 * J + 1)] = A.x;
            if (I < J) acc[tri_idx(2 * I + 1, 2 * J)] = A.y;
        }
}

// ------------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kGramStages = 6;
constexpr size_t kGramSmem = 1024 /* alignment slack */ + size_t(kGramStages) * kPartBytes + size_t(kConsumerWarps) * kTri * sizeof(float) +
                             2 * kGramStages * sizeof(uint64_t);

template <bool kRelu>
__global__ void __launch_bounds__(kThreads, 1)
gram_cl_tma_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_relu, float* __restrict__ partial,
                   int* __restrict__ slot_count, long long stages_per_sample, long long T, int nslots, int fused_tail, TailParams tp) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* stage_buf = align1024(smem_dyn);
    float* red = reinterpret_cast<float*>(stage_buf + size_t(kGramStages) * kPartBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(red + kConsumerWarps * kTri);
    uint64_t* empty = full + kGramStages;
    __shared__ IndexTables tab;
    __shared__ float wred[16];
    __shared__ int ticket_flag;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    const TileWalk walk(k, G, T, stages_per_sample, 1);
    build_index_tables(tab, tid, kThreads);
    if (tid == 0) {
        prefetch_tensormap(&tm_z);
        if (kRelu) prefetch_tensormap(&tm_relu);
#pragma unroll
        for (int s = 0; s < kGramStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();
    // z (and the workspace) may belong to the kernel in front of us (programmatic dependent launch)
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == kConsumerWarps) {
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();          // z is read exactly once
            int stage = 0;
            uint32_t phase = 0;
            for (long long b = walk.b_first; b <= walk.b_last; ++b) {
                long long t, tend;
                walk.segment(b, stages_per_sample, t, tend);
                for (; t < tend; ++t) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    const int px0 = int((t - b * stages_per_sample) * kStagePx);
                    mbar_arrive_expect_tx(&full[stage], kPartBytes);     // rows past the sample's end are zero-filled: full boxes
                    unsigned char* dst = stage_buf + size_t(stage) * kPartBytes;
#pragma unroll
                    for (int q = 0; q < kConsumerWarps; ++q)
                        tma_load_box(dst + q * kBoxBytes, &tm_z, px0 + q * kBoxPx, int(b), &full[stage], policy);
                    if (++stage == kGramStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    TailClock clk;
    float2 acc2[kPairAcc];
#pragma unroll
    for (int e = 0; e < kPairAcc; ++e) acc2[e] = make_float2(0.f, 0.f);
    int stage = 0, prev_stage = -1;
    uint32_t phase = 0;
    for (long long b = walk.b_first; b <= walk.b_last; ++b) {
        long long t, tend;
        walk.segment(b, stages_per_sample, t, tend);
        for (; t < tend; ++t) {
            mbar_wait(&full[stage], phase);
            unsigned char* sb = stage_buf + size_t(stage) * kPartBytes;
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const int p = warp * kBoxPx + j * 32 + lane;
                float4 r[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) r[q] = lds4(sb, swz(p, q));
                if (kRelu) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts4(sb, swz(p, q), relu4(r[q]));
                }
                gram_accumulate_pairs(acc2, r);
            }
            if (kRelu) {
                // the warp's 64 pixels of relu(z) leave from the stage itself; the stage goes back to the producer one
                // stage late, when this store has finished reading it
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    const int px0 = int((t - b * stages_per_sample) * kStagePx) + warp * kBoxPx;
                    tma_store_box(&tm_relu, px0, int(b), sb + size_t(warp) * kBoxPx * kC * 4);
                    tma_store_commit();
                    tma_store_wait_read<1>();
                    if (prev_stage >= 0) mbar_arrive(&empty[prev_stage]);
                }
                prev_stage = stage;
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);
            }
            if (++stage == kGramStages) { stage = 0; phase ^= 1; }
        }
        const long long first_cta = part_owner(b * stages_per_sample, T, G);
        const long long slot = k - first_cta;
        clk.mark(0);
        {
            float acc[kTri];
            unpack_pairs(acc2, acc);
            flush_gram(acc, red, warp, lane, tid, partial + (b * nslots + slot) * kTri);
        }
        clk.mark(1);
        if (fused_tail) {
            if (kRelu && lane == 0) tma_store_wait_read<0>();      // the tail's whole-batch phase reuses the stage buffers
            const int expected = int(part_owner((b + 1) * stages_per_sample - 1, T, G) - first_cta + 1);
            tail_after_flush<kConsumers>(tp, int(b), expected, tab, wred, &ticket_flag, reinterpret_cast<float*>(stage_buf), tid, 1, clk);
        } else if (tid == 0 && (b + 1) * stages_per_sample <= walk.R1) {
            slot_count[b] = int(k - first_cta + 1);
        }
#pragma unroll
        for (int e = 0; e < kPairAcc; ++e) acc2[e] = make_float2(0.f, 0.f);
    }
    if (kRelu && lane == 0) tma_store_wait<0>();                   // all of this warp's stores have left before the CTA exits
}

// ------------------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------------------
// Stage geometry of the backward (experiments: -DWTPSE_CL_APPLY_BOX=128 -DWTPSE_CL_APPLY_STAGES=3 ...).
#ifndef WTPSE_CL_APPLY_BOX
#define WTPSE_CL_APPLY_BOX 96
#endif
#ifndef WTPSE_CL_APPLY_STAGES
#define WTPSE_CL_APPLY_STAGES 4
#endif
#ifndef WTPSE_CL_APPLY_RELU_BOX
#define WTPSE_CL_APPLY_RELU_BOX 64
#endif
#ifndef WTPSE_CL_APPLY_RELU_STAGES
#define WTPSE_CL_APPLY_RELU_STAGES 3
#endif
template <bool kReluGrad>
struct ApplyCfg {
    static constexpr int kBox = kReluGrad ? WTPSE_CL_APPLY_RELU_BOX : WTPSE_CL_APPLY_BOX;   // pixels a warp owns per stage = one TMA box
    static constexpr int kBoxB = kBox * kC * 4;
    static constexpr int kStagePx = kConsumerWarps * kBox;
    static constexpr int kPartBytes = kStagePx * kC * 4;                    // one tensor's share of a stage
    static constexpr int kStageBytes = kPartBytes * (kReluGrad ? 2 : 1);    // fused: z, then grelu
    static constexpr int kStages = kReluGrad ? WTPSE_CL_APPLY_RELU_STAGES : WTPSE_CL_APPLY_STAGES;
    static constexpr int kPx = kBox / 32;                                   // pixels per lane and stage
    // input ring + one output staging buffer (a box per warp) + two matrices + barriers
    static constexpr size_t kSmem = 1024 + size_t(kStages) * kStageBytes + kPartBytes + 2 * 256 * sizeof(float) + 2 * kStages * sizeof(uint64_t);
};

template <bool kReluGrad>
__global__ void __launch_bounds__(kThreads, 1)
apply_cl_tma_kernel(const __grid_constant__ CUtensorMap tm_z, const __grid_constant__ CUtensorMap tm_g,
                    const __grid_constant__ CUtensorMap tm_dz, long long P, long long stages_per_sample, long long T, SeedArgs sa) {
    using Cfg = ApplyCfg<kReluGrad>;
    constexpr int kStages = Cfg::kStages;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* stage_buf = align1024(smem_dyn);
    unsigned char* out_buf = stage_buf + size_t(kStages) * Cfg::kStageBytes;       // [warp][box]: results on their way out
    float* msh2 = reinterpret_cast<float*>(out_buf + Cfg::kPartBytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(msh2 + 512);
    uint64_t* empty = full + kStages;
    __shared__ IndexTables tab;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    TileIter it;
    // one contiguous range of memory per CTA (stages dealt round-robin, the NCHW kernels' schedule, measured slower here:
    // 230 vs 197 us -- a stage is already one contiguous piece of memory, and every CTA then changes matrix every 4 stages)
    it.init(part_begin(k, T, G), part_begin(k + 1, T, G), 1, stages_per_sample);
    build_index_tables(tab, tid, kThreads);
    if (tid == 0) {
        prefetch_tensormap(&tm_z);
        prefetch_tensormap(&tm_dz);
        if (kReluGrad) prefetch_tensormap(&tm_g);
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kConsumerWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // z was written before the forward pass ran and is prefetched at once; grelu is the output of the kernel in front of
        // this one, which under programmatic dependent launch may still be running: wait before its first load
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();
            int stage = 0;
            uint32_t phase = 0;
            bool waited = !kReluGrad;
            int primed = 0;
            TileIter pit = it;
            // first pass over the ring: z boxes now, grelu boxes after the wait
            for (; pit.valid() && primed < kStages; pit.next(), ++primed) {
                mbar_arrive_expect_tx(&full[primed], Cfg::kStageBytes);
                unsigned char* dst = stage_buf + size_t(primed) * Cfg::kStageBytes;
                const int px0 = int(pit.tin * Cfg::kStagePx);
#pragma unroll
                for (int q = 0; q < kConsumerWarps; ++q) tma_load_box(dst + q * Cfg::kBoxB, &tm_z, px0 + q * Cfg::kBox, int(pit.b), &full[primed], policy);
            }
            if (kReluGrad) {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                waited = true;
                TileIter git = it;
                for (int s = 0; s < primed; ++s, git.next()) {
                    unsigned char* dst = stage_buf + size_t(s) * Cfg::kStageBytes + Cfg::kPartBytes;
                    const int px0 = int(git.tin * Cfg::kStagePx);
#pragma unroll
                    for (int q = 0; q < kConsumerWarps; ++q) tma_load_box(dst + q * Cfg::kBoxB, &tm_g, px0 + q * Cfg::kBox, int(git.b), &full[s], policy);
                }
            }
            (void)waited;
            stage = primed == kStages ? 0 : primed;
            phase = primed == kStages ? 1 : 0;
            for (; pit.valid(); pit.next()) {
                mbar_wait(&empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&full[stage], Cfg::kStageBytes);
                unsigned char* dst = stage_buf + size_t(stage) * Cfg::kStageBytes;
                const int px0 = int(pit.tin * Cfg::kStagePx);
#pragma unroll
                for (int q = 0; q < kConsumerWarps; ++q) tma_load_box(dst + q * Cfg::kBoxB, &tm_z, px0 + q * Cfg::kBox, int(pit.b), &full[stage], policy);
                if (kReluGrad) {
#pragma unroll
                    for (int q = 0; q < kConsumerWarps; ++q)
                        tma_load_box(dst + Cfg::kPartBytes + q * Cfg::kBoxB, &tm_g, px0 + q * Cfg::kBox, int(pit.b), &full[stage], policy);
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // saved tensors, upstream scalars and dz may belong to the kernel in front of us
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int stage = 0, cur = 0;
    uint32_t phase = 0;
    long long cur_b = it.valid() ? it.b : -1;
    SeedRegs ahead = seed_load(sa, tab, int(cur_b), tid);
    const SeedCtx sc = seed_context(sa, P);
    seed_store(ahead, sc, tab, msh2, tid);
    ahead = seed_load(sa, tab, it.valid() ? int(it.next_sample()) : -1, tid);      // one sample ahead
    named_bar_sync(1, kConsumers);
    for (; it.valid(); it.next()) {
        const long long b = it.b;
        if (b != cur_b) {
            seed_store(ahead, sc, tab, msh2 + (cur ^ 1) * 256, tid);              // the other buffer was last read two samples ago
            cur ^= 1;
            cur_b = b;
            ahead = seed_load(sa, tab, int(it.next_sample()), tid);
            named_bar_sync(1, kConsumers);
        }
        const float* msh = msh2 + cur * 256;
        mbar_wait(&full[stage], phase);
        unsigned char* sb = stage_buf + size_t(stage) * Cfg::kStageBytes;
        float4 xin[Cfg::kPx][4];
#pragma unroll
        for (int u = 0; u < Cfg::kPx; ++u) {
            const int p = warp * Cfg::kBox + u * 32 + lane;
#pragma unroll
            for (int q = 0; q < 4; ++q) xin[u][q] = lds4(sb, swz(p, q));
        }
        float4 out[Cfg::kPx][4];
#pragma unroll
        for (int u = 0; u < Cfg::kPx; ++u)
#pragma unroll
            for (int q = 0; q < 4; ++q) out[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
        // out[u][i] = sum_j M[i][j] x[u][j]: one warp-uniform LDS.128 of M feeds 4 FMAs per pixel
#pragma unroll
        for (int i = 0; i < kC; ++i) {
#pragma unroll
            for (int jq = 0; jq < 4; ++jq) {
                const float4 m = *reinterpret_cast<const float4*>(msh + i * kC + 4 * jq);
#pragma unroll
                for (int u = 0; u < Cfg::kPx; ++u) {
                    float& o = (i & 3) == 0 ? out[u][i >> 2].x : (i & 3) == 1 ? out[u][i >> 2].y : (i & 3) == 2 ? out[u][i >> 2].z : out[u][i >> 2].w;
                    o = fmaf(m.x, xin[u][jq].x, o);
                    o = fmaf(m.y, xin[u][jq].y, o);
                    o = fmaf(m.z, xin[u][jq].z, o);
                    o = fmaf(m.w, xin[u][jq].w, o);
                }
            }
        }
        if (kReluGrad) {                                       // added last: the rounding of autograd's dz_loss + dz_relu
#pragma unroll
            for (int u = 0; u < Cfg::kPx; ++u) {
                const int p = warp * Cfg::kBox + u * 32 + lane;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 xv = xin[u][q], gv = lds4(sb + Cfg::kPartBytes, swz(p, q));
                    float4& o = out[u][q];
                    o.x += xv.x <= 0.f ? 0.f : gv.x; o.y += xv.y <= 0.f ? 0.f : gv.y;
                    o.z += xv.z <= 0.f ? 0.f : gv.z; o.w += xv.w <= 0.f ? 0.f : gv.w;
                }
            }
        }
        // this warp is done with the input stage: hand it back at once (the ring keeps its full depth); the results leave
        // through the warp's own box of the output buffer, free again once the previous store has read it (issued a whole
        // stage ago: no stall)
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty[stage]);
            tma_store_wait_read<0>();
        }
        __syncwarp();
        unsigned char* ob = out_buf + size_t(warp) * Cfg::kBoxB;
#pragma unroll
        for (int u = 0; u < Cfg::kPx; ++u) {
            const int p = u * 32 + lane;                       // row inside the warp's box (boxes start on 1 KB boundaries)
#pragma unroll
            for (int q = 0; q < 4; ++q) sts4(ob, swz(p, q), out[u][q]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            tma_store_box(&tm_dz, int(it.tin * Cfg::kStagePx) + warp * Cfg::kBox, int(b), ob);
            tma_store_commit();
        }
        if (++stage == kStages) { stage = 0; phase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();
}

// ---- host side ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// {16 channels, P pixels, B samples} fp32, box 16 x box_px x 1, 64-byte swizzle
bool make_map(CUtensorMap* tm, const float* base, int B, long long P, int box_px) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {cuuint64_t(kC), cuuint64_t(P), cuuint64_t(B)};
    const cuuint64_t strides[2] = {cuuint64_t(kC) * 4, cuuint64_t(P) * kC * 4};
    const cuuint32_t box[3] = {cuuint32_t(kC), cuuint32_t(box_px), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// The tensor-map kernels take any 16-byte aligned [B][P][16] tensor whose pixel count fits the TMA coordinates.
bool cl_tma_ok(const float* a, const float* b, const float* c, long long P) {
    return g_cl_tma && encode_fn() != nullptr && P >= 1 && P < (1LL << 31) - 2048 && aligned16(a) && aligned16(b) && aligned16(c);
}

GramPlan plan_gram_cl_tma(int B, long long P, int sm_count) {
    GramPlan g{};
    g.tma = true;
    g.group = 1;
    g.tiles_per_sample = (P + kStagePx - 1) / kStagePx;
    g.T = g.tiles_per_sample * B;
    g.G = g.T < sm_count ? g.T : sm_count;
    int nslots = 1;
    for (int b = 0; b < B; ++b) {
        const long long first = part_owner((long long)b * g.tiles_per_sample, g.T, g.G);
        const long long last = part_owner((long long)(b + 1) * g.tiles_per_sample - 1, g.T, g.G);
        if (last - first + 1 > nslots) nslots = int(last - first + 1);
    }
    g.nslots = nslots;
    return g;
}

bool gram_cl_tail_fits(int B, int n_per_domain, int n_domains) {
    long long m = n_domains > 1 ? (long long)n_per_domain * n_domains : 0;
    if (m > B) m = B;
    return tail_smem_bytes(B, int(m), n_domains) <= size_t(kGramStages) * kPartBytes;
}

cudaError_t launch_gram_cl_tma(const float* z, float* relu_out, float* partial, int* slot_count, int B, long long P, const GramPlan& g,
                               cudaStream_t stream, const TailParams* tail) {
    CUtensorMap tm_z, tm_relu;
    if (!make_map(&tm_z, z, B, P, kBoxPx)) return cudaErrorInvalidValue;
    if (!make_map(&tm_relu, relu_out ? relu_out : z, B, P, kBoxPx)) return cudaErrorInvalidValue;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(g.G));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kGramSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TailParams tp{};
    if (tail) tp = *tail;
    cudaError_t e;
    if (relu_out) {
        e = cudaFuncSetAttribute(gram_cl_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGramSmem));
        if (e != cudaSuccess) return e;
        return cudaLaunchKernelEx(&cfg, gram_cl_tma_kernel<true>, tm_z, tm_relu, partial, slot_count, g.tiles_per_sample, g.T, g.nslots,
                                  tail ? 1 : 0, tp);
    }
    e = cudaFuncSetAttribute(gram_cl_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kGramSmem));
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, gram_cl_tma_kernel<false>, tm_z, tm_relu, partial, slot_count, g.tiles_per_sample, g.T, g.nslots,
                              tail ? 1 : 0, tp);
}

namespace {
template <bool kReluGrad>
cudaError_t launch_apply_cl_tma_t(const float* z, const float* grelu, const SeedArgs& seed, float* dz, int B, long long P, int sm_count,
                                  cudaStream_t stream) {
    using Cfg = ApplyCfg<kReluGrad>;
    CUtensorMap tm_z, tm_g, tm_dz;
    if (!make_map(&tm_z, z, B, P, Cfg::kBox) || !make_map(&tm_g, grelu ? grelu : z, B, P, Cfg::kBox) || !make_map(&tm_dz, dz, B, P, Cfg::kBox))
        return cudaErrorInvalidValue;
    const long long sps = (P + Cfg::kStagePx - 1) / Cfg::kStagePx;
    const long long T = sps * B;
    const long long G = T < sm_count ? T : sm_count;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(G));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaFuncSetAttribute(apply_cl_tma_kernel<kReluGrad>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(Cfg::kSmem));
    if (e != cudaSuccess) return e;
    return cudaLaunchKernelEx(&cfg, apply_cl_tma_kernel<kReluGrad>, tm_z, tm_g, tm_dz, P, sps, T, seed);
}
}  // namespace

cudaError_t launch_apply_cl_tma(const float* z, const float* grelu, const SeedArgs& seed, float* dz, int B, long long P, int sm_count,
                                cudaStream_t stream) {
    return grelu ? launch_apply_cl_tma_t<true>(z, grelu, seed, dz, B, P, sm_count, stream)
                 : launch_apply_cl_tma_t<false>(z, grelu, seed, dz, B, P, sm_count, stream);
}

}  // namespace wtpse

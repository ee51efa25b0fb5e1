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
// Default path (apply_tma_kernel<false>): the matrices come from whiten_mmat_kernel (one CTA per sample,
// whitening_epilogue.cu); this kernel is launched as its programmatic dependent, so z is already streaming
// while the matrices are derived, and the consumers wait (griddepcontrol.wait) only before the first M_b read.
//
// Alternative (apply_tma_kernel<true>, wtpse_debug_set_backward_mode(1)): every CTA derives M_b itself for the
// one or two samples of its CONTIGUOUS tile range -- one launch instead of two, but without the round-robin
// schedule's DRAM locality (188 us vs 176 us for the pair at 32x16x512x512).  Same arithmetic (mmd_device.cuh),
// bitwise-equal results (tests/test_gpu_parity.py).
#include "common.cuh"
#include "kernels.h"
#include "mmd_device.cuh"

namespace wtpse {

namespace {

constexpr int kConsumers = 256;
constexpr int kConsumerWarps = kConsumers / 32;
constexpr int kThreads = kConsumers + 32;
constexpr int kTilePx = kConsumers * 4;
constexpr int kStages = 3;
constexpr int kStageFloats = kC * kTilePx;
constexpr int kFusedMaxM = 64;     // MMD samples whose vectors fit beside the pipeline stages
constexpr size_t kSmemPipe = size_t(kStages) * kStageFloats * sizeof(float);
constexpr size_t kSmemPlain = kSmemPipe + 256 * sizeof(float) + 2 * kStages * sizeof(uint64_t);
constexpr size_t kSmemFused = kSmemPlain + (size_t(kFusedMaxM) * kVStride + kFusedMaxM) * sizeof(float) + sizeof(IndexTables) + 16;

__device__ __forceinline__ void st_stream(float* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct FusedArgs {
    const float* gram;      // [B][16][16]
    const float* rowstat;   // [B][2]
    const float *g_off, *g_diag, *g_dom;
    int B, n, K;
    int round_robin;
    int l2_hint;            // 1: TMA loads carry an L2 evict-first policy (z is read once)
};

template <bool kFused>
__global__ void __launch_bounds__(kThreads, 1)
apply_tma_kernel(const float* __restrict__ z, const float* __restrict__ mmat, float* __restrict__ dz, long long P,
                 long long tiles_per_sample, long long T, FusedArgs fa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    float* msh = stage_buf + size_t(kStages) * kStageFloats;
    uint64_t* full = reinterpret_cast<uint64_t*>(msh + 256);
    uint64_t* empty = full + kStages;
    float* vbuf = reinterpret_cast<float*>(empty + kStages);          // [kFusedMaxM][124]   (fused only)
    float* coefrow = vbuf + size_t(kFusedMaxM) * kVStride;            // [kFusedMaxM]
    IndexTables* tab = reinterpret_cast<IndexTables*>(coefrow + kFusedMaxM);

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // e.g. the next Gram kernel: it waits for us itself
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long G = gridDim.x, k = blockIdx.x;
    // tile schedule: one contiguous range per CTA, or blocks of tiles dealt round-robin over the grid (default: at any
    // moment the CTAs then stream adjacent tiles of every channel row)
    // round_robin = c > 0: blocks of c consecutive tiles dealt round-robin to the CTAs (block q -> CTA q % G)
    const long long chunk = fa.round_robin;
    const bool rr = chunk > 0;
    const long long nblocks = rr ? (T + chunk - 1) / chunk : 0;
    const long long my_blocks = rr ? (nblocks > k ? (nblocks - k + G - 1) / G : 0) : 0;
    const long long t0 = rr ? 0 : part_begin(k, T, G);
    const long long t1 = rr ? my_blocks * chunk : part_begin(k + 1, T, G);
    auto tile_of = [=](long long j) -> long long { return rr ? ((j / chunk) * G + k) * chunk + (j % chunk) : j; };

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
            const uint64_t policy = make_evict_first_policy();
            const bool g_use_hint = fa.l2_hint != 0;
            int stage = 0;
            uint32_t phase = 0;
            for (long long j = t0; j < t1; ++j) {
                const long long t = tile_of(j);
                if (t >= T) continue;
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
                for (int c = 0; c < kC; ++c) {
                        if (g_use_hint) tma_load_1d_hint(dst + c * kTilePx, src + c * P, bytes, &full[stage], policy);
                        else tma_load_1d(dst + c * kTilePx, src + c * P, bytes, &full[stage]);
                    }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // ---- fused prologue: everything that does not depend on the sample ------------------------------
    DomainInfo dom{0, 0, 0, 0};
    float g_dom = 0.f, w_off = 0.f, w_diag = 0.f;
    bool need_dom = false;
    const float denom = float(P - 1);
    if (kFused) {
        dom = make_domain(fa.B, fa.n, fa.K);
        const float g_off = fa.g_off ? __ldg(fa.g_off) : 0.f;
        const float g_diag = fa.g_diag ? __ldg(fa.g_diag) : 0.f;
        g_dom = fa.g_dom ? __ldg(fa.g_dom) : 0.f;
        need_dom = (dom.M > 0) && (g_dom != 0.f);
        w_off = g_off / (float(fa.B) * float(kOff));
        w_diag = g_diag / (float(fa.B) * float(kC));
        build_index_tables(*tab, tid, kConsumers);
        named_bar_sync(1, kConsumers);
        if (need_dom) {
            for (int idx = tid; idx < dom.M * kOff; idx += kConsumers) {
                const int b = idx / kOff, o = idx - b * kOff;
                const int ij = tab->off[o];
                vbuf[size_t(b) * kVStride + o] = __ldg(fa.gram + b * 256 + (ij >> 4) * kC + (ij & 15));
            }
        }
    }

    int stage = 0;
    uint32_t phase = 0;
    long long cur_b = -1;
    for (long long j = t0; j < t1; ++j) {
        const long long t = tile_of(j);
        if (t >= T) continue;
        const long long b = t / tiles_per_sample;
        const long long px0 = (t - b * tiles_per_sample) * kTilePx;
        const long long rem = P - px0;
        if (b != cur_b) {
            named_bar_sync(1, kConsumers);          // previous sample's matrix no longer in use; vbuf staged
            if (kFused) {
                const int bi = int(b);
                const bool in_mmd = need_dom && bi < dom.M;
                if (in_mmd) {
                    for (int c = tid; c < dom.M; c += kConsumers) {
                        const float D = mmd_distance(vbuf + size_t(bi) * kVStride, vbuf + size_t(c) * kVStride);
                        coefrow[c] = mmd_coefficient(dom, bi, c, expf(-D));
                    }
                    named_bar_sync(1, kConsumers);
                }
                if (tid < kTri) {
                    const int ij = tab->tri[tid], i = ij >> 4, j = ij & 15;
                    const float g = __ldg(fa.gram + bi * 256 + i * kC + j);
                    float dom_grad = 0.f;
                    if (in_mmd && i != j) dom_grad = g_dom * mmd_grad_entry(vbuf, coefrow, dom.M, bi, off_idx(i, j));
                    const float m = backward_matrix_entry(i, j, g, __ldg(fa.rowstat + bi * 2 + 0),
                                                          __ldg(fa.rowstat + bi * 2 + 1), w_off, w_diag, dom_grad, denom);
                    msh[i * kC + j] = m;
                    msh[j * kC + i] = m;
                }
            } else {
                // launched as a programmatic dependent of whiten_mmat_kernel: z is streaming already, the
                // matrices are only needed here (no-op when there is no programmatic dependency)
                if (cur_b < 0) asm volatile("griddepcontrol.wait;" ::: "memory");
                msh[tid] = __ldcg(mmat + b * 256 + tid);   // coherent load: an invariant (.nc) one may be hoisted above the wait
            }
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
apply_generic_kernel(const float* __restrict__ z, const float* __restrict__ mmat, const float* __restrict__ grelu,
                     float* __restrict__ dz, long long P) {
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

bool apply_can_fuse(const float* z, const float* dz, int B, long long P, int n_per_domain, int n_domains) {
    long long m = n_domains > 1 ? (long long)n_per_domain * n_domains : 0;
    if (m > B) m = B;
    return tma_ok(z, dz, P) && m <= kFusedMaxM;
}

cudaError_t launch_apply(const float* z, const float* mmat, float* dz, int B, long long P, int sm_count,
                         cudaStream_t stream, bool programmatic_dependent, const float* grelu) {
    if (grelu && apply_relu_tma_ok(z, grelu, dz, P))
        return launch_apply_relu(z, grelu, mmat, dz, B, P, sm_count, stream, programmatic_dependent);
    if (!grelu && tma_ok(z, dz, P)) {
        auto kern = apply_tma_kernel<false>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemPlain));
        if (e != cudaSuccess) return e;
        const long long tps = (P + kTilePx - 1) / kTilePx;
        const long long T = tps * B;
        const long long G = T < sm_count ? T : sm_count;
        FusedArgs fa{};
        fa.round_robin = g_apply_round_robin;   // 0 = contiguous ranges, c > 0 = round-robin blocks of c tiles
        fa.l2_hint = g_l2_evict_first;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(unsigned(G));
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = kSmemPlain;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = programmatic_dependent ? 1 : 0;
        return cudaLaunchKernelEx(&cfg, kern, z, mmat, dz, P, tps, T, fa);
    } else {
        apply_generic_kernel<<<dim3(unsigned((P + 255) / 256), unsigned(B)), 256, 0, stream>>>(z, mmat, grelu, dz, P);
    }
    return cudaGetLastError();
}

cudaError_t launch_apply_fused(const float* z, const float* gram, const float* rowstat, const float* g_off,
                               const float* g_diag, const float* g_dom, float* dz, int B, long long P, int n_per_domain,
                               int n_domains, int sm_count, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(apply_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kSmemFused));
    if (e != cudaSuccess) return e;
    const long long tps = (P + kTilePx - 1) / kTilePx;
    const long long T = tps * B;
    const long long G = T < sm_count ? T : sm_count;
    FusedArgs fa{gram, rowstat, g_off, g_diag, g_dom, B, n_per_domain, n_domains, 0, g_l2_evict_first};
    apply_tma_kernel<true><<<dim3(unsigned(G)), kThreads, kSmemFused, stream>>>(z, nullptr, dz, P, tps, T, fa);
    return cudaGetLastError();
}

}  // namespace wtpse

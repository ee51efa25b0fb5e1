// The forward tail as STAND-ALONE kernels, and the standalone MMD (compute_MMD.forward on a B x 120 input).
//
// The shipped forward runs this work inside the Gram kernel (whitening_tail.cuh: last-arriving CTA).  The kernels here
// are its fallback -- inputs the TMA pipeline cannot take (P % 4 != 0, unaligned pointers), the per-thread
// channels-last kernel, batches whose MMD working set exceeds one CTA's shared memory (global scratch variants), and
// wtpse_debug_set("fused_tail", 0) -- and the reference the in-kernel tail is compared with bit for bit
// (tests/test_gpu_parity.py).  Both call the same device code (mmd_device.cuh).
//
// Forward  (algorithms.py:1283-1307 after the bmm, compute_MMD.forward algorithms.py:102-121):
//   partial Gram slots -> f_cor = G/(P-1) + eps*I -> off_b, diag_b -> L_off, L_diag
//   -> 120-d upper-triangle vectors -> pairwise exp(-D) -> L_dom.
//   gram_reduce_kernel           one CTA per sample: slot reduction, Gram, vector, off_b, diag_b
//   whiten_epilogue_fwd_kernel   ONE CTA: instance terms, pairwise kernel values, block sums, L_dom
//   mmd_bwd_kernel               ONE CTA: d L_dom / d v (the backward's seed `domgrad`; also wtpse_mmd_backward)
//
// The work is O(B*136*slots + B^2*120) -- microseconds -- so these kernels are latency-bound and are
// built around that:
//   * a single SM issues at most 4 warp instructions per cycle: whatever is per-sample runs as one CTA per
//     sample; only the genuinely all-to-all part (the MMD) runs in one CTA, with fixed reduction orders:
//     bit-reproducible, no float atomics, no grid sync;
//   * float32 arithmetic like the reference.  B200's FP64 pipe issues ~1 warp instruction per 3
//     cycles with ~50-cycle dependent latency; a first float64 version of this file spent 25 us per
//     launch in it.  Accuracy is kept where the MMD needs it by two reformulations instead:
//       - distances use the difference form sum_e (v_a - v_c)^2, not |x|^2 + |y|^2 - 2x.y
//         (algorithms.py:65-71), so there is no cancellation in D;
//       - the kernel value is carried as u = expm1(-D) = E - 1.  Kxx + Kyy - 2Kxy is invariant under
//         E -> E - 1, so the O(1) parts cancel analytically and the result keeps full relative
//         precision even when the domains are statistically identical (SURVEY.md section 7's
//         3.8e-3 fp32-vs-fp64 gap of the reference does not arise).  Only the handful of final
//         block sums run in float64.
//   * everything these kernels read from global memory is loaded with ld.global.cg (__ldcg), never the read-only
//     (.nc / __ldg) path: they are chained by programmatic dependent launch, their inputs are written by the kernel
//     in front while they are already resident, and ptxas hoists invariant loads above griddepcontrol.wait
//     (seen in SASS: LDG.E.CONSTANT in front of ACQBULK) -- a race that showed up as one flaky gradient test;
//   * dependent loads are issued in batches (8-deep gathers) so a phase pays one L2/DRAM round trip, not one per element;
//   * pairwise distances: one LANE per unordered pair a < c (the matrix is symmetric), LDS.128 on rows
//     padded to 124 floats (conflict-free), so there is no shuffle chain per pair;
//   * index tables live in shared memory and the kernels are templated on shared-vs-global working set
//     (per-lane indexing of __constant__ memory and generic pointers to shared memory both serialise
//     in the address-divergence unit).
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "mmd_device.cuh"

namespace wtpse {

namespace {

constexpr int kEpiThreads = 1024;
constexpr int kEpiWarps = kEpiThreads / 32;
constexpr size_t kEpiSmemCap = 200 * 1024;

template <typename F>
__device__ __forceinline__ void pairwise_upper(const float* __restrict__ v, int M, int tid, F&& emit) {
    pairwise_upper_n(v, M, tid, kEpiThreads, emit);
}

// ------------------------------------------------------------------------------------------------
struct FwdParams {
    int B;
    long long P;
    int n, K;
    float* losses;
    void* scratch;      // global fallback for EpiMem
    int in_smem;
    const float* vd;    // from gram_reduce_kernel: [B][124] vectors and [B][2] row statistics
    const float* statd;
};

template <bool kSmem>
__global__ void __launch_bounds__(kEpiThreads, 1) whiten_epilogue_fwd_kernel(FwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B;
    const DomainInfo dom = make_domain(B, p.n, p.K);
    const EpiMem mem = resolve_mem<kSmem>(smem_raw, p.scratch, B, dom.M);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // dependents wait for our completion themselves

    // A. the per-sample work was done by gram_reduce_kernel (one CTA per sample): stage its results
    asm volatile("griddepcontrol.wait;" ::: "memory");
    {
        // rows are 124 floats in both places: copy them as float4, four independent loads in flight per thread
        const float4* src4 = reinterpret_cast<const float4*>(p.vd);
        float4* dst4 = reinterpret_cast<float4*>(mem.v);
        const int n4 = dom.M * (kVStride / 4);
        for (int base = 0; base < n4; base += 4 * kEpiThreads) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * kEpiThreads + tid;
                v[u] = idx < n4 ? __ldcg(src4 + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * kEpiThreads + tid;
                if (idx < n4) dst4[idx] = v[u];
            }
        }
        for (int idx = tid; idx < 2 * B; idx += kEpiThreads) mem.stat[idx] = __ldcg(p.statd + idx);
    }
    __syncthreads();

    // B. pairwise u = exp(-D) - 1
    {
        float* U = mem.U;
        const int M = dom.M;
        pairwise_upper(mem.v, M, tid, [U, M](int a, int c, float D) {
            const float u = expm1f(-D);
            U[a * M + c] = u;
            U[c * M + a] = u;
        });
        for (int a = tid; a < M; a += kEpiThreads) U[a * M + a] = expm1f(-1e-30f);   // D(a,a) = 0 -> clamp_min_(1e-30)
    }
    __syncthreads();

    // C. instance terms (warp 0), per-domain-pair block sums (warps 1..)
    if (warp == 0) {
        float so = 0.f, sd = 0.f;
        for (int b = lane; b < B; b += 32) {
            so += clamp0(mem.stat[b * 2 + 0] / float(kOff));   // clamp(off_diag_sum / 120, min=0), :1290
            sd += clamp0(mem.stat[b * 2 + 1] / float(kC));     // clamp(diag_sum / 16, min=0), :1298
        }
        so = warp_sum(so) / float(B);
        sd = warp_sum(sd) / float(B);
        if (lane == 0) {
            p.losses[0] = so;
            p.losses[1] = sd;
            p.losses[3] = so + sd;
        }
    } else if (dom.M > 0) {
        domain_block_sums(mem.U, dom, mem.blk, 1, kEpiWarps - 1, warp, lane);
    }
    __syncthreads();

    // D. L_dom
    if (warp == 0) {
        const float pen = mmd_from_blocks(mem.blk, dom, lane);
        if (lane == 0) p.losses[2] = pen;
    }
}

// ------------------------------------------------------------------------------------------------
// Forward stage 2a: the slot reduction is done by one CTA PER SAMPLE, in a fixed order, together with
// everything else that is per-sample (Gram entries, 120-d vector, off_b, diag_b).  The single-CTA epilogue
// that follows only does the MMD and the final sums.
constexpr int kReduceParts = 4;
constexpr int kReduceThreads = kTri * kReduceParts;   // 544

struct ReduceParams {
    const float* partial;     // [B][G][136]
    const int* slot_count;    // slots used per sample
    long long G;              // slots allocated per sample
    int B;
    long long P;
    int n, K;
    float margin, eps;
    float* gram;
    float* rowstat;
    float* vd;                // [B][124]
    float* statd;             // [B][2]
};

__global__ void __launch_bounds__(kReduceThreads) gram_reduce_kernel(ReduceParams p) {
    __shared__ float red[kReduceParts][kTri];
    __shared__ float wred[2][8];
    __shared__ IndexTables tab;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x;
    const int e = tid % kTri, part = tid / kTri;
    build_index_tables(tab, tid, kReduceThreads);
    asm volatile("griddepcontrol.wait;" ::: "memory");               // partials come from the Gram kernel
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int G = int(p.G);
    const int per = (G + kReduceParts - 1) / kReduceParts;
    const int k0 = part * per, k1 = (k0 + per < G) ? k0 + per : G;
    const int cnt = __ldcg(p.slot_count + b);
    const float* src = p.partial + ((long long)b * G) * kTri + e;
    float s = 0.f;
    for (int k = k0; k < k1; k += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int kk = k + u;
            v[u] = (kk < k1 && kk < cnt) ? __ldcg(src + (long long)kk * kTri) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    red[part][e] = s;
    __syncthreads();
    float off = 0.f, dg = 0.f;
    if (tid < kTri) {
        float g = ((red[0][e] + red[1][e]) + red[2][e]) + red[3][e];
        const int ij = tab.tri[e], i = ij >> 4, j = ij & 15;
        g = g / float(p.P - 1);                              // .div(HW - 1), algorithms.py:1283
        if (i == j) {
            g += p.eps;
            dg = fabsf(g - 1.0f);
            p.gram[b * 256 + i * kC + i] = g;
        } else {
            off = fabsf(g);
            p.gram[b * 256 + i * kC + j] = g;
            p.gram[b * 256 + j * kC + i] = g;
            p.vd[size_t(b) * kVStride + off_idx(i, j)] = g;
        }
    }
    if (warp < 5) {                                          // the 136 entry threads live in warps 0..4
        off = warp_sum(off);
        dg = warp_sum(dg);
        if (lane == 0) { wred[0][warp] = off; wred[1][warp] = dg; }
    }
    __syncthreads();
    if (tid == 0) {
        float so = 0.f, sd = 0.f;
        for (int w = 0; w < 5; ++w) { so += wred[0][w]; sd += wred[1][w]; }
        so -= p.margin;
        sd -= p.margin;
        p.rowstat[b * 2 + 0] = so;
        p.rowstat[b * 2 + 1] = sd;
        p.statd[b * 2 + 0] = so;
        p.statd[b * 2 + 1] = sd;
    }
}

// ---- standalone compute_MMD.forward on a B x 120 input (algorithms.py:102-121) -------------------
struct MmdParams {
    const float* v32;
    int stride;          // floats between consecutive rows of v32 (120 for the public entry points, 124 for the workspace)
    const float* gout;
    int B, n, K;
    float* loss;
    float* dv;
    void* scratch;
    int in_smem;
};

__device__ __forceinline__ void mmd_stage_vectors(const MmdParams& p, const EpiMem& mem, int M, int tid) {
    for (int idx = tid; idx < M * kOff; idx += kEpiThreads) {
        const int b = idx / kOff, o = idx - b * kOff;
        mem.v[size_t(b) * kVStride + o] = __ldcg(p.v32 + size_t(b) * p.stride + o);
    }
}

template <bool kSmem>
__global__ void __launch_bounds__(kEpiThreads, 1) mmd_fwd_kernel(MmdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DomainInfo dom = make_domain(p.B, p.n, p.K);
    const int M = dom.M;
    const EpiMem mem = resolve_mem<kSmem>(smem_raw, p.scratch, p.B, M);
    mmd_stage_vectors(p, mem, M, tid);
    __syncthreads();
    float* U = mem.U;
    pairwise_upper(mem.v, M, tid, [U, M](int a, int c, float D) {
        const float u = expm1f(-D);
        U[a * M + c] = u;
        U[c * M + a] = u;
    });
    for (int a = tid; a < M; a += kEpiThreads) U[a * M + a] = expm1f(-1e-30f);
    __syncthreads();
    if (M > 0) domain_block_sums(mem.U, dom, mem.blk, 0, kEpiWarps, warp, lane);
    __syncthreads();
    if (warp == 0) {
        const float pen = mmd_from_blocks(mem.blk, dom, lane);
        if (lane == 0) p.loss[0] = pen;
    }
}

template <bool kSmem>
__global__ void __launch_bounds__(kEpiThreads, 1) mmd_bwd_kernel(MmdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const DomainInfo dom = make_domain(p.B, p.n, p.K);
    const int M = dom.M;
    const EpiMem mem = resolve_mem<kSmem>(smem_raw, p.scratch, p.B, M);
    mmd_stage_vectors(p, mem, M, tid);
    __syncthreads();
    float* coef = mem.U;
    pairwise_upper(mem.v, M, tid, [coef, M, &dom](int a, int c, float D) {
        const float E = expf(-D);
        coef[a * M + c] = mmd_coefficient(dom, a, c, E);
        coef[c * M + a] = mmd_coefficient(dom, c, a, E);
    });
    for (int a = tid; a < M; a += kEpiThreads) coef[a * M + a] = 0.f;
    __syncthreads();
    const float g = p.gout ? __ldcg(p.gout) : 1.f;
    for (int idx = tid; idx < p.B * kOff; idx += kEpiThreads) {
        const int b = idx / kOff, o = idx - b * kOff;
        p.dv[idx] = (b < M) ? g * mmd_grad_entry(mem.v, mem.U + size_t(b) * M, M, b, o) : 0.f;
    }
}

int mmd_samples(int B, int n, int K) {
    if (K <= 1) return 0;
    const long long m = (long long)n * K;
    return int(m < B ? m : B);
}

// dynamic shared memory for the working set, or 0 when it has to live in the global scratch
size_t epi_smem(const void* fn, int B, int n, int K, int* in_smem, cudaError_t* err) {
    const size_t bytes = epi_mem_bytes(B, mmd_samples(B, n, K), K);
    *err = cudaSuccess;
    *in_smem = bytes <= kEpiSmemCap ? 1 : 0;
    if (!*in_smem) return 0;
    if (bytes > 48 * 1024) *err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    return bytes;
}

}  // namespace

size_t epilogue_scratch_bytes(int B, int K) { return epi_mem_bytes(B, B, K); }

cudaError_t launch_gram_reduce(const float* partial, const int* slot_count, int nslots, int B, long long P, int n_per_domain,
                               int n_domains, float margin, float eps, float* gram, float* rowstat, float* vd, float* statd,
                               cudaStream_t stream) {
    ReduceParams p;
    p.partial = partial; p.B = B; p.P = P;
    p.G = nslots;
    p.slot_count = slot_count;
    p.n = n_per_domain; p.K = n_domains; p.margin = margin; p.eps = eps;
    p.gram = gram; p.rowstat = rowstat; p.vd = vd; p.statd = statd;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(B));
    cfg.blockDim = dim3(kReduceThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, gram_reduce_kernel, p);
}

cudaError_t launch_whiten_epilogue_fwd(int B, long long P, int n_per_domain, int n_domains, float* losses, void* scratch,
                                       cudaStream_t stream, const float* vd, const float* statd) {
    FwdParams p;
    p.B = B; p.P = P; p.n = n_per_domain; p.K = n_domains;
    p.losses = losses; p.scratch = scratch; p.vd = vd; p.statd = statd;
    cudaError_t e;
    const size_t dyn = epi_smem((const void*)whiten_epilogue_fwd_kernel<true>, B, n_per_domain, n_domains, &p.in_smem, &e);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(kEpiThreads);
    cfg.dynamicSmemBytes = p.in_smem ? dyn : 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return p.in_smem ? cudaLaunchKernelEx(&cfg, whiten_epilogue_fwd_kernel<true>, p)
                     : cudaLaunchKernelEx(&cfg, whiten_epilogue_fwd_kernel<false>, p);
}

cudaError_t launch_mmd(const float* v, int stride, const float* gout, int B, int n_per_domain, int n_domains, float* loss, float* dv,
                       void* scratch, cudaStream_t stream) {
    MmdParams p;
    p.v32 = v; p.stride = stride; p.gout = gout; p.B = B; p.n = n_per_domain; p.K = n_domains; p.loss = loss; p.dv = dv;
    p.scratch = scratch;
    const void* fn = dv ? (const void*)mmd_bwd_kernel<true> : (const void*)mmd_fwd_kernel<true>;
    cudaError_t e;
    const size_t dyn = epi_smem(fn, B, n_per_domain, n_domains, &p.in_smem, &e);
    if (e != cudaSuccess) return e;
    if (dv) {
        if (p.in_smem) mmd_bwd_kernel<true><<<1, kEpiThreads, dyn, stream>>>(p);
        else mmd_bwd_kernel<false><<<1, kEpiThreads, 0, stream>>>(p);
    } else {
        if (p.in_smem) mmd_fwd_kernel<true><<<1, kEpiThreads, dyn, stream>>>(p);
        else mmd_fwd_kernel<false><<<1, kEpiThreads, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace wtpse

// Forward stage 2 and backward stage 1: everything between the Gram and the scalars, plus the
// standalone MMD (compute_MMD.forward on a B x 120 input).
//
// Forward  (algorithms.py:1283-1307 after the bmm, compute_MMD.forward algorithms.py:102-121):
//   partial Gram slots -> f_cor = G/(P-1) + eps*I -> off_b, diag_b -> L_off, L_diag
//   -> 120-d upper-triangle vectors -> pairwise exp(-D) -> L_dom.
//   gram_reduce_kernel           one CTA per sample: slot reduction, Gram, vector, off_b, diag_b
//   whiten_epilogue_fwd_kernel   ONE CTA: instance terms, pairwise kernel values, block sums, L_dom
//                                (also able to do the per-sample part itself: wtpse_debug_set_two_stage_epilogue(0))
// Backward: upstream grads + saved Gram -> per-sample symmetric 16x16
//   coefficient matrix M_b = (S_b + S_b^T)/(P-1) that the apply kernel multiplies into z.
//   whiten_mmat_kernel           one CTA per sample (default)
//   whiten_epilogue_bwd_kernel   ONE CTA for all samples (fallback for very many MMD samples, debug mode 2)
//
// The work is O(B*136*slots + B^2*120) -- microseconds -- so these kernels are latency-bound and are
// built around that (measured with the clock64 stamps below, tools/epilogue_phases.py):
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
//   * dependent loads are issued in batches (speculative slot loads, 8-deep gathers) so a phase pays one
//     L2/DRAM round trip, not one per element;
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

#define WTPSE_STAMP(k) do { if (p.dbg && threadIdx.x == 0) p.dbg[k] = clock64(); } while (0)

constexpr int kEpiThreads = 1024;
constexpr int kEpiWarps = kEpiThreads / 32;
constexpr size_t kEpiSmemCap = 200 * 1024;
// Working set: v [M][124] f32 | U [M][M] f32 | stat [B][2] f32 | blk [K*K] f64 ; shared memory when it fits.
struct EpiMem {
    float* v;
    float* U;
    float* stat;
    double* blk;
};

__host__ __device__ inline size_t round4(size_t x) { return (x + 3) & ~size_t(3); }
__host__ __device__ inline size_t epi_mem_bytes(int B, int M, int K) {
    const size_t kk = size_t(K > 0 ? K : 1) * size_t(K > 0 ? K : 1);
    return (round4(size_t(M) * kVStride) + round4(size_t(M) * M) + round4(size_t(B) * 2)) * sizeof(float) + kk * sizeof(double);
}

// kSmem is a template parameter so that the compiler sees shared-space pointers (LDS/STS) instead of
// generic ones: generic accesses to shared memory go through the address-divergence unit and made
// every phase of these kernels ~5x slower.
template <bool kSmem>
__device__ __forceinline__ EpiMem resolve_mem(void* smem, void* global, int B, int M) {
    float* base = reinterpret_cast<float*>(kSmem ? smem : global);
    EpiMem m;
    m.v = base;
    m.U = m.v + round4(size_t(M) * kVStride);
    m.stat = m.U + round4(size_t(M) * M);
    m.blk = reinterpret_cast<double*>(m.stat + round4(size_t(B) * 2));    // 16-byte aligned by construction
    return m;
}

// D(a, c) = max(sum_e (v_a[e] - v_c[e])^2, 1e-30) for all pairs a < c < M, one LANE per pair (no shuffle chain);
// emit(a, c, D) is expected to fill both (a, c) and (c, a).  Consecutive lanes share a and walk c, so the
// v_a loads broadcast and the v_c loads hit distinct banks (row stride 124 floats).
template <typename F>
__device__ __forceinline__ void pairwise_upper(const float* __restrict__ v, int M, int tid, F&& emit) {
    const int npairs = M * (M - 1) / 2;
    for (int pidx = tid; pidx < npairs; pidx += kEpiThreads) {
        // row a of the strict upper triangle that contains flat index pidx (row a starts at a*(2M-a-1)/2)
        const float t = float(2 * M - 1);
        int a = int((t - sqrtf(t * t - 8.0f * float(pidx))) * 0.5f);
        while (a > 0 && a * (2 * M - a - 1) / 2 > pidx) --a;
        while ((a + 1) * (2 * M - a - 2) / 2 <= pidx) ++a;
        const int c = a + 1 + (pidx - a * (2 * M - a - 1) / 2);
        emit(a, c, mmd_distance(v + size_t(a) * kVStride, v + size_t(c) * kVStride));
    }
}

// per-domain-pair sums of u = E - 1: one warp per (k <= l) block, fp32 lane partials in a fixed order,
// float64 across lanes
__device__ void domain_block_sums(const float* __restrict__ U, const DomainInfo& dom, double* __restrict__ blk,
                                  int first_warp, int nwarps, int warp, int lane) {
    const int K = dom.K;
    for (int pr = warp - first_warp; pr < K * K; pr += nwarps) {
        const int k = pr / K, l = pr - k * K;
        if (l < k) continue;
        const int a0 = chunk_lo(k, dom.n, dom.B), a1 = chunk_lo(k + 1, dom.n, dom.B);
        const int c0 = chunk_lo(l, dom.n, dom.B), c1 = chunk_lo(l + 1, dom.n, dom.B);
        const int na = a1 - a0, nc = c1 - c0;
        float s = 0.f;
        for (int q = lane; q < na * nc; q += 32) {
            const int a = a0 + q / nc, c = c0 + q % nc;
            s += U[a * dom.M + c];
        }
        const double t = warp_sum(double(s));
        if (lane == 0) blk[k * K + l] = t;
    }
}

// L_dom = sum_{k<l} (Kxx + Kyy - 2Kxy) / (K(K-1)/2)   (algorithms.py:110-116, :82-88) from the u block sums;
// executed by one warp, one lane per domain pair.  An empty chunk gives 0 * inf = NaN, as torch's
// mean() over an empty tensor does.
__device__ float mmd_from_blocks(const double* __restrict__ blk, const DomainInfo& dom, int lane) {
    const int K = dom.K;
    if (K <= 1) return 0.f;
    const int npairs = K * (K - 1) / 2;
    double acc = 0.0;
    for (int pidx = lane; pidx < npairs; pidx += 32) {
        int k = 0, r = pidx;
        while (r >= K - 1 - k) { r -= K - 1 - k; ++k; }
        const int l = k + 1 + r;
        const float nk = float(dom.size(k)), nl = float(dom.size(l));
        const double rkk = double(1.0f / (nk * nk)), rll = double(1.0f / (nl * nl)), rkl = double(1.0f / (nk * nl));
        acc += blk[k * K + k] * rkk + blk[l * K + l] * rll - 2.0 * (blk[k * K + l] * rkl);
    }
    acc = warp_sum(acc);
    return float(acc) / float(npairs);
}

// ------------------------------------------------------------------------------------------------
struct FwdParams {
    const float* partial;
    const int* slot_count;   // [B] written by the Gram kernel
    int nslots;
    int B;
    long long P;
    int n, K;
    float margin, eps;
    float* losses;
    float* gram;
    float* rowstat;
    void* scratch;      // global fallback for EpiMem
    int in_smem;
    const float* vd;    // pre-reduced mode (after gram_reduce_kernel): [B][124] vectors and [B][2] row statistics
    const float* statd;
    int pre_reduced;
    long long* dbg;     // optional phase timestamps (tools/epilogue_phases.py); nullptr in production
};

template <bool kSmem>
__global__ void __launch_bounds__(kEpiThreads, 1) whiten_epilogue_fwd_kernel(FwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B;
    const DomainInfo dom = make_domain(B, p.n, p.K);
    const EpiMem mem = resolve_mem<kSmem>(smem_raw, p.scratch, B, dom.M);
    const float denom = float(p.P - 1);
    __shared__ IndexTables tab;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // dependents wait for our completion themselves
    build_index_tables(tab, tid, kEpiThreads);
    __syncthreads();
    WTPSE_STAMP(0);

    // A. one warp per sample: slots -> Gram entries -> gram, v, off_b, diag_b.
    //    The first three slots are loaded speculatively, together with the slot count, so the warp pays
    //    one L2 round trip instead of a chain of them (unused slots hold garbage that is never added).
    if (p.pre_reduced) {
        // the per-sample work was done by gram_reduce_kernel (one CTA per sample): just stage its results
        asm volatile("griddepcontrol.wait;" ::: "memory");
        // rows are 124 floats in both places: copy them as float4, four independent loads in flight per thread
        {
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
        }
        for (int idx = tid; idx < 2 * B; idx += kEpiThreads) mem.stat[idx] = __ldcg(p.statd + idx);
    } else
    for (int b = warp; b < B; b += kEpiWarps) {
        const float* src = p.partial + ((long long)b * p.nslots) * kTri;
        float v0[5], v1[5], v2[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = lane + 32 * q;
            const bool ok = e < kTri;
            v0[q] = ok ? __ldcg(src + e) : 0.f;
            v1[q] = (ok && p.nslots > 1) ? __ldcg(src + kTri + e) : 0.f;
            v2[q] = (ok && p.nslots > 2) ? __ldcg(src + 2 * kTri + e) : 0.f;
        }
        const int cnt = __ldcg(p.slot_count + b);
        float off = 0.f, dg = 0.f;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int e = lane + 32 * q;
            if (e < kTri) {
                float s = v0[q];
                if (cnt > 1) s += v1[q];
                if (cnt > 2) s += v2[q];
                for (int sl = 3; sl < cnt; ++sl) s += __ldcg(src + (long long)sl * kTri + e);
                const int ij = tab.tri[e], i = ij >> 4, j = ij & 15;
                s = s / denom;                                   // .div(HW - 1), algorithms.py:1283
                if (i == j) {
                    s += p.eps;                                  // + eps * eye
                    dg += fabsf(s - 1.0f);                       // |f_cor_masked_diag - I|, :1297
                    p.gram[b * 256 + i * kC + i] = s;
                } else {
                    off += fabsf(s);                             // |f_cor_masked|, :1289
                    p.gram[b * 256 + i * kC + j] = s;
                    p.gram[b * 256 + j * kC + i] = s;
                    if (b < dom.M) mem.v[size_t(b) * kVStride + off_idx(i, j)] = s;
                }
            }
        }
        off = warp_sum(off) - p.margin;
        dg = warp_sum(dg) - p.margin;
        if (lane == 0) {
            mem.stat[b * 2 + 0] = off;
            mem.stat[b * 2 + 1] = dg;
            p.rowstat[b * 2 + 0] = off;
            p.rowstat[b * 2 + 1] = dg;
        }
    }
    __syncthreads();
    WTPSE_STAMP(1);

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
    WTPSE_STAMP(2);

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
    WTPSE_STAMP(3);

    // D. L_dom
    if (warp == 0) {
        const float pen = mmd_from_blocks(mem.blk, dom, lane);
        if (lane == 0) p.losses[2] = pen;
    }
    WTPSE_STAMP(4);
}

// ------------------------------------------------------------------------------------------------
// Forward stage 2a for the round-robin Gram schedule: every CTA of the Gram kernel holds a partial for
// (almost) every sample, so the slot reduction is done by one CTA PER SAMPLE, in a fixed order, together with
// everything else that is per-sample (Gram entries, 120-d vector, off_b, diag_b).  The single-CTA epilogue
// that follows only does the MMD and the final sums.
constexpr int kReduceParts = 4;
constexpr int kReduceThreads = kTri * kReduceParts;   // 544

struct ReduceParams {
    const float* partial;     // [B][G][136]
    const int* slot_count;    // contiguous schedule: slots used per sample (nullptr for round-robin)
    long long tps, T, G;      // G = slots per sample
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
    // CTA k of the Gram kernel touched this sample iff its first tile at or after the sample start lies inside it
    const long long tb = (long long)b * p.tps, te = tb + p.tps;
    const int rb = int(tb % G);
    const int cnt = p.slot_count ? __ldcg(p.slot_count + b) : 0;
    const float* src = p.partial + ((long long)b * G) * kTri + e;
    float s = 0.f;
    for (int k = k0; k < k1; k += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int kk = k + u;
            const int r = kk - rb + (kk < rb ? G : 0);
            const bool valid = p.slot_count ? (kk < cnt) : (tb + r < te);
            v[u] = (kk < k1 && valid) ? __ldcg(src + (long long)kk * kTri) : 0.f;
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

// ------------------------------------------------------------------------------------------------
struct BwdParams {
    const float* gram;
    const float* rowstat;
    const float *g_off, *g_diag, *g_dom;
    int B;
    long long P;
    int n, K;
    float* mmat;
    void* scratch;
    int in_smem;
    long long* dbg;
};

template <bool kSmem>
__global__ void __launch_bounds__(kEpiThreads, 1) whiten_epilogue_bwd_kernel(BwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int B = p.B;
    const DomainInfo dom = make_domain(B, p.n, p.K);
    const int M = dom.M;
    const EpiMem mem = resolve_mem<kSmem>(smem_raw, p.scratch, B, M);
    const float g_off = p.g_off ? __ldcg(p.g_off) : 0.f;
    const float g_diag = p.g_diag ? __ldcg(p.g_diag) : 0.f;
    const float g_dom = p.g_dom ? __ldcg(p.g_dom) : 0.f;
    const bool need_dom = (M > 0) && (g_dom != 0.f);      // block-uniform
    __shared__ IndexTables tab;
    build_index_tables(tab, tid, kEpiThreads);
    __syncthreads();
    WTPSE_STAMP(0);

    if (need_dom) {
        // 1. upper-triangle vectors of the samples that enter the MMD
        for (int idx = tid; idx < M * kOff; idx += kEpiThreads) {
            const int b = idx / kOff, o = idx - b * kOff;
            const int ij = tab.off[o];
            mem.v[size_t(b) * kVStride + o] = __ldcg(p.gram + b * 256 + (ij >> 4) * kC + (ij & 15));
        }
        __syncthreads();
        WTPSE_STAMP(1);
        // 2. symmetric coefficient matrix (stored in U's place)
        float* coef = mem.U;
        pairwise_upper(mem.v, M, tid, [coef, M, &dom](int a, int c, float D) {
            const float E = expf(-D);
            coef[a * M + c] = mmd_coefficient(dom, a, c, E);
            coef[c * M + a] = mmd_coefficient(dom, c, a, E);
        });
        for (int a = tid; a < M; a += kEpiThreads) coef[a * M + a] = 0.f;
        __syncthreads();
    }
    WTPSE_STAMP(2);

    // 3. M_b[i][j] = (S_b + S_b^T)[i][j] / (P - 1)
    const float denom = float(p.P - 1);
    const float w_off = g_off / (float(B) * float(kOff));
    const float w_diag = g_diag / (float(B) * float(kC));
    for (int idx = tid; idx < B * kTri; idx += kEpiThreads) {
        const int b = idx / kTri, e = idx - b * kTri;
        const int ij = tab.tri[e], i = ij >> 4, j = ij & 15;
        const float g = __ldcg(p.gram + b * 256 + i * kC + j);
        float dom_grad = 0.f;
        if (need_dom && b < M && i != j) dom_grad = g_dom * mmd_grad_entry(mem.v, mem.U + size_t(b) * M, M, b, off_idx(i, j));
        const float m = backward_matrix_entry(i, j, g, __ldcg(p.rowstat + b * 2 + 0), __ldcg(p.rowstat + b * 2 + 1), w_off,
                                              w_diag, dom_grad, denom);
        p.mmat[b * 256 + i * kC + j] = m;
        p.mmat[b * 256 + j * kC + i] = m;
    }
    __syncthreads();
    WTPSE_STAMP(3);
}

// ------------------------------------------------------------------------------------------------
// Backward stage 1, multi-CTA: one CTA per sample derives M_b (same arithmetic as the single-CTA
// kernel above, parallel over samples).  It triggers its dependents at once (griddepcontrol), so the
// apply kernel launched behind it with programmatic stream serialisation starts streaming z while the
// matrices are still being computed, and only waits right before it reads the first M_b.
constexpr int kMmatThreads = 256;

__global__ void __launch_bounds__(kMmatThreads) whiten_mmat_kernel(BwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x;
    const int B = p.B, b = blockIdx.x;
    const DomainInfo dom = make_domain(B, p.n, p.K);
    const int M = dom.M;
    float* vbuf = reinterpret_cast<float*>(smem_raw);                 // [M][124]
    float* coefrow = vbuf + round4(size_t(M) * kVStride);            // [M]
    __shared__ IndexTables tab;
    build_index_tables(tab, tid, kMmatThreads);
    asm volatile("griddepcontrol.wait;" ::: "memory");             // inputs may come from the kernel right before us
    const float g_off = p.g_off ? __ldcg(p.g_off) : 0.f;
    const float g_diag = p.g_diag ? __ldcg(p.g_diag) : 0.f;
    const float g_dom = p.g_dom ? __ldcg(p.g_dom) : 0.f;
    const bool in_mmd = (M > 0) && (g_dom != 0.f) && (b < M);         // block-uniform
    __syncthreads();
    if (in_mmd) {
        // stage the vectors of all MMD samples: batches of 8 independent gathers per thread (one DRAM/L2 round trip
        // per batch instead of one per element)
        for (int base = 0; base < M * kOff; base += 8 * kMmatThreads) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * kMmatThreads + tid;
                if (idx < M * kOff) {
                    const int c = idx / kOff, o = idx - c * kOff;
                    const int ij = tab.off[o];
                    v[u] = __ldcg(p.gram + c * 256 + (ij >> 4) * kC + (ij & 15));
                } else {
                    v[u] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * kMmatThreads + tid;
                if (idx < M * kOff) {
                    const int c = idx / kOff, o = idx - c * kOff;
                    vbuf[size_t(c) * kVStride + o] = v[u];
                }
            }
        }
        __syncthreads();
        for (int c = tid; c < M; c += kMmatThreads) {
            const float D = mmd_distance(vbuf + size_t(b) * kVStride, vbuf + size_t(c) * kVStride);
            coefrow[c] = mmd_coefficient(dom, b, c, expf(-D));
        }
        __syncthreads();
    }
    const float denom = float(p.P - 1);
    const float w_off = g_off / (float(B) * float(kOff));
    const float w_diag = g_diag / (float(B) * float(kC));
    if (tid < kTri) {
        const int ij = tab.tri[tid], i = ij >> 4, j = ij & 15;
        const float g = __ldcg(p.gram + b * 256 + i * kC + j);
        float dom_grad = 0.f;
        if (in_mmd && i != j) dom_grad = g_dom * mmd_grad_entry(vbuf, coefrow, M, b, off_idx(i, j));
        const float m = backward_matrix_entry(i, j, g, __ldcg(p.rowstat + b * 2 + 0), __ldcg(p.rowstat + b * 2 + 1), w_off,
                                              w_diag, dom_grad, denom);
        p.mmat[b * 256 + i * kC + j] = m;
        p.mmat[b * 256 + j * kC + i] = m;
    }
}

// ---- standalone compute_MMD.forward on a B x 120 input (algorithms.py:102-121) -------------------
struct MmdParams {
    const float* v32;
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
        mem.v[size_t(b) * kVStride + o] = __ldcg(p.v32 + idx);
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

int g_epilogue_repeat = 1;             // diagnostics: launch the epilogues this many times back to back
long long* g_epilogue_dbg = nullptr;   // set through wtpse_debug_set_stamp_buffer (16 x int64 device buffer)

size_t epilogue_scratch_bytes(int B, int K) { return epi_mem_bytes(B, B, K); }

cudaError_t launch_gram_reduce(const float* partial, const int* slot_count, const GramPlan& g, int B, long long P, int n_per_domain, int n_domains,
                               float margin, float eps, float* gram, float* rowstat, float* vd, float* statd,
                               cudaStream_t stream) {
    ReduceParams p;
    p.partial = partial; p.tps = g.tiles_per_sample; p.T = g.T; p.B = B; p.P = P;
    p.G = g.nslots;                                   // == grid size for the round-robin schedule
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

cudaError_t launch_whiten_epilogue_fwd(const float* partial, const int* slot_count, int nslots, int B, long long P,
                                       int n_per_domain, int n_domains, float margin, float eps, float* losses,
                                       float* gram, float* rowstat, void* scratch, cudaStream_t stream, const float* vd,
                                       const float* statd, bool pre_reduced) {
    FwdParams p;
    p.partial = partial; p.slot_count = slot_count; p.nslots = nslots;
    p.B = B; p.P = P; p.n = n_per_domain; p.K = n_domains; p.margin = margin; p.eps = eps;
    p.losses = losses; p.gram = gram; p.rowstat = rowstat;
    p.scratch = scratch; p.dbg = g_epilogue_dbg;
    p.vd = vd; p.statd = statd; p.pre_reduced = pre_reduced ? 1 : 0;
    cudaError_t e;
    const size_t dyn = epi_smem((const void*)whiten_epilogue_fwd_kernel<true>, B, n_per_domain, n_domains, &p.in_smem, &e);
    if (e != cudaSuccess) return e;
    if (pre_reduced) {
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
    for (int r = 0; r < g_epilogue_repeat; ++r) {
        if (p.in_smem) whiten_epilogue_fwd_kernel<true><<<1, kEpiThreads, dyn, stream>>>(p);
        else whiten_epilogue_fwd_kernel<false><<<1, kEpiThreads, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

cudaError_t launch_whiten_epilogue_bwd(const float* gram, const float* rowstat, const float* g_off, const float* g_diag,
                                       const float* g_dom, int B, long long P, int n_per_domain, int n_domains,
                                       float* mmat, void* scratch, cudaStream_t stream) {
    BwdParams p;
    p.gram = gram; p.rowstat = rowstat; p.g_off = g_off; p.g_diag = g_diag; p.g_dom = g_dom;
    p.B = B; p.P = P; p.n = n_per_domain; p.K = n_domains; p.mmat = mmat;
    p.scratch = scratch; p.dbg = g_epilogue_dbg ? g_epilogue_dbg + 8 : nullptr;
    cudaError_t e;
    const size_t dyn = epi_smem((const void*)whiten_epilogue_bwd_kernel<true>, B, n_per_domain, n_domains, &p.in_smem, &e);
    if (e != cudaSuccess) return e;
    for (int r = 0; r < g_epilogue_repeat; ++r) {
        if (p.in_smem) whiten_epilogue_bwd_kernel<true><<<1, kEpiThreads, dyn, stream>>>(p);
        else whiten_epilogue_bwd_kernel<false><<<1, kEpiThreads, 0, stream>>>(p);
    }
    return cudaGetLastError();
}

bool mmat_multi_cta_ok(int B, int n_per_domain, int n_domains) {
    const size_t M = size_t(mmd_samples(B, n_per_domain, n_domains));
    return (round4(M * kVStride) + round4(M)) * sizeof(float) <= kEpiSmemCap;
}

cudaError_t launch_whiten_mmat(const float* gram, const float* rowstat, const float* g_off, const float* g_diag,
                               const float* g_dom, int B, long long P, int n_per_domain, int n_domains, float* mmat,
                               cudaStream_t stream) {
    BwdParams p;
    p.gram = gram; p.rowstat = rowstat; p.g_off = g_off; p.g_diag = g_diag; p.g_dom = g_dom;
    p.B = B; p.P = P; p.n = n_per_domain; p.K = n_domains; p.mmat = mmat;
    p.scratch = nullptr; p.in_smem = 1; p.dbg = nullptr;
    const size_t M = size_t(mmd_samples(B, n_per_domain, n_domains));
    const size_t dyn = (round4(M * kVStride) + round4(M)) * sizeof(float);
    if (dyn > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(whiten_mmat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dyn));
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(unsigned(B));
    cfg.blockDim = dim3(kMmatThreads);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, whiten_mmat_kernel, p);
}

cudaError_t launch_mmd(const float* v, const float* gout, int B, int n_per_domain, int n_domains, float* loss, float* dv,
                       void* scratch, cudaStream_t stream) {
    MmdParams p;
    p.v32 = v; p.gout = gout; p.B = B; p.n = n_per_domain; p.K = n_domains; p.loss = loss; p.dv = dv;
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

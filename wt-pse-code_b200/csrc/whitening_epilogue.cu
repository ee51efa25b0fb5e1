// Forward stage 2 and backward stage 1: everything between the Gram and the scalars, plus the
// standalone MMD (compute_MMD.forward on a B x 120 input).
//
// Forward  (algorithms.py:1283-1307 after the bmm, compute_MMD.forward algorithms.py:102-121):
//   partial Gram slots -> f_cor = G/(P-1) + eps*I -> off_b, diag_b -> L_off, L_diag
//   -> 120-d upper-triangle vectors -> pairwise exp(-D) -> L_dom.
// Backward: upstream grads + saved Gram -> per-sample symmetric 16x16
//   coefficient matrix M_b = (S_b + S_b^T)/(P-1) that the apply kernel multiplies into z.
//
// The work is O(B*136*slots + B^2*120): microseconds.  It runs as ONE CTA so that every reduction
// has a fixed order (bit-reproducible run to run, no float atomics) and is evaluated in float64,
// which keeps the MMD's Kxx + Kyy - 2Kxy cancellation closer to the
// reference's own float64 result than the reference's float32 path is.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kEpiThreads = 1024;
constexpr int kEpiWarps = kEpiThreads / 32;
constexpr size_t kEpiSmemCap = 200 * 1024;

__device__ __forceinline__ void tri_decode(int e, int& i, int& j) {
    int r = 0, len = kC;
    while (e >= len) { e -= len; --len; ++r; }
    i = r;
    j = r + e;
}
// index of (i,j), i<j, in torch.triu_indices(16,16,1) order (algorithms.py:1305)
__device__ __forceinline__ int off_idx(int i, int j) { return tri_idx(i, j) - (i + 1); }

// torch.clamp(x, min=0): NaN propagates
__device__ __forceinline__ double clamp0(double x) { return (x < 0.0) ? 0.0 : x; }
// clamp_min_(1e-30): NaN propagates
__device__ __forceinline__ double clamp_tiny(double x) { return (x < 1e-30) ? 1e-30 : x; }

__device__ __forceinline__ int chunk_lo(int k, int n, int B) {
    const long long v = (long long)n * k;
    return int(v < B ? v : B);
}

// features[k] = inputs[n*k : n*(k+1)] with python slice truncation (algorithms.py:107)
struct DomainInfo {
    int M;  // samples that enter the MMD: min(B, K*n), 0 when K <= 1
    int K, n, B;
    __device__ int domain_of(int a) const { return n > 0 ? a / n : 0; }
    __device__ int size(int k) const { return chunk_lo(k + 1, n, B) - chunk_lo(k, n, B); }
};

__device__ __forceinline__ DomainInfo make_domain(int B, int n, int K) {
    DomainInfo dom{0, K, n, B};
    const long long m = (long long)K * n;
    dom.M = K > 1 ? int(m < B ? m : B) : 0;
    return dom;
}

// E[a][c] = exp(-max(|v_a - v_c|^2, 1e-30)) for a,c < M ; v has row stride 120 (shared or global)
__device__ void pairwise_kernel(const double* __restrict__ v, double* __restrict__ E, int M, int warp, int lane) {
    for (int a = warp; a < M; a += kEpiWarps) {
        double va[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = lane + 32 * q;
            va[q] = e < kOff ? v[a * kOff + e] : 0.0;
        }
        for (int c = a; c < M; ++c) {
            double d = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int e = lane + 32 * q;
                const double diff = e < kOff ? va[q] - v[c * kOff + e] : 0.0;
                d = fma(diff, diff, d);
            }
            d = warp_sum(d);
            if (lane == 0) {
                const double val = exp(-clamp_tiny(d));
                E[a * M + c] = val;
                E[c * M + a] = val;
            }
        }
    }
}

// per-domain-pair sums of E, one warp per (k <= l) block, fixed lane-strided order
__device__ void domain_block_sums(const double* __restrict__ E, const DomainInfo& dom, double* __restrict__ blk,
                                  int first_warp, int nwarps, int warp, int lane) {
    const int K = dom.K;
    for (int pr = warp - first_warp; pr < K * K; pr += nwarps) {
        const int k = pr / K, l = pr - k * K;
        if (l < k) continue;
        const int a0 = chunk_lo(k, dom.n, dom.B), a1 = chunk_lo(k + 1, dom.n, dom.B);
        const int c0 = chunk_lo(l, dom.n, dom.B), c1 = chunk_lo(l + 1, dom.n, dom.B);
        const int na = a1 - a0, nc = c1 - c0;
        double s = 0.0;
        for (int q = lane; q < na * nc; q += 32) {
            const int a = a0 + q / nc, c = c0 + q % nc;
            s += E[a * dom.M + c];
        }
        s = warp_sum(s);
        if (lane == 0) blk[k * K + l] = s;
    }
}

// L_dom = sum_{k<l} (Kxx + Kyy - 2Kxy) / (K(K-1)/2)      algorithms.py:110-116, :82-88
__device__ double mmd_from_blocks(const double* __restrict__ blk, const DomainInfo& dom) {
    double pen = 0.0;
    const int K = dom.K;
    if (K > 1) {
        for (int k = 0; k < K; ++k)
            for (int l = k + 1; l < K; ++l) {
                const int sk = dom.size(k), sl = dom.size(l);
                const double nk = double(sk), nl = double(sl);
                // empty chunk: 0/0 = NaN, as torch's mean() over an empty tensor
                const double kxx = (sk > 0 ? blk[k * K + k] : 0.0) / (nk * nk);
                const double kyy = (sl > 0 ? blk[l * K + l] : 0.0) / (nl * nl);
                const double kxy = ((sk > 0 && sl > 0) ? blk[k * K + l] : 0.0) / (nk * nl);
                pen += kxx + kyy - 2.0 * kxy;
            }
        pen /= double(K) * double(K - 1) / 2.0;
    }
    return pen;
}

// coef_ac = dL/dD_ac + dL/dD_ca  (zero on the diagonal, where clamp_min_(1e-30) is active)
__device__ void mmd_coefficients(const double* __restrict__ E, const DomainInfo& dom, double* __restrict__ coef, int tid) {
    const int M = dom.M;
    const double npairs = double(dom.K) * double(dom.K - 1) / 2.0;
    for (int idx = tid; idx < M * M; idx += kEpiThreads) {
        const int a = idx / M, c = idx - a * M;
        const int ka = dom.domain_of(a), kc = dom.domain_of(c);
        double w;
        if (ka == kc) {
            const double nk = double(dom.size(ka));
            w = -2.0 * double(dom.K - 1) / (nk * nk);
        } else {
            w = 2.0 / (double(dom.size(ka)) * double(dom.size(kc)));
        }
        coef[idx] = (a == c) ? 0.0 : E[idx] * w / npairs;
    }
}

// d L_dom / d v_b[o] = 2 * sum_c coef_bc (v_b[o] - v_c[o])
__device__ __forceinline__ double mmd_grad_entry(const double* __restrict__ v, const double* __restrict__ coef, int M,
                                                 int b, int o) {
    const double vb = v[b * kOff + o];
    double acc = 0.0;
    for (int c = 0; c < M; ++c) acc = fma(coef[b * M + c], vb - v[c * kOff + o], acc);
    return 2.0 * acc;
}

// ------------------------------------------------------------------------------------------------
struct FwdParams {
    const float* partial;
    int tma;
    long long tps, T, G;
    int nslots;
    int B;
    long long P;
    int n, K;
    float margin, eps;
    float* losses;
    float* gram;
    float* rowstat;
    double *gd, *stat, *E, *blk, *vd;
    int v_in_smem;
};

__global__ void __launch_bounds__(kEpiThreads, 1) whiten_epilogue_fwd_kernel(FwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* vs = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B;
    const DomainInfo dom = make_domain(B, p.n, p.K);
    const double inv = 1.0 / double(p.P - 1);

    // 1. slots -> scaled Gram (+eps on the diagonal)
    for (int idx = tid; idx < B * kTri; idx += kEpiThreads) {
        const int b = idx / kTri, e = idx - b * kTri;
        int cnt = p.nslots;
        if (p.tma) {
            const long long first = part_owner((long long)b * p.tps, p.T, p.G);
            const long long last = part_owner((long long)(b + 1) * p.tps - 1, p.T, p.G);
            cnt = int(last - first + 1);
        }
        double s = 0.0;
        const float* src = p.partial + ((long long)b * p.nslots) * kTri + e;
        for (int sl = 0; sl < cnt; ++sl) s += double(src[(long long)sl * kTri]);
        int i, j;
        tri_decode(e, i, j);
        s *= inv;
        if (i == j) s += double(p.eps);
        p.gd[idx] = s;
        p.gram[b * 256 + i * kC + j] = float(s);
        p.gram[b * 256 + j * kC + i] = float(s);
        if (i != j) {
            p.vd[b * kOff + off_idx(i, j)] = s;
            if (p.v_in_smem && b < dom.M) vs[b * kOff + off_idx(i, j)] = s;
        }
    }
    __syncthreads();

    // 2. per-sample sums (one warp per sample)
    for (int b = warp; b < B; b += kEpiWarps) {
        double off = 0.0, dg = 0.0;
        for (int e = lane; e < kTri; e += 32) {
            int i, j;
            tri_decode(e, i, j);
            const double g = p.gd[b * kTri + e];
            if (i == j) dg += fabs(g - 1.0);
            else off += fabs(g);
        }
        off = warp_sum(off);
        dg = warp_sum(dg);
        if (lane == 0) {
            off -= double(p.margin);
            dg -= double(p.margin);
            p.stat[b * 4 + 0] = off;
            p.stat[b * 4 + 1] = dg;
            p.rowstat[b * 2 + 0] = float(off);
            p.rowstat[b * 2 + 1] = float(dg);
        }
    }
    // 3. pairwise gaussian kernel values (independent of step 2)
    pairwise_kernel(p.v_in_smem ? vs : p.vd, p.E, dom.M, warp, lane);
    __syncthreads();

    // 4. instance terms (warp 0), per-domain-pair block sums of E (warps 1..)
    if (warp == 0) {
        double so = 0.0, sd = 0.0;
        for (int b = lane; b < B; b += 32) {
            so += clamp0(p.stat[b * 4 + 0] / double(kOff));
            sd += clamp0(p.stat[b * 4 + 1] / double(kC));
        }
        so = warp_sum(so) / double(B);
        sd = warp_sum(sd) / double(B);
        if (lane == 0) {
            p.losses[0] = float(so);
            p.losses[1] = float(sd);
            p.losses[3] = float(so + sd);
        }
    } else if (dom.M > 0) {
        domain_block_sums(p.E, dom, p.blk, 1, kEpiWarps - 1, warp, lane);
    }
    __syncthreads();

    // 5. L_dom
    if (tid == 0) p.losses[2] = float(mmd_from_blocks(p.blk, dom));
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
    double *E, *vd, *coef;
    int v_in_smem;
};

__global__ void __launch_bounds__(kEpiThreads, 1) whiten_epilogue_bwd_kernel(BwdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int B = p.B;
    const DomainInfo dom = make_domain(B, p.n, p.K);
    const int M = dom.M;
    const double g_off = p.g_off ? double(*p.g_off) : 0.0;
    const double g_diag = p.g_diag ? double(*p.g_diag) : 0.0;
    const double g_dom = p.g_dom ? double(*p.g_dom) : 0.0;
    double* v = p.v_in_smem ? reinterpret_cast<double*>(smem_raw) : p.vd;

    // 1. upper-triangle vectors of the samples that enter the MMD
    for (int idx = tid; idx < M * kOff; idx += kEpiThreads) {
        const int b = idx / kOff, e = idx - b * kOff;
        int i = 0, r = e, len = kC - 1;
        while (r >= len) { r -= len; --len; ++i; }
        const int j = i + 1 + r;
        v[idx] = double(p.gram[b * 256 + i * kC + j]);
    }
    __syncthreads();
    // 2. E, then the symmetric coefficient matrix
    pairwise_kernel(v, p.E, M, warp, lane);
    __syncthreads();
    if (M > 0) mmd_coefficients(p.E, dom, p.coef, tid);
    __syncthreads();

    // 3. M_b[i][j]
    const double inv = 1.0 / double(p.P - 1);
    for (int idx = tid; idx < B * kTri; idx += kEpiThreads) {
        const int b = idx / kTri, e = idx - b * kTri;
        int i, j;
        tri_decode(e, i, j);
        const float g = p.gram[b * 256 + i * kC + j];
        double s;
        if (i == j) {
            const float d = g - 1.0f;   // f_cor_masked_diag - diagonal_matrix in fp32, algorithms.py:1297
            const double sgn = (d > 0.f) ? 1.0 : ((d < 0.f) ? -1.0 : 0.0);
            const bool act = (p.rowstat[b * 2 + 1] / float(kC)) >= 0.f;     // clamp(min=0) passes grad at x >= 0
            s = act ? g_diag * sgn / (double(B) * double(kC)) : 0.0;
            p.mmat[b * 256 + i * kC + i] = float(2.0 * s * inv);
        } else {
            const double sgn = (g > 0.f) ? 1.0 : ((g < 0.f) ? -1.0 : 0.0);
            const bool act = (p.rowstat[b * 2 + 0] / float(kOff)) >= 0.f;
            s = act ? g_off * sgn / (double(B) * double(kOff)) : 0.0;
            if (b < M && g_dom != 0.0) s += g_dom * mmd_grad_entry(v, p.coef, M, b, off_idx(i, j));
            const float m = float(s * inv);
            p.mmat[b * 256 + i * kC + j] = m;
            p.mmat[b * 256 + j * kC + i] = m;
        }
    }
}

// ---- standalone compute_MMD.forward on a B x 120 input (algorithms.py:102-121) -------------------
struct MmdParams {
    const float* v32;
    const float* gout;
    int B, n, K;
    float* loss;
    float* dv;
    double *vd, *E, *coef, *blk;
    int v_in_smem;
};

__device__ __forceinline__ double* mmd_stage_vectors(const MmdParams& p, const DomainInfo& dom, double* vs, int tid) {
    double* v = p.v_in_smem ? vs : p.vd;
    for (int idx = tid; idx < dom.M * kOff; idx += kEpiThreads) v[idx] = double(p.v32[idx]);
    return v;
}

__global__ void __launch_bounds__(kEpiThreads, 1) mmd_fwd_kernel(MmdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DomainInfo dom = make_domain(p.B, p.n, p.K);
    double* v = mmd_stage_vectors(p, dom, reinterpret_cast<double*>(smem_raw), tid);
    __syncthreads();
    pairwise_kernel(v, p.E, dom.M, warp, lane);
    __syncthreads();
    if (dom.M > 0) domain_block_sums(p.E, dom, p.blk, 0, kEpiWarps, warp, lane);
    __syncthreads();
    if (tid == 0) p.loss[0] = float(mmd_from_blocks(p.blk, dom));
}

__global__ void __launch_bounds__(kEpiThreads, 1) mmd_bwd_kernel(MmdParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const DomainInfo dom = make_domain(p.B, p.n, p.K);
    const int M = dom.M;
    double* v = mmd_stage_vectors(p, dom, reinterpret_cast<double*>(smem_raw), tid);
    __syncthreads();
    pairwise_kernel(v, p.E, M, warp, lane);
    __syncthreads();
    if (M > 0) mmd_coefficients(p.E, dom, p.coef, tid);
    __syncthreads();
    const double g = p.gout ? double(*p.gout) : 1.0;
    for (int idx = tid; idx < p.B * kOff; idx += kEpiThreads) {
        const int b = idx / kOff, e = idx - b * kOff;
        p.dv[idx] = (b < M) ? float(g * mmd_grad_entry(v, p.coef, M, b, e)) : 0.f;
    }
}

int mmd_samples(int B, int n, int K) {
    if (K <= 1) return 0;
    const long long m = (long long)n * K;
    return int(m < B ? m : B);
}

// dynamic shared memory for the double-precision vectors, or 0 when they stay in the global scratch
size_t vector_smem(const void* fn, int B, int n, int K, int* in_smem, cudaError_t* err) {
    const size_t bytes = size_t(mmd_samples(B, n, K)) * kOff * sizeof(double);
    *err = cudaSuccess;
    *in_smem = bytes <= kEpiSmemCap ? 1 : 0;
    if (!*in_smem) return 0;
    if (bytes > 48 * 1024) *err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    return bytes;
}

}  // namespace

static size_t blk_doubles(int K) { return size_t(K > 0 ? K : 1) * size_t(K > 0 ? K : 1); }

size_t epilogue_scratch_doubles(int B, int K) {
    return size_t(B) * kTri + size_t(B) * 4 + 2 * size_t(B) * B + blk_doubles(K) + size_t(B) * kOff;
}

EpilogueScratch carve_epilogue_scratch(double* base, int B, int K) {
    EpilogueScratch s;
    s.gd = base;
    s.stat = s.gd + size_t(B) * kTri;
    s.E = s.stat + size_t(B) * 4;
    s.coef = s.E + size_t(B) * B;
    s.blk = s.coef + size_t(B) * B;
    s.vd = s.blk + blk_doubles(K);
    return s;
}

cudaError_t launch_whiten_epilogue_fwd(const float* partial, const GramPlan& g, int B, long long P, int n_per_domain,
                                       int n_domains, float margin, float eps, float* losses, float* gram,
                                       float* rowstat, const EpilogueScratch& s, cudaStream_t stream) {
    FwdParams p;
    p.partial = partial;
    p.tma = g.tma ? 1 : 0;
    p.tps = g.tiles_per_sample; p.T = g.T; p.G = g.G; p.nslots = g.nslots;
    p.B = B; p.P = P; p.n = n_per_domain; p.K = n_domains; p.margin = margin; p.eps = eps;
    p.losses = losses; p.gram = gram; p.rowstat = rowstat;
    p.gd = s.gd; p.stat = s.stat; p.E = s.E; p.blk = s.blk; p.vd = s.vd;
    cudaError_t e;
    const size_t dyn = vector_smem((const void*)whiten_epilogue_fwd_kernel, B, n_per_domain, n_domains, &p.v_in_smem, &e);
    if (e != cudaSuccess) return e;
    whiten_epilogue_fwd_kernel<<<1, kEpiThreads, dyn, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_whiten_epilogue_bwd(const float* gram, const float* rowstat, const float* g_off, const float* g_diag,
                                       const float* g_dom, int B, long long P, int n_per_domain, int n_domains,
                                       float* mmat, const EpilogueScratch& s, cudaStream_t stream) {
    BwdParams p;
    p.gram = gram; p.rowstat = rowstat; p.g_off = g_off; p.g_diag = g_diag; p.g_dom = g_dom;
    p.B = B; p.P = P; p.n = n_per_domain; p.K = n_domains; p.mmat = mmat;
    p.E = s.E; p.vd = s.vd; p.coef = s.coef;
    cudaError_t e;
    const size_t dyn = vector_smem((const void*)whiten_epilogue_bwd_kernel, B, n_per_domain, n_domains, &p.v_in_smem, &e);
    if (e != cudaSuccess) return e;
    whiten_epilogue_bwd_kernel<<<1, kEpiThreads, dyn, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_mmd(const float* v, const float* gout, int B, int n_per_domain, int n_domains, float* loss, float* dv,
                       const EpilogueScratch& s, cudaStream_t stream) {
    MmdParams p;
    p.v32 = v; p.gout = gout; p.B = B; p.n = n_per_domain; p.K = n_domains; p.loss = loss; p.dv = dv;
    p.vd = s.vd; p.E = s.E; p.coef = s.coef; p.blk = s.blk;
    const void* fn = dv ? (const void*)mmd_bwd_kernel : (const void*)mmd_fwd_kernel;
    cudaError_t e;
    const size_t dyn = vector_smem(fn, B, n_per_domain, n_domains, &p.v_in_smem, &e);
    if (e != cudaSuccess) return e;
    if (dv) mmd_bwd_kernel<<<1, kEpiThreads, dyn, stream>>>(p);
    else mmd_fwd_kernel<<<1, kEpiThreads, dyn, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace wtpse

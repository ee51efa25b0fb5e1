// Device helpers shared by the epilogue kernels (whitening_epilogue.cu) and the fused backward
// (whitening_apply.cu): index tables, domain chunking (algorithms.py:107), NaN-propagating clamps and
// the MMD gradient coefficient.  One definition, so both backward paths produce identical bits.
#pragma once
#include "common.cuh"

namespace wtpse {

constexpr int kVStride = 124;   // 120 padded to a multiple of 4 with 124/4 odd: LDS.128 by row is conflict-free

// (i << 4) | j lookup tables for the packed upper-triangle index e (136, diagonal included) and for the
// strict-upper-triangle index o in torch.triu_indices(16,16,1) order (120, algorithms.py:1305).
// They live in SHARED memory: indexing a __constant__ table with a per-lane index serialises in the
// address-divergence unit (measured: 15-20k cycles per phase for this kernel).
struct IndexTables {
    unsigned char tri[kTri];
    unsigned char off[kOff];
};

// Fill the tables with `nthreads` cooperating threads (tid in [0, nthreads)); the caller synchronises.
__device__ __forceinline__ void build_index_tables(IndexTables& t, int tid, int nthreads) {
    for (int q = tid; q < kTri + kOff; q += nthreads) {
        if (q < kTri) {
            int e = q, r = 0, len = kC;
            while (e >= len) { e -= len; --len; ++r; }
            t.tri[q] = (unsigned char)((r << 4) | (r + e));
        } else {
            int e = q - kTri, r = 0, len = kC - 1;
            while (e >= len) { e -= len; --len; ++r; }
            t.off[q - kTri] = (unsigned char)((r << 4) | (r + 1 + e));
        }
    }
}

__device__ __forceinline__ int off_idx(int i, int j) { return tri_idx(i, j) - (i + 1); }

// torch.clamp(x, min=0): NaN propagates
__device__ __forceinline__ float clamp0(float x) { return (x < 0.f) ? 0.f : x; }
// clamp_min_(1e-30): NaN propagates
__device__ __forceinline__ float clamp_tiny(float x) { return (x < 1e-30f) ? 1e-30f : x; }

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

// dL/dD_ac + dL/dD_ca for the kernel value E_ac (zero on the diagonal, where clamp_min_(1e-30) is active)
__device__ __forceinline__ float mmd_coefficient(const DomainInfo& dom, int a, int c, float E) {
    if (a == c) return 0.f;
    const int ka = dom.domain_of(a), kc = dom.domain_of(c);
    const float npairs = float(dom.K) * float(dom.K - 1) * 0.5f;
    float w;
    if (ka == kc) {
        const float nk = float(dom.size(ka));
        w = -2.0f * float(dom.K - 1) / (nk * nk);
    } else {
        w = 2.0f / (float(dom.size(ka)) * float(dom.size(kc)));
    }
    return E * w / npairs;
}

// d L_dom / d v_b[o] = 2 * sum_c coef_row[c] (v_b[o] - v_c[o]),  coef_row = row b of the coefficient matrix
__device__ __forceinline__ float mmd_grad_entry(const float* __restrict__ v, const float* __restrict__ coef_row, int M,
                                                int b, int o) {
    const float vb = v[size_t(b) * kVStride + o];
    float a0 = 0.f, a1 = 0.f;
    int c = 0;
    for (; c + 1 < M; c += 2) {
        a0 = fmaf(coef_row[c], vb - v[size_t(c) * kVStride + o], a0);
        a1 = fmaf(coef_row[c + 1], vb - v[size_t(c + 1) * kVStride + o], a1);
    }
    if (c < M) a0 = fmaf(coef_row[c], vb - v[size_t(c) * kVStride + o], a0);
    return 2.0f * (a0 + a1);
}

// distance row: D(b, c) = max(sum_e (v_b[e] - v_c[e])^2, 1e-30); rows are kVStride floats, 16-byte aligned
__device__ __forceinline__ float mmd_distance(const float* __restrict__ vb, const float* __restrict__ vc) {
    const float4* x = reinterpret_cast<const float4*>(vb);
    const float4* y = reinterpret_cast<const float4*>(vc);
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll 5
    for (int e = 0; e < kOff / 4; ++e) {
        const float4 a = x[e], c = y[e];
        float d;
        d = a.x - c.x; p0 = fmaf(d, d, p0);
        d = a.y - c.y; p1 = fmaf(d, d, p1);
        d = a.z - c.z; p2 = fmaf(d, d, p2);
        d = a.w - c.w; p3 = fmaf(d, d, p3);
    }
    return clamp_tiny((p0 + p1) + (p2 + p3));
}

// (S_b + S_b^T)[i][j] / (P - 1) for one packed entry e = (i, j), i <= j  (SURVEY.md appendix A.2).
//   g        gram[b][i][j]
//   dom_grad g_dom * dL_dom/dv_b[(i,j)]  (0 when the sample is outside the MMD or i == j)
__device__ __forceinline__ float backward_matrix_entry(int i, int j, float g, float off_b, float diag_b, float w_off,
                                                       float w_diag, float dom_grad, float denom) {
    if (i == j) {
        const float d = g - 1.0f;   // f_cor_masked_diag - diagonal_matrix, algorithms.py:1297
        const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
        const bool act = (diag_b / float(kC)) >= 0.f;   // clamp(min=0) passes the gradient at x >= 0
        const float s = act ? w_diag * sgn : 0.f;
        return 2.0f * s / denom;
    }
    const float sgn = (g > 0.f) ? 1.f : ((g < 0.f) ? -1.f : 0.f);
    const bool act = (off_b / float(kOff)) >= 0.f;
    const float s = (act ? w_off * sgn : 0.f) + dom_grad;
    return s / denom;
}

}  // namespace wtpse

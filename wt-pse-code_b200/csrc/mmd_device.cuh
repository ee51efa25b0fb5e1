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

// Four consecutive entries o = 4*q .. 4*q+3 of the same gradient row at once: one LDS.128 per sample instead of four
// scalar loads.  Per entry the operations and their order are those of mmd_grad_entry (even samples into one
// accumulator, odd ones into the other): identical bits.
__device__ __forceinline__ float4 mmd_grad_entry4(const float* __restrict__ v, const float* __restrict__ coef_row, int M,
                                                  int b, int q) {
    const float4 vb = *reinterpret_cast<const float4*>(v + size_t(b) * kVStride + 4 * q);
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    int c = 0;
#pragma unroll 4
    for (; c + 1 < M; c += 2) {
        const float4 x0 = *reinterpret_cast<const float4*>(v + size_t(c) * kVStride + 4 * q);
        const float4 x1 = *reinterpret_cast<const float4*>(v + size_t(c + 1) * kVStride + 4 * q);
        const float k0 = coef_row[c], k1 = coef_row[c + 1];
        a0.x = fmaf(k0, vb.x - x0.x, a0.x); a0.y = fmaf(k0, vb.y - x0.y, a0.y);
        a0.z = fmaf(k0, vb.z - x0.z, a0.z); a0.w = fmaf(k0, vb.w - x0.w, a0.w);
        a1.x = fmaf(k1, vb.x - x1.x, a1.x); a1.y = fmaf(k1, vb.y - x1.y, a1.y);
        a1.z = fmaf(k1, vb.z - x1.z, a1.z); a1.w = fmaf(k1, vb.w - x1.w, a1.w);
    }
    if (c < M) {
        const float4 x0 = *reinterpret_cast<const float4*>(v + size_t(c) * kVStride + 4 * q);
        const float k0 = coef_row[c];
        a0.x = fmaf(k0, vb.x - x0.x, a0.x); a0.y = fmaf(k0, vb.y - x0.y, a0.y);
        a0.z = fmaf(k0, vb.z - x0.z, a0.z); a0.w = fmaf(k0, vb.w - x0.w, a0.w);
    }
    return make_float4(2.0f * (a0.x + a1.x), 2.0f * (a0.y + a1.y), 2.0f * (a0.z + a1.z), 2.0f * (a0.w + a1.w));
}

// The same entries for a BLOCK of kBlk consecutive samples b0 .. b0 + kBlk - 1 at one q: every v_c piece is loaded once and
// used kBlk times.  One thread per (b, q) re-reads all of v for every b -- 0.5 MB through one SM's shared memory, which is what
// bounded the forward tail's largest phase (2 us of its 4); here it is 36 instead of 80 bytes per four entries.  Per entry the
// operations and their order are those of mmd_grad_entry (even samples into one accumulator, odd ones into the other):
// identical bits.  Rows b >= M are skipped.
template <int kBlk>
__device__ __forceinline__ void mmd_grad_block4(const float* __restrict__ v, const float* __restrict__ coef, int M, int b0, int q,
                                                float* __restrict__ out) {
    float4 vb[kBlk], a0[kBlk], a1[kBlk];
    const float* crow[kBlk];
#pragma unroll
    for (int u = 0; u < kBlk; ++u) {
        const int b = b0 + u < M ? b0 + u : M - 1;
        vb[u] = *reinterpret_cast<const float4*>(v + size_t(b) * kVStride + 4 * q);
        a0[u] = a1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        crow[u] = coef + size_t(b) * M;
    }
    int c = 0;
#pragma unroll 2
    for (; c + 1 < M; c += 2) {
        const float4 x0 = *reinterpret_cast<const float4*>(v + size_t(c) * kVStride + 4 * q);
        const float4 x1 = *reinterpret_cast<const float4*>(v + size_t(c + 1) * kVStride + 4 * q);
#pragma unroll
        for (int u = 0; u < kBlk; ++u) {
            const float k0 = crow[u][c], k1 = crow[u][c + 1];
            a0[u].x = fmaf(k0, vb[u].x - x0.x, a0[u].x); a0[u].y = fmaf(k0, vb[u].y - x0.y, a0[u].y);
            a0[u].z = fmaf(k0, vb[u].z - x0.z, a0[u].z); a0[u].w = fmaf(k0, vb[u].w - x0.w, a0[u].w);
            a1[u].x = fmaf(k1, vb[u].x - x1.x, a1[u].x); a1[u].y = fmaf(k1, vb[u].y - x1.y, a1[u].y);
            a1[u].z = fmaf(k1, vb[u].z - x1.z, a1[u].z); a1[u].w = fmaf(k1, vb[u].w - x1.w, a1[u].w);
        }
    }
    if (c < M) {
        const float4 x0 = *reinterpret_cast<const float4*>(v + size_t(c) * kVStride + 4 * q);
#pragma unroll
        for (int u = 0; u < kBlk; ++u) {
            const float k0 = crow[u][c];
            a0[u].x = fmaf(k0, vb[u].x - x0.x, a0[u].x); a0[u].y = fmaf(k0, vb[u].y - x0.y, a0[u].y);
            a0[u].z = fmaf(k0, vb[u].z - x0.z, a0[u].z); a0[u].w = fmaf(k0, vb[u].w - x0.w, a0[u].w);
        }
    }
#pragma unroll
    for (int u = 0; u < kBlk; ++u)
        if (b0 + u < M)
            *reinterpret_cast<float4*>(out + size_t(b0 + u) * kOff + 4 * q) =
                make_float4(2.0f * (a0[u].x + a1[u].x), 2.0f * (a0[u].y + a1[u].y), 2.0f * (a0[u].z + a1[u].z), 2.0f * (a0[u].w + a1[u].w));
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

// ------------------------------------------------------------------------------------------------
// Working set of the MMD part (single CTA): v [M][124] f32 | U [M][M] f32 | stat [B][2] f32 | blk [K*K] f64.
// Shared by the stand-alone epilogue kernels (whitening_epilogue.cu) and the in-kernel tail of the Gram kernels
// (whitening_tail.cuh): one definition, identical bits.
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

// D(a, c) = max(sum_e (v_a[e] - v_c[e])^2, 1e-30) for all pairs a < c < M; emit(a, c, D) is expected to fill both (a, c)
// and (c, a).  nthreads cooperate.
//
// Register-blocked: one thread owns kPairRows rows a0 .. (a0 a multiple of kPairRows) against ONE column c, so a step of four
// components costs kPairRows + 1 LDS.128 (broadcast rows + one column) instead of two per pair, and carries 4 kPairRows
// independent FMA chains -- the one-pair-per-lane version ran at 0.35 instructions per cycle and scheduler with the 6-7
// warps a tail has (3.3 us for 435 pairs; this one: tools/tail_phases.py).  Consecutive lanes walk c: the row loads
// broadcast, the column loads hit distinct banks (row stride 124 floats).  Per pair the operations and their order are
// exactly mmd_distance's, so D -- and everything derived from it -- keeps its bits.
constexpr int kPairRows = 3;                                   // rows per task: M = 30 gives 155 tasks -- one round for 192 or 224 threads
__device__ __forceinline__ int pair_tasks(int M) {            // sum over row blocks of the columns c > a0
    int n = 0;
    for (int a0 = 0; a0 + 1 < M; a0 += kPairRows) n += M - 1 - a0;
    return n;
}

template <typename F>
__device__ __forceinline__ void pairwise_upper_n(const float* __restrict__ v, int M, int tid, int nthreads, F&& emit) {
    const int ntasks = pair_tasks(M);
    for (int t = tid; t < ntasks; t += nthreads) {
        int a0 = 0, r = t;
        while (r >= M - 1 - a0) { r -= M - 1 - a0; a0 += kPairRows; }
        const int c = a0 + 1 + r;
        const float4* xc = reinterpret_cast<const float4*>(v + size_t(c) * kVStride);
        const float4* xa[kPairRows];
#pragma unroll
        for (int u = 0; u < kPairRows; ++u) xa[u] = reinterpret_cast<const float4*>(v + size_t(a0 + u < M ? a0 + u : M - 1) * kVStride);
        float p[kPairRows][4];
#pragma unroll
        for (int u = 0; u < kPairRows; ++u)
#pragma unroll
            for (int w = 0; w < 4; ++w) p[u][w] = 0.f;
#pragma unroll 3
        for (int e = 0; e < kOff / 4; ++e) {
            const float4 y = xc[e];
#pragma unroll
            for (int u = 0; u < kPairRows; ++u) {
                const float4 x = xa[u][e];
                float d;
                d = x.x - y.x; p[u][0] = fmaf(d, d, p[u][0]);
                d = x.y - y.y; p[u][1] = fmaf(d, d, p[u][1]);
                d = x.z - y.z; p[u][2] = fmaf(d, d, p[u][2]);
                d = x.w - y.w; p[u][3] = fmaf(d, d, p[u][3]);
            }
        }
#pragma unroll
        for (int u = 0; u < kPairRows; ++u)
            if (a0 + u < c) emit(a0 + u, c, clamp_tiny((p[u][0] + p[u][1]) + (p[u][2] + p[u][3])));
    }
}

// per-domain-pair sums of u = E - 1: one warp per (k <= l) block.  Lane j sums column c0 + j (+ 32, ...) of the block top to
// bottom in fp32 -- a fixed order with no index arithmetic in the loop -- and the lanes are combined in float64 (the
// result does not depend on how many warps share the blocks).
__device__ inline void domain_block_sums(const float* __restrict__ U, const DomainInfo& dom, double* __restrict__ blk,
                                         int first_warp, int nwarps, int warp, int lane) {
    const int K = dom.K;
    // the K (K + 1) / 2 blocks with k <= l, dealt evenly to the warps (row-major over the upper triangle)
    for (int pr = warp - first_warp; pr < K * (K + 1) / 2; pr += nwarps) {
        int k = 0, l = pr;
        while (l >= K - k) { l -= K - k; ++k; }
        l += k;
        const int a0 = chunk_lo(k, dom.n, dom.B), a1 = chunk_lo(k + 1, dom.n, dom.B);
        const int c0 = chunk_lo(l, dom.n, dom.B), c1 = chunk_lo(l + 1, dom.n, dom.B);
        float s = 0.f;
        for (int c = c0 + lane; c < c1; c += 32)
            for (int a = a0; a < a1; ++a) s += U[a * dom.M + c];
        const double t = warp_sum(double(s));
        if (lane == 0) blk[k * K + l] = t;
    }
}

// L_dom = sum_{k<l} (Kxx + Kyy - 2Kxy) / (K(K-1)/2)   (algorithms.py:110-116, :82-88) from the u block sums;
// executed by one warp, one lane per domain pair.  An empty chunk gives 0 * inf = NaN, as torch's
// mean() over an empty tensor does.
__device__ inline float mmd_from_blocks(const double* __restrict__ blk, const DomainInfo& dom, int lane) {
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
    // butterfly over the lanes that hold a pair only (float64 adds have ~50 cycles of latency each: 2 levels for 3 domains
    // instead of 5); lanes beyond hold 0
#pragma unroll
    for (int m = 16; m > 0; m >>= 1)
        if (m < 2 * npairs) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    return float(acc) / float(npairs);
}

}  // namespace wtpse

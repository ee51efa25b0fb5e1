// Backward: the per-sample coefficient matrix M_b = (S_b + S_b^T)/(P-1)   derived INSIDE the
// apply kernels from what the forward saved -- gram, rowstat and the MMD gradient seed `domgrad` (whitening_tail.cuh) --
// and the three upstream scalars, which exist only at backward time.  136 threads, three independent L2 loads and a
// handful of flops per sample change: the backward pass is one launch for every shape and layout.  Checked against
// autograd's derivation from the reference's statements (tests/test_gpu_parity.py).
#pragma once
#include "mmd_device.cuh"

namespace wtpse {

struct SeedArgs {
    const float* gram;       // [B][16][16]
    const float* rowstat;    // [B][2]
    const float* domgrad;    // [B][120]
    const float *g_off, *g_diag, *g_dom;   // device scalars, nullptr == 0
    int B, n, K;
};

struct SeedCtx {
    float g_dom, w_off, w_diag, denom;
    int M;
    bool need_dom;
};

// Call after griddepcontrol.wait (the scalars and the saved tensors may come from the kernel right before us).
__device__ __forceinline__ SeedCtx seed_context(const SeedArgs& a, long long P) {
    SeedCtx c;
    const float g_off = a.g_off ? __ldcg(a.g_off) : 0.f;
    const float g_diag = a.g_diag ? __ldcg(a.g_diag) : 0.f;
    c.g_dom = a.g_dom ? __ldcg(a.g_dom) : 0.f;
    const long long m = (long long)a.K * a.n;
    c.M = a.K > 1 ? int(m < a.B ? m : a.B) : 0;
    c.need_dom = (c.M > 0) && (c.g_dom != 0.f);
    c.w_off = g_off / (float(a.B) * float(kOff));
    c.w_diag = g_diag / (float(a.B) * float(kC));
    c.denom = float(P - 1);
    return c;
}

// The four values thread tid < 136 needs for its entry of M_b.  Loaded one sample AHEAD (the loads have a whole sample's
// worth of tiles to land), so a sample change costs one barrier and no memory latency.
struct SeedRegs {
    float g, off_b, diag_b, dom;
};

// (Does not depend on the upstream scalars: the kernel's first seed_load is issued together with their loads.)
__device__ __forceinline__ SeedRegs seed_load(const SeedArgs& a, const IndexTables& tab, int b, int tid) {
    SeedRegs r{0.f, 0.f, 0.f, 0.f};
    if (tid >= kTri || b < 0 || b >= a.B) return r;
    const int ij = tab.tri[tid], i = ij >> 4, j = ij & 15;
    const long long m = (long long)a.K * a.n;
    const int M = a.K > 1 ? int(m < a.B ? m : a.B) : 0;
    r.g = __ldcg(a.gram + b * 256 + i * kC + j);
    r.off_b = __ldcg(a.rowstat + b * 2 + 0);
    r.diag_b = __ldcg(a.rowstat + b * 2 + 1);
    if (b < M && i != j) r.dom = __ldcg(a.domgrad + b * kOff + off_idx(i, j));
    return r;
}

// Threads tid < 136 fill msh[16][16] from the values seed_load fetched; the caller synchronises afterwards.
__device__ __forceinline__ void seed_store(const SeedRegs& r, const SeedCtx& c, const IndexTables& tab, float* msh, int tid) {
    if (tid >= kTri) return;
    const int ij = tab.tri[tid], i = ij >> 4, j = ij & 15;
    const float dom_grad = (c.need_dom && r.dom != 0.f) ? c.g_dom * r.dom : 0.f;   // no upstream gradient / outside the MMD: exactly 0
    const float m = backward_matrix_entry(i, j, r.g, r.off_b, r.diag_b, c.w_off, c.w_diag, dom_grad, c.denom);
    msh[i * kC + j] = m;
    msh[j * kC + i] = m;
}

// Threads tid < 136 fill msh[16][16] for sample b; the caller synchronises before and after.
__device__ __forceinline__ void seed_matrix(const SeedArgs& a, const SeedCtx& c, const IndexTables& tab, int b, float* msh, int tid) {
    seed_store(seed_load(a, tab, b, tid), c, tab, msh, tid);
}

}  // namespace wtpse

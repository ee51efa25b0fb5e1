// Track W, db2: the streamed levels of the fused plan as persistent TMA pipelines that take ONE or TWO levels per pass,
// with the 4-tap filter bank in factored form.  PARITY UNPINNED (oracle/wavelet_np.py is this repository's own spec).
//
// Factored db2.  With a_j = x[2j], b_j = x[2j+1] and h = [h0 h1 h2 h3] the orthonormal db2 filter, the ratios
// h0/h1 = -h3/h2 = 1/sqrt3 and h1/h0 = -h2/h3 = sqrt3 give
//     p_j = b_j + a_j / sqrt3,   q_j = b_j - sqrt3 a_j,
//     lo_j = h1 (p_j + (h3/h1) q_{j+1}),        hi_j = -h2 (p_j + (h0/h2) q_{j+1})
// i.e. 4 fused multiply-adds per (lo, hi) pair instead of 8, and the scales h1, -h2 never have to be applied: the L1
// loss takes |.| and sign(.) of the detail bands (the scales are positive constants folded into the level weight, the
// two sign flips into the synthesis constants), and only the low-low band that leaves the pass is multiplied by h1^2.
// The adjoint runs the same two stages backwards: pbar_j = h1 lo_j - h2 hi_j, qbar_{j+1} = h3 lo_j - h0 hi_j,
// a_j = (pbar_j - 3 qbar_j) / sqrt3, b_j = pbar_j + qbar_j (5 operations per pair instead of 10).
//
//   db2_analysis_kernel   strip of input rows (TMA ring) -> level 1 [-> low-low rows in shared memory -> level 2]
//                         -> low-low band of the last level (global), one sign byte per site and level, |d| partial sums.
//                         A thread owns two adjacent sites and walks down the rows of its segment.
//   db2_synthesis_kernel  piece of coefficient rows: [gradient of LL2 + signs of level 2 -> gradient of LL1 in shared
//                         memory ->] + signs of level 1 -> 2 output rows per coefficient row, 128-bit coalesced stores.
//                         A WARP owns whole coefficient rows (lane l: sites 64k + 2l, 64k + 2l + 1), so the one value a
//                         site needs from its left neighbour comes through a shuffle and the periodic wrap is lane 31 ->
//                         lane 0; the row above is carried in registers.
// Two levels per pass keep LL1 (a quarter of the map, written and read twice by the one-level plan) out of memory: the
// pass moves 4.6 B per element instead of 5.25 + 1.3, and the resident stage that follows works on 1/16 of the map.
// The level-1 and level-2 work of a two-level pass run on separate warp groups of the CTA (hand-over buffers + mbarriers),
// so level 2 of strip n overlaps level 1 of strip n + 1.
//
// Haar goes through the same pipelines (kHaar / haar_synthesis_kernel): its filters do not overlap, so there are no halo
// rows, no carried row and no neighbour exchange -- sums / differences in the analysis, a 2 x 2 Hadamard butterfly per site
// in the synthesis.
//
// Measured (B200, 64 x 2 x 1024^2, J = 2): Haar analysis 84.5 us (the HBM read rate), db2 analysis 123 us (issue-bound: 25 %
// of the level-1 rows are overlap rows at this width), synthesis 103 us for both (bound by the write stream).
#include <math.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr double kS3d = 1.7320508075688772935;
constexpr double kS2d = 1.4142135623730950488;
constexpr double kH0d = (1.0 + kS3d) / (4.0 * kS2d), kH1d = (3.0 + kS3d) / (4.0 * kS2d);
constexpr double kH2d = (3.0 - kS3d) / (4.0 * kS2d), kH3d = (1.0 - kS3d) / (4.0 * kS2d);
constexpr float kI3 = float(1.0 / kS3d);            // 1 / sqrt3
constexpr float kR3 = float(kS3d);                  // sqrt3
constexpr float kKl = float(kH3d / kH1d);           // h3 / h1  (< 0)
constexpr float kKh = float(kH0d / kH2d);           // h0 / h2
constexpr float kH11 = float(kH1d * kH1d);          // scale of LL
constexpr float kH12 = float(kH1d * kH2d);          // |scale| of LH, HL
constexpr float kH22 = float(kH2d * kH2d);          // scale of HH

constexpr int kDb2MaxThreads = 544;                 // 16 consumer warps + the producer warp
constexpr int kDb2Smem = 226 * 1024;
constexpr uint32_t kDb2Chunk = 32 * 1024;

__device__ __forceinline__ void bulk_rows(unsigned char* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy, bool hint) {
    const char* s = static_cast<const char*>(src);
    for (uint32_t o = 0; o < bytes; o += kDb2Chunk) {
        const uint32_t n = min(kDb2Chunk, bytes - o);
        if (hint) tma_load_1d_hint(dst + o, s + o, n, bar, policy);
        else tma_load_1d(dst + o, s + o, n, bar);
    }
}

// rows [a, a + n) of a periodic plane of `hrows` rows (a may be negative, a + n <= hrows + (a < 0 ? 0 : ...)): at most two
// contiguous pieces.  Returns the bytes queued.
__device__ __forceinline__ uint32_t bulk_rows_wrapped(unsigned char* dst, const unsigned char* plane, uint32_t row_bytes, int a, int n,
                                                      int hrows, uint64_t* bar, uint64_t policy, bool hint) {
    int done = 0;
    while (done < n) {
        int r = a + done;
        r %= hrows;
        if (r < 0) r += hrows;
        const int len = min(n - done, hrows - r);
        bulk_rows(dst + size_t(done) * row_bytes, plane + size_t(r) * row_bytes, uint32_t(len) * row_bytes, bar, policy, hint);
        done += len;
    }
    return uint32_t(n) * row_bytes;
}

// ------------------------------------------------------------------------------------------------------------------
// analysis
// ------------------------------------------------------------------------------------------------------------------

// sign code of a detail value as HALF the 2-bit code: 0 (negative), 0.5 (zero), 1 (positive) -- one saturating FMA.
// (|v| < 2^-126 counts as zero; NaN gives 0.)
__device__ __forceinline__ float half_code(float v) { return __saturatef(fmaf(v, 4.2535295865117308e37f, 0.5f)); }

// along W: unscaled low / high pass of one row for the two sites at input columns c0 .. c0+3 (c4 = c0 + 4, wrapped)
__device__ __forceinline__ void db2_row(const float* row, int c0, int c4, float& l0, float& g0, float& l1, float& g1) {
    const float4 v = *reinterpret_cast<const float4*>(row + c0);
    const float2 u = *reinterpret_cast<const float2*>(row + c4);
    const float p0 = fmaf(v.x, kI3, v.y), p1 = fmaf(v.z, kI3, v.w);
    const float q1 = fmaf(v.z, -kR3, v.w), q2 = fmaf(u.x, -kR3, u.y);
    l0 = fmaf(q1, kKl, p0); g0 = fmaf(q1, kKh, p0);
    l1 = fmaf(q2, kKl, p1); g1 = fmaf(q2, kKh, p1);
}

struct Db2FwdArgs {
    const float* x;         // [nmaps][H][W]
    float* ll;              // low-low band of the LAST level of the pass: [nmaps][H/2][W/2] or [nmaps][H/4][W/4]
    unsigned char* sg1;     // [nmaps][H/2][W/2]
    unsigned char* sg2;     // [nmaps][H/4][W/4] (two levels)
    int H, W, nmaps, R, stages, pdl_wait, nw2, seg1, seg2;   // R: rows of the last level per strip; nw2: warps of the level-2 group; segN: rows per thread task
    float sc1, sc2;         // w_j / (3 * sites of level j * nmaps)
    double* partial;        // one per CTA
};

// One level on a strip in shared memory.  in: rows_out * 2 + 2 rows of width w (row stride w) at byte offset in_off.
// Output row i < n_own: sign byte + |d| (rows beyond are halo rows of the next level: low-low only).
//   kLLSmem: low-low rows (times h1^2) -> shared memory at ll_off, row stride w / 2; else -> ll_g (row stride w / 2).
template <bool kLLSmem, bool kGrad, bool kHaar>
__device__ __forceinline__ void db2_fwd_level(int in_off, int w, int rows_out, int n_own, int seg, int ll_off, float* __restrict__ ll_g,
                                              unsigned char* __restrict__ sg_g, float sc, float& ab, int tid, int nthreads) {
    extern __shared__ __align__(128) unsigned char smem[];
    const float* in = reinterpret_cast<const float*>(smem + in_off);
    float* ll_s = reinterpret_cast<float*>(smem + (kLLSmem ? ll_off : 0));
    const bool store_ll = ll_g != nullptr;                  // the deepest level's low-low band is not needed by anybody
    const int w2 = w >> 1, pairs = w >> 2;
    const int nseg = (rows_out + seg - 1) / seg;
    const int ntasks = nseg * pairs;
    for (int task = tid; task < ntasks; task += nthreads) {
        const int jj = task % pairs, si = task / pairs;
        const int i0 = si * seg, i1 = min(i0 + seg, rows_out), c0 = 4 * jj;
        int c4 = c0 + 4;
        if (c4 >= w) c4 -= w;
        // running pointers (byte arithmetic once per task, not per row)
        const float* r0 = in + (2 * i0) * w + c0;                   // columns c0 .. c0 + 3 of the current even row
        const float* r4 = in + (2 * i0) * w + c4;                   // columns c0 + 4, c0 + 5 (wrapped)
        // Haar (kHaar): the filters do not overlap -- output row i reads input rows 2 i, 2 i + 1 only, nothing is carried, and the
        // unscaled bands are plain sums / differences (scale 1/2 each, folded like db2's)
        float P[4] = {0.f, 0.f, 0.f, 0.f};
        if (!kHaar) {
            float a[4], b[4];
            db2_row(r0, 0, int(r4 - r0), a[0], a[1], a[2], a[3]);
            db2_row(r0 + w, 0, int(r4 - r0), b[0], b[1], b[2], b[3]);
#pragma unroll
            for (int u = 0; u < 4; ++u) P[u] = fmaf(a[u], kI3, b[u]);
        } else {
            r0 -= 2 * w;                                            // the loop advances before it reads
        }
        const int d4 = int(r4 - r0);
        float* llp_s = ll_s + i0 * w2 + 2 * jj;
        float* llp_g = ll_g + (long long)i0 * w2 + 2 * jj;
        unsigned char* sgp = sg_g + (long long)i0 * w2 + 2 * jj;
        float s1 = 0.f, s2 = 0.f;
        for (int i = i0; i < i1; ++i) {
            r0 += 2 * w;
            float a[4], b[4], lo[4], hi[4];
            if (!kHaar) {
                db2_row(r0, 0, d4, a[0], a[1], a[2], a[3]);
                db2_row(r0 + w, 0, d4, b[0], b[1], b[2], b[3]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float Pn = fmaf(a[u], kI3, b[u]), Qn = fmaf(a[u], -kR3, b[u]);
                    lo[u] = fmaf(Qn, kKl, P[u]);
                    hi[u] = fmaf(Qn, kKh, P[u]);
                    P[u] = Pn;
                }
            } else {
                const float4 va = *reinterpret_cast<const float4*>(r0), vb = *reinterpret_cast<const float4*>(r0 + w);
                a[0] = va.x + va.y; a[1] = va.x - va.y; a[2] = va.z + va.w; a[3] = va.z - va.w;      // along W: low, high of sites 0, 1
                b[0] = vb.x + vb.y; b[1] = vb.x - vb.y; b[2] = vb.z + vb.w; b[3] = vb.z - vb.w;
#pragma unroll
                for (int u = 0; u < 4; ++u) { lo[u] = a[u] + b[u]; hi[u] = a[u] - b[u]; }                // along H
            }
            // arrays: 0 = low along W of site 0, 1 = high along W of site 0, 2 / 3 = the same for site 1
            // unscaled bands: LL = lo[0], HL = hi[0] (high along H), LH = lo[1], HH = hi[1]
            const float2 LL = make_float2(lo[0] * (kHaar ? 0.5f : kH11), lo[2] * (kHaar ? 0.5f : kH11));
            if (kLLSmem) *reinterpret_cast<float2*>(llp_s) = LL;
            else if (store_ll) *reinterpret_cast<float2*>(llp_g) = LL;
            if (i < n_own) {
                s1 += (fabsf(lo[1]) + fabsf(hi[0])) + (fabsf(lo[3]) + fabsf(hi[2]));
                s2 += fabsf(hi[1]) + fabsf(hi[3]);
                if (kGrad) {
                    // 2-bit codes (0 negative, 1 zero, 2 positive) of LH | HL << 2 | HH << 4 per site, site 1 in the high byte,
                    // accumulated in the integer part of a float (1.5 * 2^23 + n has bit pattern 0x4B400000 + n)
                    float acc = 12582912.0f;
                    acc = fmaf(half_code(lo[1]), 2.0f, acc);
                    acc = fmaf(half_code(hi[0]), 8.0f, acc);
                    acc = fmaf(half_code(hi[1]), 32.0f, acc);
                    acc = fmaf(half_code(lo[3]), 512.0f, acc);
                    acc = fmaf(half_code(hi[2]), 2048.0f, acc);
                    acc = fmaf(half_code(hi[3]), 8192.0f, acc);
                    *reinterpret_cast<unsigned short*>(sgp) = static_cast<unsigned short>(__float_as_uint(acc));
                }
            }
            llp_s += w2;
            llp_g += w2;
            sgp += w2;
        }
        ab += kHaar ? sc * (0.5f * (s1 + s2)) : sc * fmaf(kH12, s1, kH22 * s2);
    }
}

template <bool kTwo, bool kGrad, bool kHaar>
__global__ void __launch_bounds__(kDb2MaxThreads, 1) db2_analysis_kernel(Db2FwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int H = a.H, W = a.W, R = a.R, S = a.stages;
    // warps: [0, NW1) level 1, [NW1, NW1 + NW2) level 2 (two-level pass only), then the producer warp.  The two consumer
    // groups work on DIFFERENT strips at any time (level 2 of strip n overlaps level 1 of strip n + 1), handing the LL1 rows
    // over through two shared-memory buffers guarded by mbarriers.
    const int NW2 = kTwo ? a.nw2 : 0;
    const int NW1 = int(blockDim.x) / 32 - 1 - NW2;
    const int hl = kTwo ? H >> 2 : H >> 1;                  // rows of the last level
    constexpr int HALO = kHaar ? 0 : 2;                     // overlap rows of one level (db2: 2, Haar: none)
    const int rows_in = kTwo ? 4 * R + 3 * HALO : 2 * R + HALO;
    const int n1 = kTwo ? 2 * R + HALO : R;                 // level-1 rows a strip computes
    const uint32_t stage_bytes = uint32_t(rows_in) * W * 4u;
    const int ll1_bytes = kTwo ? n1 * (W >> 1) * 4 : 0;     // one of the two LL1 buffers
    const int ll1_off = int(S * stage_bytes);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + ll1_off + 2 * ll1_bytes);
    uint64_t* empty = full + S;
    uint64_t* ll_full = empty + S;                          // [2]
    uint64_t* ll_empty = ll_full + 2;                       // [2]
    const int spm = hl / R;                                 // strips per map
    const long long T = (long long)a.nmaps * spm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&ll_full[b], NW1);
            mbar_init(&ll_empty[b], NW2 > 0 ? NW2 : 1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (a.pdl_wait) asm volatile("griddepcontrol.wait;" ::: "memory");     // input written by the pass in front of this one

    double acc = 0.0;
    const int h2 = H >> 1, w2 = W >> 1, w4 = W >> 2;
    if (warp == NW1 + NW2) {
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();
            int n = 0;
            for (long long t = blockIdx.x; t < T; t += gridDim.x, ++n) {
                const int s = n % S;
                if (n >= S) mbar_wait(&empty[s], ((n / S) & 1) ^ 1);
                const long long m = t / spm;
                const int r0 = (kTwo ? 4 : 2) * R * int(t % spm);
                const unsigned char* plane = reinterpret_cast<const unsigned char*>(a.x + m * (long long)H * W);
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                bulk_rows_wrapped(smem + size_t(s) * stage_bytes, plane, uint32_t(W) * 4u, r0, rows_in, H, &full[s], policy, true);
            }
        }
    } else if (warp < NW1) {
        const int tid = threadIdx.x, NC1 = NW1 * 32;
        float ab = 0.f;
        int n = 0;
        for (long long t = blockIdx.x; t < T; t += gridDim.x, ++n) {
            const int s = n % S;
            const long long m = t / spm;
            const int st = int(t % spm);
            mbar_wait(&full[s], (n / S) & 1);
            if (!kTwo) {
                const long long row0 = m * h2 + (long long)R * st;
                db2_fwd_level<false, kGrad, kHaar>(int(s * stage_bytes), W, R, R, a.seg1, 0, a.ll ? a.ll + row0 * w2 : nullptr, a.sg1 + row0 * w2, a.sc1, ab, tid, NC1);
            } else {
                const int b = n & 1, k = n >> 1;
                if (k > 0) mbar_wait(&ll_empty[b], (k & 1) ^ 1);                // level 2 is done with this buffer's previous strip
                const long long row1 = m * h2 + (long long)2 * R * st;          // first level-1 row of the strip
                db2_fwd_level<true, kGrad, kHaar>(int(s * stage_bytes), W, n1, 2 * R, a.seg1, ll1_off + b * ll1_bytes, nullptr, a.sg1 + row1 * w2, a.sc1, ab, tid,
                                           NC1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&ll_full[b]);                        // release: the LL1 rows this warp wrote
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);                              // the input rows are free
            acc += double(ab);
            ab = 0.f;
        }
    } else {
        // level 2 of the two-level pass
        const int tid = threadIdx.x - NW1 * 32, NC2 = NW2 * 32;
        float ab = 0.f;
        int n = 0;
        for (long long t = blockIdx.x; t < T; t += gridDim.x, ++n) {
            const long long m = t / spm;
            const int st = int(t % spm);
            const int b = n & 1, k = n >> 1;
            mbar_wait(&ll_full[b], k & 1);
            const long long row2 = m * (H >> 2) + (long long)R * st;
            db2_fwd_level<false, kGrad, kHaar>(ll1_off + b * ll1_bytes, w2, R, R, a.seg2, 0, a.ll ? a.ll + row2 * w4 : nullptr, a.sg2 + row2 * w4, a.sc2, ab, tid, NC2);
            __syncwarp();
            if (lane == 0) mbar_arrive(&ll_empty[b]);
            acc += double(ab);
            ab = 0.f;
        }
    }

    __shared__ double red[kDb2MaxThreads / 32];
    const double sum = warp_sum(acc);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int q = 0; q < int(blockDim.x) / 32; ++q) tot += red[q];
        a.partial[blockIdx.x] = tot;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// synthesis
// ------------------------------------------------------------------------------------------------------------------

// 2-bit code at bits [s, s + 1] of b -> -1 / 0 / +1 in two instructions and no shift: (b & mask) | bits(2^(23 - s)) puts the code
// where the mantissa has weight 1, then subtract 2^(23 - s) + 1 (exact).  The OR-ed constant must sit in a REGISTER for the and-or
// to be one LOP3 (two immediates make two): the six constants travel in SynConst.
template <int s>
__device__ __forceinline__ float code_float(unsigned b, unsigned magic) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(b), "n"(3u << s), "r"(magic));      // (b & mask) | magic
    return __uint_as_float(r);
}
template <int s>
__host__ __device__ __forceinline__ constexpr unsigned code_magic() { return unsigned(150 - s) << 23; }
template <int s>
__device__ __forceinline__ constexpr float code_bias() { return -(float(1u << (23 - s)) + 1.0f); }

// Synthesis constants.  With kappa = 3 h3 / h1 and rho = (h0 / h2) / (h3 / h1) the carried neighbour term is kept as
// qt = 3 qbar / kappa, so that it costs ONE fma (qt_next = lo' + rho hi') and the outputs are a = cA qt + pbar / sqrt3,
// b = cB qt + pbar with cA = -kappa / sqrt3, cB = kappa / 3.
constexpr float kRho = float((kH0d / kH2d) / (kH3d / kH1d));
constexpr float kCA = float(-(3.0 * kH3d / kH1d) / kS3d);
constexpr float kCB = float(kH3d / kH1d);
struct SynConst {
    float cLL, c1, c1r, c2, c2r;        // h1^2 g, h1 h2 sc g, rho h1 h2 sc g, h2^2 sc g, rho h2^2 sc g   (g = upstream gradient)
    unsigned m0, m2, m4, m8, m10, m12;  // code_magic<s>() in registers
};

__device__ __forceinline__ SynConst make_syn_const(float sc, float gs, const unsigned* __restrict__ magic) {
    SynConst c;
    c.cLL = kH11 * gs;
    c.c1 = kH12 * sc * gs;
    c.c1r = kRho * c.c1;
    c.c2 = kH22 * sc * gs;
    c.c2r = kRho * c.c2;
    c.m0 = magic[0]; c.m2 = magic[1]; c.m4 = magic[2]; c.m8 = magic[3]; c.m10 = magic[4]; c.m12 = magic[5];
    return c;
}

// Synthesis along H of ONE site: its LL gradient g and sign byte bs, the carried qt of the coefficient row above.
// Outputs (when kOut): the column-synthesised values of the even (A) and odd (B) output row for the "low along W" array
// (already times h1) and the "high along W" array (already times -h2); always: the carries of this row.
template <bool kHasLL, bool kOut>
__device__ __forceinline__ void syn_vert_site(float g, unsigned bs, const SynConst& c, float qlo_in, float qhi_in, float& qlo_out,
                                              float& qhi_out, float& Alo, float& Blo, float& Ahi, float& Bhi) {
    const float sLH = code_float<0>(bs, c.m0) + code_bias<0>(), sHL = code_float<2>(bs, c.m2) + code_bias<2>(),
                sHH = code_float<4>(bs, c.m4) + code_bias<4>();
    const float e = c.c1 * sHL, er = c.c1r * sHL;
    const float pb = kHasLL ? fmaf(c.cLL, g, e) : e;
    const float qn = kHasLL ? fmaf(c.cLL, g, er) : er;
    const float lam = c.c1 * sLH;
    const float pbh = fmaf(c.c2, sHH, lam);
    const float qnh = fmaf(c.c2r, sHH, lam);
    if (kOut) {
        Alo = fmaf(qlo_in, kCA, kI3 * pb);
        Blo = fmaf(qlo_in, kCB, pb);
        Ahi = fmaf(qhi_in, kCA, kI3 * pbh);
        Bhi = fmaf(qhi_in, kCB, pbh);
    }
    qlo_out = qn;
    qhi_out = qnh;
}

// The same for a lane's PAIR of adjacent sites with packed fp32 arithmetic (fma.rn.f32x2 / add / mul, SASS FFMA2 FADD2 FMUL2):
// the two sites run identical operations on adjacent registers (the LL gradients arrive as one 64-bit shared load), so the
// vertical stage costs 13 packed instructions per pair instead of 26 -- these kernels are bound by instruction issue.
struct SynPair {
    float2 Alo, Blo, Ahi, Bhi;
};
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

template <bool kHasLL, bool kOut>
__device__ __forceinline__ void syn_vert_pair(float2 g, unsigned b, const SynConst& c, float2& qlo, float2& qhi, SynPair& o) {
    const float2 sLH = __fadd2_rn(make_float2(code_float<0>(b, c.m0), code_float<8>(b, c.m8)), make_float2(code_bias<0>(), code_bias<8>()));
    const float2 sHL = __fadd2_rn(make_float2(code_float<2>(b, c.m2), code_float<10>(b, c.m10)), make_float2(code_bias<2>(), code_bias<10>()));
    const float2 sHH = __fadd2_rn(make_float2(code_float<4>(b, c.m4), code_float<12>(b, c.m12)), make_float2(code_bias<4>(), code_bias<12>()));
    const float2 e = __fmul2_rn(splat(c.c1), sHL), er = __fmul2_rn(splat(c.c1r), sHL);
    const float2 pb = kHasLL ? __ffma2_rn(splat(c.cLL), g, e) : e;
    const float2 qn = kHasLL ? __ffma2_rn(splat(c.cLL), g, er) : er;
    const float2 lam = __fmul2_rn(splat(c.c1), sLH);
    const float2 pbh = __ffma2_rn(splat(c.c2), sHH, lam);
    const float2 qnh = __ffma2_rn(splat(c.c2r), sHH, lam);
    if (kOut) {
        o.Alo = __ffma2_rn(qlo, splat(kCA), __fmul2_rn(splat(kI3), pb));
        o.Blo = __ffma2_rn(qlo, splat(kCB), pb);
        o.Ahi = __ffma2_rn(qhi, splat(kCA), __fmul2_rn(splat(kI3), pbh));
        o.Bhi = __ffma2_rn(qhi, splat(kCB), pbh);
    }
    qlo = qn;
    qhi = qnh;
}

// One level, executed by one warp on the coefficient rows [ra, rb) of a shared-memory buffer whose row ra - 1 is the row
// above (ra >= 1).  g rows: fp32, row stride wj = 64 K; sign rows: bytes, row stride wj.  Output row 2 (r - 1) + pr:
//   kToSmem: shared memory at out_off + (2 (r - 1) + pr) * 2 wj floats;   else out_g + (2 (r - 1) + pr) * out_ld.
template <int K, bool kHasLL, bool kToSmem>
__device__ __forceinline__ void db2_inv_rows(int g_off, int sg_off, int ra, int rb, const SynConst& c, int out_off, float* __restrict__ out_g,
                                             long long out_ld, int lane) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int wj = 64 * K;
    const float* gbase = reinterpret_cast<const float*>(smem + (kHasLL ? g_off : 0));
    const unsigned char* sbase = smem + sg_off;
    float* outs = reinterpret_cast<float*>(smem + (kToSmem ? out_off : 0));
    float2 q3lo[K], q3hi[K];                                  // per chunk: carried qt of (site 0, site 1), low / high along W
    {   // the row above: carries only
        const float* grow = gbase + (ra - 1) * wj + 2 * lane;
        const unsigned char* srow = sbase + (ra - 1) * wj + 2 * lane;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float2 g2 = make_float2(0.f, 0.f);
            if (kHasLL) g2 = *reinterpret_cast<const float2*>(grow + 64 * k);
            const unsigned b = *reinterpret_cast<const unsigned short*>(srow + 64 * k);
            SynPair d;
            syn_vert_pair<kHasLL, false>(g2, b, c, q3lo[k], q3hi[k], d);
        }
    }
    const float* grow = gbase + ra * wj + 2 * lane;
    const unsigned char* srow = sbase + ra * wj + 2 * lane;
    float* orow_s = outs + (2 * (ra - 1)) * (2 * wj) + 4 * lane;
    float* orow_g = out_g + (long long)(2 * (ra - 1)) * out_ld + 4 * lane;
    for (int r = ra; r < rb; ++r) {
        // the value lane 0 needs for its first site of chunk 0: 3 * qbar (along W) of the row's LAST site (lane 31, chunk K - 1,
        // site 1), for both output rows -- computed by every lane for its own last site (pure, the carries are not touched)
        float left[2];
        {
            float g = 0.f;
            if (kHasLL) g = grow[64 * (K - 1) + 1];
            const unsigned bs = srow[64 * (K - 1) + 1];
            float qa, qb, Alo, Blo, Ahi, Bhi;
            syn_vert_site<kHasLL, true>(g, bs, c, q3lo[K - 1].y, q3hi[K - 1].y, qa, qb, Alo, Blo, Ahi, Bhi);
            const float w0 = fmaf(kRho, Ahi, Alo), w1 = fmaf(kRho, Bhi, Blo);
            left[0] = __shfl_sync(0xffffffffu, w0, 31);
            left[1] = __shfl_sync(0xffffffffu, w1, 31);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float2 g2 = make_float2(0.f, 0.f);
            if (kHasLL) g2 = *reinterpret_cast<const float2*>(grow + 64 * k);
            const unsigned b = *reinterpret_cast<const unsigned short*>(srow + 64 * k);
            SynPair v;
            syn_vert_pair<kHasLL, true>(g2, b, c, q3lo[k], q3hi[k], v);
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                const float2 Xlo = pr ? v.Blo : v.Alo, Xhi = pr ? v.Bhi : v.Ahi;
                const float2 pb = __fadd2_rn(Xlo, Xhi);
                const float2 qn = __ffma2_rn(splat(kRho), Xhi, Xlo);     // qt: .x for site 1, .y for the next lane's site 0
                const float recv = __shfl_sync(0xffffffffu, qn.y, (lane + 31) & 31);
                const float ql = lane == 0 ? left[pr] : recv;
                left[pr] = recv;                              // lane 0: lane 31's value of THIS chunk = its neighbour in the next
                float4 o;
                o.x = fmaf(ql, kCA, kI3 * pb.x);
                o.y = fmaf(ql, kCB, pb.x);
                o.z = fmaf(qn.x, kCA, kI3 * pb.y);
                o.w = fmaf(qn.x, kCB, pb.y);
                if (kToSmem) *reinterpret_cast<float4*>(orow_s + pr * (2 * wj) + 128 * k) = o;
                else __stcs(reinterpret_cast<float4*>(orow_g + pr * out_ld + 128 * k), o);     // written once, read by somebody else later: streaming store
            }
        }
        grow += wj;
        srow += wj;
        orow_s += 2 * (2 * wj);
        orow_g += 2 * out_ld;
    }
}

// The same for rows of only 32 sites (level 2 of a pass over 128-wide planes): a warp takes TWO rows at a time, one per half-warp
// (lane l: sites 2 (l & 15), 2 (l & 15) + 1 of the row of half l >> 4), each half with its own row segment, carried row above and
// periodic wrap (lane 15 -> lane 0 of the same half).  Both halves run the same instruction stream; a half that has run out of
// rows repeats its last row without storing.
template <bool kHasLL, bool kToSmem>
__device__ __forceinline__ void db2_inv_rows_half(int g_off, int sg_off, int ra, int rb, const SynConst& c, int out_off, float* __restrict__ out_g,
                                                  long long out_ld, int lane) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int wj = 32;
    const int l16 = lane & 15, half = lane >> 4;
    const int mid = ra + (rb - ra + 1) / 2;
    const int r_first = half ? mid : ra, r_end = half ? rb : mid;            // this half's rows [r_first, r_end)
    const int iters = mid - ra;                                                // >= the other half's count
    const float* gbase = reinterpret_cast<const float*>(smem + (kHasLL ? g_off : 0));
    const unsigned char* sbase = smem + sg_off;
    float* outs = reinterpret_cast<float*>(smem + (kToSmem ? out_off : 0));
    float2 q3lo, q3hi;
    {
        const int r0 = (r_first < r_end ? r_first : ra) - 1;                  // an empty half still executes (on valid memory)
        float2 g2 = make_float2(0.f, 0.f);
        if (kHasLL) g2 = *reinterpret_cast<const float2*>(gbase + r0 * wj + 2 * l16);
        const unsigned b = *reinterpret_cast<const unsigned short*>(sbase + r0 * wj + 2 * l16);
        SynPair d;
        syn_vert_pair<kHasLL, false>(g2, b, c, q3lo, q3hi, d);
    }
    for (int i = 0; i < iters; ++i) {
        const bool valid = r_first + i < r_end;
        const int r = valid ? r_first + i : (r_first < r_end ? r_end - 1 : ra);
        const float* grow = gbase + r * wj + 2 * l16;
        const unsigned char* srow = sbase + r * wj + 2 * l16;
        float2 g2 = make_float2(0.f, 0.f);
        if (kHasLL) g2 = *reinterpret_cast<const float2*>(grow);
        const unsigned b = *reinterpret_cast<const unsigned short*>(srow);
        // the row's last site (lane 15 of this half, site 1) supplies the first site's left neighbour: compute this lane's own
        // site-1 neighbour terms first (they are needed anyway), pass lane 15's to lane 0
        float2 cq_lo = q3lo, cq_hi = q3hi;
        SynPair v;
        syn_vert_pair<kHasLL, true>(g2, b, c, cq_lo, cq_hi, v);
        if (valid) { q3lo = cq_lo; q3hi = cq_hi; }
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            const float2 Xlo = pr ? v.Blo : v.Alo, Xhi = pr ? v.Bhi : v.Ahi;
            const float2 pb = __fadd2_rn(Xlo, Xhi);
            const float2 qn = __ffma2_rn(splat(kRho), Xhi, Xlo);               // qt: .x for site 1, .y for the next lane's site 0
            const float ql = __shfl_sync(0xffffffffu, qn.y, (lane & 16) | ((l16 + 15) & 15));     // lane 0 of a half: from its lane 15 (wrap)
            float4 o;
            o.x = fmaf(ql, kCA, kI3 * pb.x);
            o.y = fmaf(ql, kCB, pb.x);
            o.z = fmaf(qn.x, kCA, kI3 * pb.y);
            o.w = fmaf(qn.x, kCB, pb.y);
            if (valid) {
                if (kToSmem) *reinterpret_cast<float4*>(outs + (2 * (r - 1) + pr) * (2 * wj) + 4 * l16) = o;
                else __stcs(reinterpret_cast<float4*>(out_g + (long long)(2 * (r - 1) + pr) * out_ld + 4 * l16), o);
            }
        }
    }
}

struct Db2InvArgs {
    const float* g;             // gradient of the low-low band that enters the pass (ignored when !kHasLL)
    const unsigned char* sg1;   // [nmaps][H/2][W/2]
    const unsigned char* sg2;   // [nmaps][H/4][W/4] (two levels)
    float* out;                 // [nmaps][H][W]
    int H, W, nmaps, R, stages, nw2, rr; // R: most level-1 coefficient rows per piece; nw2: warps of the level-2 group; rr: pieces dealt round-robin
    float sc1, sc2;
    const float* upstream;      // device scalar multiplied into out (nullptr: 1)
    const double* partial;      // loss partials of the preceding kernels, summed in fixed order by CTA 0 ...
    int n_partials;
    float* loss;                // ... into loss (nullptr: somebody else does it)
    unsigned magic[6];          // code_magic<0, 2, 4, 8, 10, 12>: kernel parameters, so that they live in registers (see code_float)
};

// contiguous, balanced ranges of the nmaps * h2 coefficient rows, cut into pieces of at most R rows that stay inside a map
struct Db2Pieces {
    long long r, e;             // contiguous mode: this CTA's row range
    long long p, T;             // round-robin mode: next piece, number of pieces
    int h2, R, ppm;             // rows per map, rows per piece, pieces per map
    bool rr;
    // rr: pieces of R rows (the last one of a map shorter) dealt round-robin to the CTAs -- at any moment the grid then writes ONE
    // contiguous window of the output (148 adjacent pieces) instead of 148 scattered ones; contiguous: balanced row ranges
    __device__ Db2Pieces(long long total, int h2_, int R_, bool rr_) : h2(h2_), R(R_), rr(rr_) {
        r = total * blockIdx.x / gridDim.x;
        e = total * (blockIdx.x + 1) / gridDim.x;
        ppm = (h2 + R - 1) / R;
        p = blockIdx.x;
        T = (total / h2) * ppm;
    }
    __device__ bool next(long long& m, int& i_first, int& len) {
        if (rr) {
            if (p >= T) return false;
            m = p / ppm;
            i_first = int(p - m * ppm) * R;
            len = min(R, h2 - i_first);
            p += gridDim.x;
            return true;
        }
        if (r >= e) return false;
        m = r / h2;
        i_first = int(r - m * h2);
        len = int(min((long long)min(R, h2 - i_first), e - r));
        r += len;
        return true;
    }
};

__device__ __forceinline__ int floor_half(int v) { return v >= 0 ? v >> 1 : -((1 - v) >> 1); }

// Shared-memory carve of one stage (bytes).  One level: [g (R + 1) rows][sg1 (R + 1) rows].
// Two levels: [g2 (R / 2 + 3) rows][sg2 (R / 2 + 3) rows][sg1 (R + 1) rows].
struct Db2InvLayout {
    int g, s2, s1, stage, g1buf, g1bytes;
};
__host__ __device__ inline Db2InvLayout db2_inv_layout(int W, int R, bool two, bool has_ll, int S, bool haar = false) {
    Db2InvLayout o;
    const int w2 = W >> 1, w4 = W >> 2;
    const int above = haar ? 0 : 1;                 // coefficient rows above a piece (db2 overlap)
    int off = 0;
    o.g = off;
    if (two) {
        const int n2 = R / 2 + 1 + 2 * above;
        if (has_ll) off += n2 * w4 * 4;
        o.s2 = off;
        off += (n2 * w4 + 15) & ~15;
    } else {
        if (has_ll) off += (R + above) * w2 * 4;
        o.s2 = off;
    }
    o.s1 = off;
    off += ((R + above) * w2 + 15) & ~15;
    o.stage = (off + 127) & ~127;
    o.g1buf = S * o.stage;
    o.g1bytes = two ? (R + 2 + 4 * above) * w2 * 4 : 0;     // rows of dL/dLL1 one piece needs, one of two buffers
    return o;
}

template <int K1, bool kTwo, bool kHasLL>
__global__ void __launch_bounds__(kDb2MaxThreads, 1) db2_synthesis_kernel(Db2InvArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // warps: [0, NW1) level 1 -> global, [NW1, NW) level 2 -> dL/dLL1 in shared memory (two-level pass only), warp NW = producer.
    // Level 2 of piece n + 1 overlaps level 1 of piece n: two dL/dLL1 buffers, handed over through mbarriers.
    constexpr int NW = kDb2MaxThreads / 32 - 1;
    constexpr int K2 = (kTwo && K1 >= 2) ? K1 / 2 : 1;       // K1 == 1 (128-wide planes): level 2 has 32 sites per row -> db2_inv_rows_half
    const int NW2 = kTwo ? a.nw2 : 0, NW1 = NW - NW2;
    const int H = a.H, W = a.W, h2 = H >> 1, w2 = W >> 1, h4 = H >> 2, w4 = W >> 2, R = a.R, S = a.stages;
    const Db2InvLayout lay = db2_inv_layout(W, R, kTwo, kHasLL, S);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + lay.g1buf + 2 * lay.g1bytes);
    uint64_t* empty = full + S;
    uint64_t* g1_full = empty + S;          // [2]
    uint64_t* g1_empty = g1_full + 2;       // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&g1_full[b], NW2 > 0 ? NW2 : 1);
            mbar_init(&g1_empty[b], NW1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    Db2Pieces pieces((long long)a.nmaps * h2, h2, R, a.rr != 0);
    long long m;
    int i_first, len;
    if (warp == NW) {
        if (lane == 0) {
            for (int n = 0; pieces.next(m, i_first, len); ++n) {
                const int s = n % S;
                if (n >= S) mbar_wait(&empty[s], ((n / S) & 1) ^ 1);
                unsigned char* dst = smem + size_t(s) * lay.stage;
                uint32_t bytes = uint32_t(len + 1) * w2;
                if (kTwo) {
                    const int lo2 = floor_half(i_first - 1) - 1, n2 = floor_half(i_first + len - 1) - lo2 + 1;
                    bytes += uint32_t(n2) * w4 * (kHasLL ? 5u : 1u);
                    mbar_arrive_expect_tx(&full[s], bytes);
                    if (kHasLL)
                        bulk_rows_wrapped(dst + lay.g, reinterpret_cast<const unsigned char*>(a.g + m * (long long)h4 * w4), uint32_t(w4) * 4u, lo2, n2,
                                          h4, &full[s], 0, false);
                    bulk_rows_wrapped(dst + lay.s2, a.sg2 + m * (long long)h4 * w4, uint32_t(w4), lo2, n2, h4, &full[s], 0, false);
                } else {
                    if (kHasLL) bytes += uint32_t(len + 1) * w2 * 4u;
                    mbar_arrive_expect_tx(&full[s], bytes);
                    if (kHasLL)
                        bulk_rows_wrapped(dst + lay.g, reinterpret_cast<const unsigned char*>(a.g + m * (long long)h2 * w2), uint32_t(w2) * 4u,
                                          i_first - 1, len + 1, h2, &full[s], 0, false);
                }
                bulk_rows_wrapped(dst + lay.s1, a.sg1 + m * (long long)h2 * w2, uint32_t(w2), i_first - 1, len + 1, h2, &full[s], 0, false);
            }
        }
        __syncwarp();
        if (a.loss && blockIdx.x == 0) {                    // fixed-order sum of the loss partials, once this CTA's loads are queued
            double s = 0.0;
            for (int i = lane; i < a.n_partials; i += 32) s += a.partial[i];
            s = warp_sum(s);
            if (lane == 0) a.loss[0] = float(s);
        }
    } else if (warp < NW1) {
        const float gs = a.upstream ? __ldg(a.upstream) : 1.0f;
        const SynConst c1 = make_syn_const(a.sc1, gs, a.magic);
        for (int n = 0; pieces.next(m, i_first, len); ++n) {
            const int s = n % S;
            const int st = s * lay.stage;
            mbar_wait(&full[s], (n / S) & 1);
            float* obase = a.out + m * (long long)H * W + (long long)(2 * i_first) * W;       // output row of coefficient row i_first
            const int seg1 = (len + NW1 - 1) / NW1;
            const int ra1 = 1 + warp * seg1, rb1 = min(ra1 + seg1, len + 1);
            if (!kTwo) {
                if (ra1 < rb1) db2_inv_rows<K1, kHasLL, false>(st + lay.g, st + lay.s1, ra1, rb1, c1, 0, obase, W, lane);
            } else {
                const int b = n & 1, k = n >> 1;
                mbar_wait(&g1_full[b], k & 1);
                // buffer row 0 of dL/dLL1 is row 2 (lo2 + 1); the row above the piece, i_first - 1, sits at offset
                // i_first - 1 - 2 (lo2 + 1) in {0, 1}
                const int lo2 = floor_half(i_first - 1) - 1;
                const int shift = i_first - 1 - 2 * (lo2 + 1);
                if (ra1 < rb1)
                    db2_inv_rows<K1, true, false>(lay.g1buf + b * lay.g1bytes + shift * w2 * 4, st + lay.s1, ra1, rb1, c1, 0, obase, W, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&g1_empty[b]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    } else if (kTwo) {
        // level 2: coefficient rows lo2 + 1 .. hi2 -> dL/dLL1 rows 2 (lo2 + 1) .. 2 hi2 + 1 in this piece's buffer
        const SynConst c2 = make_syn_const(a.sc2, 1.0f, a.magic);
        const int w = warp - NW1;
        for (int n = 0; pieces.next(m, i_first, len); ++n) {
            const int s = n % S;
            const int st = s * lay.stage;
            const int b = n & 1, k = n >> 1;
            mbar_wait(&full[s], (n / S) & 1);
            if (k > 0) mbar_wait(&g1_empty[b], (k & 1) ^ 1);
            const int lo2 = floor_half(i_first - 1) - 1, rows2 = floor_half(i_first + len - 1) - lo2;
            const int nw2d = kTwo ? NW2 : 1;
            const int seg2 = (rows2 + nw2d - 1) / nw2d;
            const int ra2 = 1 + w * seg2, rb2 = min(ra2 + seg2, rows2 + 1);
            if (ra2 < rb2) {
                if constexpr (K1 >= 2) db2_inv_rows<K2, kHasLL, true>(st + lay.g, st + lay.s2, ra2, rb2, c2, lay.g1buf + b * lay.g1bytes, nullptr, 0, lane);
                else db2_inv_rows_half<kHasLL, true>(st + lay.g, st + lay.s2, ra2, rb2, c2, lay.g1buf + b * lay.g1bytes, nullptr, 0, lane);
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&g1_full[b]);
                mbar_arrive(&empty[s]);
            }
        }
    }
}

// ---- Haar synthesis: no overlap, so no carried row, no neighbour exchange -- a 2 x 2 output block is the Hadamard butterfly of
// its site's four gradient coefficients:  x[2i + pr][2j + pc] = 1/2 (gLL + (-1)^pc gLH + (-1)^pr gHL + (-1)^(pr + pc) gHH).
// Same warp-per-row layout, sign bytes and packed site pairs as the db2 rows above.
template <int K, bool kHasLL, bool kToSmem>
__device__ __forceinline__ void haar_inv_rows(int g_off, int sg_off, int ra, int rb, float cLL, float cD, const SynConst& c, int out_off,
                                              float* __restrict__ out_g, long long out_ld, int lane) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int wj = 64 * K;
    const float* grow = reinterpret_cast<const float*>(smem + (kHasLL ? g_off : 0)) + ra * wj + 2 * lane;
    const unsigned char* srow = smem + sg_off + ra * wj + 2 * lane;
    float* orow_s = reinterpret_cast<float*>(smem + (kToSmem ? out_off : 0)) + (2 * ra) * (2 * wj) + 4 * lane;
    float* orow_g = out_g + (long long)(2 * ra) * out_ld + 4 * lane;
    for (int r = ra; r < rb; ++r) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float2 t0 = make_float2(0.f, 0.f);
            if (kHasLL) t0 = __fmul2_rn(splat(cLL), *reinterpret_cast<const float2*>(grow + 64 * k));
            const unsigned b = *reinterpret_cast<const unsigned short*>(srow + 64 * k);
            const float2 sLH = __fadd2_rn(make_float2(code_float<0>(b, c.m0), code_float<8>(b, c.m8)), make_float2(code_bias<0>(), code_bias<8>()));
            const float2 sHL = __fadd2_rn(make_float2(code_float<2>(b, c.m2), code_float<10>(b, c.m10)), make_float2(code_bias<2>(), code_bias<10>()));
            const float2 sHH = __fadd2_rn(make_float2(code_float<4>(b, c.m4), code_float<12>(b, c.m12)), make_float2(code_bias<4>(), code_bias<12>()));
            const float2 eHL = __fmul2_rn(splat(cD), sHL);
            const float2 p = __ffma2_rn(splat(cD), sHL, t0), m = __fadd2_rn(t0, make_float2(-eHL.x, -eHL.y));        // rows 2i, 2i + 1
            const float2 eHH = __fmul2_rn(splat(cD), sHH);
            const float2 q = __ffma2_rn(splat(cD), sLH, eHH), rr = __ffma2_rn(splat(cD), sLH, make_float2(-eHH.x, -eHH.y));
            const float4 o0 = make_float4(p.x + q.x, p.x - q.x, p.y + q.y, p.y - q.y);
            const float4 o1 = make_float4(m.x + rr.x, m.x - rr.x, m.y + rr.y, m.y - rr.y);
            if (kToSmem) {
                *reinterpret_cast<float4*>(orow_s + 128 * k) = o0;
                *reinterpret_cast<float4*>(orow_s + 2 * wj + 128 * k) = o1;
            } else {
                __stcs(reinterpret_cast<float4*>(orow_g + 128 * k), o0);
                __stcs(reinterpret_cast<float4*>(orow_g + out_ld + 128 * k), o1);
            }
        }
        grow += wj;
        srow += wj;
        orow_s += 2 * (2 * wj);
        orow_g += 2 * out_ld;
    }
}

template <int K1, bool kTwo, bool kHasLL>
__global__ void __launch_bounds__(kDb2MaxThreads, 1) haar_synthesis_kernel(Db2InvArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int NW = kDb2MaxThreads / 32 - 1;
    constexpr int K2 = kTwo ? K1 / 2 : 1;
    const int NW2 = kTwo ? a.nw2 : 0, NW1 = NW - NW2;
    const int H = a.H, W = a.W, h2 = H >> 1, w2 = W >> 1, h4 = H >> 2, w4 = W >> 2, R = a.R, S = a.stages;
    const Db2InvLayout lay = db2_inv_layout(W, R, kTwo, kHasLL, S, true);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + lay.g1buf + 2 * lay.g1bytes);
    uint64_t* empty = full + S;
    uint64_t* g1_full = empty + S;
    uint64_t* g1_empty = g1_full + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], NW);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&g1_full[b], NW2 > 0 ? NW2 : 1);
            mbar_init(&g1_empty[b], NW1);
        }
        fence_mbar_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    Db2Pieces pieces((long long)a.nmaps * h2, h2, R, a.rr != 0);
    long long m;
    int i_first, len;
    if (warp == NW) {
        if (lane == 0) {
            for (int n = 0; pieces.next(m, i_first, len); ++n) {
                const int s = n % S;
                if (n >= S) mbar_wait(&empty[s], ((n / S) & 1) ^ 1);
                unsigned char* dst = smem + size_t(s) * lay.stage;
                uint32_t bytes = uint32_t(len) * w2;
                if (kTwo) {
                    const int lo2 = i_first >> 1, n2 = ((i_first + len - 1) >> 1) - lo2 + 1;
                    bytes += uint32_t(n2) * w4 * (kHasLL ? 5u : 1u);
                    mbar_arrive_expect_tx(&full[s], bytes);
                    if (kHasLL)
                        bulk_rows(dst + lay.g, a.g + (m * (long long)h4 + lo2) * w4, uint32_t(n2) * w4 * 4u, &full[s], 0, false);
                    bulk_rows(dst + lay.s2, a.sg2 + (m * (long long)h4 + lo2) * w4, uint32_t(n2) * w4, &full[s], 0, false);
                } else {
                    if (kHasLL) bytes += uint32_t(len) * w2 * 4u;
                    mbar_arrive_expect_tx(&full[s], bytes);
                    if (kHasLL) bulk_rows(dst + lay.g, a.g + (m * (long long)h2 + i_first) * w2, uint32_t(len) * w2 * 4u, &full[s], 0, false);
                }
                bulk_rows(dst + lay.s1, a.sg1 + (m * (long long)h2 + i_first) * w2, uint32_t(len) * w2, &full[s], 0, false);
            }
        }
        __syncwarp();
        if (a.loss && blockIdx.x == 0) {                    // fixed-order sum of the loss partials, once this CTA's loads are queued
            double s = 0.0;
            for (int i = lane; i < a.n_partials; i += 32) s += a.partial[i];
            s = warp_sum(s);
            if (lane == 0) a.loss[0] = float(s);
        }
    } else if (warp < NW1) {
        const float gs = a.upstream ? __ldg(a.upstream) : 1.0f;
        const SynConst c = make_syn_const(a.sc1, gs, a.magic);
        const float cLL = 0.5f * gs, cD = 0.5f * a.sc1 * gs;
        for (int n = 0; pieces.next(m, i_first, len); ++n) {
            const int s = n % S;
            const int st = s * lay.stage;
            mbar_wait(&full[s], (n / S) & 1);
            float* obase = a.out + m * (long long)H * W + (long long)(2 * i_first) * W;
            const int seg1 = (len + NW1 - 1) / NW1;
            const int ra1 = warp * seg1, rb1 = min(ra1 + seg1, len);
            if (!kTwo) {
                if (ra1 < rb1) haar_inv_rows<K1, kHasLL, false>(st + lay.g, st + lay.s1, ra1, rb1, cLL, cD, c, 0, obase, W, lane);
            } else {
                const int b = n & 1, k = n >> 1;
                mbar_wait(&g1_full[b], k & 1);
                const int shift = i_first - 2 * (i_first >> 1);                 // the piece's first row inside its dL/dLL1 buffer
                if (ra1 < rb1)
                    haar_inv_rows<K1, true, false>(lay.g1buf + b * lay.g1bytes + shift * w2 * 4, st + lay.s1, ra1, rb1, cLL, cD, c, 0, obase, W, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&g1_empty[b]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    } else if (kTwo) {
        const SynConst c = make_syn_const(a.sc2, 1.0f, a.magic);
        const float cLL = 0.5f, cD = 0.5f * a.sc2;
        const int w = warp - NW1;
        for (int n = 0; pieces.next(m, i_first, len); ++n) {
            const int s = n % S;
            const int st = s * lay.stage;
            const int b = n & 1, k = n >> 1;
            mbar_wait(&full[s], (n / S) & 1);
            if (k > 0) mbar_wait(&g1_empty[b], (k & 1) ^ 1);
            const int rows2 = ((i_first + len - 1) >> 1) - (i_first >> 1) + 1;
            const int nw2d = kTwo ? NW2 : 1;
            const int seg2 = (rows2 + nw2d - 1) / nw2d;
            const int ra2 = w * seg2, rb2 = min(ra2 + seg2, rows2);
            if (ra2 < rb2) haar_inv_rows<K2, kHasLL, true>(st + lay.g, st + lay.s2, ra2, rb2, cLL, cD, c, lay.g1buf + b * lay.g1bytes, nullptr, 0, lane);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&g1_full[b]);
                mbar_arrive(&empty[s]);
            }
        }
    }
}

template <typename Kernel, typename Args>
cudaError_t launch_db2(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t stream, const Args& args, bool pdl) {
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> opted_in;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_pair(reinterpret_cast<const void*>(kernel), dev);
        auto it = opted_in.find(key);
        if (it == opted_in.end() || it->second < smem) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            opted_in[key] = smem;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args);
}

int divisor_le(int n, int cap) {
    for (int d = std::min(n, cap); d > 1; --d)
        if (n % d == 0) return d;
    return 1;
}

size_t db2_fwd_smem(int W, int R, int S, bool two, bool haar = false) {
    const int halo = haar ? 0 : 2;
    const size_t rows_in = two ? 4 * R + 3 * halo : 2 * R + halo;
    const size_t ll1 = two ? size_t(2 * R + halo) * (W / 2) * 4 : 0;
    return S * rows_in * W * 4 + 2 * ll1 + size_t(2 * S + 4) * sizeof(uint64_t);
}

}  // namespace

int g_wavelet_db2 = 1;          // diagnostics: 0 = the round-1 level kernels (wavelet_tiles.cu) for db2 as well
int g_wavelet_db2_two = 1;      // diagnostics: 0 = one level per pass only
int g_wavelet_haar_passes = 1;   // Haar: streamed levels through the same passes (1) or the round-1 level kernels / band kernel only (0)
int g_wavelet_db2_deep = 0;     // 1: keep peeling levels with the factored passes as long as the band's width allows, resident stage only for the rest

int g_wavelet_db2_rf = 0, g_wavelet_db2_ri = 0, g_wavelet_db2_nw2 = 0;
int g_wavelet_db2_rr = 1;       // synthesis pieces dealt round-robin (1) or one contiguous row range per CTA (0)     // diagnostics: overrides of the geometry below (0 = automatic)

namespace {
// rows per thread task that minimise rounds * (rows + ~1.5 rows of task prologue) for `rows` rows x `pairs` column groups on `threads` threads
int best_seg(int rows, int pairs, int threads) {
    int best = rows;
    double best_cost = 1e30;
    for (int seg = 1; seg <= rows; ++seg) {
        const int tasks = ((rows + seg - 1) / seg) * pairs;
        const int rounds = (tasks + threads - 1) / threads;
        const double cost = rounds * (seg + 1.5);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = seg; }
    }
    return best;
}
}  // namespace

// Geometry of a pass over H x W planes; false: shape not taken (the caller falls back to wavelet_tiles.cu / wavelet_stream.cu).
//   R_fwd: rows of the pass's last level per analysis strip, S_fwd ring depth, NC_fwd consumer threads (both groups)
//   R_inv: level-1 coefficient rows per synthesis piece, S_inv ring depth
bool wavelet_db2_pass(int H, int W, bool two, bool has_ll, int* R_fwd, int* S_fwd, int* NC_fwd, int* R_inv, int* S_inv, bool haar) {
    if (!g_wavelet_db2 || (two && !g_wavelet_db2_two)) return false;
    if (haar && !g_wavelet_haar_passes) return false;
    if (W != 128 && W != 256 && W != 512 && W != 1024) return false;    // warp-per-row synthesis: W / 128 chunks of 64 sites, unrolled
    if (haar && two && W < 256) return false;                           // (32-site rows exist for db2 only)
    if (H % (two ? 4 : 2) || H < (two ? 16 : 4)) return false;
    const int hl = two ? H / 4 : H / 2;
    int R = divisor_le(hl, g_wavelet_db2_rf > 0 ? g_wavelet_db2_rf : (two ? 8 : 16));
    while (R > 1 && db2_fwd_smem(W, R, 2, two, haar) > size_t(kDb2Smem)) R = divisor_le(hl, R - 1);
    if (db2_fwd_smem(W, R, 2, two, haar) > size_t(kDb2Smem)) return false;
    int S = 2;
    while (S < 4 && db2_fwd_smem(W, R, S + 1, two, haar) <= size_t(kDb2Smem)) ++S;
    *R_fwd = R; *S_fwd = S; *NC_fwd = kDb2MaxThreads - 32;
    // measured at 64 x 2 x 1024^2, J = 2, db2: pieces of 24 (what fits) / 20 / 16 / 12 rows: 219.6 / 215.8 / 214.8 / 234.1 us
    int Ri = g_wavelet_db2_ri > 0 ? g_wavelet_db2_ri : (two ? (W >= 1024 && !haar ? 16 : 48) : 64);
    Ri = std::min(Ri, H / 2);
    auto inv_smem = [&](int r, int s) {
        const Db2InvLayout lay = db2_inv_layout(W, r, two, has_ll, s, haar);
        return size_t(lay.g1buf) + 2 * size_t(lay.g1bytes) + size_t(2 * s + 4) * sizeof(uint64_t);
    };
    while (Ri > 4 && inv_smem(Ri, 3) > size_t(kDb2Smem)) Ri -= 4;
    int Si = 4;
    while (Si > 2 && inv_smem(Ri, Si) > size_t(kDb2Smem)) --Si;
    if (inv_smem(Ri, Si) > size_t(kDb2Smem)) return false;
    *R_inv = Ri; *S_inv = Si;
    return true;
}

cudaError_t launch_db2_analysis(const float* x, float* ll, unsigned char* sg1, unsigned char* sg2, int nmaps, int H, int W, bool two,
                                float sc1, float sc2, bool grad, bool pdl_wait, double* partial, int sm_count, cudaStream_t stream,
                                int* n_partials, bool haar) {
    int Rf, Sf, NCf, Ri, Si;
    if (!wavelet_db2_pass(H, W, two, true, &Rf, &Sf, &NCf, &Ri, &Si, haar)) return cudaErrorInvalidValue;
    Db2FwdArgs a;
    a.x = x; a.ll = ll; a.sg1 = sg1; a.sg2 = sg2; a.H = H; a.W = W; a.nmaps = nmaps; a.R = Rf; a.stages = Sf; a.pdl_wait = pdl_wait ? 1 : 0;
    a.sc1 = sc1; a.sc2 = sc2; a.partial = partial;
    const int nw = NCf / 32, halo = haar ? 0 : 2;
    a.nw2 = two ? (g_wavelet_db2_nw2 > 0 ? std::min(g_wavelet_db2_nw2, nw - 1) : (W >= 1024 && !haar ? 6 : 4)) : 0;     // measured: db2 1024^2 J=2 236 -> 227 us with 6; Haar: 203 -> 196 us with 4
    a.seg1 = best_seg(two ? 2 * Rf + halo : Rf, W / 4, (nw - a.nw2) * 32);
    a.seg2 = two ? best_seg(Rf, W / 8, a.nw2 * 32) : 1;
    const long long T = (long long)nmaps * ((two ? H / 4 : H / 2) / Rf);
    const int grid = int(std::min<long long>(sm_count, T));
    const size_t smem = db2_fwd_smem(W, Rf, Sf, two, haar);
    *n_partials = grid;
    const int threads = NCf + 32;
#define WTPSE_DB2_FWD(TWO, GRAD)                                                                                              \
    (haar ? launch_db2(db2_analysis_kernel<TWO, GRAD, true>, grid, threads, smem, stream, a, pdl_wait)                         \
          : launch_db2(db2_analysis_kernel<TWO, GRAD, false>, grid, threads, smem, stream, a, pdl_wait))
    if (two) return grad ? WTPSE_DB2_FWD(true, true) : WTPSE_DB2_FWD(true, false);
    return grad ? WTPSE_DB2_FWD(false, true) : WTPSE_DB2_FWD(false, false);
#undef WTPSE_DB2_FWD
}

namespace {
template <int K1>
cudaError_t launch_db2_synthesis_k(const Db2InvArgs& a, bool two, bool has_ll, bool haar, int grid, size_t smem, cudaStream_t stream) {
    if (haar) {
        if (two) {
            if constexpr (K1 >= 2) {
                return has_ll ? launch_db2(haar_synthesis_kernel<K1, true, true>, grid, kDb2MaxThreads, smem, stream, a, true)
                              : launch_db2(haar_synthesis_kernel<K1, true, false>, grid, kDb2MaxThreads, smem, stream, a, true);
            } else {
                return cudaErrorInvalidValue;
            }
        }
        return has_ll ? launch_db2(haar_synthesis_kernel<K1, false, true>, grid, kDb2MaxThreads, smem, stream, a, true)
                      : launch_db2(haar_synthesis_kernel<K1, false, false>, grid, kDb2MaxThreads, smem, stream, a, true);
    }
    if (two)
        return has_ll ? launch_db2(db2_synthesis_kernel<K1, true, true>, grid, kDb2MaxThreads, smem, stream, a, true)
                      : launch_db2(db2_synthesis_kernel<K1, true, false>, grid, kDb2MaxThreads, smem, stream, a, true);
    return has_ll ? launch_db2(db2_synthesis_kernel<K1, false, true>, grid, kDb2MaxThreads, smem, stream, a, true)
                  : launch_db2(db2_synthesis_kernel<K1, false, false>, grid, kDb2MaxThreads, smem, stream, a, true);
}
}  // namespace

// g: gradient of the low-low band entering the pass (level 2's when `two`), nullptr / has_ll == false: zero
cudaError_t launch_db2_synthesis(const float* g, const unsigned char* sg1, const unsigned char* sg2, float* out, int nmaps, int H, int W,
                                 bool two, bool has_ll, float sc1, float sc2, const float* upstream, const double* partial, int n_partials,
                                 float* loss, int sm_count, cudaStream_t stream, bool haar) {
    int Rf, Sf, NCf, Ri, Si;
    if (!wavelet_db2_pass(H, W, two, has_ll, &Rf, &Sf, &NCf, &Ri, &Si, haar)) return cudaErrorInvalidValue;
    Db2InvArgs a;
    a.g = g; a.sg1 = sg1; a.sg2 = sg2; a.out = out; a.H = H; a.W = W; a.nmaps = nmaps; a.R = Ri; a.stages = Si; a.sc1 = sc1; a.sc2 = sc2;
    a.upstream = upstream; a.partial = partial; a.n_partials = n_partials; a.loss = loss;
    a.magic[0] = code_magic<0>(); a.magic[1] = code_magic<2>(); a.magic[2] = code_magic<4>();
    a.magic[3] = code_magic<8>(); a.magic[4] = code_magic<10>(); a.magic[5] = code_magic<12>();
    a.nw2 = two ? (g_wavelet_db2_nw2 > 0 ? std::min(g_wavelet_db2_nw2, kDb2MaxThreads / 32 - 2) : (W >= 1024 && !haar ? 6 : 4)) : 0;
    // measured (same box): one-level passes, 1024^2 J = 1: 192.8 -> 187.2 us with round-robin pieces; two-level passes: 512^2 J = 4 52.0 -> 53.6 us,
    // 1024^2 J = 2 222.7 -> 226.9 us (their pieces carry recomputed overlap rows of level 2, and the hand-over favours long runs)
    // ... and only with enough pieces for the rounds to come out even (512^2 J = 1: 256 pieces on 148 CTAs, 35.9 -> 37.0 us)
    a.rr = (g_wavelet_db2_rr && !two && (long long)nmaps * ((H / 2 + Ri - 1) / Ri) >= 6ll * sm_count) ? 1 : 0;
    const long long rows = (long long)nmaps * (H / 2);
    const int grid = int(std::min<long long>(sm_count, std::max<long long>(1, rows / 4)));
    const Db2InvLayout lay = db2_inv_layout(W, Ri, two, has_ll, Si, haar);
    const size_t smem = size_t(lay.g1buf) + 2 * size_t(lay.g1bytes) + size_t(2 * Si + 4) * sizeof(uint64_t);
    switch (W / 128) {
        case 1: return launch_db2_synthesis_k<1>(a, two, has_ll, haar, grid, smem, stream);
        case 2: return launch_db2_synthesis_k<2>(a, two, has_ll, haar, grid, smem, stream);
        case 4: return launch_db2_synthesis_k<4>(a, two, has_ll, haar, grid, smem, stream);
        case 8: return launch_db2_synthesis_k<8>(a, two, has_ll, haar, grid, smem, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace wtpse

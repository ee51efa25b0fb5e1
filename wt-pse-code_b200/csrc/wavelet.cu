// Track W: multi-level 2-D orthonormal DWT / IDWT (Haar, db2; periodic extension) and an L1
// detail-coefficient shape loss with its backward.  PARITY UNPINNED: the reference contains no wavelet code
// (SURVEY.md section 0); the specification these kernels implement is this repository's own
// (oracle/wavelet_np.py) and is checked by mathematical identities, not against the reference.
//
// One kernel per level, each level one pass: separable low/high-pass filtering along W and H plus the 2x
// decimation are fused -- a thread reads the TAPS x TAPS input patch of one half-resolution site (64-bit
// coalesced loads; the 4x patch overlap of db2 is served by L1/L2) and writes the four sub-band values.
// HBM-bound: level j moves h_j*w_j floats in and out, 4/3 * (read + write) of the map over all levels.
// The loss variant never stores coefficients: it writes w_j * sign(d) / norm straight into the gradient-
// coefficient buffer and accumulates |d| into per-block partials (fixed-order final sum), so the backward
// is ONE inverse transform (the synthesis bank is the adjoint) scaled by the upstream gradient.
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "wavelet_bank.cuh"

namespace wtpse {

namespace {

constexpr int kBx = 32, kBy = 8;

struct LevelArgs {
    const float* in;  long long in_map; int in_ld;      // h x w block of every map
    float* ll;        long long ll_map; int ll_ld;      // (h/2) x (w/2) low-low output
    float* det;       long long det_map; int det_ld;    // Mallat buffer whose top-left h x w block receives LH / HL / HH
    int h, w, nmaps;
    float det_scale;        // loss mode: w_j / (3 * h/2 * w/2 * nmaps)
    int zero_ll;            // loss mode, last level: the LL quadrant of the gradient buffer is zero
    double* partial;        // loss mode: per-block sums of |detail| * det_scale
    int partial_base;
};

// One thread -> TWO horizontally adjacent half-resolution sites (j, j+1): their input patches overlap in TAPS-2
// columns, so a row costs one 128-bit load (+ one 64-bit load for db2) instead of two / four 64-bit loads, and every
// sub-band is written as a float2.  Requires w % 4 == 0 (the launcher falls back to dwt_level_kernel otherwise).
template <int TAPS, bool kLoss>
__global__ void __launch_bounds__(kBx * kBy) dwt_level2_kernel(LevelArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int h2 = a.h >> 1, w2 = a.w >> 1;
    const int j = 2 * (blockIdx.x * kBx + threadIdx.x), i = blockIdx.y * kBy + threadIdx.y, m = blockIdx.z;
    float absum = 0.f;
    if (i < h2 && j < w2) {
        const float* src = a.in + (long long)m * a.in_map;
        float lo0[TAPS], hi0[TAPS], lo1[TAPS], hi1[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            int r = 2 * i + k;
            if (r >= a.h) r -= a.h;
            const float* row = src + (long long)r * a.in_ld;
            float x[TAPS + 2];
            const float4 v = __ldg(reinterpret_cast<const float4*>(row + 2 * j));      // 2j % 4 == 0, w % 4 == 0
            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
            if (TAPS == 4) {
                int c = 2 * j + 4;
                if (c >= a.w) c -= a.w;
                const float2 u = __ldg(reinterpret_cast<const float2*>(row + c));
                x[TAPS] = u.x; x[TAPS + 1] = u.y;
            }
            float s0 = 0.f, d0 = 0.f, s1 = 0.f, d1 = 0.f;
#pragma unroll
            for (int l = 0; l < TAPS; ++l) {
                s0 = fmaf(Bank<TAPS>::h(l), x[l], s0); d0 = fmaf(Bank<TAPS>::g(l), x[l], d0);
                s1 = fmaf(Bank<TAPS>::h(l), x[l + 2], s1); d1 = fmaf(Bank<TAPS>::g(l), x[l + 2], d1);
            }
            lo0[k] = s0; hi0[k] = d0; lo1[k] = s1; hi1[k] = d1;
        }
        float2 LL = make_float2(0.f, 0.f), LH = LL, HL = LL, HH = LL;
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            LL.x = fmaf(Bank<TAPS>::h(k), lo0[k], LL.x); LL.y = fmaf(Bank<TAPS>::h(k), lo1[k], LL.y);
            LH.x = fmaf(Bank<TAPS>::h(k), hi0[k], LH.x); LH.y = fmaf(Bank<TAPS>::h(k), hi1[k], LH.y);
            HL.x = fmaf(Bank<TAPS>::g(k), lo0[k], HL.x); HL.y = fmaf(Bank<TAPS>::g(k), lo1[k], HL.y);
            HH.x = fmaf(Bank<TAPS>::g(k), hi0[k], HH.x); HH.y = fmaf(Bank<TAPS>::g(k), hi1[k], HH.y);
        }
        if (a.ll) *reinterpret_cast<float2*>(a.ll + (long long)m * a.ll_map + (long long)i * a.ll_ld + j) = LL;
        float* det = a.det + (long long)m * a.det_map;
        if (kLoss) {
            const float sc = a.det_scale;
            absum = (fabsf(LH.x) + fabsf(LH.y) + fabsf(HL.x) + fabsf(HL.y) + fabsf(HH.x) + fabsf(HH.y)) * sc;
            auto sg = [&](float v) { return v > 0.f ? sc : (v < 0.f ? -sc : 0.f); };
            LH = make_float2(sg(LH.x), sg(LH.y));
            HL = make_float2(sg(HL.x), sg(HL.y));
            HH = make_float2(sg(HH.x), sg(HH.y));
            if (a.zero_ll) *reinterpret_cast<float2*>(det + (long long)i * a.det_ld + j) = make_float2(0.f, 0.f);
        }
        *reinterpret_cast<float2*>(det + (long long)i * a.det_ld + w2 + j) = LH;
        *reinterpret_cast<float2*>(det + (long long)(h2 + i) * a.det_ld + j) = HL;
        *reinterpret_cast<float2*>(det + (long long)(h2 + i) * a.det_ld + w2 + j) = HH;
    }
    if (kLoss) {
        __shared__ double red[kBx * kBy / 32];
        double s = warp_sum(double(absum));
        const int t = threadIdx.y * kBx + threadIdx.x;
        if ((t & 31) == 0) red[t >> 5] = s;
        __syncthreads();
        if (t == 0) {
            double tot = 0.0;
#pragma unroll
            for (int q = 0; q < kBx * kBy / 32; ++q) tot += red[q];
            a.partial[a.partial_base + (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
        }
    }
}

template <int TAPS, bool kLoss>
__global__ void __launch_bounds__(kBx * kBy) dwt_level_kernel(LevelArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int h2 = a.h >> 1, w2 = a.w >> 1;
    const int j = blockIdx.x * kBx + threadIdx.x, i = blockIdx.y * kBy + threadIdx.y, m = blockIdx.z;
    float absum = 0.f;
    if (i < h2 && j < w2) {
        const float* src = a.in + (long long)m * a.in_map;
        float lo[TAPS], hi[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            int r = 2 * i + k;
            if (r >= a.h) r -= a.h;
            const float* row = src + (long long)r * a.in_ld;
            float x[TAPS];
#pragma unroll
            for (int l = 0; l < TAPS; l += 2) {
                int c = 2 * j + l;
                if (c >= a.w) c -= a.w;                          // w is even: the pair never straddles the wrap
                const float2 v = __ldg(reinterpret_cast<const float2*>(row + c));
                x[l] = v.x;
                x[l + 1] = v.y;
            }
            float s = 0.f, d = 0.f;
#pragma unroll
            for (int l = 0; l < TAPS; ++l) { s = fmaf(Bank<TAPS>::h(l), x[l], s); d = fmaf(Bank<TAPS>::g(l), x[l], d); }
            lo[k] = s;
            hi[k] = d;
        }
        float LL = 0.f, LH = 0.f, HL = 0.f, HH = 0.f;
#pragma unroll
        for (int k = 0; k < TAPS; ++k) {
            LL = fmaf(Bank<TAPS>::h(k), lo[k], LL);
            LH = fmaf(Bank<TAPS>::h(k), hi[k], LH);
            HL = fmaf(Bank<TAPS>::g(k), lo[k], HL);
            HH = fmaf(Bank<TAPS>::g(k), hi[k], HH);
        }
        if (a.ll) a.ll[(long long)m * a.ll_map + (long long)i * a.ll_ld + j] = LL;
        float* det = a.det + (long long)m * a.det_map;
        if (kLoss) {
            absum = (fabsf(LH) + fabsf(HL) + fabsf(HH)) * a.det_scale;
            auto sg = [&](float v) { return v > 0.f ? a.det_scale : (v < 0.f ? -a.det_scale : 0.f); };
            det[(long long)i * a.det_ld + w2 + j] = sg(LH);
            det[(long long)(h2 + i) * a.det_ld + j] = sg(HL);
            det[(long long)(h2 + i) * a.det_ld + w2 + j] = sg(HH);
            if (a.zero_ll) det[(long long)i * a.det_ld + j] = 0.f;
        } else {
            det[(long long)i * a.det_ld + w2 + j] = LH;
            det[(long long)(h2 + i) * a.det_ld + j] = HL;
            det[(long long)(h2 + i) * a.det_ld + w2 + j] = HH;
        }
    }
    if (kLoss) {
        __shared__ double red[kBx * kBy / 32];
        double s = warp_sum(double(absum));
        const int t = threadIdx.y * kBx + threadIdx.x;
        if ((t & 31) == 0) red[t >> 5] = s;
        __syncthreads();
        if (t == 0) {
            double tot = 0.0;
#pragma unroll
            for (int q = 0; q < kBx * kBy / 32; ++q) tot += red[q];
            a.partial[a.partial_base + (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = tot;
        }
    }
}

struct SynthArgs {
    const float* ll;  long long ll_map; int ll_ld;      // (h/2) x (w/2)
    const float* det; long long det_map; int det_ld;    // Mallat buffer, detail bands of the h x w block
    float* out;       long long out_map; int out_ld;    // h x w
    int h, w, nmaps;
    const float* scale;     // optional device scalar multiplied into the output (upstream gradient)
};

// adjoint of dwt_level_kernel: out[2i'+pr][2j'+pc] = sum_{m,n} f_r[2m+pr] f_c[2n+pc] * band[i'-m][j'-n]
template <int TAPS>
__global__ void __launch_bounds__(kBx * kBy) idwt_level_kernel(SynthArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int h2 = a.h >> 1, w2 = a.w >> 1;
    const int jp = blockIdx.x * kBx + threadIdx.x, ip = blockIdx.y * kBy + threadIdx.y, mp = blockIdx.z;
    if (ip >= h2 || jp >= w2) return;
    const float* ll = a.ll + (long long)mp * a.ll_map;
    const float* det = a.det + (long long)mp * a.det_map;
    float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;
#pragma unroll
    for (int m = 0; m < TAPS / 2; ++m) {
        int i = ip - m;
        if (i < 0) i += h2;
#pragma unroll
        for (int n = 0; n < TAPS / 2; ++n) {
            int j = jp - n;
            if (j < 0) j += w2;
            const float LL = __ldg(ll + (long long)i * a.ll_ld + j);
            const float LH = __ldg(det + (long long)i * a.det_ld + w2 + j);
            const float HL = __ldg(det + (long long)(h2 + i) * a.det_ld + j);
            const float HH = __ldg(det + (long long)(h2 + i) * a.det_ld + w2 + j);
            const float hr0 = Bank<TAPS>::h(2 * m), hr1 = Bank<TAPS>::h(2 * m + 1);
            const float gr0 = Bank<TAPS>::g(2 * m), gr1 = Bank<TAPS>::g(2 * m + 1);
            const float hc0 = Bank<TAPS>::h(2 * n), hc1 = Bank<TAPS>::h(2 * n + 1);
            const float gc0 = Bank<TAPS>::g(2 * n), gc1 = Bank<TAPS>::g(2 * n + 1);
            // along columns first: t_L = low-row content, t_H = high-row content for the two column parities
            const float tL0 = fmaf(hc0, LL, gc0 * LH), tL1 = fmaf(hc1, LL, gc1 * LH);
            const float tH0 = fmaf(hc0, HL, gc0 * HH), tH1 = fmaf(hc1, HL, gc1 * HH);
            o00 += fmaf(hr0, tL0, gr0 * tH0);
            o01 += fmaf(hr0, tL1, gr0 * tH1);
            o10 += fmaf(hr1, tL0, gr1 * tH0);
            o11 += fmaf(hr1, tL1, gr1 * tH1);
        }
    }
    const float sc = a.scale ? __ldg(a.scale) : 1.0f;
    float* out = a.out + (long long)mp * a.out_map;
    *reinterpret_cast<float2*>(out + (long long)(2 * ip) * a.out_ld + 2 * jp) = make_float2(o00 * sc, o01 * sc);
    *reinterpret_cast<float2*>(out + (long long)(2 * ip + 1) * a.out_ld + 2 * jp) = make_float2(o10 * sc, o11 * sc);
}

__global__ void __launch_bounds__(1024) wavelet_loss_final_kernel(const double* __restrict__ partial, int n,
                                                                  float* __restrict__ loss) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __shared__ double red[32];
    // four independent accumulators per thread: the FP64 add has ~50 cycles of dependent latency on B200
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int i = threadIdx.x;
    for (; i + 3 * 1024 < n; i += 4 * 1024) {
        s0 += partial[i]; s1 += partial[i + 1024]; s2 += partial[i + 2048]; s3 += partial[i + 3072];
    }
    for (; i < n; i += 1024) s0 += partial[i];
    double s = warp_sum((s0 + s1) + (s2 + s3));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = warp_sum(red[threadIdx.x]);
        if (threadIdx.x == 0) loss[0] = float(s);
    }
}

template <typename Kernel, typename Args>
cudaError_t launch_pss(Kernel kernel, dim3 grid, dim3 block, cudaStream_t stream, const Args& args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args);
}

dim3 level_grid(int h, int w, int nmaps) { return dim3((w / 2 + kBx - 1) / kBx, (h / 2 + kBy - 1) / kBy, nmaps); }

}  // namespace

size_t wavelet_scratch_floats(long long nmaps, int H, int W) {
    // per-level path: ping-pong low-low buffers (H/2 x W/2 and H/4 x W/4 per map).  Streamed plan: the low-low bands
    // of all peeled levels (< N/3 floats) + their packed sign planes (< N/3 bytes) + alignment.
    const size_t n = size_t(nmaps) * H * W;
    return n / 3 + n / 12 + 16 * 256 + size_t(nmaps) * (size_t(H / 2) * (W / 2) + size_t(H / 4) * (W / 4)) / 64;
}

size_t wavelet_partial_doubles(long long nmaps, int H, int W, int J) {
    size_t n = 0;
    for (int j = 0; j < J; ++j) {
        const dim3 g = level_grid(H >> j, W >> j, int(nmaps));
        n += size_t(g.x) * g.y * g.z;
    }
    return n;
}

// loss == nullptr: coefficients -> coef (Mallat layout, same shape as x).  Otherwise loss + gradient coefficients -> coef.
cudaError_t launch_dwt(const float* x, int nmaps, int H, int W, int taps, int J, float* coef, float* scratch,
                       const float* weights_host, float* loss, double* partial, cudaStream_t stream) {
    const long long map = (long long)H * W;
    float* s0 = scratch;
    float* s1 = scratch + size_t(nmaps) * (H / 2) * (W / 2);
    const bool loss_mode = loss != nullptr;
    int pbase = 0;
    for (int j = 0; j < J; ++j) {
        const int h = H >> j, w = W >> j;
        LevelArgs a;
        if (j == 0) { a.in = x; a.in_map = map; a.in_ld = W; }
        else {
            const float* prev = ((j - 1) % 2 == 0) ? s0 : s1;
            a.in = prev; a.in_map = (long long)h * w; a.in_ld = w;
        }
        const bool last = (j == J - 1);
        if (last && !loss_mode) { a.ll = coef; a.ll_map = map; a.ll_ld = W; }
        else if (last) { a.ll = nullptr; a.ll_map = 0; a.ll_ld = 0; }
        else { a.ll = (j % 2 == 0) ? s0 : s1; a.ll_map = (long long)(h / 2) * (w / 2); a.ll_ld = w / 2; }
        a.det = coef; a.det_map = map; a.det_ld = W;
        a.h = h; a.w = w; a.nmaps = nmaps;
        a.det_scale = loss_mode ? weights_host[j] / (3.0f * float(h / 2) * float(w / 2) * float(nmaps)) : 0.f;
        a.zero_ll = (loss_mode && last) ? 1 : 0;
        a.partial = partial; a.partial_base = pbase;
        const bool wide = (w % 4 == 0) && (w >= 8);                                   // two sites per thread
        const dim3 b(kBx, kBy);
        const dim3 g = wide ? dim3((w / 4 + kBx - 1) / kBx, (h / 2 + kBy - 1) / kBy, nmaps) : level_grid(h, w, nmaps);
        cudaError_t e;
        if (loss_mode) {
            if (wide) e = taps == 2 ? launch_pss(dwt_level2_kernel<2, true>, g, b, stream, a) : launch_pss(dwt_level2_kernel<4, true>, g, b, stream, a);
            else e = taps == 2 ? launch_pss(dwt_level_kernel<2, true>, g, b, stream, a) : launch_pss(dwt_level_kernel<4, true>, g, b, stream, a);
            pbase += int(g.x * g.y * g.z);
        } else {
            if (wide) e = taps == 2 ? launch_pss(dwt_level2_kernel<2, false>, g, b, stream, a) : launch_pss(dwt_level2_kernel<4, false>, g, b, stream, a);
            else e = taps == 2 ? launch_pss(dwt_level_kernel<2, false>, g, b, stream, a) : launch_pss(dwt_level_kernel<4, false>, g, b, stream, a);
        }
        if (e != cudaSuccess) return e;
    }
    if (loss_mode) wavelet_loss_final_kernel<<<1, 1024, 0, stream>>>(partial, pbase, loss);
    return cudaGetLastError();
}

cudaError_t launch_wavelet_loss_final(const double* partial, int n, float* loss, cudaStream_t stream) {
    wavelet_loss_final_kernel<<<1, 1024, 0, stream>>>(partial, n, loss);
    return cudaGetLastError();
}

// Inverse (== adjoint) transform: coef (Mallat layout) -> x, optionally scaled by a device scalar.
cudaError_t launch_idwt(const float* coef, int nmaps, int H, int W, int taps, int J, float* x, float* scratch,
                        const float* scale, cudaStream_t stream) {
    const long long map = (long long)H * W;
    float* s0 = scratch;
    float* s1 = scratch + size_t(nmaps) * (H / 2) * (W / 2);
    for (int j = J - 1; j >= 0; --j) {
        const int h = H >> j, w = W >> j;
        SynthArgs a;
        if (j == J - 1) { a.ll = coef; a.ll_map = map; a.ll_ld = W; }
        else {
            const float* prev = (j % 2 == 0) ? s0 : s1;          // written by level j+1 below
            a.ll = prev; a.ll_map = (long long)(h / 2) * (w / 2); a.ll_ld = w / 2;
        }
        a.det = coef; a.det_map = map; a.det_ld = W;
        if (j == 0) { a.out = x; a.out_map = map; a.out_ld = W; a.scale = scale; }
        else {
            float* dst = ((j - 1) % 2 == 0) ? s0 : s1;
            a.out = dst; a.out_map = (long long)h * w; a.out_ld = w; a.scale = nullptr;
        }
        a.h = h; a.w = w; a.nmaps = nmaps;
        const dim3 g = level_grid(h, w, nmaps), b(kBx, kBy);
        const cudaError_t e = taps == 2 ? launch_pss(idwt_level_kernel<2>, g, b, stream, a) : launch_pss(idwt_level_kernel<4>, g, b, stream, a);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

}  // namespace wtpse

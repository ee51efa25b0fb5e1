// Training-mode BatchNorm2d (+ the ReLU that follows it) for the U-Net stages, channels-last, forward and backward.
//
// ConvD / ConvU / DoubleConv (algorithms.py:877-962, 398-413) are conv -> BatchNorm2d -> ReLU chains; in the train step
// cuDNN's NHWC batch-norm kernels run at about a third of the HBM roofline for these shapes (C = 16..256; 370 us for
// 15x16x512x512 where the three passes over 252 MB take 126 us at the measured peak), and ATen adds a separate in-place
// ReLU pass forward and a threshold_backward pass backward.  Here:
//   forward : pass 1  per-channel shifted sums  S1 = sum(x - K), S2 = sum((x - K)^2), K = the channel's first value
//                     (shifted-data variance: no cancellation when |mean| >> std), per-CTA partials, fixed order
//             final   mean, biased variance (float64 combine), invstd, running statistics (unbiased variance, momentum;
//                     the running mean may be offset by a folded convolution bias, see segmentation._conv_bn)
//             pass 2  y = relu?((x - mean) * (gamma * invstd) + beta)
//   backward: pass 1  sum(dy'), sum(dy' * (x - mean)),  dy' = dy * [y > 0]  with y recomputed from x exactly as in pass 2
//             final   dgamma, dbeta and the three per-channel coefficients of dx
//             pass 2  dx = gamma * invstd * (dy' - mean(dy') - (x - mean) * invstd^2 * mean(dy' * (x - mean)))
// Every thread owns one 128-bit channel quad (thread t -> quad t % (C/4)), so all accesses are coalesced 128-bit;
// reductions are two deterministic stages (no atomics).  HBM traffic: forward 2 reads + 1 write, backward 4 reads + 1 write.
#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ void acc4(float4& a, const float4& v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
__device__ __forceinline__ void fma4(float4& a, const float4& u, const float4& v) {
    a.x = fmaf(u.x, v.x, a.x); a.y = fmaf(u.y, v.y, a.y); a.z = fmaf(u.z, v.z, a.z); a.w = fmaf(u.w, v.w, a.w);
}

// y = (x - mean) * scale + beta, then the optional ReLU; written once so that forward and backward agree bit for bit
__device__ __forceinline__ float4 bn_out(const float4& x, const float4& mean, const float4& scale, const float4& beta) {
    return make_float4(fmaf(x.x - mean.x, scale.x, beta.x), fmaf(x.y - mean.y, scale.y, beta.y),
                       fmaf(x.z - mean.z, scale.z, beta.z), fmaf(x.w - mean.w, scale.w, beta.w));
}

// Block-level tree over the `lanes` threads that share a channel quad; result valid in the threads with r == 0.
__device__ __forceinline__ float4 quad_tree(float4 v, float4* red, int tid, int C4, int lanes, int r) {
    red[tid] = v;
    __syncthreads();
    for (int s = lanes >> 1; s > 0; s >>= 1) {
        if (r < s) acc4(red[tid], red[tid + s * C4]);
        __syncthreads();
    }
    const float4 out = red[tid];
    __syncthreads();
    return out;
}

// partial layout: [2][nblocks][C]
template <bool kBackward>
__global__ void __launch_bounds__(kThreads)
bn_partial_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long npix, int C4,
                  const float* __restrict__ mean, const float* __restrict__ scale, const float* __restrict__ beta, int relu,
                  float* __restrict__ partial) {
    __shared__ float4 red[kThreads];
    const int tid = threadIdx.x;
    const int lanes = kThreads / C4;
    const int q = tid % C4, r = tid / C4;
    // grid-stride over pixel rows: at any moment the whole grid reads one contiguous window per stream (contiguous
    // per-CTA ranges made 2 x 592 scattered streams: 142 us instead of 94 us for the backward sums at 15x16x512x512)
    const long long stride = (long long)gridDim.x * lanes;
    const long long p1 = npix;
    // forward: the shift K = first pixel of the channel; backward: the batch mean
    const float4 ctr = kBackward ? ldg4(mean + 4 * q) : ldg4(x + 4 * q);
    float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), bt = sc;
    if (kBackward) { sc = ldg4(scale + 4 * q); bt = ldg4(beta + 4 * q); }
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1, b1 = a1, b2 = a1;
    auto step = [&](long long p, float4& s1, float4& s2) {
        const float4 xv = ldg4(x + (p * C4 + q) * 4);
        const float4 d = sub4(xv, ctr);
        if (kBackward) {
            float4 g = ldg4(dy + (p * C4 + q) * 4);
            if (relu) {
                const float4 y = bn_out(xv, ctr, sc, bt);
                g.x = y.x <= 0.f ? 0.f : g.x; g.y = y.y <= 0.f ? 0.f : g.y; g.z = y.z <= 0.f ? 0.f : g.z; g.w = y.w <= 0.f ? 0.f : g.w;
            }
            acc4(s1, g);
            fma4(s2, g, d);
        } else {
            acc4(s1, d);
            fma4(s2, d, d);
        }
    };
    // four independent pixel rows per iteration: four 128-bit loads (eight in the backward) in flight per thread
    float4 c1 = a1, c2 = a1, d1 = a1, d2 = a1;
    long long p = (long long)blockIdx.x * lanes + r;
    for (; p + 3 * stride < p1; p += 4 * stride) {
        step(p, a1, a2);
        step(p + stride, b1, b2);
        step(p + 2 * stride, c1, c2);
        step(p + 3 * stride, d1, d2);
    }
    for (; p < p1; p += stride) step(p, a1, a2);
    acc4(b1, d1); acc4(b2, d2);
    acc4(a1, c1); acc4(a2, c2);
    acc4(a1, b1);
    acc4(a2, b2);
    const float4 t1 = quad_tree(a1, red, tid, C4, lanes, r);
    const float4 t2 = quad_tree(a2, red, tid, C4, lanes, r);
    if (r == 0) {
        const long long nb = gridDim.x;
        reinterpret_cast<float4*>(partial)[(long long)blockIdx.x * C4 + q] = t1;
        reinterpret_cast<float4*>(partial)[(nb + blockIdx.x) * C4 + q] = t2;
    }
}

// forward statistics
__global__ void bn_fwd_final_kernel(const float* __restrict__ partial, int nblocks, int C, long long npix, const float* __restrict__ x,
                                    const float* __restrict__ gamma, const float* __restrict__ mean_shift, float eps, float momentum,
                                    float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean_out,
                                    float* __restrict__ invstd_out, float* __restrict__ scale_out) {
    // one warp per channel: lanes stride over the per-CTA partials, fixed-order butterfly in float64
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
    for (int b = lane; b < nblocks; b += 32) {
        s1 += double(partial[(long long)b * C + c]);
        s2 += double(partial[(long long)(nblocks + b) * C + c]);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane != 0) return;
    const double n = double(npix);
    const double m1 = s1 / n;
    double var = s2 / n - m1 * m1;                       // biased variance of the batch (shift-invariant)
    if (var < 0.0) var = 0.0;
    const float mean = float(double(x[c]) + m1);
    const float invstd = float(1.0 / sqrt(var + double(eps)));
    mean_out[c] = mean;
    invstd_out[c] = invstd;
    scale_out[c] = (gamma ? gamma[c] : 1.0f) * invstd;
    if (running_mean) {
        const float tracked = mean + (mean_shift ? mean_shift[c] : 0.0f);
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * tracked;
    }
    if (running_var) {
        const float unbiased = float(npix > 1 ? var * n / (n - 1.0) : var);
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * unbiased;
    }
}

__global__ void __launch_bounds__(kThreads)
bn_fwd_apply_kernel(const float* __restrict__ x, long long npix, int C4, const float* __restrict__ mean,
                    const float* __restrict__ scale, const float* __restrict__ beta, int relu, float* __restrict__ y) {
    const long long total = npix * C4;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int q = int(idx % C4);
        float4 v = bn_out(ldg4(x + idx * 4), ldg4(mean + 4 * q), ldg4(scale + 4 * q), ldg4(beta + 4 * q));
        if (relu) { v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y; v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w; }
        reinterpret_cast<float4*>(y)[idx] = v;
    }
}

// backward sums -> dgamma, dbeta and coef[3][C] = (gamma*invstd, mean(dy'), invstd^2 * mean(dy' (x - mean)))
__global__ void bn_bwd_final_kernel(const float* __restrict__ partial, int nblocks, int C, long long npix,
                                    const float* __restrict__ gamma, const float* __restrict__ invstd, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, float* __restrict__ coef) {
    // one warp per channel: lanes stride over the per-CTA partials, fixed-order butterfly in float64
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
    for (int b = lane; b < nblocks; b += 32) {
        s1 += double(partial[(long long)b * C + c]);
        s2 += double(partial[(long long)(nblocks + b) * C + c]);
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane != 0) return;
    const double n = double(npix), is = double(invstd[c]);
    if (dbeta) dbeta[c] = float(s1);
    if (dgamma) dgamma[c] = float(s2 * is);
    coef[c] = (gamma ? gamma[c] : 1.0f) * invstd[c];
    coef[C + c] = float(s1 / n);
    coef[2 * C + c] = float(is * is * s2 / n);
}

__global__ void __launch_bounds__(kThreads)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, long long npix, int C4, int C,
                    const float* __restrict__ mean, const float* __restrict__ scale, const float* __restrict__ beta, int relu,
                    const float* __restrict__ coef, float* __restrict__ dx) {
    const long long total = npix * C4;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int q = int(idx % C4);
        const float4 m = ldg4(mean + 4 * q);
        const float4 xv = ldg4(x + idx * 4);
        float4 g = ldg4(dy + idx * 4);
        if (relu) {
            const float4 y = bn_out(xv, m, ldg4(scale + 4 * q), ldg4(beta + 4 * q));
            g.x = y.x <= 0.f ? 0.f : g.x; g.y = y.y <= 0.f ? 0.f : g.y; g.z = y.z <= 0.f ? 0.f : g.z; g.w = y.w <= 0.f ? 0.f : g.w;
        }
        const float4 a = ldg4(coef + 4 * q), b = ldg4(coef + C + 4 * q), c = ldg4(coef + 2 * C + 4 * q);
        float4 o;
        o.x = a.x * (g.x - b.x - (xv.x - m.x) * c.x);
        o.y = a.y * (g.y - b.y - (xv.y - m.y) * c.y);
        o.z = a.z * (g.z - b.z - (xv.z - m.z) * c.z);
        o.w = a.w * (g.w - b.w - (xv.w - m.w) * c.w);
        reinterpret_cast<float4*>(dx)[idx] = o;
    }
}

int stat_blocks(long long npix, int C, int sm_count) {
    const long long lanes = kThreads / (C / 4);
    long long b = (npix + 4 * lanes - 1) / (4 * lanes);          // at least ~4 pixels per thread row
    const long long cap = 4LL * sm_count;                          // 4 resident CTAs per SM, 4 loads per thread in flight
    if (b > cap) b = cap;
    return int(b < 1 ? 1 : b);
}

int apply_blocks(long long items, int sm_count) {
    long long b = (items + kThreads - 1) / kThreads;
    const long long cap = 16LL * sm_count;
    if (b > cap) b = cap;
    return int(b < 1 ? 1 : b);
}

}  // namespace

bool batchnorm_supported(int C) { return C >= 4 && C <= 4 * kThreads && (C % 4) == 0 && (kThreads % (C / 4)) == 0; }

// floats: partial sums [2][blocks][C] + coef [3][C] (+ padding)
size_t batchnorm_workspace_floats(long long npix, int C, int sm_count) {
    return size_t(2) * size_t(stat_blocks(npix, C, sm_count)) * C + size_t(4) * C;
}

cudaError_t launch_batchnorm_fwd(const float* x, long long npix, int C, const float* gamma, const float* beta,
                                 const float* mean_shift, float eps, float momentum, bool relu, float* running_mean,
                                 float* running_var, float* y, float* save_mean, float* save_invstd, float* save_scale,
                                 float* workspace, int sm_count, cudaStream_t stream) {
    const int nb = stat_blocks(npix, C, sm_count);
    bn_partial_kernel<false><<<nb, kThreads, 0, stream>>>(x, nullptr, npix, C / 4, nullptr, nullptr, nullptr, 0, workspace);
    bn_fwd_final_kernel<<<(C + 3) / 4, 128, 0, stream>>>(workspace, nb, C, npix, x, gamma, mean_shift, eps, momentum, running_mean,
                                                             running_var, save_mean, save_invstd, save_scale);
    bn_fwd_apply_kernel<<<apply_blocks(npix * (C / 4), sm_count), kThreads, 0, stream>>>(x, npix, C / 4, save_mean, save_scale, beta,
                                                                                         relu ? 1 : 0, y);
    return cudaGetLastError();
}

cudaError_t launch_batchnorm_bwd(const float* x, const float* dy, long long npix, int C, const float* gamma, const float* beta,
                                 const float* save_mean, const float* save_invstd, const float* save_scale, bool relu, float* dx,
                                 float* dgamma, float* dbeta, float* workspace, int sm_count, cudaStream_t stream) {
    const int nb = stat_blocks(npix, C, sm_count);
    float* coef = workspace + size_t(2) * nb * C;
    bn_partial_kernel<true><<<nb, kThreads, 0, stream>>>(x, dy, npix, C / 4, save_mean, save_scale, beta, relu ? 1 : 0, workspace);
    bn_bwd_final_kernel<<<(C + 3) / 4, 128, 0, stream>>>(workspace, nb, C, npix, gamma, save_invstd, dgamma, dbeta, coef);
    bn_bwd_apply_kernel<<<apply_blocks(npix * (C / 4), sm_count), kThreads, 0, stream>>>(x, dy, npix, C / 4, C, save_mean, save_scale,
                                                                                         beta, relu ? 1 : 0, coef, dx);
    return cudaGetLastError();
}

}  // namespace wtpse

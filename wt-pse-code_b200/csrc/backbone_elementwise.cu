// Channels-last element-wise kernels of the U-Net backbone's train step: bilinear x2 up-sampling of the decoder stages
// (forward and adjoint) and the convolution bias (+ ReLU) pass.
//
// ConvU.forward (algorithms.py:947: `F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)`) runs 8
// times per U-Net pass; ATen's NHWC kernel reaches ~0.6 TB/s on it (1.04 ms for 15x32x256x256 -> 512x512, 26 ms of
// the 242 ms train iteration forward + backward).  The operation is a fixed 2-tap filter per axis:
//     out[2k]   = 0.25 in[k-1] + 0.75 in[k]      (out[0]    = in[0])
//     out[2k+1] = 0.75 in[k]   + 0.25 in[k+1]    (out[2n-1] = in[n-1])
// evaluated exactly in ATen's order, h0*(w0*a + w1*b) + h1*(w0*c + w1*d) with (lambda0, lambda1) = (0.75, 0.25) or
// (0.25, 0.75) and clamped neighbours, so results agree with F.interpolate to the last bit or two.
//
// Layout: x [N][H][W][C] (channels-last memory of an N x C x H x W tensor), C % 4 == 0; a thread owns one 128-bit
// channel quad of one INPUT pixel: forward reads its 3x3 neighbourhood (L1/L2 hits) and writes the 2x2 output block;
// the adjoint gathers the 4x4 output-gradient neighbourhood and writes one quad (no atomics, deterministic).
// HBM-bound: forward 4 + 16 B per input element, backward 16 + 4.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float4 lerp2(float h0, float h1, float w0, float w1, const float4& a, const float4& b,
                                        const float4& c, const float4& d) {
    float4 r;
    r.x = h0 * (w0 * a.x + w1 * b.x) + h1 * (w0 * c.x + w1 * d.x);
    r.y = h0 * (w0 * a.y + w1 * b.y) + h1 * (w0 * c.y + w1 * d.y);
    r.z = h0 * (w0 * a.z + w1 * b.z) + h1 * (w0 * c.z + w1 * d.z);
    r.w = h0 * (w0 * a.w + w1 * b.w) + h1 * (w0 * c.w + w1 * d.w);
    return r;
}

__global__ void __launch_bounds__(kThreads)
upsample2x_nhwc_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long N, int H, int W, int C4) {
    const long long total = N * H * W * C4;
    const long long C = 4LL * C4;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int c = int(idx % C4);
        long long r = idx / C4;
        const int j = int(r % W); r /= W;
        const int i = int(r % H);
        const long long n = r / H;
        const int im = i > 0 ? i - 1 : 0, ip = i < H - 1 ? i + 1 : H - 1;
        const int jm = j > 0 ? j - 1 : 0, jp = j < W - 1 ? j + 1 : W - 1;
        const float* base = x + n * H * W * C + 4 * c;
        auto at = [&](int ii, int jj) { return ld4(base + ((long long)ii * W + jj) * C); };
        const float4 a00 = at(im, jm), a01 = at(im, j), a02 = at(im, jp);
        const float4 a10 = at(i, jm), a11 = at(i, j), a12 = at(i, jp);
        const float4 a20 = at(ip, jm), a21 = at(ip, j), a22 = at(ip, jp);
        // even output row/col 2k: source k - 0.25 -> (k-1, k) with lambda (0.25, 0.75); at k == 0 the source clamps to 0:
        // (0, 1) with lambda (1, 0).  odd 2k+1: source k + 0.25 -> (k, k+1 clamped) with lambda (0.75, 0.25).
        const float he0 = i > 0 ? 0.25f : 1.0f, he1 = i > 0 ? 0.75f : 0.0f;
        const float we0 = j > 0 ? 0.25f : 1.0f, we1 = j > 0 ? 0.75f : 0.0f;
        // operands of the even position at the border: (in[0], in[1]) -- in[1] is `ip` / `jp`
        const float4 e_r0c0 = i > 0 ? (j > 0 ? a00 : a01) : (j > 0 ? a10 : a11);      // (row0, col0) of the even/even stencil
        const float4 e_r0c1 = i > 0 ? (j > 0 ? a01 : a02) : (j > 0 ? a11 : a12);
        const float4 e_r1c0 = i > 0 ? (j > 0 ? a10 : a11) : (j > 0 ? a20 : a21);
        const float4 e_r1c1 = i > 0 ? (j > 0 ? a11 : a12) : (j > 0 ? a21 : a22);
        // even row, odd col: rows as above, cols (j, jp) with (0.75, 0.25)
        const float4 eo_r0c0 = i > 0 ? a01 : a11, eo_r0c1 = i > 0 ? a02 : a12;
        const float4 eo_r1c0 = i > 0 ? a11 : a21, eo_r1c1 = i > 0 ? a12 : a22;
        // odd row: rows (i, ip) with (0.75, 0.25)
        const float4 oe_r0c0 = j > 0 ? a10 : a11, oe_r0c1 = j > 0 ? a11 : a12;
        const float4 oe_r1c0 = j > 0 ? a20 : a21, oe_r1c1 = j > 0 ? a21 : a22;
        float* out = y + ((n * 2 * H + 2 * i) * 2 * W + 2 * j) * C + 4 * c;
        const long long row = 2LL * W * C;
        *reinterpret_cast<float4*>(out) = lerp2(he0, he1, we0, we1, e_r0c0, e_r0c1, e_r1c0, e_r1c1);
        *reinterpret_cast<float4*>(out + C) = lerp2(he0, he1, 0.75f, 0.25f, eo_r0c0, eo_r0c1, eo_r1c0, eo_r1c1);
        *reinterpret_cast<float4*>(out + row) = lerp2(0.75f, 0.25f, we0, we1, oe_r0c0, oe_r0c1, oe_r1c0, oe_r1c1);
        *reinterpret_cast<float4*>(out + row + C) = lerp2(0.75f, 0.25f, 0.75f, 0.25f, a11, a12, a21, a22);
    }
}

// 1-D adjoint weights: input k receives from outputs 2k-1 (0.25), 2k (0.75, or 1 at k == 0), 2k+1 (0.75, or 1 at k == n-1:
// there in[n-1] is both operands), 2k+2 (0.25); out-of-range outputs contribute nothing.
__device__ __forceinline__ void adjoint_taps(int k, int n, float (&w)[4]) {
    w[0] = k > 0 ? 0.25f : 0.f;
    w[1] = k > 0 ? 0.75f : 1.0f;
    w[2] = k < n - 1 ? 0.75f : 1.0f;
    w[3] = k < n - 1 ? 0.25f : 0.f;
}

__global__ void __launch_bounds__(kThreads)
upsample2x_nhwc_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, long long N, int H, int W, int C4) {
    const long long total = N * H * W * C4;
    const long long C = 4LL * C4;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int c = int(idx % C4);
        long long r = idx / C4;
        const int j = int(r % W); r /= W;
        const int i = int(r % H);
        const long long n = r / H;
        float wr[4], wc[4];
        adjoint_taps(i, H, wr);
        adjoint_taps(j, W, wc);
        const float* base = gy + n * 4 * H * W * C + 4 * c;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const int oi = 2 * i - 1 + p;
            if (oi < 0 || oi >= 2 * H) continue;
            float4 rowacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int oj = 2 * j - 1 + q;
                if (oj < 0 || oj >= 2 * W) continue;
                const float4 g = ld4(base + ((long long)oi * 2 * W + oj) * C);
                rowacc.x = fmaf(wc[q], g.x, rowacc.x); rowacc.y = fmaf(wc[q], g.y, rowacc.y);
                rowacc.z = fmaf(wc[q], g.z, rowacc.z); rowacc.w = fmaf(wc[q], g.w, rowacc.w);
            }
            acc.x = fmaf(wr[p], rowacc.x, acc.x); acc.y = fmaf(wr[p], rowacc.y, acc.y);
            acc.z = fmaf(wr[p], rowacc.z, acc.z); acc.w = fmaf(wr[p], rowacc.w, acc.w);
        }
        *reinterpret_cast<float4*>(gx + ((n * H + i) * W + j) * C + 4 * c) = acc;
    }
}

// y[n][h][w][c] = act(y + bias[c]) in place, channels-last, C % 4 == 0: the bias add ATen runs as a separate broadcast
// `add_` after a cuDNN convolution (non-vectorised for channels-last operands: ~3 TB/s of traffic), merged with the
// ReLU that follows it in the 1x1 heads and the DeepWT blocks (algorithms.py:416-428, 1019-1030): one pass instead of two.
__global__ void __launch_bounds__(kThreads)
bias_act_nhwc_kernel(float* __restrict__ y, const float* __restrict__ bias, long long npix, int C4, int relu) {
    const long long total = npix * C4;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int c = int(idx % C4);
        const float4 b = ld4(bias + 4 * c);
        float4 v = reinterpret_cast<float4*>(y)[idx];
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        if (relu) {          // ATen's relu is clamp_min(0): NaN propagates
            v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y; v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
        }
        reinterpret_cast<float4*>(y)[idx] = v;
    }
}

// out[c] = sum over pixels of g[p][c] (channels-last): the bias gradient of a convolution, `g.sum((0, 2, 3))`.
// Two deterministic stages: every CTA sums a fixed grid-strided set of pixels per channel (thread t owns channel quad
// t % C4, 128-bit loads, fixed-order shared-memory tree), then one warp per channel adds the per-CTA partials.
__global__ void __launch_bounds__(kThreads)
channel_sum_partial_kernel(const float* __restrict__ g, long long npix, int C4, float* __restrict__ partial) {
    __shared__ float4 red[kThreads];
    const int tid = threadIdx.x;
    const int lanes = kThreads / C4;                 // pixels in flight per CTA iteration (C4 divides 256)
    const int q = tid % C4, r = tid / C4;
    // grid-stride over pixel rows: the whole grid reads one contiguous window at a time (DRAM locality)
    const long long stride = (long long)gridDim.x * lanes;
    const long long p1 = npix;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
    long long p = (long long)blockIdx.x * lanes + r;
    for (; p + 3 * stride < p1; p += 4 * stride) {    // four independent accumulators: four 128-bit loads in flight
        const float4 v0 = ld4(g + (p * C4 + q) * 4), v1 = ld4(g + ((p + stride) * C4 + q) * 4);
        const float4 v2 = ld4(g + ((p + 2 * stride) * C4 + q) * 4), v3 = ld4(g + ((p + 3 * stride) * C4 + q) * 4);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
        a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
        a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
    }
    for (; p < p1; p += stride) {
        const float4 v0 = ld4(g + (p * C4 + q) * 4);
        a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    }
    a0.x += a2.x; a0.y += a2.y; a0.z += a2.z; a0.w += a2.w;
    a1.x += a3.x; a1.y += a3.y; a1.z += a3.z; a1.w += a3.w;
    red[tid] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
    __syncthreads();
    for (int s = lanes >> 1; s > 0; s >>= 1) {
        if (r < s) {
            const float4 o = red[tid + s * C4];
            float4 m = red[tid];
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
            red[tid] = m;
        }
        __syncthreads();
    }
    if (r == 0) reinterpret_cast<float4*>(partial)[(long long)blockIdx.x * C4 + q] = red[tid];
}

// Backward of bias + ReLU in one pass: gx = [out > 0] * g (ATen's threshold_backward on the saved output) and the
// per-CTA channel sums of gx (the bias gradient) -- instead of threshold_backward followed by a second read for the sum.
// Same schedule and reduction order as channel_sum_partial_kernel.
__global__ void __launch_bounds__(kThreads)
relu_bwd_channel_sum_kernel(const float* __restrict__ g, const float* __restrict__ out, float* __restrict__ gx, long long npix, int C4,
                            float* __restrict__ partial) {
    __shared__ float4 red[kThreads];
    const int tid = threadIdx.x;
    const int lanes = kThreads / C4;
    const int q = tid % C4, r = tid / C4;
    const long long stride = (long long)gridDim.x * lanes;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    auto step = [&](long long p, float4& a) {
        const long long e = (p * C4 + q) * 4;
        float4 v = ld4(g + e);
        const float4 o = ld4(out + e);
        v.x = o.x <= 0.f ? 0.f : v.x; v.y = o.y <= 0.f ? 0.f : v.y; v.z = o.z <= 0.f ? 0.f : v.z; v.w = o.w <= 0.f ? 0.f : v.w;
        *reinterpret_cast<float4*>(gx + e) = v;
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    };
    long long p = (long long)blockIdx.x * lanes + r;
    for (; p + stride < npix; p += 2 * stride) {
        step(p, a0);
        step(p + stride, a1);
    }
    if (p < npix) step(p, a0);
    red[tid] = make_float4(a0.x + a1.x, a0.y + a1.y, a0.z + a1.z, a0.w + a1.w);
    __syncthreads();
    for (int s = lanes >> 1; s > 0; s >>= 1) {
        if (r < s) {
            const float4 o = red[tid + s * C4];
            float4 m = red[tid];
            m.x += o.x; m.y += o.y; m.z += o.z; m.w += o.w;
            red[tid] = m;
        }
        __syncthreads();
    }
    if (r == 0) reinterpret_cast<float4*>(partial)[(long long)blockIdx.x * C4 + q] = red[tid];
}

// one warp per channel: lanes stride over the per-CTA partials, fixed-order butterfly
__global__ void channel_sum_final_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ out) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    float s = 0.f;
    for (int b = lane; b < nblocks; b += 32) s += partial[(long long)b * C + c];
    s = warp_sum(s);
    if (lane == 0) out[c] = s;
}

// 2x2 / stride-2 max pooling of the encoder stages (F.max_pool2d(x, 2), algorithms.py:897), channels-last.  The argmax
// is kept as one byte per output element (ATen keeps an int64: 8 B) and the backward writes every input element
// exactly once (the windows do not overlap), so it needs no zero-fill and no atomics.  Selection rule as ATen's:
// scan the window row-major, take v if v > current or v is NaN.
__global__ void __launch_bounds__(kThreads)
maxpool2_nhwc_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, unsigned char* __restrict__ arg, long long N, int Ho,
                         int Wo, int C4) {
    const long long total = N * Ho * Wo * C4;
    const long long C = 4LL * C4;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int c = int(idx % C4);
        long long r = idx / C4;
        const int j = int(r % Wo); r /= Wo;
        const int i = int(r % Ho);
        const long long n = r / Ho;
        const float* base = x + ((n * 2 * Ho + 2 * i) * 2 * Wo + 2 * j) * C + 4 * c;
        const long long row = 2LL * Wo * C;
        const float4 v[4] = {ld4(base), ld4(base + C), ld4(base + row), ld4(base + row + C)};
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        unsigned char a[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int l = 0; l < 4; ++l)
                if (e[l] > m[l] || e[l] != e[l]) { m[l] = e[l]; a[l] = (unsigned char)k; }
        }
        reinterpret_cast<float4*>(y)[idx] = make_float4(m[0], m[1], m[2], m[3]);
        reinterpret_cast<uchar4*>(arg)[idx] = make_uchar4(a[0], a[1], a[2], a[3]);
    }
}

__global__ void __launch_bounds__(kThreads)
maxpool2_nhwc_bwd_kernel(const float* __restrict__ gy, const unsigned char* __restrict__ arg, float* __restrict__ gx, long long N,
                         int Ho, int Wo, int C4) {
    const long long total = N * Ho * Wo * C4;
    const long long C = 4LL * C4;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const int c = int(idx % C4);
        long long r = idx / C4;
        const int j = int(r % Wo); r /= Wo;
        const int i = int(r % Ho);
        const long long n = r / Ho;
        const float4 g = ld4(gy + idx * 4);
        const uchar4 a = reinterpret_cast<const uchar4*>(arg)[idx];
        float* base = gx + ((n * 2 * Ho + 2 * i) * 2 * Wo + 2 * j) * C + 4 * c;
        const long long row = 2LL * Wo * C;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 o = make_float4(a.x == k ? g.x : 0.f, a.y == k ? g.y : 0.f, a.z == k ? g.z : 0.f, a.w == k ? g.w : 0.f);
            *reinterpret_cast<float4*>(base + (k >> 1) * row + (k & 1) * C) = o;
        }
    }
}

int grid_for(long long items, int sm_count) {
    long long b = (items + kThreads - 1) / kThreads;
    const long long cap = 16LL * sm_count;
    if (b > cap) b = cap;
    return int(b < 1 ? 1 : b);
}

}  // namespace

cudaError_t launch_upsample2x_nhwc(const float* in, float* out, long long N, int H, int W, int C, bool adjoint, int sm_count,
                                   cudaStream_t stream) {
    const long long items = N * H * W * (C / 4);
    if (adjoint) upsample2x_nhwc_bwd_kernel<<<grid_for(items, sm_count), kThreads, 0, stream>>>(in, out, N, H, W, C / 4);
    else upsample2x_nhwc_fwd_kernel<<<grid_for(items, sm_count), kThreads, 0, stream>>>(in, out, N, H, W, C / 4);
    return cudaGetLastError();
}

cudaError_t launch_bias_act_nhwc(float* y, const float* bias, long long npix, int C, bool relu, int sm_count, cudaStream_t stream) {
    bias_act_nhwc_kernel<<<grid_for(npix * (C / 4), sm_count), kThreads, 0, stream>>>(y, bias, npix, C / 4, relu ? 1 : 0);
    return cudaGetLastError();
}

// Ho, Wo: pooled sizes; the input is N x 2Ho x 2Wo x C.  backward: in = gy, out = gx.
cudaError_t launch_maxpool2_nhwc(const float* in, float* out, unsigned char* arg, long long N, int Ho, int Wo, int C, bool backward,
                                 int sm_count, cudaStream_t stream) {
    const long long items = N * Ho * Wo * (C / 4);
    if (backward) maxpool2_nhwc_bwd_kernel<<<grid_for(items, sm_count), kThreads, 0, stream>>>(in, arg, out, N, Ho, Wo, C / 4);
    else maxpool2_nhwc_fwd_kernel<<<grid_for(items, sm_count), kThreads, 0, stream>>>(in, out, arg, N, Ho, Wo, C / 4);
    return cudaGetLastError();
}

int channel_sum_blocks(long long npix, int C, int sm_count) {
    const long long lanes = kThreads / (C / 4);
    long long b = (npix + 4 * lanes - 1) / (4 * lanes);        // about four pixels per thread row at least
    const long long cap = 4LL * sm_count;
    if (b > cap) b = cap;
    return int(b < 1 ? 1 : b);
}

bool channel_sum_supported(int C) { return C >= 4 && C <= 4 * kThreads && (C % 4) == 0 && (kThreads % (C / 4)) == 0; }

cudaError_t launch_channel_sum_nhwc(const float* g, long long npix, int C, float* out, float* partial, int sm_count,
                                    cudaStream_t stream) {
    const int blocks = channel_sum_blocks(npix, C, sm_count);
    channel_sum_partial_kernel<<<blocks, kThreads, 0, stream>>>(g, npix, C / 4, partial);
    channel_sum_final_kernel<<<(C + 3) / 4, 128, 0, stream>>>(partial, blocks, C, out);
    return cudaGetLastError();
}

cudaError_t launch_relu_bwd_channel_sum_nhwc(const float* g, const float* out, float* gx, long long npix, int C, float* bias_grad,
                                             float* partial, int sm_count, cudaStream_t stream) {
    const int blocks = channel_sum_blocks(npix, C, sm_count);
    relu_bwd_channel_sum_kernel<<<blocks, kThreads, 0, stream>>>(g, out, gx, npix, C / 4, partial);
    channel_sum_final_kernel<<<(C + 3) / 4, 128, 0, stream>>>(partial, blocks, C, bias_grad);
    return cudaGetLastError();
}

}  // namespace wtpse

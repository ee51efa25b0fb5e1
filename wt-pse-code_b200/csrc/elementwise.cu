// Streaming element-wise kernels either side of the whitening loss:
//   * prepare_batch   uint8 image + raw mask -> fp32 CHW image in [-1,1] and the {0,1} OD / OC labels
//                     (custom_transforms.py:466-499 Normalize_tf + :581-599 ToTensor)        -- integer path, bit-exact
//   * od_roi          od_pred = sigmoid(logits) > 0.75 ; image += 1 ; roi = image * od_pred - 1 ; pos-weight sums
//                     (Trainer.py:842-853, 865-867)                                          -- compare path, bit-exact
//   * attention_fuse  att = sigmoid(w * z_post + b) ; mask = att > 0.75 ; fuse = coef * emb + att * emb
//                     (algorithms.py:1243-1249 with attention_layer = Conv2d(1,1,1) + Sigmoid, :1120-1129), fwd + bwd
// All are HBM-bound one-pass kernels: 128-bit coalesced accesses, grid sized to a multiple of the SM
// count, no shared-memory staging (no reuse).  Arithmetic that must match the reference bit for bit
// uses explicit round-to-nearest intrinsics so nvcc cannot contract it into FMAs.
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf_ref(float x) {
    // ATen's CUDA sigmoid: one / (one + std::exp(-a)) in fp32 with the accurate expf and IEEE division
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

int blocks_for(long long work_items, int sm_count, int per_sm = 8) {
    long long b = (work_items + kThreads - 1) / kThreads;
    const long long cap = (long long)per_sm * sm_count;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return int(b);
}

// ---------------------------------------------------------------------------------------------
// prepare_batch: one thread per pixel (reads 3 + 1 (+1) bytes, writes 3 + 2 floats)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
prepare_batch_kernel(const uint8_t* __restrict__ img, const uint8_t* __restrict__ raw_od, const uint8_t* __restrict__ raw_oc,
                     int B, long long HW, float* __restrict__ image, float* __restrict__ label_od,
                     float* __restrict__ label_oc) {
    const long long total = (long long)B * HW;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const long long b = i / HW, p = i - b * HW;
        if (img) {
            const uint8_t* px = img + i * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c)   // img /= 127.5 ; img -= 1.0   (custom_transforms.py:471-472)
                image[(b * 3 + c) * HW + p] = __fsub_rn(__fdiv_rn(float(px[c]), 127.5f), 1.0f);
        }
        // tri-level image of the OD raw mask (:473-477): 255 above 200, 128 in (50, 201), else 0
        const unsigned r = raw_od[i];
        const int tri = r > 200 ? 255 : ((r > 50 && r < 201) ? 128 : 0);
        // :480-481  mask_od[tri < 255] = 1 ; mask_od[tri == 255] = 0
        label_od[i] = tri < 255 ? 1.0f : 0.0f;
        // :493-494  mask_oc[tri_od > 0] = 0 ; mask_oc[tri_od == 0] = 1   (the OD tri-level image, a reference quirk;
        // raw_oc only supplies the buffer that gets fully overwritten)
        (void)raw_oc;
        label_oc[i] = tri == 0 ? 1.0f : 0.0f;
    }
}

// ---------------------------------------------------------------------------------------------
// od_roi
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
od_roi_kernel(const float* __restrict__ logits, const float* __restrict__ target_oc, float* __restrict__ image,
              float* __restrict__ od_pred, float* __restrict__ image_roi, int B, int C, long long HW, float thr,
              unsigned long long* __restrict__ counts) {
    __shared__ unsigned int red[2][kThreads / 32];
    const long long total = (long long)B * HW;
    unsigned int n_pred = 0, n_both = 0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const long long b = i / HW, p = i - b * HW;
        const float pred = sigmoidf_ref(logits[i]) > thr ? 1.0f : 0.0f;      // Trainer.py:842
        od_pred[i] = pred;
        n_pred += pred != 0.0f;
        if (target_oc) n_both += (__fmul_rn(pred, target_oc[i]) != 0.0f);   // od_pred * target_oc, :865
        for (int c = 0; c < C; ++c) {
            const long long q = (b * C + c) * HW + p;
            const float t = __fadd_rn(image[q], 1.0f);                       // image += 1        :850
            image[q] = t;
            image_roi[q] = __fsub_rn(__fmul_rn(t, pred), 1.0f);              // image * od_pred ; -= 1   :851-852
        }
    }
    n_pred = __reduce_add_sync(0xffffffffu, n_pred);
    n_both = __reduce_add_sync(0xffffffffu, n_both);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = n_pred; red[1][threadIdx.x >> 5] = n_both; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long a = 0, c = 0;
        for (int w = 0; w < kThreads / 32; ++w) { a += red[0][w]; c += red[1][w]; }
        atomicAdd(&counts[0], a);     // integer atomics: order-independent, exact
        atomicAdd(&counts[1], c);
    }
}

// sums -> torch.sum(od_pred) / torch.sum(od_pred * target_oc), 1 when inf/nan (Trainer.py:865-867); the sums of
// 0/1 floats are exact in fp32 below 2^24 like torch's, and are reproduced from exact integer counts above it.
__global__ void pos_weight_kernel(const unsigned long long* __restrict__ counts, float* __restrict__ out) {
    const float s_pred = float(counts[0]), s_both = float(counts[1]);
    float w = __fdiv_rn(s_pred, s_both);
    if (isinf(w) || isnan(w)) w = 1.0f;
    out[0] = s_pred;
    out[1] = s_both;
    out[2] = w;
}

// ---------------------------------------------------------------------------------------------
// attention fuse
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
fuse_fwd_kernel(const float* __restrict__ emb, const float* __restrict__ zp, const float* __restrict__ wb, float coef,
                int B, int Ce, long long P, float thr, float* __restrict__ fuse, float* __restrict__ mask,
                float* __restrict__ att_out) {
    const float w = __ldg(wb), bias = __ldg(wb + 1);
    const long long total = (long long)B * P;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const long long b = i / P, p = i - b * P;
        const float att = sigmoidf_ref(fmaf(w, zp[i], bias));          // Conv2d(1,1,1) then Sigmoid, algorithms.py:1126-1128
        att_out[i] = att;
        mask[i] = att > thr ? 1.0f : 0.0f;                             // (att > 0.75).float(), :1244-1245
        for (int c = 0; c < Ce; ++c) {
            const long long q = (b * Ce + c) * P + p;
            const float e = emb[q];
            fuse[q] = __fadd_rn(__fmul_rn(coef, e), __fmul_rn(att, e));   // coef * emb + att * emb, :1248-1249
        }
    }
}

// backward: d_emb = g * (coef + att) ; d_att = sum_c g_c * emb_c ; d_pre = d_att * att * (1 - att) ;
//           d_zp = d_pre * w ; d_w = sum d_pre * zp ; d_b = sum d_pre   (per-block partials, fixed-order final sum)
__global__ void __launch_bounds__(kThreads)
fuse_bwd_kernel(const float* __restrict__ g, const float* __restrict__ emb, const float* __restrict__ zp,
                const float* __restrict__ att_in, const float* __restrict__ wb, float coef, int B, int Ce, long long P,
                float* __restrict__ d_emb, float* __restrict__ d_zp, double* __restrict__ partial) {
    __shared__ double red[2][kThreads / 32];
    const float w = __ldg(wb);
    const long long total = (long long)B * P;
    float acc_w = 0.f, acc_b = 0.f;
    double dw = 0.0, db = 0.0;
    int since = 0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
        const long long b = i / P, p = i - b * P;
        const float att = att_in[i];
        const float scale = coef + att;
        float d_att = 0.f;
        for (int c = 0; c < Ce; ++c) {
            const long long q = (b * Ce + c) * P + p;
            const float gq = g[q];
            if (d_emb) d_emb[q] = gq * scale;
            d_att = fmaf(gq, emb[q], d_att);
        }
        const float d_pre = d_att * att * (1.0f - att);
        if (d_zp) d_zp[i] = d_pre * w;
        acc_w = fmaf(d_pre, zp[i], acc_w);
        acc_b += d_pre;
        if (++since == 32) { dw += double(acc_w); db += double(acc_b); acc_w = acc_b = 0.f; since = 0; }
    }
    dw += double(acc_w);
    db += double(acc_b);
    dw = warp_sum(dw);
    db = warp_sum(db);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dw; red[1][threadIdx.x >> 5] = db; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int q = 0; q < kThreads / 32; ++q) { a += red[0][q]; c += red[1][q]; }
        partial[2 * blockIdx.x + 0] = a;
        partial[2 * blockIdx.x + 1] = c;
    }
}

__global__ void __launch_bounds__(32) fuse_bwd_final_kernel(const double* __restrict__ partial, int nblocks,
                                                            float* __restrict__ d_wb) {
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { a += partial[2 * i]; c += partial[2 * i + 1]; }
    a = warp_sum(a);
    c = warp_sum(c);
    if (threadIdx.x == 0) { d_wb[0] = float(a); d_wb[1] = float(c); }
}

}  // namespace

cudaError_t launch_prepare_batch(const uint8_t* img, const uint8_t* raw_od, const uint8_t* raw_oc, int B, long long HW,
                                 float* image, float* label_od, float* label_oc, int sm_count, cudaStream_t stream) {
    prepare_batch_kernel<<<blocks_for((long long)B * HW, sm_count), kThreads, 0, stream>>>(img, raw_od, raw_oc, B, HW, image,
                                                                                         label_od, label_oc);
    return cudaGetLastError();
}

size_t od_roi_workspace_bytes() { return 256; }

cudaError_t launch_od_roi(const float* logits, const float* target_oc, float* image, float* od_pred, float* image_roi, int B,
                          int C, long long HW, float thr, float* sums, void* workspace, int sm_count, cudaStream_t stream) {
    unsigned long long* counts = static_cast<unsigned long long*>(workspace);
    cudaError_t e = cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess) return e;
    od_roi_kernel<<<blocks_for((long long)B * HW, sm_count), kThreads, 0, stream>>>(logits, target_oc, image, od_pred, image_roi,
                                                                                  B, C, HW, thr, counts);
    if (sums) pos_weight_kernel<<<1, 1, 0, stream>>>(counts, sums);
    return cudaGetLastError();
}

cudaError_t launch_fuse_fwd(const float* emb, const float* zp, const float* wb, float coef, int B, int Ce, long long P,
                            float thr, float* fuse, float* mask, float* att, int sm_count, cudaStream_t stream) {
    fuse_fwd_kernel<<<blocks_for((long long)B * P, sm_count), kThreads, 0, stream>>>(emb, zp, wb, coef, B, Ce, P, thr, fuse, mask, att);
    return cudaGetLastError();
}

size_t fuse_bwd_partial_doubles(int B, long long P, int sm_count) { return 2 * size_t(blocks_for((long long)B * P, sm_count)); }

cudaError_t launch_fuse_bwd(const float* g, const float* emb, const float* zp, const float* att, const float* wb, float coef,
                            int B, int Ce, long long P, float* d_emb, float* d_zp, float* d_wb, double* partial, int sm_count,
                            cudaStream_t stream) {
    const int blocks = blocks_for((long long)B * P, sm_count);
    fuse_bwd_kernel<<<blocks, kThreads, 0, stream>>>(g, emb, zp, att, wb, coef, B, Ce, P, d_emb, d_zp, partial);
    if (d_wb) fuse_bwd_final_kernel<<<1, 32, 0, stream>>>(partial, blocks, d_wb);
    return cudaGetLastError();
}

}  // namespace wtpse

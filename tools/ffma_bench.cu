// Microbenchmark: FP32 FMA issue rate on sm_100a, scalar FFMA vs packed fma.rn.f32x2 (FFMA2),
// in the register-blocked outer-product pattern the Gram/apply kernels use.  Prints TFLOP/s.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
    unsigned long long da = *reinterpret_cast<unsigned long long*>(&d);
    const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(da) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&da);
}

template <int ITERS>
__global__ void __launch_bounds__(256) k_scalar(float* out, float seed) {
    float a[8], b[8], acc[64];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x * 1e-3f; b[i] = seed - i; }
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(a[i], b[j], acc[i * 8 + j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i] += 1e-7f; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ITERS>
__global__ void __launch_bounds__(256) k_packed(float* out, float seed) {
    float2 a[8], b[4], acc[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed + i + threadIdx.x * 1e-3f, seed + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = make_float2(seed - i, seed - 2 * i);
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = make_float2(0.f, 0.f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ffma2(acc[i * 4 + j], a[i], b[j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a[i].x += 1e-7f; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    constexpr int ITERS = 20000;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2) {
        for (int variant = 0; variant < 2; ++variant) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (variant == 0) k_scalar<ITERS><<<sms * blocks_per_sm, 256>>>(out, 1.0f);
                else k_packed<ITERS><<<sms * blocks_per_sm, 256>>>(out, 1.0f);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            const double fma = double(sms) * blocks_per_sm * 256.0 * ITERS * 64.0;
            printf("%s blocks/SM=%d  %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM @1.9GHz)\n", variant ? "ffma2 " : "scalar", blocks_per_sm,
                   best, 2.0 * fma / best / 1e9, fma / (best * 1e-3) / sms / 1.9e9);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

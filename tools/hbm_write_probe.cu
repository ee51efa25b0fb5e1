// Pure-write (and pure-read, copy) bandwidth with INCOMPRESSIBLE data: every thread stores values derived from a hash of its index, so
// neither cudaMemset's zeros nor a constant fill flatter the number.  The ceilings the write-dominated kernels (wavelet synthesis) should
// be read against.  Build and run on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/hbm_write_probe.cu -o /tmp/hwp && /tmp/hwp
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float hashf(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return __uint_as_float((x & 0x007fffffu) | 0x3f800000u);
}

template <int MODE>      // 0: st.global, 1: st.global.cs, 2: read (sum), 3: copy
__global__ void __launch_bounds__(256) probe(float4* __restrict__ dst, const float4* __restrict__ src, long long n4, float* sink, unsigned salt) {
    float acc = 0.f;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
        if (MODE <= 1) {
            const unsigned k = unsigned(i) * 4u + salt;
            const float4 v = make_float4(hashf(k), hashf(k + 1), hashf(k + 2), hashf(k + 3));
            if (MODE == 0) dst[i] = v;
            else __stcs(dst + i, v);
        } else if (MODE == 2) {
            const float4 v = __ldcs(src + i);
            acc += v.x + v.y + v.z + v.w;
        } else {
            __stcs(dst + i, __ldcs(src + i));
        }
    }
    if (MODE == 2 && acc == 123.456f) *sink = acc;
}

template <int MODE>
double run(float4* a, float4* b, long long n4, float* sink, int grid) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) probe<MODE><<<grid, 256>>>(a, b, n4, sink, i);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) probe<MODE><<<grid, 256>>>(a, b, n4, sink, 7 + i);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return n4 * 16.0 / (ms / 10 * 1e-3) / 1e9;
}

int main() {
    const long long n4 = 1ll << 27;                      // 2 GiB
    float4 *a, *b;
    float* sink;
    cudaMalloc(&a, n4 * 16); cudaMalloc(&b, n4 * 16); cudaMalloc(&sink, 4);
    for (int grid : {148 * 4, 148 * 8, 148 * 16}) {
        printf("grid %5d: write st.global %.0f GB/s | write st.global.cs %.0f GB/s | read %.0f GB/s | copy %.0f GB/s (read + write)\n", grid,
               run<0>(a, b, n4, sink, grid), run<1>(a, b, n4, sink, grid), run<2>(a, a, n4, sink, grid), 2 * run<3>(a, b, n4, sink, grid));
    }
    return 0;
}

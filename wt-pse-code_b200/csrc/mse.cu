// KD loss: nn.MSELoss(reduction='mean') of ShapeVariationalDist_x.wasser_distance
// (shape_networks.py:434,596-597), forward and backward.  Pure streaming: 8 B/element forward,
// 8 B read + 8 B written backward.  Two launches forward (block partials in float64, then a
// fixed-order final sum) keep the result bit-reproducible without atomics.
#include "common.cuh"
#include "kernels.h"

namespace wtpse {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ld_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}

__global__ void __launch_bounds__(kThreads)
mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long N, int vec_ok,
                   double* __restrict__ partial) {
    __shared__ double red[kThreads / 32];
    const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * kThreads;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    double acc = 0.0;
    if (vec_ok) {
        const long long n4 = N >> 2;
        int since = 0;
        for (long long i = tid; i < n4; i += nthreads) {
            const float4 x = ld_stream(a + 4 * i), y = ld_stream(b + 4 * i);
            const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
            s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
            if (++since == 64) { acc += double(s0) + double(s1) + double(s2) + double(s3); s0 = s1 = s2 = s3 = 0.f; since = 0; }
        }
        for (long long i = (n4 << 2) + tid; i < N; i += nthreads) { const float d = a[i] - b[i]; s0 = fmaf(d, d, s0); }
    } else {
        int since = 0;
        for (long long i = tid; i < N; i += nthreads) {
            const float d = a[i] - b[i];
            s0 = fmaf(d, d, s0);
            if (++since == 256) { acc += double(s0); s0 = 0.f; since = 0; }
        }
    }
    acc += double(s0) + double(s1) + double(s2) + double(s3);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += red[w];
        partial[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(32) mse_final_kernel(const double* __restrict__ partial, int nblocks, long long N,
                                                       float* __restrict__ loss) {
    double s = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) s += partial[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) loss[0] = float(s / double(N));
}

__global__ void __launch_bounds__(kThreads)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gout, long long N,
               int vec_ok, float* __restrict__ da, float* __restrict__ db) {
    const float scale = 2.0f * (gout ? *gout : 1.0f) / float(N);   // torch: grad * 2 * (a - b) / N
    const long long tid = (long long)blockIdx.x * kThreads + threadIdx.x;
    const long long nthreads = (long long)gridDim.x * kThreads;
    if (vec_ok) {
        const long long n4 = N >> 2;
        for (long long i = tid; i < n4; i += nthreads) {
            const float4 x = ld_stream(a + 4 * i), y = ld_stream(b + 4 * i);
            float4 g;
            g.x = (x.x - y.x) * scale; g.y = (x.y - y.y) * scale; g.z = (x.z - y.z) * scale; g.w = (x.w - y.w) * scale;
            if (da) *reinterpret_cast<float4*>(da + 4 * i) = g;
            if (db) *reinterpret_cast<float4*>(db + 4 * i) = make_float4(-g.x, -g.y, -g.z, -g.w);
        }
        for (long long i = (n4 << 2) + tid; i < N; i += nthreads) {
            const float g = (a[i] - b[i]) * scale;
            if (da) da[i] = g;
            if (db) db[i] = -g;
        }
    } else {
        for (long long i = tid; i < N; i += nthreads) {
            const float g = (a[i] - b[i]) * scale;
            if (da) da[i] = g;
            if (db) db[i] = -g;
        }
    }
}

int grid_for(long long N, int sm_count) {
    long long blocks = (N / 4 + kThreads - 1) / kThreads;
    const long long cap = 8LL * sm_count;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return int(blocks);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

size_t mse_partial_doubles(long long N, int sm_count) { return size_t(grid_for(N, sm_count)); }

cudaError_t launch_mse_fwd(const float* a, const float* b, long long N, float* loss, double* partial, int sm_count,
                           cudaStream_t stream) {
    const int g = grid_for(N, sm_count);
    const int vec = aligned16(a) && aligned16(b);
    mse_partial_kernel<<<g, kThreads, 0, stream>>>(a, b, N, vec, partial);
    mse_final_kernel<<<1, 32, 0, stream>>>(partial, g, N, loss);
    return cudaGetLastError();
}

cudaError_t launch_mse_bwd(const float* a, const float* b, const float* gout, long long N, float* da, float* db,
                           int sm_count, cudaStream_t stream) {
    const int g = grid_for(N, sm_count);
    const int vec = aligned16(a) && aligned16(b) && (!da || aligned16(da)) && (!db || aligned16(db));
    mse_bwd_kernel<<<g, kThreads, 0, stream>>>(a, b, gout, N, vec, da, db);
    return cudaGetLastError();
}

}  // namespace wtpse

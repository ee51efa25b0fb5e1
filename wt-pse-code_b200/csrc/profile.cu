#include "profile.cuh"

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/wtpse_b200.h"
#include "../../include/wtpse_b200_debug.h"

namespace wtpse {

namespace {

struct Pair {
    cudaEvent_t a, b;
    int id;
};

std::mutex g_mu;
bool g_on = false;
std::vector<Pair> g_pairs;      // recorded this session
std::vector<Pair> g_free;       // recycled events
std::atomic<long long> g_launches[kKernCount];
constexpr size_t kMaxPairs = 1 << 16;

const char* kNames[kKernCount] = {"gram_tma_kernel", "whiten_epilogue_fwd_kernel", "whiten_epilogue_bwd_kernel",
                                  "apply_tma_kernel", "mmd_fwd_kernel", "mmd_bwd_kernel", "mse_fwd_kernels",
                                  "mse_bwd_kernel", "fuse_kernels", "label_kernels", "wavelet_fwd_kernels",
                                  "wavelet_bwd_kernels", "gram_reduce_kernel", "whiten_mmat_kernel", "backbone_elementwise_kernels"};

thread_local Pair t_open = {nullptr, nullptr, -1};

}  // namespace

const char* kernel_name(int id) { return (id >= 0 && id < kKernCount) ? kNames[id] : "?"; }

void profile_record_begin(int id, cudaStream_t s) {
    g_launches[id].fetch_add(1, std::memory_order_relaxed);
    t_open.id = -1;
    if (!g_on) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_pairs.size() >= kMaxPairs) return;
    Pair p;
    if (!g_free.empty()) {
        p = g_free.back();
        g_free.pop_back();
    } else if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) {
        return;
    }
    p.id = id;
    cudaEventRecord(p.a, s);
    t_open = p;
}

void profile_count_kernel(int id) { g_launches[id].fetch_add(1, std::memory_order_relaxed); }

void profile_record_end(int id, cudaStream_t s) {
    if (t_open.id != id) return;
    cudaEventRecord(t_open.b, s);
    std::lock_guard<std::mutex> lk(g_mu);
    g_pairs.push_back(t_open);
    t_open.id = -1;
}

}  // namespace wtpse

using namespace wtpse;

extern "C" {

void wtpse_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_on = on != 0;
}

void wtpse_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& p : g_pairs) g_free.push_back(p);
    g_pairs.clear();
    for (int i = 0; i < kKernCount; ++i) g_launches[i].store(0);
}

int wtpse_profile_kernel_count(void) { return kKernCount; }
const char* wtpse_profile_kernel_name(int id) { return kernel_name(id); }

long long wtpse_profile_launches(int id) {
    if (id < 0) {
        long long t = 0;
        for (int i = 0; i < kKernCount; ++i) t += g_launches[i].load();
        return t;
    }
    return id < kKernCount ? g_launches[id].load() : 0;
}

int wtpse_profile_read(int id, long long* timed_launches, double* total_ms) {
    std::lock_guard<std::mutex> lk(g_mu);
    long long n = 0;
    double ms = 0.0;
    for (auto& p : g_pairs) {
        if (p.id != id) continue;
        if (cudaEventSynchronize(p.b) != cudaSuccess) return WTPSE_ERR_CUDA;
        float t = 0.f;
        if (cudaEventElapsedTime(&t, p.a, p.b) != cudaSuccess) return WTPSE_ERR_CUDA;
        ms += t;
        ++n;
    }
    if (timed_launches) *timed_launches = n;
    if (total_ms) *total_ms = ms;
    return WTPSE_OK;
}

}  // extern "C"

// ---- stress helper for the programmatic-dependent-launch paths (tests/test_gpu_fusion.py) ---------------------------------------
// A producer that behaves the worst legal way for its dependents: it signals griddepcontrol.launch_dependents at once, then spins
// for `spin_cycles` clocks, and only then copies src -> dst.  A dependent kernel that reads dst before its own
// griddepcontrol.wait sees the bytes dst held before (the test fills it with NaN), one that waits sees src.
namespace wtpse {
namespace {
__global__ void __launch_bounds__(256) pdl_slow_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n,
                                                            long long spin_cycles) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long t0 = clock64();
    while (clock64() - t0 < spin_cycles) {
    }
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) dst[i] = src[i];
}
}  // namespace
}  // namespace wtpse

extern "C" int wtpse_debug_pdl_slow_copy(float* dst, const float* src, long long n, long long spin_cycles, void* stream) {
    if (!dst || !src || n <= 0) return WTPSE_ERR_INVALID;
    // few CTAs: the dependent grid must find free SMs to become resident while this one is still spinning
    wtpse::pdl_slow_copy_kernel<<<16, 256, 0, static_cast<cudaStream_t>(stream)>>>(dst, src, (long long)n, (long long)spin_cycles);
    return cudaGetLastError() == cudaSuccess ? WTPSE_OK : WTPSE_ERR_CUDA;
}

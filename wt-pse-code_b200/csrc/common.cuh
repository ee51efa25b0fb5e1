// Shared device helpers for the sm_100a kernels: mbarrier + 1-D TMA bulk copies, warp reductions,
// and the persistent tile partition used by the Gram (forward) and apply (backward) kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace wtpse {

constexpr int kC = 16;                       // channels (self.dim, algorithms.py:1157)
constexpr int kTri = kC * (kC + 1) / 2;      // 136 unique Gram entries (i <= j)
constexpr int kOff = kC * (kC - 1) / 2;      // 120 strict upper-triangle entries

// packed index of (i, j), i <= j, row-major over the upper triangle including the diagonal
__host__ __device__ constexpr int tri_idx(int i, int j) { return i * kC - (i * (i - 1)) / 2 + (j - i); }

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA (cp.async.bulk) primitives
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// global -> shared bulk copy, completion counted on `bar` (bytes % 16 == 0, both addresses 16B aligned)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// same, with an L2 eviction-priority hint (policy from make_evict_first_policy): z is read exactly once
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar,
                                                 uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    return policy;
}

// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// make generic-proxy smem writes visible to the async proxy (before a TMA store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// warp reductions
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}

// ------------------------------------------------------------------------------------------------
// persistent tile partition: T tiles over G CTAs, CTA k owns [floor(k*T/G), floor((k+1)*T/G)).
// A sample's tiles may straddle CTAs; CTA k writes its partial for sample b into slot
// k - owner(first tile of b).  The epilogue re-derives the same slot counts, so the reduction over
// slots has a fixed order (deterministic, no float atomics).
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ long long part_begin(long long k, long long T, long long G) { return (k * T) / G; }
__host__ __device__ __forceinline__ long long part_owner(long long t, long long T, long long G) {
    return ((t + 1) * G - 1) / T;
}


// Tile schedule.  The G CTAs form G/g groups of g; each group owns one contiguous range of tiles and deals it
// round-robin to its members (member r takes tiles r, r+g, ... of the range).  g = 1: one contiguous range per
// CTA; g = G: pure round-robin.  Every member flushes one partial for EVERY sample its group's range touches
// (zeros if it happened to get no tile of it), so the slot bookkeeping is uniform:
//   slot(b, CTA) = (group - first group touching b) * g + r,   slot_count[b] = (#groups touching b) * g.
struct TileWalk {
    long long R0, R1;      // group range
    long long g, r;        // group size, member index
    long long b_first, b_last;
    __device__ TileWalk(long long k, long long G, long long T, long long tps, long long group) {
        g = group;
        const long long Gg = G / g, grp = k / g;
        r = k - grp * g;
        R0 = part_begin(grp, T, Gg);
        R1 = part_begin(grp + 1, T, Gg);
        b_first = R0 / tps;
        b_last = (R1 - 1) / tps;
    }
    // [first, end) of this CTA's tiles inside sample b, step g
    __device__ void segment(long long b, long long tps, long long& first, long long& end) const {
        const long long s0 = b * tps > R0 ? b * tps : R0;
        const long long s1 = (b + 1) * tps < R1 ? (b + 1) * tps : R1;
        long long d = (r - (s0 - R0)) % g;
        if (d < 0) d += g;
        first = s0 + d;
        end = s1;
    }
};

// Walks one CTA's tiles without a division per tile: round-robin (tile k, k + G, ...) or one contiguous range.
struct TileIter {
    long long t, tin, b;        // tile, tile index inside its sample, sample
    long long step, tend, tps;
    __device__ __forceinline__ void init(long long first, long long end, long long stride, long long tiles_per_sample) {
        t = first; tend = end; step = stride; tps = tiles_per_sample;
        b = first / tps;
        tin = first - b * tps;
    }
    __device__ __forceinline__ bool valid() const { return t < tend; }
    __device__ __forceinline__ void next() {
        t += step;
        tin += step;
        while (tin >= tps) { tin -= tps; ++b; }
    }
    // sample of the first later tile that belongs to another sample (-1: none)
    __device__ __forceinline__ long long next_sample() const {
        TileIter it = *this;
        for (it.next(); it.valid(); it.next())
            if (it.b != b) return it.b;
        return -1;
    }
};

}  // namespace wtpse

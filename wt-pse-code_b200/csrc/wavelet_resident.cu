// Track W, cluster-resident path: loss AND gradient of the J-level detail-coefficient shape loss in ONE kernel,
// one HBM read of the map and one HBM write of its gradient (8 B per element; the per-level kernels of wavelet.cu
// move ~21 B per element).  PARITY UNPINNED (see wavelet.cu / oracle/wavelet_np.py).
//
// A whole H x W map lives in the distributed shared memory of one thread-block cluster: CTA r of a cluster of cs
// holds the band of rows [r*H/cs, (r+1)*H/cs) (full width, so the periodic wrap along W is local), loaded by 1-D
// TMA bulk copies.  Every level is then computed shared-to-shared:
//   forward level j : band of LL_{j-1} -> band of LL_j (fp32) + the SIGNS of the three detail sub-bands, 2 bits each,
//                     one byte per site (the L1 loss only needs sign(d) for its gradient); |d| is accumulated;
//   inverse level j : gLL_j + w_j * sign(d_j) / norm_j -> gLL_{j-1}, written over LL_{j-1}; level 1 writes the
//                     gradient straight to global memory.
// The only inter-CTA traffic is the filter overlap: TAPS-2 rows below the band for the analysis (taken from the next
// CTA's shared memory; for the input map itself they are simply loaded by TMA as well) and TAPS/2-1 coefficient rows
// above the band for the synthesis (previous CTA), both copied through DSMEM after a cluster barrier.  Haar has no
// overlap, so its CTAs never synchronise with each other.  The band buffer of the input map is free once level 1 is
// done, so the TMA load of the cluster's next map overlaps levels 2..J and the whole inverse chain.
// Both level passes walk down the rows of a column strip so the row-filtered values (analysis) / column-synthesised
// values (synthesis) of the overlapping rows are computed once and carried in registers: 32 FMA per site per pass.
#include <cooperative_groups.h>
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "kernels.h"
#include "wavelet_level.cuh"

namespace cg = cooperative_groups;

namespace wtpse {

namespace {

constexpr int kResThreads = 512;
constexpr int kResMaxJ = 6;
constexpr int kResSmemLimit = 227 * 1024;
constexpr int kResChunk = 32 * 1024;            // bytes per bulk copy
constexpr int kResLoadChunks = 4;               // the band of the input map arrives in this many row chunks, one mbarrier each

struct ResidentArgs {
    const float* x;
    float* grad;                // nullptr in the loss-only instantiation
    const float* upstream;      // optional device scalar multiplied into grad
    int H, W, J, nmaps, cs;
    float scale[kResMaxJ];      // w_j / (3 * (H>>j) * (W>>j) * nmaps), j = 1..J
    double* partial;            // one double per CTA
};

// Byte offsets of the per-CTA shared-memory carve.  Rb = band rows of the input map.
//   X    : (Rb + halo) x W fp32                      own rows, then the halo rows of the next band
//   L[j] : (1 + Rb>>j + halo) x (W>>j) fp32, j < J   row 0 = last row of the previous band (synthesis), then own rows,
//                                                    then the first rows of the next band (analysis)
//   S[j] : (1 + Rb>>j) x (W>>j) bytes, j <= J        packed detail signs, row 0 = last row of the previous band
struct ResLayout {
    int x, l[kResMaxJ + 1], s[kResMaxJ + 1], red, bar, total;
};

__host__ __device__ inline ResLayout res_layout(int Rb, int W, int J, int halo) {
    ResLayout o;
    int off = 0;
    o.x = off;
    off += (Rb + halo) * W * 4;
    for (int j = 1; j <= kResMaxJ; ++j) {
        o.l[j] = off;
        if (j < J) off += (1 + (Rb >> j) + halo) * (W >> j) * 4;
        off = (off + 15) & ~15;
    }
    for (int j = 1; j <= kResMaxJ; ++j) {
        o.s[j] = off;
        if (j <= J) off += (1 + (Rb >> j)) * (W >> j);
        off = (off + 15) & ~15;
    }
    o.red = off;
    off += (kResThreads / 32) * 8;
    o.bar = off;
    off += 8 * kResLoadChunks;
    o.total = off;
    return o;
}

// low/high-pass along W of one input row for the two half-resolution sites starting at input column c0 (c0 % 4 == 0)
template <int TAPS>
__device__ __forceinline__ void row_pair(const float* row, int c0, int w, float& lo0, float& hi0, float& lo1, float& hi1) {
    float x[TAPS + 2];
    const float4 v = *reinterpret_cast<const float4*>(row + c0);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    if (TAPS == 4) {
        int c = c0 + 4;
        if (c >= w) c -= w;
        const float2 u = *reinterpret_cast<const float2*>(row + c);
        x[TAPS] = u.x; x[TAPS + 1] = u.y;
    }
    row_filter<TAPS>(x, lo0, hi0, lo1, hi1);
}

// Analysis of one level, shared to shared.  in: own rows from row 0, the TAPS-2 overlap rows right after them.
// ll_off < 0: the low-low band is not needed (deepest level).  sg: own rows only (the caller skips the halo row).
// wait_bar != nullptr (level 1 only): the input arrives in row chunks of wait_rpc rows, chunk c on wait_bar[c]; a task
// starts as soon as the chunks holding its rows have landed instead of waiting for the whole band.
template <int TAPS, bool kGrad>
__device__ __forceinline__ void fwd_level(int in_off, int w, int rows_out, int ll_off, int sg_off, float sc, double& acc,
                                          uint64_t* wait_bar = nullptr, int wait_rpc = 1, uint32_t wait_phase = 0) {
    extern __shared__ __align__(128) unsigned char smem[];
    const float* in = reinterpret_cast<const float*>(smem + in_off);
    float* ll = reinterpret_cast<float*>(smem + (ll_off < 0 ? 0 : ll_off));
    unsigned char* sg = smem + sg_off;
    const int w2 = w >> 1, pairs = w >> 2;
    int seg = (rows_out * pairs) / kResThreads;
    if (seg < 1) seg = 1;
    const int nseg = (rows_out + seg - 1) / seg;
    const int ntasks = nseg * pairs;
    for (int t = threadIdx.x; t < ntasks; t += kResThreads) {
        const int jj = t % pairs, si = t / pairs;
        const int i0 = si * seg, i1 = min(i0 + seg, rows_out), c0 = 4 * jj;
        if (wait_bar) {
            const int c_hi = min(kResLoadChunks - 1, (2 * i1 + TAPS - 3) / wait_rpc);      // overlap rows: last chunk
            for (int c = (2 * i0) / wait_rpc; c <= c_hi; ++c) mbar_wait(&wait_bar[c], wait_phase);
        }
        float lo0[TAPS], hi0[TAPS], lo1[TAPS], hi1[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS - 2; ++k) row_pair<TAPS>(in + (2 * i0 + k) * w, c0, w, lo0[k], hi0[k], lo1[k], hi1[k]);
        float ab = 0.f;
        for (int i = i0; i < i1; ++i) {
            row_pair<TAPS>(in + (2 * i + TAPS - 2) * w, c0, w, lo0[TAPS - 2], hi0[TAPS - 2], lo1[TAPS - 2], hi1[TAPS - 2]);
            row_pair<TAPS>(in + (2 * i + TAPS - 1) * w, c0, w, lo0[TAPS - 1], hi0[TAPS - 1], lo1[TAPS - 1], hi1[TAPS - 1]);
            float2 LL, LH, HL, HH;
            col_filter<TAPS>(lo0, hi0, lo1, hi1, LL, LH, HL, HH);
            if (ll_off >= 0) *reinterpret_cast<float2*>(ll + i * w2 + 2 * jj) = LL;
            ab += abs_sum(LH, HL, HH) * sc;
            if (kGrad) *reinterpret_cast<unsigned short*>(sg + i * w2 + 2 * jj) = static_cast<unsigned short>(sign_pack2(LH, HL, HH));
#pragma unroll
            for (int k = 0; k < TAPS - 2; ++k) {
                lo0[k] = lo0[k + 2]; hi0[k] = hi0[k + 2]; lo1[k] = lo1[k + 2]; hi1[k] = hi1[k + 2];
            }
        }
        acc += double(ab);
    }
}

// Column synthesis of one coefficient row (buffer row r; row 0 is the halo) for the two output sites 2q, 2q+1:
// tL / tH [A parity 0, A parity 1, B parity 0, B parity 1] = low-row / high-row content of the four output columns.
template <int TAPS>
__device__ __forceinline__ void col_synth(const float* gll, bool has_ll, const unsigned char* sgp, int wj, int r, int q, float sc,
                                          float (&tL)[4], float (&tH)[4]) {
    const int c0 = 2 * q;
    float2 l01 = make_float2(0.f, 0.f);
    if (has_ll) l01 = *reinterpret_cast<const float2*>(gll + r * wj + c0);
    const unsigned b01 = *reinterpret_cast<const unsigned short*>(sgp + r * wj + c0);
    float lm = 0.f;
    unsigned bm = 0;
    if (TAPS == 4) {
        const int cm = c0 ? c0 - 1 : wj - 1;
        if (has_ll) lm = gll[r * wj + cm];
        bm = sgp[r * wj + cm];
    }
    col_synth_vals<TAPS>(l01, lm, b01, bm, sc, tL, tH);
}

// Synthesis (== adjoint) of one level.  gll_off < 0: the incoming low-low gradient is zero (deepest level).
// Output rows 2i, 2i+1 of the own band go to shared memory (out_off, row stride 2*wj) or, kGlobal, to gout.
template <int TAPS, bool kGlobal>
__device__ __forceinline__ void inv_level(int gll_off, int sg_off, int wj, int rj, float sc, int out_off, float* __restrict__ gout,
                                          int out_ld, float gs) {
    extern __shared__ __align__(128) unsigned char smem[];
    const bool has_ll = gll_off >= 0;
    const float* gll = reinterpret_cast<const float*>(smem + (has_ll ? gll_off : 0));
    const unsigned char* sgp = smem + sg_off;
    float* outs = reinterpret_cast<float*>(smem + (kGlobal ? 0 : out_off));
    const int pairs = wj >> 1;
    int seg = (rj * pairs) / kResThreads;
    if (seg < 1) seg = 1;
    const int nseg = (rj + seg - 1) / seg;
    const int ntasks = nseg * pairs;
    for (int t = threadIdx.x; t < ntasks; t += kResThreads) {
        const int q = t % pairs, si = t / pairs;
        const int i0 = si * seg, i1 = min(i0 + seg, rj);
        float pL[4] = {0.f, 0.f, 0.f, 0.f}, pH[4] = {0.f, 0.f, 0.f, 0.f};
        if (TAPS == 4) col_synth<TAPS>(gll, has_ll, sgp, wj, i0, q, sc, pL, pH);       // coefficient row i0-1 = buffer row i0
        for (int i = i0; i < i1; ++i) {
            float cL[4], cH[4];
            col_synth<TAPS>(gll, has_ll, sgp, wj, i + 1, q, sc, cL, cH);
#pragma unroll
            for (int pr = 0; pr < 2; ++pr) {
                float o[4];
                row_synth<TAPS>(cL, cH, pL, pH, pr, gs, o);
                if (kGlobal) {
                    *reinterpret_cast<float4*>(gout + (long long)(2 * i + pr) * out_ld + 4 * q) =
                        make_float4(o[0], o[1], o[2], o[3]);
                } else {
                    *reinterpret_cast<float4*>(outs + (2 * i + pr) * out_ld + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { pL[k] = cL[k]; pH[k] = cH[k]; }
        }
    }
}

// bytes from the same offset of another CTA of the cluster (16-byte words when everything is aligned, else 2-byte)
__device__ __forceinline__ void copy_from_rank(cg::cluster_group& cluster, unsigned rank, int dst_off, int src_off, int bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    if (((dst_off | src_off | bytes) & 15) == 0) {
        const uint4* src = cluster.map_shared_rank(reinterpret_cast<uint4*>(smem + src_off), rank);
        uint4* dst = reinterpret_cast<uint4*>(smem + dst_off);
        for (int i = threadIdx.x; i < (bytes >> 4); i += kResThreads) dst[i] = src[i];
    } else {
        const unsigned short* src = cluster.map_shared_rank(reinterpret_cast<unsigned short*>(smem + src_off), rank);
        unsigned short* dst = reinterpret_cast<unsigned short*>(smem + dst_off);
        for (int i = threadIdx.x; i < (bytes >> 1); i += kResThreads) dst[i] = src[i];
    }
}

template <int TAPS, bool kGrad>
__global__ void __launch_bounds__(kResThreads, 1) wavelet_resident_kernel(ResidentArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int HALO = TAPS - 2;
    cg::cluster_group cluster = cg::this_cluster();
    // db2: a.cs CTAs of one cluster share a map (bands coupled through the filter overlap), work item = map.
    // Haar: the filters do not overlap, so a band of rows (a multiple of 2^J) is a work item of its own: no cluster,
    // a.cs = bands per map, every CTA walks items (map, band) with the grid as stride.
    constexpr bool kBands = (TAPS == 2);
    const unsigned cs = kBands ? unsigned(a.cs) : cluster.num_blocks();
    const unsigned rank = kBands ? 0u : cluster.block_rank();
    const unsigned nxt = (rank + 1) % cs, prv = (rank + cs - 1) % cs;
    const int first = kBands ? int(blockIdx.x) : int(blockIdx.x / cs), stride = kBands ? int(gridDim.x) : int(gridDim.x / cs);
    const int n_items = kBands ? a.nmaps * int(cs) : a.nmaps;
    const int H = a.H, W = a.W, J = a.J, Rb = H / int(cs);
    const ResLayout lay = res_layout(Rb, W, J, HALO);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + lay.bar);
    const long long map_elems = (long long)H * W;
    const bool sync_cluster = (TAPS > 2);          // Haar bands are independent

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int c = 0; c < kResLoadChunks; ++c) mbar_init(&bar[c], 1);
        fence_mbar_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");      // launched programmatically behind the level-1 analysis kernel

    const uint64_t policy = make_evict_first_policy();
    const int rpc = (Rb + kResLoadChunks - 1) / kResLoadChunks;        // rows per load chunk
    auto issue_load = [&](int item) {              // one thread
        const int map = kBands ? item / int(cs) : item;
        const unsigned r = kBands ? unsigned(item) % cs : rank;
        const float* band = a.x + map * map_elems + (long long)r * Rb * W;
        for (int c = 0; c < kResLoadChunks; ++c) {
            const int r0 = min(c * rpc, Rb), r1 = min(r0 + rpc, Rb);
            const bool last = (c == kResLoadChunks - 1);
            const uint32_t bytes = uint32_t(r1 - r0) * W * 4u, halo = last ? uint32_t(HALO) * W * 4u : 0u;
            mbar_arrive_expect_tx(&bar[c], bytes + halo);
            const char* src = reinterpret_cast<const char*>(band + (long long)r0 * W);
            unsigned char* dst = smem + lay.x + size_t(r0) * W * 4u;
            for (uint32_t o = 0; o < bytes; o += kResChunk)
                tma_load_1d_hint(dst + o, src + o, min(uint32_t(kResChunk), bytes - o), &bar[c], policy);
            if (halo) {
                const int hr = int(((r + 1) * Rb) % H);
                tma_load_1d_hint(smem + lay.x + size_t(Rb) * W * 4u, a.x + map * map_elems + (long long)hr * W, halo, &bar[c], policy);
            }
        }
    };
    if (threadIdx.x == 0 && first < n_items) issue_load(first);

    const float gs = (kGrad && a.upstream) ? __ldg(a.upstream) : 1.0f;
    double acc = 0.0;
    uint32_t phase = 0;
    for (int item = first; item < n_items; item += stride) {
        const int map = kBands ? item / int(cs) : item;
        const unsigned r = kBands ? unsigned(item) % cs : rank;
        // the neighbours may still be copying halo rows of the previous map out of this CTA's level buffers
        if (item != first) {
            if (sync_cluster) cluster.sync();
            else __syncthreads();
        }
        // ---- analysis ----
        fwd_level<TAPS, kGrad>(lay.x, W, Rb >> 1, J > 1 ? lay.l[1] + (W >> 1) * 4 : -1, lay.s[1] + (W >> 1), a.scale[0], acc,
                               bar, rpc, phase);
        phase ^= 1;
        __syncthreads();                            // the band buffer is free: prefetch the cluster's next map
        if (threadIdx.x == 0 && item + stride < n_items) issue_load(item + stride);
        for (int j = 2; j <= J; ++j) {
            const int wi = W >> (j - 1), ri = Rb >> (j - 1);           // input band of this level
            if (sync_cluster) {
                cluster.sync();                     // LL_{j-1} complete in every CTA
                copy_from_rank(cluster, nxt, lay.l[j - 1] + (1 + ri) * wi * 4, lay.l[j - 1] + wi * 4, HALO * wi * 4);
                __syncthreads();
            }
            fwd_level<TAPS, kGrad>(lay.l[j - 1] + wi * 4, wi, ri >> 1, j < J ? lay.l[j] + (wi >> 1) * 4 : -1,
                                   lay.s[j] + (wi >> 1), a.scale[j - 1], acc);
            if (!sync_cluster) __syncthreads();
        }
        // ---- synthesis of the gradient ----
        if (kGrad) {
            float* gout = a.grad + map * map_elems + (long long)r * Rb * W;
            for (int j = J; j >= 1; --j) {
                const int wj = W >> j, rj = Rb >> j;
                if (sync_cluster) {
                    cluster.sync();                 // signs (and gLL_j) of this level complete in every CTA
                    copy_from_rank(cluster, prv, lay.s[j], lay.s[j] + rj * wj, wj);
                    if (j < J) copy_from_rank(cluster, prv, lay.l[j], lay.l[j] + rj * wj * 4, wj * 4);
                }
                __syncthreads();
                if (j > 1)
                    inv_level<TAPS, false>(j < J ? lay.l[j] : -1, lay.s[j], wj, rj, a.scale[j - 1], lay.l[j - 1] + 2 * wj * 4, nullptr,
                                           2 * wj, 1.0f);
                else
                    inv_level<TAPS, true>(j < J ? lay.l[j] : -1, lay.s[j], wj, rj, a.scale[j - 1], 0, gout, W, gs);
            }
        }
    }

    // fixed-order loss partial of this CTA
    double s = warp_sum(acc);
    double* red = reinterpret_cast<double*>(smem + lay.red);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int q = 0; q < kResThreads / 32; ++q) tot += red[q];
        a.partial[blockIdx.x] = tot;
    }
    if (sync_cluster) cluster.sync();              // nobody exits while a neighbour may still read its shared memory
}

// data *= *scale unless *scale == 1 (decided on the device: no host synchronisation, and the common case -- the loss is
// the root of the backward pass or enters it with weight 1 -- costs one launch and no memory traffic)
__global__ void __launch_bounds__(256) scale_unless_one_kernel(float* __restrict__ data, long long n4, long long n,
                                                               const float* __restrict__ scale) {
    const float s = __ldg(scale);
    if (s == 1.0f) return;
    float4* d4 = reinterpret_cast<float4*>(data);
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += gridDim.x * 256ll) {
        float4 v = d4[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        d4[i] = v;
    }
    for (long long i = 4 * n4 + blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) data[i] *= s;
}

template <int TAPS, bool kGrad>
cudaError_t launch_resident_t(const ResidentArgs& a, int smem, cudaStream_t stream, int* grid_out) {
    auto kernel = wavelet_resident_kernel<TAPS, kGrad>;
    // the shared-memory opt-in and the cluster occupancy are queried once per (device, cluster size, footprint)
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, int> clusters_that_fit;
    static std::map<int, int> opted_in;             // device -> largest dynamic shared-memory size opted into
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const auto key = std::make_tuple(dev, a.cs, smem);
    int ncl = 0;
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = clusters_that_fit.find(key);
        if (it != clusters_that_fit.end()) ncl = it->second;
    }
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(kResThreads);
    const int cluster = TAPS == 2 ? 1 : a.cs;      // Haar: a.cs independent bands per map, no cluster
    cfg.gridDim = dim3(cluster);
    cfg.dynamicSmemBytes = size_t(smem);
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (ncl == 0) {
        std::lock_guard<std::mutex> lock(mu);
        if (opted_in[dev] < smem) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return e;
            opted_in[dev] = smem;
        }
        e = cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg);
        if (e != cudaSuccess) return e;
        if (ncl < 1) return cudaErrorLaunchOutOfResources;
        clusters_that_fit[key] = ncl;
    }
    const long long items = TAPS == 2 ? (long long)a.nmaps * a.cs : a.nmaps;
    cfg.gridDim = dim3(unsigned(min((long long)ncl, items)) * cluster);
    cfg.numAttrs = 2;
    *grid_out = int(cfg.gridDim.x);
    return cudaLaunchKernelEx(&cfg, kernel, a);
}

}  // namespace

int g_wavelet_resident = 1;
int g_wavelet_cluster_max = 8;      // diagnostics: largest cluster size the planner may pick

// Cluster size for H x W maps: the SMALLEST cluster whose bands fit (fewer, larger bands: every level costs two cluster
// barriers and a DSMEM copy whatever the band size -- measured at 32 x 2 maps of 256 x 256, J = 3: clusters of 2 / 4 / 8
// -> 79.8 / 82 / 91 us for the whole streamed plan; widening the cluster for small batches did not pay either: 8 x 2 maps
// of 256 x 256, clusters of 2 / 4 -> 17.3 / 17.8 us).
int wavelet_resident_cluster(int H, int W, int taps, int J, int /*nmaps*/) {
    if (J < 1 || J > kResMaxJ || (W % (1 << (J + 1))) || (H % (1 << J))) return 0;
    if ((long long)H * W > (1ll << 24)) return 0;
    if (taps == 2) {
        // Haar: number of independent row bands per map.  The largest band (a multiple of 2^J rows dividing H) whose
        // buffers leave room for two CTAs per SM; failing that the smallest band, if it fits at all.
        int best = 0;
        for (int nb = 1; nb <= (H >> J); ++nb) {
            if (H % nb) continue;
            const int Rb = H / nb;
            if (Rb % (1 << J)) continue;
            const long long tot = (long long)Rb * W * 4 * 3 / 2;          // coarse bound, avoids int overflow below
            if (tot > (1ll << 28)) continue;
            const int smem = res_layout(Rb, W, J, 0).total;
            if (smem <= kResSmemLimit) best = nb;                           // keeps shrinking while it still fits ...
            if (smem <= 110 * 1024) break;                                  // ... until two CTAs fit an SM
        }
        return best;
    }
    for (int cs = 1; cs <= 8 && cs <= g_wavelet_cluster_max; cs <<= 1) {
        if (H % cs) continue;
        const int Rb = H / cs;
        if (Rb % (1 << J)) continue;
        if ((long long)(Rb + taps - 2) * W * 4 > kResSmemLimit) continue;
        if (res_layout(Rb, W, J, taps - 2).total > kResSmemLimit) continue;
        return cs;
    }
    return 0;
}

// loss == nullptr: the caller runs the final reduction itself over partial[0 .. *n_partials).  grad may alias x (every
// CTA overwrites only band rows it -- and, for the halo, its predecessor -- has already read).
cudaError_t launch_wavelet_resident(const float* x, int nmaps, int H, int W, int taps, int J, const float* weights_host,
                                    const float* upstream, float* loss, float* grad, double* partial, cudaStream_t stream,
                                    int* n_partials) {
    ResidentArgs a;
    a.x = x; a.grad = grad; a.upstream = upstream; a.H = H; a.W = W; a.J = J; a.nmaps = nmaps;
    a.cs = wavelet_resident_cluster(H, W, taps, J, nmaps);
    if (a.cs == 0) return cudaErrorInvalidValue;
    for (int j = 1; j <= J; ++j) a.scale[j - 1] = weights_host[j - 1] / (3.0f * float(H >> j) * float(W >> j) * float(nmaps));
    a.partial = partial;
    const int smem = res_layout(H / a.cs, W, J, taps - 2).total;
    cudaError_t e;
    int grid = 0;
    if (grad) e = taps == 2 ? launch_resident_t<2, true>(a, smem, stream, &grid) : launch_resident_t<4, true>(a, smem, stream, &grid);
    else e = taps == 2 ? launch_resident_t<2, false>(a, smem, stream, &grid) : launch_resident_t<4, false>(a, smem, stream, &grid);
    if (e != cudaSuccess) return e;
    if (n_partials) *n_partials = grid;
    if (!loss) return cudaSuccess;
    return launch_wavelet_loss_final(partial, grid, loss, stream);       // fixed order over the CTAs
}

cudaError_t launch_scale_unless_one(float* data, long long n, const float* scale, int sm_count, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    const long long n4 = (reinterpret_cast<uintptr_t>(data) & 15) ? 0 : n / 4;
    const int grid = int(min((long long)sm_count * 8, max(1ll, (n + 1023) / 1024)));
    scale_unless_one_kernel<<<grid, 256, 0, stream>>>(data, n4, n, scale);
    return cudaGetLastError();
}

}  // namespace wtpse

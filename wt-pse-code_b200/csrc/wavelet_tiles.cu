// Track W, level-1 kernels of the streamed plan as persistent TMA pipelines.  PARITY UNPINNED (see wavelet.cu).
//
// The first versions of these two kernels (wavelet_stream.cu, kept as the fallback for shapes the pipelines do not
// take) read their operands with per-thread global loads.  Measured on B200 (32 x 2 x 512 x 512, db2): both sit at
// ~22 us, half of the HBM rate, stalled on the loads however deep the register prefetch is (1.5 eligible warps per
// scheduler out of 6).  Here a producer lane streams strips of rows into a shared-memory ring with 1-D TMA bulk
// copies and 16 consumer warps only ever wait on an mbarrier:
//   dwt1_tile_kernel   strip = R output rows of one map = 2R (+ TAPS-2 overlap) input rows, full width
//                      -> LL1 (float2 stores), packed detail signs (16-bit stores), |d| partial sums
//   idwt1_tile_kernel  strip = R coefficient rows (+ TAPS/2-1 above) of dL/dLL1 and of the packed signs
//                      -> 2R rows of dL/dx (128-bit stores), times the upstream gradient
// Strips are dealt round-robin to the CTAs (one per SM) so that concurrently streamed strips are adjacent in DRAM.
#include <math.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"
#include "kernels.h"
#include "wavelet_level.cuh"

namespace wtpse {

namespace {

constexpr int kTileConsumers = 512;
constexpr int kTileThreads = kTileConsumers + 32;
constexpr int kInvConsumers = 512;
constexpr int kInvThreads = kInvConsumers + 32;
constexpr int kTileSmem = 216 * 1024;
constexpr uint32_t kTileChunk = 32 * 1024;

__device__ __forceinline__ void tile_bulk(unsigned char* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy, bool hint) {
    const char* s = static_cast<const char*>(src);
    for (uint32_t o = 0; o < bytes; o += kTileChunk) {
        const uint32_t n = min(kTileChunk, bytes - o);
        if (hint) tma_load_1d_hint(dst + o, s + o, n, bar, policy);
        else tma_load_1d(dst + o, s + o, n, bar);
    }
}

struct TileFwdArgs {
    const float* x;         // [nmaps][H][W]
    float* ll;              // [nmaps][H/2][W/2]
    unsigned char* sg;      // [nmaps][H/2][W/2]
    int H, W, nmaps, R, stages;
    float sc;
    double* partial;        // one per CTA
};

template <int TAPS, bool kGrad>
__global__ void __launch_bounds__(kTileThreads, 1) dwt1_tile_kernel(TileFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int HALO = TAPS - 2;
    const int H = a.H, W = a.W, h2 = H >> 1, w2 = W >> 1, R = a.R, S = a.stages;
    const int rows_in = 2 * R + HALO;
    const uint32_t stage_bytes = uint32_t(rows_in) * W * 4u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(S) * stage_bytes);
    uint64_t* empty = full + S;
    const int spm = h2 / R;                                 // strips per map
    const long long T = (long long)a.nmaps * spm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kTileConsumers / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();

    float ab = 0.f;
    double acc = 0.0;
    if (warp == kTileConsumers / 32) {
        if (lane == 0) {
            const uint64_t policy = make_evict_first_policy();
            int n = 0;
            for (long long t = blockIdx.x; t < T; t += gridDim.x, ++n) {
                const int s = n % S;
                if (n >= S) mbar_wait(&empty[s], ((n / S) & 1) ^ 1);
                const long long m = t / spm;
                const int r0 = 2 * R * int(t % spm);
                const int rows_main = min(rows_in, H - r0);
                const float* map = a.x + m * (long long)H * W;
                unsigned char* dst = smem + size_t(s) * stage_bytes;
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                tile_bulk(dst, map + (long long)r0 * W, uint32_t(rows_main) * W * 4u, &full[s], policy, true);
                if (rows_main < rows_in)                    // last strip of the map: the overlap rows wrap to the top
                    tile_bulk(dst + size_t(rows_main) * W * 4u, map, uint32_t(rows_in - rows_main) * W * 4u, &full[s], policy, false);
            }
        }
    } else {
        const int pairs = W >> 2;
        int seg = (R * pairs) / kTileConsumers;
        if (seg < 1) seg = 1;
        const int nseg = (R + seg - 1) / seg;
        const int ntasks = nseg * pairs;
        int n = 0;
        for (long long t = blockIdx.x; t < T; t += gridDim.x, ++n) {
            const int s = n % S;
            mbar_wait(&full[s], (n / S) & 1);
            const float* in = reinterpret_cast<const float*>(smem + size_t(s) * stage_bytes);
            const long long m = t / spm;
            const long long orow = m * h2 + (long long)R * (t % spm);     // first output row of the strip
            for (int task = threadIdx.x; task < ntasks; task += kTileConsumers) {
                const int jj = task % pairs, si = task / pairs;
                const int i0 = si * seg, i1 = min(i0 + seg, R), c0 = 4 * jj;
                int c4 = c0 + 4;
                if (c4 >= W) c4 -= W;
                auto filter = [&](int r, float& lo0, float& hi0, float& lo1, float& hi1) {
                    const float* row = in + r * W;
                    float x[TAPS + 2];
                    const float4 v = *reinterpret_cast<const float4*>(row + c0);
                    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
                    if (TAPS == 4) {
                        const float2 u = *reinterpret_cast<const float2*>(row + c4);
                        x[TAPS] = u.x; x[TAPS + 1] = u.y;
                    }
                    row_filter<TAPS>(x, lo0, hi0, lo1, hi1);
                };
                float lo0[TAPS], hi0[TAPS], lo1[TAPS], hi1[TAPS];
#pragma unroll
                for (int k = 0; k < TAPS - 2; ++k) filter(2 * i0 + k, lo0[k], hi0[k], lo1[k], hi1[k]);
                float* ll = a.ll + (orow + i0) * w2 + 2 * jj;
                unsigned char* sg = a.sg + (orow + i0) * w2 + 2 * jj;
                for (int i = i0; i < i1; ++i) {
                    filter(2 * i + TAPS - 2, lo0[TAPS - 2], hi0[TAPS - 2], lo1[TAPS - 2], hi1[TAPS - 2]);
                    filter(2 * i + TAPS - 1, lo0[TAPS - 1], hi0[TAPS - 1], lo1[TAPS - 1], hi1[TAPS - 1]);
                    float2 LL, LH, HL, HH;
                    col_filter<TAPS>(lo0, hi0, lo1, hi1, LL, LH, HL, HH);
                    *reinterpret_cast<float2*>(ll) = LL;
                    ab += abs_sum(LH, HL, HH) * a.sc;
                    if (kGrad) *reinterpret_cast<unsigned short*>(sg) = static_cast<unsigned short>(sign_pack2(LH, HL, HH));
                    ll += w2;
                    sg += w2;
#pragma unroll
                    for (int k = 0; k < TAPS - 2; ++k) {
                        lo0[k] = lo0[k + 2]; hi0[k] = hi0[k + 2]; lo1[k] = lo1[k + 2]; hi1[k] = hi1[k + 2];
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            acc += double(ab);
            ab = 0.f;
        }
    }

    __shared__ double red[kTileThreads / 32];
    const double sum = warp_sum(acc);
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int q = 0; q < kTileThreads / 32; ++q) tot += red[q];
        a.partial[blockIdx.x] = tot;
    }
}

struct TileInvArgs {
    const float* gll;           // [nmaps][H/2][W/2] (ignored when !kHasLL)
    const unsigned char* sg;    // [nmaps][H/2][W/2]
    float* out;                 // [nmaps][H][W]
    int H, W, nmaps, R, stages;
    float sc;
    const float* upstream;
    const double* partial;      // loss partials of the preceding kernels, summed in fixed order by CTA 0 ...
    int n_partials;
    float* loss;                // ... into loss (nullptr: somebody else does it)
};

// The nmaps * H/2 coefficient rows are cut into one contiguous range per CTA (balanced to a row; with 2 CTAs per SM
// whole strips would leave a quarter of the SMs one strip short), each range into pieces of at most R rows that do
// not cross a map.
struct RowPieces {
    long long r, e;
    int h2, R;
    __device__ RowPieces(long long total, int h2_, int R_) : h2(h2_), R(R_) {
        r = total * blockIdx.x / gridDim.x;
        e = total * (blockIdx.x + 1) / gridDim.x;
    }
    __device__ bool next(long long& m, int& i_first, int& len) {
        if (r >= e) return false;
        m = r / h2;
        i_first = int(r - m * h2);
        len = int(min((long long)min(R, h2 - i_first), e - r));
        r += len;
        return true;
    }
};

template <int TAPS, bool kHasLL>
__global__ void __launch_bounds__(kInvThreads, 1) idwt1_tile_kernel(TileInvArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    constexpr int TOP = TAPS / 2 - 1;                       // coefficient rows above the piece
    const int H = a.H, W = a.W, h2 = H >> 1, w2 = W >> 1, R = a.R, S = a.stages;
    const uint32_t ll_cap = kHasLL ? uint32_t(R + TOP) * w2 * 4u : 0u;
    const uint32_t stage_bytes = ll_cap + ((uint32_t(R + TOP) * w2 + 15u) & ~15u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(S) * stage_bytes);
    uint64_t* empty = full + S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kInvConsumers / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");

    RowPieces pieces((long long)a.nmaps * h2, h2, R);
    long long m;
    int i_first, len;
    if (warp == kInvConsumers / 32) {
        if (lane == 0) {
            for (int n = 0; pieces.next(m, i_first, len); ++n) {
                const int s = n % S;
                if (n >= S) mbar_wait(&empty[s], ((n / S) & 1) ^ 1);
                const int rows_in = len + TOP;
                const int r0 = i_first - TOP;               // first coefficient row needed (-1: wraps to the bottom)
                const int wrap = r0 < 0 ? -r0 : 0;
                unsigned char* dst = smem + size_t(s) * stage_bytes;
                const float* gl = a.gll + m * (long long)h2 * w2;
                const unsigned char* sp = a.sg + m * (long long)h2 * w2;
                mbar_arrive_expect_tx(&full[s], uint32_t(rows_in) * w2 * (kHasLL ? 5u : 1u));
                if (wrap) {
                    if (kHasLL) tile_bulk(dst, gl + (long long)(h2 - wrap) * w2, uint32_t(wrap) * w2 * 4u, &full[s], 0, false);
                    tile_bulk(dst + ll_cap, sp + (long long)(h2 - wrap) * w2, uint32_t(wrap) * w2, &full[s], 0, false);
                }
                if (kHasLL)
                    tile_bulk(dst + size_t(wrap) * w2 * 4u, gl + (long long)(r0 + wrap) * w2, uint32_t(rows_in - wrap) * w2 * 4u, &full[s], 0, false);
                tile_bulk(dst + ll_cap + size_t(wrap) * w2, sp + (long long)(r0 + wrap) * w2, uint32_t(rows_in - wrap) * w2, &full[s], 0, false);
            }
        }
        __syncwarp();
        if (a.loss && blockIdx.x == 0) {                    // fixed-order sum of the loss partials, once this CTA's loads are queued
            double s = 0.0;
            for (int i = lane; i < a.n_partials; i += 32) s += a.partial[i];
            s = warp_sum(s);
            if (lane == 0) a.loss[0] = float(s);
        }
    } else {
        const float gs = a.upstream ? __ldg(a.upstream) : 1.0f;
        const int pairs = w2 >> 1;
        for (int n = 0; pieces.next(m, i_first, len); ++n) {
            const int s = n % S;
            const int groups = max(1, kInvConsumers / pairs);             // row groups that fit the consumer threads
            const int seg = (len + groups - 1) / groups;
            const int nseg = (len + seg - 1) / seg;
            const int ntasks = nseg * pairs;
            mbar_wait(&full[s], (n / S) & 1);
            const float* gl = reinterpret_cast<const float*>(smem + size_t(s) * stage_bytes);           // row 0 = row above the piece (db2)
            const unsigned char* sp = smem + size_t(s) * stage_bytes + ll_cap;
            float* obase = a.out + m * (long long)H * W + (long long)(2 * i_first) * W;
            for (int task = threadIdx.x; task < ntasks; task += kInvConsumers) {
                const int q = task % pairs, si = task / pairs;
                const int i0 = si * seg, i1 = min(i0 + seg, len);
                const int c0 = 2 * q, cm = c0 ? c0 - 1 : w2 - 1;
                auto synth = [&](int r, float (&tL)[4], float (&tH)[4]) {       // r: buffer row
                    float2 l01 = make_float2(0.f, 0.f);
                    if (kHasLL) l01 = *reinterpret_cast<const float2*>(gl + r * w2 + c0);
                    const unsigned b01 = *reinterpret_cast<const unsigned short*>(sp + r * w2 + c0);
                    float lm = 0.f;
                    unsigned bm = 0;
                    if (TAPS == 4) {
                        if (kHasLL) lm = gl[r * w2 + cm];
                        bm = sp[r * w2 + cm];
                    }
                    col_synth_vals<TAPS, kHasLL>(l01, lm, b01, bm, a.sc, tL, tH);
                };
                float pL[4] = {0.f, 0.f, 0.f, 0.f}, pH[4] = {0.f, 0.f, 0.f, 0.f};
                if (TAPS == 4) synth(i0, pL, pH);                               // coefficient row i0-1 = buffer row i0
                float* out = obase + (long long)(2 * i0) * W + 4 * q;
                for (int i = i0; i < i1; ++i) {
                    float cL[4], cH[4];
                    synth(i + TOP, cL, cH);
#pragma unroll
                    for (int pr = 0; pr < 2; ++pr) {
                        float o[4];
                        row_synth<TAPS>(cL, cH, pL, pH, pr, gs, o);
                        *reinterpret_cast<float4*>(out + (long long)pr * W) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                    out += 2 * W;
#pragma unroll
                    for (int k = 0; k < 4; ++k) { pL[k] = cL[k]; pH[k] = cH[k]; }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
    }
}

// largest divisor of n that is <= cap (>= 1)
int divisor_at_most(int n, int cap) {
    for (int d = min(n, cap); d > 1; --d)
        if (n % d == 0) return d;
    return 1;
}

template <typename Kernel, typename Args>
cudaError_t launch_tile(Kernel kernel, int grid, int threads, size_t smem, cudaStream_t stream, const Args& args, bool pdl) {
    // shared-memory opt-in: once per (kernel, device, footprint).  Keyed by the function pointer: instantiations with
    // the same signature share this template instance.
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> opted_in;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    {
        std::lock_guard<std::mutex> lock(mu);
        const auto key = std::make_pair(reinterpret_cast<const void*>(kernel), dev);
        auto it = opted_in.find(key);
        if (it == opted_in.end() || it->second < smem) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return e;
            opted_in[key] = smem;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args);
}

}  // namespace

int g_wavelet_tiles = 1;        // diagnostics: 0 = per-thread global loads (wavelet_stream.cu) for level 1 of the streamed plan

// Strip height (output / coefficient rows) and ring depth of the two pipelines; R = 0: shape not taken.
void wavelet_tile_plan(int H, int W, int taps, bool has_ll, int* R_fwd, int* S_fwd, int* R_inv, int* S_inv) {
    *R_fwd = *S_fwd = *R_inv = *S_inv = 0;
    if (!g_wavelet_tiles || (W % 32) || (H % 2) || W > 4096) return;
    const int h2 = H / 2, w2 = W / 2;
    {   // analysis: ~4 rows per consumer thread, at least 2 stages
        const int pairs = W / 4;
        int R = divisor_at_most(h2, max(1, 4 * kTileConsumers / pairs));
        while (R > 1 && size_t(2) * (2 * R + taps - 2) * W * 4 > size_t(kTileSmem)) R = divisor_at_most(h2, R - 1);
        const size_t stage = size_t(2 * R + taps - 2) * W * 4;
        const int S = int(std::min<size_t>(4, kTileSmem / stage));
        if (S >= 2 && (taps == 2 || 2 * R >= taps - 2)) { *R_fwd = R; *S_fwd = S; }
    }
    {   // synthesis: ~8 rows per consumer thread
        const int pairs = w2 / 2;
        int R = divisor_at_most(h2, max(1, 8 * kInvConsumers / pairs));
        auto stage_of = [&](int r) {
            const size_t rows = size_t(r + taps / 2 - 1);
            return (has_ll ? rows * w2 * 4 : 0) + ((rows * w2 + 15) & ~size_t(15));
        };
        const size_t budget = size_t(kTileSmem);
        while (R > 1 && 2 * stage_of(R) > budget) R = divisor_at_most(h2, R - 1);
        const int S = int(std::min<size_t>(4, budget / stage_of(R)));
        if (S >= 2) { *R_inv = R; *S_inv = S; }
    }
}

cudaError_t launch_dwt1_tiles(const float* x, float* ll, unsigned char* sg, int nmaps, int H, int W, int taps, int R, int S,
                              float sc, bool grad, double* partial, int sm_count, cudaStream_t stream, int* n_partials) {
    TileFwdArgs a;
    a.x = x; a.ll = ll; a.sg = sg; a.H = H; a.W = W; a.nmaps = nmaps; a.R = R; a.stages = S; a.sc = sc; a.partial = partial;
    const long long T = (long long)nmaps * (H / 2 / R);
    const int grid = int(std::min<long long>(sm_count, T));
    const size_t smem = size_t(S) * (2 * R + taps - 2) * W * 4 + size_t(2 * S) * sizeof(uint64_t);
    *n_partials = grid;
    if (grad) return taps == 2 ? launch_tile(dwt1_tile_kernel<2, true>, grid, kTileThreads, smem, stream, a, false) : launch_tile(dwt1_tile_kernel<4, true>, grid, kTileThreads, smem, stream, a, false);
    return taps == 2 ? launch_tile(dwt1_tile_kernel<2, false>, grid, kTileThreads, smem, stream, a, false) : launch_tile(dwt1_tile_kernel<4, false>, grid, kTileThreads, smem, stream, a, false);
}

cudaError_t launch_idwt1_tiles(const float* gll, const unsigned char* sg, float* out, int nmaps, int H, int W, int taps, int R,
                               int S, float sc, const float* upstream, bool has_ll, const double* partial, int n_partials,
                               float* loss, int sm_count, cudaStream_t stream) {
    TileInvArgs a;
    a.gll = gll; a.sg = sg; a.out = out; a.H = H; a.W = W; a.nmaps = nmaps; a.R = R; a.stages = S; a.sc = sc; a.upstream = upstream;
    a.partial = partial; a.n_partials = n_partials; a.loss = loss;
    const int w2 = W / 2;
    const long long rows = (long long)nmaps * (H / 2);
    const int grid = int(std::min<long long>(sm_count, std::max<long long>(1, rows / 8)));
    const size_t srows = size_t(R + taps / 2 - 1);
    const size_t stage = (has_ll ? srows * w2 * 4 : 0) + ((srows * w2 + 15) & ~size_t(15));
    const size_t smem = size_t(S) * stage + size_t(2 * S) * sizeof(uint64_t);
    if (has_ll) return taps == 2 ? launch_tile(idwt1_tile_kernel<2, true>, grid, kInvThreads, smem, stream, a, true) : launch_tile(idwt1_tile_kernel<4, true>, grid, kInvThreads, smem, stream, a, true);
    return taps == 2 ? launch_tile(idwt1_tile_kernel<2, false>, grid, kInvThreads, smem, stream, a, true) : launch_tile(idwt1_tile_kernel<4, false>, grid, kInvThreads, smem, stream, a, true);
}

}  // namespace wtpse

// Track W, streaming level-1 kernels + the plan that combines them with the cluster-resident kernel.
// PARITY UNPINNED (see wavelet.cu / oracle/wavelet_np.py).
//
// A 512 x 512 map needs a cluster of 8 CTAs to be resident, so only ~15 maps are in flight on 148 SMs and every level
// costs two cluster barriers.  Peeling the FIRST level off fixes both: level 1 holds 3/4 of all coefficients, needs
// no inter-CTA exchange at all when it streams from / to global memory, and what is left (the low-low band, a quarter
// of the map) is small enough for clusters of 2..4 with several clusters per SM.
//   A  dwt1_stream_kernel   x -> LL1 (fp32, N/4) + packed signs of the level-1 detail bands (1 byte per site, N/4 bytes)
//                           + partial sums of |d1|;                     reads 4N, writes 1.25N bytes
//   B  wavelet_resident_kernel on LL1 (levels 2..J, in place: LL1 -> dL/dLL1), everything in shared memory
//   C  idwt1_stream_kernel  dL/dLL1 + signs -> dL/dx (times the upstream gradient);  reads 1.25N, writes 4N bytes
// A and C walk down the rows of a 4-column strip with the overlapping rows carried in registers (wavelet_level.cuh),
// 128-bit coalesced loads / stores, no shared memory, every SM busy.
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include "wavelet_level.cuh"

namespace wtpse {

namespace {

constexpr int kStreamThreads = 128;

struct StreamFwdArgs {
    const float* x;         // [nmaps][H][W]
    float* ll;              // [nmaps][H/2][W/2]
    unsigned char* sg;      // [nmaps][H/2][W/2]
    int H, W, seg;
    float sc;               // w_1 / (3 * H/2 * W/2 * nmaps)
    double* partial;        // one per CTA
};

template <int TAPS, bool kGrad>
__global__ void __launch_bounds__(kStreamThreads, 7) dwt1_stream_kernel(StreamFwdArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int H = a.H, W = a.W, h2 = H >> 1, w2 = W >> 1, pairs = W >> 2;
    const int nseg = (h2 + a.seg - 1) / a.seg;
    const int t = blockIdx.x * kStreamThreads + threadIdx.x, m = blockIdx.y;
    float ab = 0.f;
    if (t < nseg * pairs) {
        const int jj = t % pairs, si = t / pairs;
        const int i0 = si * a.seg, i1 = min(i0 + a.seg, h2), c0 = 4 * jj;
        int c4 = c0 + 4;
        if (c4 >= W) c4 -= W;
        const long long map_elems = (long long)H * W;
        const float* src = a.x + (long long)m * map_elems;
        float* ll = a.ll + (long long)m * h2 * w2 + (long long)i0 * w2 + 2 * jj;
        unsigned char* sg = a.sg + (long long)m * h2 * w2 + (long long)i0 * w2 + 2 * jj;
        // raw columns 4jj .. 4jj+TAPS+1 of one input row; rows past the end wrap to the top of the map
        struct Raw { float4 v; float2 u; };
        auto load_row = [&](const float* row) {
            Raw r;
            r.v = __ldg(reinterpret_cast<const float4*>(row + c0));
            r.u = TAPS == 4 ? __ldg(reinterpret_cast<const float2*>(row + c4)) : make_float2(0.f, 0.f);
            return r;
        };
        auto filter = [&](const Raw& r, float& lo0, float& hi0, float& lo1, float& hi1) {
            float x[TAPS + 2];
            x[0] = r.v.x; x[1] = r.v.y; x[2] = r.v.z; x[3] = r.v.w;
            if (TAPS == 4) { x[TAPS] = r.u.x; x[TAPS + 1] = r.u.y; }
            row_filter<TAPS>(x, lo0, hi0, lo1, hi1);
        };
        const float* rp = src + (long long)(2 * i0) * W;           // 2*i0 + TAPS-3 < H: the carried rows never wrap
        float lo0[TAPS], hi0[TAPS], lo1[TAPS], hi1[TAPS];
#pragma unroll
        for (int k = 0; k < TAPS - 2; ++k) filter(load_row(rp + (long long)k * W), lo0[k], hi0[k], lo1[k], hi1[k]);
        rp += (long long)(TAPS - 2) * W;                           // the two NEW rows of output row i0
        auto new_rows = [&](int i, Raw& ra, Raw& rb) {
            const float* p = rp;
            if (TAPS == 4 && i == h2 - 1) p -= map_elems;           // rows H, H+1 -> 0, 1
            ra = load_row(p);
            rb = load_row(p + W);
        };
        // software pipeline: the loads of output rows i+1 .. i+kDepth are in flight while row i is computed.  One SM
        // needs ~90 KB in flight to cover the loaded HBM latency at its share of the bandwidth; a thread holds
        // kDepth * 48 B, so depth 1 (what the compiler does by itself) stalls at half the bandwidth.
        constexpr int kDepth = 3;
        Raw ra[kDepth] = {}, rb[kDepth] = {};
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            if (i0 + d < i1) new_rows(i0 + d, ra[d], rb[d]);
            rp += 2 * W;
        }
        for (int ib = i0; ib < i1; ib += kDepth) {
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                const int i = ib + d;
                if (i < i1) {
                    const Raw ca = ra[d], cb = rb[d];
                    if (i + kDepth < i1) new_rows(i + kDepth, ra[d], rb[d]);
                    rp += 2 * W;
                    filter(ca, lo0[TAPS - 2], hi0[TAPS - 2], lo1[TAPS - 2], hi1[TAPS - 2]);
                    filter(cb, lo0[TAPS - 1], hi0[TAPS - 1], lo1[TAPS - 1], hi1[TAPS - 1]);
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
        }
    }
    __shared__ double red[kStreamThreads / 32];
    const double s = warp_sum(double(ab));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
#pragma unroll
        for (int q = 0; q < kStreamThreads / 32; ++q) tot += red[q];
        a.partial[blockIdx.y * gridDim.x + blockIdx.x] = tot;
    }
}

struct StreamInvArgs {
    const float* gll;           // [nmaps][H/2][W/2] gradient wrt the low-low band (ignored when !kHasLL)
    const unsigned char* sg;    // [nmaps][H/2][W/2]
    float* out;                 // [nmaps][H][W]
    int H, W, seg;
    float sc;
    const float* upstream;      // optional device scalar
};

template <int TAPS, bool kHasLL>
__global__ void __launch_bounds__(kStreamThreads) idwt1_stream_kernel(StreamInvArgs a) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int H = a.H, W = a.W, h2 = H >> 1, w2 = W >> 1, pairs = w2 >> 1;
    const int nseg = (h2 + a.seg - 1) / a.seg;
    const int t = blockIdx.x * kStreamThreads + threadIdx.x, m = blockIdx.y;
    if (t >= nseg * pairs) return;
    const int q = t % pairs, si = t / pairs;
    const int i0 = si * a.seg, i1 = min(i0 + a.seg, h2);
    const int c0 = 2 * q, cm = c0 ? c0 - 1 : w2 - 1;
    const float* gl = a.gll + (long long)m * h2 * w2;
    const unsigned char* sp = a.sg + (long long)m * h2 * w2;
    float* out = a.out + (long long)m * H * W + (long long)(2 * i0) * W + 4 * q;
    const float gs = a.upstream ? __ldg(a.upstream) : 1.0f;
    // raw operands of one coefficient row: low-low gradient and packed signs of columns 2q, 2q+1 and (db2) 2q-1
    struct Raw { float2 l01; float lm; unsigned b01, bm; };
    auto load_row = [&](int r) {
        Raw v;
        if (r < 0) r += h2;
        const long long o = (long long)r * w2;
        v.l01 = kHasLL ? __ldg(reinterpret_cast<const float2*>(gl + o + c0)) : make_float2(0.f, 0.f);
        v.b01 = __ldg(reinterpret_cast<const unsigned short*>(sp + o + c0));
        v.lm = (TAPS == 4 && kHasLL) ? __ldg(gl + o + cm) : 0.f;
        v.bm = TAPS == 4 ? unsigned(__ldg(sp + o + cm)) : 0u;
        return v;
    };
    float pL[4] = {0.f, 0.f, 0.f, 0.f}, pH[4] = {0.f, 0.f, 0.f, 0.f};
    if (TAPS == 4) {
        const Raw v = load_row(i0 - 1);
        col_synth_vals<TAPS>(v.l01, v.lm, v.b01, v.bm, a.sc, pL, pH);
    }
    // software pipeline, two coefficient rows ahead (a row is five registers)
    Raw r0 = load_row(i0), r1 = load_row(min(i0 + 1, i1 - 1));
    for (int i = i0; i < i1; ++i) {
        const Raw r2 = load_row(min(i + 2, i1 - 1));
        float cL[4], cH[4];
        col_synth_vals<TAPS>(r0.l01, r0.lm, r0.b01, r0.bm, a.sc, cL, cH);
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            float o[4];
            row_synth<TAPS>(cL, cH, pL, pH, pr, gs, o);
            *reinterpret_cast<float4*>(out + (long long)pr * W) = make_float4(o[0], o[1], o[2], o[3]);
        }
        out += 2 * W;
#pragma unroll
        for (int k = 0; k < 4; ++k) { pL[k] = cL[k]; pH[k] = cH[k]; }
        r0 = r1;
        r1 = r2;
    }
}

template <typename Kernel, typename Args>
cudaError_t launch_stream(Kernel kernel, dim3 grid, cudaStream_t stream, const Args& args, bool pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kStreamThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args);
}

int stream_seg(int h2) { return h2 >= 64 ? 16 : (h2 >= 16 ? 8 : h2); }

}  // namespace

int g_wavelet_split = -1;       // -1: automatic, 0: whole map resident (when it fits), 1: level 1 streamed (when possible)
int g_wavelet_haar_min_log2px = 16;     // Haar: smallest map (log2 pixels) that takes the pass kernels instead of the band kernel
int g_wavelet_peel_max = 8;     // diagnostics: most levels the streamed plan may peel off before the resident stage

static bool level_streamable(int H, int W, int taps) {
    return !(W % 8) && !(H % 2) && H >= 4 && W >= 8 && (taps == 2 || (H >= 4 && W >= 8));
}
static bool level_tiled(int H, int W, int taps, bool has_ll) {
    int Rf, Sf, Ri, Si;
    wavelet_tile_plan(H, W, taps, has_ll, &Rf, &Sf, &Ri, &Si);
    return Rf > 0 && Ri > 0;
}

// Factored pass kernels (wavelet_db2.cu): the streamed levels as passes of two (or one) levels.  Fills pass[] with the
// number of levels of each pass and returns the number of passes (0: the pass kernels do not take this shape);
// *k = levels streamed, *cs = cluster size (Haar: bands per map) of the resident stage that follows (1 when k == J).
static int db2_pass_plan(int H, int W, int taps, int J, int nmaps, int pass[16], int* k_out, int* cs) {
    const bool haar = taps == 2;
    int k = 0, np = 0;
    *cs = 0;
    while (k < J && k < g_wavelet_peel_max) {
        const int Hc = H >> k, Wc = W >> k, rem = J - k;
        int r0, r1, r2, r3, r4;
        if (k > 0 && !g_wavelet_db2_deep) {
            // the band fits a cluster of <= 2: go resident -- unless two more levels can be taken by a two-level pass, which beats the
            // resident stage on bands of 256^2 and larger (1024^2 maps, J = 4: 265.5 -> 259.5 us).  On 128^2 bands a pass pair
            // costs what the resident stage costs (each extra kernel boundary is ~3-4 us of drain and fill: 512^2 J = 4 51.4 vs 52.2 us,
            // J = 5 57.5 vs 54.2 us), so the band goes resident.  Haar: the same rule with its band kernel (no clusters) as the
            // resident stage.
            const int c = wavelet_resident_cluster(Hc, Wc, taps, rem, nmaps);
            const bool two_more = rem >= 2 && Wc >= 256 && k + 2 <= g_wavelet_peel_max && wavelet_db2_pass(Hc, Wc, true, rem > 2, &r0, &r1, &r2, &r3, &r4, haar);
            if (c > 0 && (haar || c <= 2) && !two_more) { *cs = haar ? 1 : c; break; }
        }
        if (rem >= 2 && k + 2 <= g_wavelet_peel_max && wavelet_db2_pass(Hc, Wc, true, rem > 2, &r0, &r1, &r2, &r3, &r4, haar)) {
            pass[np++] = 2;
            k += 2;
        } else if (wavelet_db2_pass(Hc, Wc, false, rem > 1, &r0, &r1, &r2, &r3, &r4, haar)) {
            pass[np++] = 1;
            k += 1;
        } else {
            break;
        }
    }
    if (k == 0) return 0;
    if (k < J) {
        const int c = wavelet_resident_cluster(H >> k, W >> k, taps, J - k, nmaps);
        if (c == 0) return 0;
        *cs = haar ? 1 : c;
    } else {
        *cs = 1;
    }
    *k_out = k;
    return np;
}

// Streamed plan: the first k levels run global-to-global (TMA tile pipelines; level 1 may fall back to the per-thread
// kernels), the remaining J-k levels in the cluster-resident kernel on the k-th low-low band.  k grows while that band
// would still need a cluster of more than 2 CTAs (a 1024 x 1024 map peels two levels: 63 us of resident stage on
// 512 x 512 bands in clusters of 8 become a second streamed level plus clusters of 2).  k == J: nothing resident.
// Returns k, 0 if the plan does not exist; *cs = cluster size of the resident stage (1 when k == J).
int wavelet_stream_levels(int H, int W, int taps, int J, int nmaps, int* cs) {
    *cs = 0;
    {
        int pass[16], k = 0;
        if (db2_pass_plan(H, W, taps, J, nmaps, pass, &k, cs) > 0) return k;
        *cs = 0;
    }
    if (!level_streamable(H, W, taps)) return 0;
    int best_k = 0, best_cs = 0;
    for (int k = 1; k <= J && k <= g_wavelet_peel_max; ++k) {
        // levels 2..k go through the tile pipelines only
        if (k > 1 && !(level_streamable(H >> (k - 1), W >> (k - 1), taps) && level_tiled(H >> (k - 1), W >> (k - 1), taps, k < J))) break;
        if (k == J) { best_k = k; best_cs = 1; break; }
        const int c = wavelet_resident_cluster(H >> k, W >> k, taps, J - k, nmaps);
        if (c > 0) { best_k = k; best_cs = taps == 2 ? 1 : c; }
        if (c > 0 && (taps == 2 || c <= 2)) break;
    }
    *cs = best_cs;
    return best_k;
}

// 0: no fused path; 1: whole map resident; 2: streamed plan
int wavelet_fused_plan(int H, int W, int taps, int J, int nmaps) {
    const int whole = wavelet_resident_cluster(H, W, taps, J, nmaps);
    int cs;
    const bool split = wavelet_stream_levels(H, W, taps, J, nmaps, &cs) > 0;
    if (g_wavelet_split == 0) return whole ? 1 : 0;
    if (g_wavelet_split == 1) return split ? 2 : (whole ? 1 : 0);
    // automatic.  db2: a map that needs a cluster of more than 2 CTAs leaves SMs idle and pays two cluster barriers per
    // level -> streamed plan.  Haar: row bands are independent work items (no cluster, any map size), one kernel and
    // 8 B per element, but its deep levels have little parallelism per band.  Measured, fused band kernel vs streamed
    // plan, 32 x 2 maps: 1024^2, J = 1 / 3 / 5: 106 / 146 / 177 us vs 120 / 165 / 177; 512^2: 35 / 47 / 55 vs 37 / 46 / 48.
    if (whole && taps == 2) {
        const long long px = (long long)H * W;
        // Haar through the two-level passes (wavelet_db2.cu) from J = 2 on large maps: the band kernel reaches 0.69 / 0.59 / 0.52 of the
        // roofline at 1024^2 for J = 2 / 3 / 4 (its deep levels have little parallelism per band)
        // (a handful of small maps is launch-bound and keeps the single band kernel: 8 x 2 maps of 256^2, J = 3: 17.1 vs 21.6 us)
        const bool enough = nmaps == 0 || (long long)nmaps * px >= (1ll << 22);
        if (g_wavelet_haar_passes && g_wavelet_db2 && split && enough && J >= 2 && px >= (1ll << g_wavelet_haar_min_log2px)) return 2;
        if (J <= 2 || px >= (1ll << 20) || px <= (1ll << 16) || !split) return 1;
    } else if (whole && whole <= 2) {
        // db2 maps that fit a cluster of <= 2 (256^2 and smaller): the pass kernels beat the whole-map-resident kernel by a third
        // (128 x 2 maps of 256^2: J = 2 / 3 / 4 70.1 / 80.2 / 89.0 -> 45.4 / 56.0 / 60.2 us; 512 x 2 maps of 128^2, J = 2: 76.5 -> 65.9 us),
        // and still by a few microseconds when a handful of maps leaves everything launch-bound (4 x 2 maps of 256^2, J = 3: 25.0 -> 21.4 us)
        int pass[16], kk = 0, c2 = 0;
        if (split && g_wavelet_db2 && db2_pass_plan(H, W, taps, J, nmaps, pass, &kk, &c2) > 0) return 2;
        return 1;
    }
    return split ? 2 : (whole ? 1 : 0);
}

size_t wavelet_stream_partials(int nmaps, int H, int W) {
    const int h2 = H / 2, seg = stream_seg(h2);
    const int nseg = (h2 + seg - 1) / seg;
    const size_t per_map = (size_t(nseg) * (W / 4) + kStreamThreads - 1) / kStreamThreads;
    return per_map * nmaps + 16 * 1024;          // + one per CTA of every tile-pipeline launch
}

// scratch (floats): low-low bands LL_1 .. LL_k back to back, then the packed sign planes S_1 .. S_k (bytes)
cudaError_t launch_wavelet_loss_split(const float* x, int nmaps, int H, int W, int taps, int J, const float* weights_host,
                                      const float* upstream, float* loss, float* grad, float* scratch, double* partial,
                                      int sm_count, cudaStream_t stream) {
    int cs;
    const int k = wavelet_stream_levels(H, W, taps, J, nmaps, &cs);
    if (k < 1) return cudaErrorInvalidValue;
    float* ll[16];
    unsigned char* sg[16];
    {
        float* p = scratch;
        for (int i = 1; i <= k; ++i) { ll[i] = p; p += size_t(nmaps) * (H >> i) * (W >> i); }
        unsigned char* q = reinterpret_cast<unsigned char*>(p);
        for (int i = 1; i <= k; ++i) { sg[i] = q; q += (size_t(nmaps) * (H >> i) * (W >> i) + 255) & ~size_t(255); }
    }
    auto scale_of = [&](int i) { return weights_host[i - 1] / (3.0f * float(H >> i) * float(W >> i) * float(nmaps)); };
    cudaError_t e = cudaSuccess;
    int np = 0;
    {
        // pass kernels (wavelet_db2.cu; factored db2 or Haar): 2 (or 1) levels per pass down, resident stage, the same passes up
        const bool haar = taps == 2;
        int pass[16], kk = 0, cs2 = 0;
        const int npass = db2_pass_plan(H, W, taps, J, nmaps, pass, &kk, &cs2);
        if (npass > 0 && kk == k) {
            int lvl = 0;                                            // levels done
            for (int q = 0; q < npass; ++q) {
                const bool two = pass[q] == 2;
                const int Hi = H >> lvl, Wi = W >> lvl, last = lvl + pass[q];
                int n = 0;
                e = launch_db2_analysis(lvl == 0 ? x : ll[lvl], last < J ? ll[last] : nullptr, sg[lvl + 1], two ? sg[lvl + 2] : nullptr, nmaps, Hi, Wi, two,
                                        scale_of(lvl + 1), two ? scale_of(lvl + 2) : 0.f, grad != nullptr, q > 0, partial + np, sm_count, stream, &n, haar);
                if (e != cudaSuccess) return e;
                np += n;
                lvl = last;
            }
            if (k < J) {
                int nb = 0;
                e = launch_wavelet_resident(ll[k], nmaps, H >> k, W >> k, taps, J - k, weights_host + k, nullptr, nullptr, grad ? ll[k] : nullptr,
                                            partial + np, stream, &nb);
                if (e != cudaSuccess) return e;
                np += nb;
            }
            if (!grad) return launch_wavelet_loss_final(partial, np, loss, stream);
            for (int q = npass - 1; q >= 0; --q) {
                const bool two = pass[q] == 2;
                lvl -= pass[q];
                const int Hi = H >> lvl, Wi = W >> lvl, last = lvl + pass[q];
                e = launch_db2_synthesis(ll[last], sg[lvl + 1], two ? sg[lvl + 2] : nullptr, lvl == 0 ? grad : ll[lvl], nmaps, Hi, Wi, two,
                                         last < J, scale_of(lvl + 1), two ? scale_of(lvl + 2) : 0.f, lvl == 0 ? upstream : nullptr, partial, np,
                                         lvl == 0 ? loss : nullptr, sm_count, stream, haar);
                if (e != cudaSuccess) return e;
            }
            return cudaSuccess;
        }
    }
    // ---- analysis, levels 1..k ----
    for (int i = 1; i <= k; ++i) {
        const int Hi = H >> (i - 1), Wi = W >> (i - 1);             // input of this level
        const float* in = i == 1 ? x : ll[i - 1];
        int Rf, Sf, Ri, Si, n = 0;
        wavelet_tile_plan(Hi, Wi, taps, i < J, &Rf, &Sf, &Ri, &Si);
        if (Rf) {
            e = launch_dwt1_tiles(in, ll[i], sg[i], nmaps, Hi, Wi, taps, Rf, Sf, scale_of(i), grad != nullptr, partial + np, sm_count, stream, &n);
        } else {
            const int h2 = Hi / 2, seg = stream_seg(h2), nseg = (h2 + seg - 1) / seg;
            StreamFwdArgs fa;
            fa.x = in; fa.ll = ll[i]; fa.sg = sg[i]; fa.H = Hi; fa.W = Wi; fa.seg = seg; fa.sc = scale_of(i); fa.partial = partial + np;
            const dim3 gf(unsigned((size_t(nseg) * (Wi / 4) + kStreamThreads - 1) / kStreamThreads), nmaps);
            if (grad) e = taps == 2 ? launch_stream(dwt1_stream_kernel<2, true>, gf, stream, fa, false) : launch_stream(dwt1_stream_kernel<4, true>, gf, stream, fa, false);
            else e = taps == 2 ? launch_stream(dwt1_stream_kernel<2, false>, gf, stream, fa, false) : launch_stream(dwt1_stream_kernel<4, false>, gf, stream, fa, false);
            n = int(gf.x * gf.y);
        }
        if (e != cudaSuccess) return e;
        np += n;
    }
    // ---- levels k+1..J resident, in place on LL_k ----
    if (k < J) {
        int nb = 0;
        e = launch_wavelet_resident(ll[k], nmaps, H >> k, W >> k, taps, J - k, weights_host + k, nullptr, nullptr, grad ? ll[k] : nullptr,
                                    partial + np, stream, &nb);
        if (e != cudaSuccess) return e;
        np += nb;
    }
    if (!grad) return launch_wavelet_loss_final(partial, np, loss, stream);
    // ---- synthesis, levels k..1: dL/dLL_i + signs_i -> dL/dLL_{i-1} (written over LL_{i-1}) ... -> dL/dx ----
    bool loss_done = false;
    for (int i = k; i >= 1; --i) {
        const int Hi = H >> (i - 1), Wi = W >> (i - 1);             // output of this level
        float* out = i == 1 ? grad : ll[i - 1];
        const float* up = i == 1 ? upstream : nullptr;
        const bool has_ll = i < J;
        int Rf, Sf, Ri, Si;
        wavelet_tile_plan(Hi, Wi, taps, has_ll, &Rf, &Sf, &Ri, &Si);
        if (Ri) {
            // the last synthesis kernel sums the loss partials itself (one launch less)
            const bool fold = (i == 1);
            e = launch_idwt1_tiles(ll[i], sg[i], out, nmaps, Hi, Wi, taps, Ri, Si, scale_of(i), up, has_ll, partial, np,
                                   fold ? loss : nullptr, sm_count, stream);
            loss_done = loss_done || fold;
        } else {
            const int h2 = Hi / 2, seg = stream_seg(h2), nseg = (h2 + seg - 1) / seg;
            StreamInvArgs ia;
            ia.gll = ll[i]; ia.sg = sg[i]; ia.out = out; ia.H = Hi; ia.W = Wi; ia.seg = seg; ia.sc = scale_of(i); ia.upstream = up;
            const dim3 gi(unsigned((size_t(nseg) * (Wi / 4) + kStreamThreads - 1) / kStreamThreads), nmaps);
            if (has_ll) e = taps == 2 ? launch_stream(idwt1_stream_kernel<2, true>, gi, stream, ia, false) : launch_stream(idwt1_stream_kernel<4, true>, gi, stream, ia, false);
            else e = taps == 2 ? launch_stream(idwt1_stream_kernel<2, false>, gi, stream, ia, false) : launch_stream(idwt1_stream_kernel<4, false>, gi, stream, ia, false);
        }
        if (e != cudaSuccess) return e;
    }
    if (!loss_done) e = launch_wavelet_loss_final(partial, np, loss, stream);
    return e;
}

}  // namespace wtpse

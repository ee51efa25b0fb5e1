// C ABI (include/wtpse_b200.h) over the kernels.  No exceptions, no torch types, no host syncs on
// the device-pointer entry points.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/wtpse_b200.h"
#include "../../include/wtpse_b200_debug.h"
#include "common.cuh"
#include "kernels.h"
#include "profile.cuh"

using namespace wtpse;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(WTPSE_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

int sm_count_cached() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev] = n;
    }
    return cache[dev];
}

// Diagnostic switches (include/wtpse_b200_debug.h, wtpse_debug_set): process-wide, read when a call is enqueued.
int g_cl_tma_launches = 0;               // read-only counter: channels-last calls that took the tensor-map TMA kernels
int g_tail_stamps = 0;                   // 1: the in-kernel tail records phase timestamps in the last 128 bytes of the workspace
int g_fused_tail = 1;                     // 0: forward/backward as chains of separate kernels (the fallback path) for every shape

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WhitenWorkspace {
    int* ticket;             // [B + 1], FIRST in the workspace: zero on entry, zero on exit (wtpse_whitening_ticket_bytes)
    float* partial;
    int* slot_count;
    float* vd;
    float* statd;
    void* scratch;
    long long* stamps;       // 16 x int64, diagnostics only
    size_t total;
};

// K is not known when the workspace is sized: reserve the epilogue scratch for K*K <= B*B + 64 domain pairs
// (check_common rejects more; only domains that can be non-empty matter).
size_t scratch_bytes_any_k(int B) { return align_up(epilogue_scratch_bytes(B, 1) + (size_t(B) * B + 64) * sizeof(double), 256); }

WhitenWorkspace carve(void* base, int B, long long P, int sms) {
    WhitenWorkspace w;
    size_t off = 0;
    const size_t partial_bytes = align_up(gram_partial_floats(B, P, sms) * sizeof(float), 256);
    const size_t count_bytes = align_up(size_t(B) * sizeof(int), 256);
    char* p = static_cast<char*>(base);
    w.ticket = reinterpret_cast<int*>(p + off); off += align_up(size_t(B + 1) * sizeof(int), 256);
    w.partial = reinterpret_cast<float*>(p + off); off += partial_bytes;
    w.slot_count = reinterpret_cast<int*>(p + off); off += count_bytes;
    w.vd = reinterpret_cast<float*>(p + off); off += align_up(size_t(B) * 124 * sizeof(float), 256);
    w.statd = reinterpret_cast<float*>(p + off); off += align_up(size_t(B) * 2 * sizeof(float), 256);
    w.scratch = p + off; off += scratch_bytes_any_k(B);
    w.stamps = reinterpret_cast<long long*>(p + off); off += 128;
    w.total = off;
    return w;
}

int check_common(const void* z, int B, int C, long long P, int n, int K) {
    if (!z) return fail(WTPSE_ERR_INVALID, "null input pointer");
    if (C != WTPSE_CHANNELS) return fail(WTPSE_ERR_INVALID, "whitening loss is defined for C == 16 channels, got %d", C);
    if (B <= 0 || P <= 1) return fail(WTPSE_ERR_INVALID, "need B >= 1 and H*W >= 2 (got B=%d, P=%lld)", B, P);
    if (n < 0 || K < 0) return fail(WTPSE_ERR_INVALID, "negative domain configuration (n=%d, K=%d)", n, K);
    if ((long long)K * K > (long long)B * B + 64) return fail(WTPSE_ERR_INVALID, "n_domains=%d too large for B=%d", K, B);
    return WTPSE_OK;
}

}  // namespace

extern "C" {

int wtpse_abi_version(void) { return WTPSE_ABI_VERSION; }
const char* wtpse_last_error(void) { return g_err; }
int wtpse_sm_count(void) { return sm_count_cached(); }

size_t wtpse_whitening_workspace_bytes(int B, int64_t P) {
    if (B <= 0 || P <= 0) return 0;
    return carve(nullptr, B, P, sm_count_cached()).total;
}

size_t wtpse_whitening_ticket_bytes(int B) {
    if (B <= 0) return 0;
    return align_up(size_t(B + 1) * sizeof(int), 256);
}

// One launch (Gram + in-kernel tail) when the batch's MMD fits the last CTA's pipeline buffers; otherwise the tail runs
// as a chain of small kernels.  Either way the forward leaves the same outputs, `domgrad` included.
static bool use_fused_tail(int B, int n, int K) { return g_fused_tail && gram_tail_fits(B, n, K); }

static int forward_tail_kernels(const WhitenWorkspace& w, int nslots, int B, int64_t P, int n, int K, float margin, float eps,
                                float* losses, float* gram, float* rowstat, float* domgrad, cudaStream_t s) {
    cudaError_t e;
    {
        LaunchScope scope(kKernEpilogueFwd, s);       // per-sample reduce (B CTAs) + single-CTA MMD, chained programmatically
        e = launch_gram_reduce(w.partial, w.slot_count, nslots, B, P, n, K, margin, eps, gram, rowstat, w.vd, w.statd, s);
        if (e != cudaSuccess) return cuda_fail(e, "gram reduce launch");
        profile_count_kernel(kKernGramReduce);
        e = launch_whiten_epilogue_fwd(B, P, n, K, losses, w.scratch, s, w.vd, w.statd);
    }
    if (e != cudaSuccess) return cuda_fail(e, "forward epilogue launch");
    {
        LaunchScope scope(kKernMmdBwd, s);            // the backward's MMD seed: d L_dom / d v_b
        e = launch_mmd(w.vd, kVStride, nullptr, B, n, K, nullptr, domgrad, w.scratch, s);
    }
    if (e != cudaSuccess) return cuda_fail(e, "MMD gradient launch");
    return WTPSE_OK;
}

static int whitening_forward_impl(const float* z, float* relu_out, int B, int C, int64_t P, int n_per_domain, int n_domains,
                                  float margin, float eps, float* losses, float* gram, float* rowstat, float* domgrad,
                                  void* workspace, size_t workspace_bytes, wtpse_stream_t stream, bool channels_last = false) {
    if (int rc = check_common(z, B, C, P, n_per_domain, n_domains)) return rc;
    if (!losses || !gram || !rowstat || !domgrad || !workspace) return fail(WTPSE_ERR_INVALID, "null output/workspace pointer");
    const int sms = sm_count_cached();
    const WhitenWorkspace w = carve(workspace, B, P, sms);
    if (workspace_bytes < w.total) return fail(WTPSE_ERR_WORKSPACE, "workspace %zu < required %zu bytes", workspace_bytes, w.total);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (channels_last && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(relu_out)) & 15u))
        return fail(WTPSE_ERR_INVALID, "channels-last tensors must be 16-byte aligned");
    TailParams tp{};
    tp.ticket = w.ticket; tp.partial = w.partial; tp.B = B; tp.P = P; tp.n = n_per_domain; tp.K = n_domains;
    tp.margin = margin; tp.eps = eps; tp.gram = gram; tp.rowstat = rowstat; tp.vd = w.vd; tp.losses = losses; tp.domgrad = domgrad;
    tp.stamps = g_tail_stamps ? w.stamps : nullptr;
    const bool fused = use_fused_tail(B, n_per_domain, n_domains);
    const bool cl_tma = channels_last && cl_tma_ok(z, relu_out, nullptr, P);      // tensor-map TMA pipeline, else per-thread loads
    const GramPlan g = cl_tma ? plan_gram_cl_tma(B, P, sms) : channels_last ? plan_gram_cl(B, P, sms) : plan_gram(z, B, P, sms, relu_out);
    tp.nslots = g.nslots;
    const bool in_kernel = fused && g.tma && (!channels_last || gram_cl_tail_fits(B, n_per_domain, n_domains));
    cudaError_t e;
    {
        LaunchScope scope(kKernGram, s);
        if (cl_tma) ++g_cl_tma_launches;
        e = cl_tma         ? launch_gram_cl_tma(z, relu_out, w.partial, w.slot_count, B, P, g, s, in_kernel ? &tp : nullptr)
            : channels_last ? launch_gram_cl(z, relu_out, w.partial, w.slot_count, B, P, g, s)
                            : launch_gram(z, w.partial, w.slot_count, B, P, g, s, relu_out, in_kernel ? &tp : nullptr);
    }
    if (e != cudaSuccess) return cuda_fail(e, "gram launch");
    if (in_kernel) return WTPSE_OK;
    return forward_tail_kernels(w, g.nslots, B, P, n_per_domain, n_domains, margin, eps, losses, gram, rowstat, domgrad, s);
}

int wtpse_whitening_forward(const float* z, int B, int C, int64_t P, int n_per_domain, int n_domains, float margin,
                            float eps, float* losses, float* gram, float* rowstat, float* domgrad, void* workspace,
                            size_t workspace_bytes, wtpse_stream_t stream) {
    return whitening_forward_impl(z, nullptr, B, C, P, n_per_domain, n_domains, margin, eps, losses, gram, rowstat, domgrad,
                                  workspace, workspace_bytes, stream);
}

int wtpse_whitening_relu_forward(const float* z, float* relu_out, int B, int C, int64_t P, int n_per_domain, int n_domains,
                                 float margin, float eps, float* losses, float* gram, float* rowstat, float* domgrad,
                                 void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (!relu_out) return fail(WTPSE_ERR_INVALID, "null relu_out pointer");
    if (relu_out == z) return fail(WTPSE_ERR_INVALID, "relu_out must not alias z (the backward pass re-reads z)");
    return whitening_forward_impl(z, relu_out, B, C, P, n_per_domain, n_domains, margin, eps, losses, gram, rowstat, domgrad,
                                  workspace, workspace_bytes, stream);
}

static int whitening_backward_impl(const float* z, const float* grelu, const float* gram, const float* rowstat,
                                   const float* domgrad, const float* g_off, const float* g_diag, const float* g_dom, int B,
                                   int C, int64_t P, int n_per_domain, int n_domains, float* dz, wtpse_stream_t stream,
                                   bool channels_last = false) {
    if (int rc = check_common(z, B, C, P, n_per_domain, n_domains)) return rc;
    if (!gram || !rowstat || !domgrad || !dz) return fail(WTPSE_ERR_INVALID, "null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (channels_last && ((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(grelu) | reinterpret_cast<uintptr_t>(dz)) & 15u))
        return fail(WTPSE_ERR_INVALID, "channels-last tensors must be 16-byte aligned");
    // ONE launch: every apply kernel derives M_b itself from (gram, rowstat, domgrad) and the upstream scalars
    const SeedArgs seed{gram, rowstat, domgrad, g_off, g_diag, g_dom, B, n_per_domain, n_domains};
    const int sms = sm_count_cached();
    cudaError_t e;
    {
        LaunchScope scope(kKernApply, s);
        const bool cl_tma = channels_last && cl_tma_ok(z, grelu, dz, P);
        if (cl_tma) ++g_cl_tma_launches;
        e = !channels_last ? launch_apply(z, seed, dz, B, P, sms, s, grelu)
            : cl_tma       ? launch_apply_cl_tma(z, grelu, seed, dz, B, P, sms, s)
                           : launch_apply_cl(z, grelu, seed, dz, B, P, sms, s);
    }
    if (e != cudaSuccess) return cuda_fail(e, "apply launch");
    return WTPSE_OK;
}

int wtpse_whitening_backward(const float* z, const float* gram, const float* rowstat, const float* domgrad, const float* g_off,
                             const float* g_diag, const float* g_dom, int B, int C, int64_t P, int n_per_domain,
                             int n_domains, float* dz, wtpse_stream_t stream) {
    return whitening_backward_impl(z, nullptr, gram, rowstat, domgrad, g_off, g_diag, g_dom, B, C, P, n_per_domain, n_domains, dz,
                                   stream);
}

int wtpse_whitening_forward_cl(const float* z, float* relu_out, int B, int C, int64_t P, int n_per_domain, int n_domains,
                               float margin, float eps, float* losses, float* gram, float* rowstat, float* domgrad,
                               void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (relu_out && relu_out == z) return fail(WTPSE_ERR_INVALID, "relu_out must not alias z (the backward pass re-reads z)");
    return whitening_forward_impl(z, relu_out, B, C, P, n_per_domain, n_domains, margin, eps, losses, gram, rowstat, domgrad,
                                  workspace, workspace_bytes, stream, /*channels_last=*/true);
}

int wtpse_whitening_backward_cl(const float* z, const float* grad_relu, const float* gram, const float* rowstat,
                                const float* domgrad, const float* g_off, const float* g_diag, const float* g_dom, int B, int C,
                                int64_t P, int n_per_domain, int n_domains, float* dz, wtpse_stream_t stream) {
    return whitening_backward_impl(z, grad_relu, gram, rowstat, domgrad, g_off, g_diag, g_dom, B, C, P, n_per_domain, n_domains,
                                   dz, stream, /*channels_last=*/true);
}

int wtpse_whitening_relu_backward(const float* z, const float* grad_relu, const float* gram, const float* rowstat,
                                  const float* domgrad, const float* g_off, const float* g_diag, const float* g_dom, int B,
                                  int C, int64_t P, int n_per_domain, int n_domains, float* dz, wtpse_stream_t stream) {
    if (!grad_relu) return fail(WTPSE_ERR_INVALID, "null grad_relu pointer (use wtpse_whitening_backward)");
    return whitening_backward_impl(z, grad_relu, gram, rowstat, domgrad, g_off, g_diag, g_dom, B, C, P, n_per_domain, n_domains,
                                   dz, stream);
}

// ---- diagnostic switches (include/wtpse_b200_debug.h; not part of the product ABI) -----------------------------------------
namespace {
struct Knob { const char* name; int* value; int lo, hi; };
const Knob* knobs(int* count) {
    static const Knob table[] = {
        {"fused_tail", &g_fused_tail, 0, 1},
        {"tail_stamps", &g_tail_stamps, 0, 1},                    // phase clocks of the in-kernel tail -> last 128 workspace bytes                      // 0: forward tail as separate kernels for every shape
        {"apply_round_robin", &g_apply_round_robin, 0, 1},        // NCHW apply kernel: tiles dealt round-robin (1) or contiguous ranges (0)
        {"l2_hint", &g_l2_evict_first, 0, 1},                     // L2 evict-first policy on the TMA loads of z
        {"cl_tma", &g_cl_tma, 0, 1},
        {"cl_tma_launches", &g_cl_tma_launches, 0, 0},            // counter (setting it resets it to 0)                              // channels-last kernels: tensor-map TMA pipelines (1) or per-thread loads (0)
        {"wavelet_resident", &g_wavelet_resident, 0, 1},
        {"wavelet_tiles", &g_wavelet_tiles, 0, 1},
        {"wavelet_db2", &g_wavelet_db2, 0, 1},
        {"wavelet_db2_two", &g_wavelet_db2_two, 0, 1},
        {"wavelet_haar_passes", &g_wavelet_haar_passes, 0, 1},
        {"wavelet_haar_min_log2px", &g_wavelet_haar_min_log2px, 0, 40},
        {"wavelet_db2_deep", &g_wavelet_db2_deep, 0, 1},
        {"wavelet_db2_rf", &g_wavelet_db2_rf, 0, 64},
        {"wavelet_db2_ri", &g_wavelet_db2_ri, 0, 128},
        {"wavelet_db2_nw2", &g_wavelet_db2_nw2, 0, 12},
        {"wavelet_db2_rr", &g_wavelet_db2_rr, 0, 1},
        {"wavelet_peel_max", &g_wavelet_peel_max, 1, 16},
        {"wavelet_split", &g_wavelet_split, -1, 1},
        {"wavelet_cluster_max", &g_wavelet_cluster_max, 1, 8},
    };
    *count = int(sizeof(table) / sizeof(table[0]));
    return table;
}
}  // namespace

int wtpse_debug_set(const char* name, int value) {
    int n = 0;
    const Knob* t = knobs(&n);
    for (int i = 0; i < n; ++i)
        if (name && strcmp(name, t[i].name) == 0) {
            *t[i].value = value < t[i].lo ? t[i].lo : (value > t[i].hi ? t[i].hi : value);
            return WTPSE_OK;
        }
    return fail(WTPSE_ERR_INVALID, "unknown debug switch '%s'", name ? name : "(null)");
}

int wtpse_debug_get(const char* name, int* value) {
    int n = 0;
    const Knob* t = knobs(&n);
    for (int i = 0; i < n; ++i)
        if (name && value && strcmp(name, t[i].name) == 0) {
            *value = *t[i].value;
            return WTPSE_OK;
        }
    return fail(WTPSE_ERR_INVALID, "unknown debug switch '%s'", name ? name : "(null)");
}

size_t wtpse_mmd_workspace_bytes(int B) {
    if (B <= 0) return 0;
    return scratch_bytes_any_k(B);
}

static int check_mmd(const void* v, int B, int D, int n, int K) {
    if (!v) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (D != 120) return fail(WTPSE_ERR_INVALID, "MMD input must be B x 120 (upper triangle of a 16x16 Gram), got D=%d", D);
    if (B <= 0 || n < 0 || K < 0 || (long long)K * K > (long long)B * B + 64)
        return fail(WTPSE_ERR_INVALID, "bad MMD configuration (B=%d, n=%d, K=%d)", B, n, K);
    return WTPSE_OK;
}

int wtpse_mmd_forward(const float* v, int B, int D, int n_per_domain, int n_domains, float* loss, void* workspace,
                      size_t workspace_bytes, wtpse_stream_t stream) {
    if (int rc = check_mmd(v, B, D, n_per_domain, n_domains)) return rc;
    if (!loss || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (workspace_bytes < wtpse_mmd_workspace_bytes(B)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernMmdFwd, s); e = launch_mmd(v, kOff, nullptr, B, n_per_domain, n_domains, loss, nullptr, workspace, s); }
    if (e != cudaSuccess) return cuda_fail(e, "mmd forward launch");
    return WTPSE_OK;
}

int wtpse_mmd_backward(const float* v, const float* gout, int B, int D, int n_per_domain, int n_domains, float* dv,
                       void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (int rc = check_mmd(v, B, D, n_per_domain, n_domains)) return rc;
    if (!dv || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (workspace_bytes < wtpse_mmd_workspace_bytes(B)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernMmdBwd, s); e = launch_mmd(v, kOff, gout, B, n_per_domain, n_domains, nullptr, dv, workspace, s); }
    if (e != cudaSuccess) return cuda_fail(e, "mmd backward launch");
    return WTPSE_OK;
}

size_t wtpse_mse_workspace_bytes(int64_t N) {
    if (N <= 0) return 0;
    return align_up(mse_partial_doubles(N, sm_count_cached()) * sizeof(double), 256);
}

int wtpse_mse_forward(const float* a, const float* b, int64_t N, float* loss, void* workspace, size_t workspace_bytes,
                      wtpse_stream_t stream) {
    if (!a || !b || !loss || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (N <= 0) return fail(WTPSE_ERR_INVALID, "empty input (N=%lld)", (long long)N);
    if (workspace_bytes < wtpse_mse_workspace_bytes(N)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernMseFwd, s); e = launch_mse_fwd(a, b, N, loss, static_cast<double*>(workspace), sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "mse forward launch");
    return WTPSE_OK;
}

int wtpse_mse_backward(const float* a, const float* b, const float* gout, int64_t N, float* da, float* db,
                       wtpse_stream_t stream) {
    if (!a || !b) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (N <= 0) return fail(WTPSE_ERR_INVALID, "empty input (N=%lld)", (long long)N);
    if (!da && !db) return WTPSE_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernMseBwd, s); e = launch_mse_bwd(a, b, gout, N, da, db, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "mse backward launch");
    return WTPSE_OK;
}

// ---------------------------------------------------------------------------------------------
// element-wise kernels either side of the loss
// ---------------------------------------------------------------------------------------------
int wtpse_prepare_batch(const unsigned char* img_hwc, const unsigned char* raw_od, const unsigned char* raw_oc, int B, int H,
                        int W, float* image_chw, float* label_od, float* label_oc, wtpse_stream_t stream) {
    if (!raw_od || !label_od || !label_oc) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (img_hwc && !image_chw) return fail(WTPSE_ERR_INVALID, "image output missing");
    if (B <= 0 || H <= 0 || W <= 0) return fail(WTPSE_ERR_INVALID, "bad shape %dx%dx%d", B, H, W);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernLabels, s); e = launch_prepare_batch(img_hwc, raw_od, raw_oc, B, (long long)H * W, image_chw, label_od, label_oc, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "prepare_batch launch");
    return WTPSE_OK;
}

size_t wtpse_od_roi_workspace_bytes(void) { return od_roi_workspace_bytes(); }

int wtpse_od_roi(const float* logits, const float* target_oc, float* image, float* od_pred, float* image_roi, int B, int C,
                 int64_t HW, float threshold, float* sums, void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (!logits || !image || !od_pred || !image_roi || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (B <= 0 || C <= 0 || HW <= 0) return fail(WTPSE_ERR_INVALID, "bad shape");
    if (workspace_bytes < od_roi_workspace_bytes()) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernLabels, s); e = launch_od_roi(logits, target_oc, image, od_pred, image_roi, B, C, HW, threshold, sums, workspace, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "od_roi launch");
    return WTPSE_OK;
}

size_t wtpse_attention_fuse_workspace_bytes(int B, int64_t P) {
    if (B <= 0 || P <= 0) return 0;
    return align_up(fuse_bwd_partial_doubles(B, P, sm_count_cached()) * sizeof(double), 256);
}

int wtpse_attention_fuse_forward(const float* emb, const float* z_post, const float* weight_bias, float coef, int B, int Ce,
                                 int64_t P, float threshold, float* fuse, float* att_mask, float* att,
                                 wtpse_stream_t stream) {
    if (!emb || !z_post || !weight_bias || !fuse || !att_mask || !att) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (B <= 0 || Ce <= 0 || P <= 0) return fail(WTPSE_ERR_INVALID, "bad shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernFuse, s); e = launch_fuse_fwd(emb, z_post, weight_bias, coef, B, Ce, P, threshold, fuse, att_mask, att, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "attention_fuse forward launch");
    return WTPSE_OK;
}

int wtpse_attention_fuse_backward(const float* grad_fuse, const float* emb, const float* z_post, const float* att,
                                  const float* weight_bias, float coef, int B, int Ce, int64_t P, float* d_emb, float* d_z_post,
                                  float* d_weight_bias, void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (!grad_fuse || !emb || !z_post || !att || !weight_bias || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (B <= 0 || Ce <= 0 || P <= 0) return fail(WTPSE_ERR_INVALID, "bad shape");
    if (workspace_bytes < wtpse_attention_fuse_workspace_bytes(B, P)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernFuse, s); e = launch_fuse_bwd(grad_fuse, emb, z_post, att, weight_bias, coef, B, Ce, P, d_emb, d_z_post, d_weight_bias, static_cast<double*>(workspace), sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "attention_fuse backward launch");
    return WTPSE_OK;
}

int wtpse_upsample2x_nhwc(const float* in, float* out, int64_t N, int H, int W, int C, int adjoint, wtpse_stream_t stream) {
    if (!in || !out) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (N <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 4) != 0) return fail(WTPSE_ERR_INVALID, "need N, H, W >= 1 and C a positive multiple of 4 (got C=%d)", C);
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) return fail(WTPSE_ERR_INVALID, "pointers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernUpsample, s); e = launch_upsample2x_nhwc(in, out, N, H, W, C, adjoint != 0, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "upsample launch");
    return WTPSE_OK;
}

int wtpse_bias_act_nhwc(float* y, const float* bias, int64_t npix, int C, int relu, wtpse_stream_t stream) {
    if (!y || !bias) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (npix <= 0 || C <= 0 || (C % 4) != 0) return fail(WTPSE_ERR_INVALID, "need npix >= 1 and C a positive multiple of 4 (got C=%d)", C);
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15u) return fail(WTPSE_ERR_INVALID, "pointers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernUpsample, s); e = launch_bias_act_nhwc(y, bias, npix, C, relu != 0, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "bias_act launch");
    return WTPSE_OK;
}

size_t wtpse_channel_sum_workspace_bytes(int64_t npix, int C) {
    if (npix <= 0 || !channel_sum_supported(C)) return 0;
    return align_up(size_t(channel_sum_blocks(npix, C, sm_count_cached())) * C * sizeof(float), 256);
}

int wtpse_channel_sum_nhwc(const float* g, int64_t npix, int C, float* out, void* workspace, size_t workspace_bytes,
                           wtpse_stream_t stream) {
    if (!g || !out || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (npix <= 0 || !channel_sum_supported(C)) return fail(WTPSE_ERR_INVALID, "need npix >= 1 and C in {4, 8, 16, ..., 1024} (got C=%d)", C);
    if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(workspace)) & 15u) return fail(WTPSE_ERR_INVALID, "pointers must be 16-byte aligned");
    if (workspace_bytes < wtpse_channel_sum_workspace_bytes(npix, C)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernUpsample, s); e = launch_channel_sum_nhwc(g, npix, C, out, static_cast<float*>(workspace), sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "channel_sum launch");
    return WTPSE_OK;
}

int wtpse_relu_backward_channel_sum_nhwc(const float* g, const float* out, int64_t npix, int C, float* gx, float* bias_grad,
                                         void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (!g || !out || !gx || !bias_grad || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (npix <= 0 || !channel_sum_supported(C)) return fail(WTPSE_ERR_INVALID, "need npix >= 1 and C a power of two in [4, 1024] (got C=%d)", C);
    if ((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(gx) | reinterpret_cast<uintptr_t>(workspace)) & 15u)
        return fail(WTPSE_ERR_INVALID, "pointers must be 16-byte aligned");
    if (workspace_bytes < wtpse_channel_sum_workspace_bytes(npix, C)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernUpsample, s); e = launch_relu_bwd_channel_sum_nhwc(g, out, gx, npix, C, bias_grad, static_cast<float*>(workspace), sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "relu_backward_channel_sum launch");
    return WTPSE_OK;
}

int wtpse_maxpool2_nhwc(const float* in, float* out, unsigned char* argmax, int64_t N, int Ho, int Wo, int C, int backward,
                        wtpse_stream_t stream) {
    if (!in || !out || !argmax) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (N <= 0 || Ho <= 0 || Wo <= 0 || C <= 0 || (C % 4) != 0) return fail(WTPSE_ERR_INVALID, "need N, Ho, Wo >= 1 and C a positive multiple of 4 (got C=%d)", C);
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) || (reinterpret_cast<uintptr_t>(argmax) & 3u))
        return fail(WTPSE_ERR_INVALID, "float pointers must be 16-byte aligned, argmax 4-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernUpsample, s); e = launch_maxpool2_nhwc(in, out, argmax, N, Ho, Wo, C, backward != 0, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "maxpool launch");
    return WTPSE_OK;
}

size_t wtpse_batchnorm_workspace_bytes(int64_t npix, int C) {
    if (npix <= 0 || !batchnorm_supported(C)) return 0;
    return align_up(batchnorm_workspace_floats(npix, C, sm_count_cached()) * sizeof(float), 256);
}

static int check_bn(const void* x, int64_t npix, int C, const void* gamma, const void* beta, const void* ws, size_t ws_bytes) {
    if (!x || !gamma || !beta || !ws) return fail(WTPSE_ERR_INVALID, "null pointer (affine BatchNorm only)");
    if (npix <= 0 || !batchnorm_supported(C)) return fail(WTPSE_ERR_INVALID, "need npix >= 1 and C a power of two in [4, 1024] (got C=%d)", C);
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(ws)) & 15u)
        return fail(WTPSE_ERR_INVALID, "pointers must be 16-byte aligned");
    if (ws_bytes < wtpse_batchnorm_workspace_bytes(npix, C)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    return WTPSE_OK;
}

int wtpse_batchnorm_relu_forward(const float* x, int64_t npix, int C, const float* gamma, const float* beta, const float* mean_shift,
                                 float eps, float momentum, int relu, float* running_mean, float* running_var, float* y,
                                 float* save_stats, void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (int rc = check_bn(x, npix, C, gamma, beta, workspace, workspace_bytes)) return rc;
    if (!y || !save_stats) return fail(WTPSE_ERR_INVALID, "null output pointer");
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(save_stats)) & 15u) return fail(WTPSE_ERR_INVALID, "pointers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    {
        LaunchScope scope(kKernUpsample, s);
        e = launch_batchnorm_fwd(x, npix, C, gamma, beta, mean_shift, eps, momentum, relu != 0, running_mean, running_var, y, save_stats,
                                 save_stats + C, save_stats + 2 * C, static_cast<float*>(workspace), sm_count_cached(), s);
    }
    if (e != cudaSuccess) return cuda_fail(e, "batchnorm forward launch");
    return WTPSE_OK;
}

int wtpse_batchnorm_relu_backward(const float* x, const float* dy, int64_t npix, int C, const float* gamma, const float* beta,
                                  const float* save_stats, int relu, float* dx, float* dgamma, float* dbeta, void* workspace,
                                  size_t workspace_bytes, wtpse_stream_t stream) {
    if (int rc = check_bn(x, npix, C, gamma, beta, workspace, workspace_bytes)) return rc;
    if (!dy || !save_stats || !dx) return fail(WTPSE_ERR_INVALID, "null pointer");
    if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(save_stats)) & 15u) return fail(WTPSE_ERR_INVALID, "pointers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    {
        LaunchScope scope(kKernUpsample, s);
        e = launch_batchnorm_bwd(x, dy, npix, C, gamma, beta, save_stats, save_stats + C, save_stats + 2 * C, relu != 0, dx, dgamma, dbeta,
                                 static_cast<float*>(workspace), sm_count_cached(), s);
    }
    if (e != cudaSuccess) return cuda_fail(e, "batchnorm backward launch");
    return WTPSE_OK;
}

// ---------------------------------------------------------------------------------------------
// Track W: wavelet transform + L1 detail loss (parity unpinned, see include/wtpse_b200.h)
// ---------------------------------------------------------------------------------------------
static int check_wavelet(const void* p, int nmaps, int H, int W, int wavelet, int J) {
    if (!p) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (wavelet != 0 && wavelet != 1) return fail(WTPSE_ERR_INVALID, "wavelet must be 0 (haar) or 1 (db2)");
    if (nmaps <= 0 || H <= 0 || W <= 0 || J < 1 || J > 16) return fail(WTPSE_ERR_INVALID, "bad shape / level count");
    if ((H % (1 << J)) || (W % (1 << J))) return fail(WTPSE_ERR_INVALID, "H=%d and W=%d must be divisible by 2^J=%d", H, W, 1 << J);
    if (wavelet == 1 && ((H >> (J - 1)) < 4 || (W >> (J - 1)) < 4)) return fail(WTPSE_ERR_INVALID, "db2 needs at least 4 samples at the coarsest level");
    return WTPSE_OK;
}

size_t wtpse_wavelet_workspace_bytes(int nmaps, int H, int W, int J) {
    if (nmaps <= 0 || H <= 0 || W <= 0 || J < 1) return 0;
    // + one double per CTA of the resident kernel (at most one CTA per SM)
    return align_up(wavelet_scratch_floats(nmaps, H, W) * sizeof(float), 256) +
           align_up((wavelet_partial_doubles(nmaps, H, W, J) + wavelet_stream_partials(nmaps, H, W) + 1024) * sizeof(double), 256);
}

static float* wavelet_scratch(void* ws) { return static_cast<float*>(ws); }
static double* wavelet_partials(void* ws, int nmaps, int H, int W) {
    return reinterpret_cast<double*>(static_cast<char*>(ws) + align_up(wavelet_scratch_floats(nmaps, H, W) * sizeof(float), 256));
}

int wtpse_dwt2d_forward(const float* x, int nmaps, int H, int W, int wavelet, int J, float* coef, void* workspace,
                        size_t workspace_bytes, wtpse_stream_t stream) {
    if (int rc = check_wavelet(x, nmaps, H, W, wavelet, J)) return rc;
    if (!coef || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (workspace_bytes < wtpse_wavelet_workspace_bytes(nmaps, H, W, J)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernWaveletFwd, s); e = launch_dwt(x, nmaps, H, W, wavelet ? 4 : 2, J, coef, wavelet_scratch(workspace), nullptr, nullptr, nullptr, s); }
    if (e != cudaSuccess) return cuda_fail(e, "dwt launch");
    return WTPSE_OK;
}

int wtpse_dwt2d_inverse(const float* coef, int nmaps, int H, int W, int wavelet, int J, float* x, const float* scale,
                        void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (int rc = check_wavelet(coef, nmaps, H, W, wavelet, J)) return rc;
    if (!x || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (workspace_bytes < wtpse_wavelet_workspace_bytes(nmaps, H, W, J)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernWaveletBwd, s); e = launch_idwt(coef, nmaps, H, W, wavelet ? 4 : 2, J, x, wavelet_scratch(workspace), scale, s); }
    if (e != cudaSuccess) return cuda_fail(e, "idwt launch");
    return WTPSE_OK;
}

int wtpse_wavelet_loss_forward(const float* x, int nmaps, int H, int W, int wavelet, int J, const float* level_weights,
                               float* loss, float* grad_coef, void* workspace, size_t workspace_bytes, wtpse_stream_t stream) {
    if (int rc = check_wavelet(x, nmaps, H, W, wavelet, J)) return rc;
    if (!loss || !grad_coef || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (workspace_bytes < wtpse_wavelet_workspace_bytes(nmaps, H, W, J)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    float w[16];
    for (int j = 0; j < J; ++j) w[j] = level_weights ? level_weights[j] : 1.0f;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernWaveletFwd, s); e = launch_dwt(x, nmaps, H, W, wavelet ? 4 : 2, J, grad_coef, wavelet_scratch(workspace), w, loss, wavelet_partials(workspace, nmaps, H, W), s); }
    if (e != cudaSuccess) return cuda_fail(e, "wavelet loss launch");
    return WTPSE_OK;
}

int wtpse_wavelet_resident_cluster(int H, int W, int wavelet, int J) {
    if (!g_wavelet_resident || (wavelet != 0 && wavelet != 1) || H <= 0 || W <= 0 || J < 1 || J > 16) return 0;
    if ((H % (1 << J)) || (W % (1 << J))) return 0;
    if (wavelet == 1 && ((H >> (J - 1)) < 4 || (W >> (J - 1)) < 4)) return 0;
    const int taps = wavelet ? 4 : 2;
    switch (wavelet_fused_plan(H, W, taps, J)) {
        case 1: return taps == 2 ? 1 : wavelet_resident_cluster(H, W, taps, J);       // Haar: independent bands, no cluster
        case 2: { int cs = 0; wavelet_stream_levels(H, W, taps, J, 0, &cs); return cs; }
        default: return 0;
    }
}

int wtpse_wavelet_loss_resident(const float* x, int nmaps, int H, int W, int wavelet, int J, const float* level_weights,
                                const float* upstream, float* loss, float* grad_x, void* workspace, size_t workspace_bytes,
                                wtpse_stream_t stream) {
    if (int rc = check_wavelet(x, nmaps, H, W, wavelet, J)) return rc;
    if (!loss || !workspace) return fail(WTPSE_ERR_INVALID, "null pointer");
    if (wtpse_wavelet_resident_cluster(H, W, wavelet, J) == 0)
        return fail(WTPSE_ERR_INVALID, "a %d x %d map (J=%d) does not fit the cluster-resident path; use wtpse_wavelet_loss_forward", H, W, J);
    if (workspace_bytes < wtpse_wavelet_workspace_bytes(nmaps, H, W, J)) return fail(WTPSE_ERR_WORKSPACE, "workspace too small");
    float w[16];
    for (int j = 0; j < J; ++j) w[j] = level_weights ? level_weights[j] : 1.0f;
    const int taps = wavelet ? 4 : 2;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    {
        LaunchScope scope(kKernWaveletFwd, s);
        if (wavelet_fused_plan(H, W, taps, J, nmaps) == 2)
            e = launch_wavelet_loss_split(x, nmaps, H, W, taps, J, w, upstream, loss, grad_x, wavelet_scratch(workspace),
                                          wavelet_partials(workspace, nmaps, H, W), sm_count_cached(), s);
        else
            e = launch_wavelet_resident(x, nmaps, H, W, taps, J, w, upstream, loss, grad_x, wavelet_partials(workspace, nmaps, H, W), s);
    }
    if (e != cudaSuccess) return cuda_fail(e, "fused wavelet loss launch");
    return WTPSE_OK;
}

int wtpse_scale_unless_one(float* data, int64_t n, const float* scale, wtpse_stream_t stream) {
    if (n < 0 || (n > 0 && (!data || !scale))) return fail(WTPSE_ERR_INVALID, "null pointer / negative size");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e;
    { LaunchScope scope(kKernWaveletBwd, s); e = launch_scale_unless_one(data, n, scale, sm_count_cached(), s); }
    if (e != cudaSuccess) return cuda_fail(e, "scale launch");
    return WTPSE_OK;
}

// ---------------------------------------------------------------------------------------------
// host-buffer plan
// ---------------------------------------------------------------------------------------------
// Two slots, each with its own device buffers and stream: consecutive submissions alternate slots, so the D2H of
// step i (slot A) and the H2D of step i+1 (slot B) share the full-duplex PCIe link instead of queueing behind
// each other -- the path is PCIe-bound (1.07 GB per step against 0.29 ms of kernels).
struct HostSlot {
    float *z, *dz, *gram, *rowstat, *domgrad, *losses, *gvec;
    void* ws;
    float* pinned;            // [8] page-locked: losses (4) + upstream gradients (4); copies to/from it are truly asynchronous
    float* user_losses;       // where the step's losses go once it has left the device (filled by drain_slot)
    cudaStream_t stream;
    cudaEvent_t done;
    bool pending;
};

struct wtpse_host_plan {
    int B;
    long long P;
    size_t ws_bytes;
    HostSlot slot[2];
    unsigned long long submitted;
};

static void free_slot(HostSlot& s) {
    cudaFree(s.z); cudaFree(s.dz); cudaFree(s.gram); cudaFree(s.rowstat); cudaFree(s.domgrad); cudaFree(s.losses); cudaFree(s.gvec); cudaFree(s.ws);
    if (s.pinned) cudaFreeHost(s.pinned);
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
}

// Wait until the slot's step has left the device and hand its losses to the caller's buffer.
static int drain_slot(HostSlot& s) {
    if (!s.pending) return WTPSE_OK;
    cudaError_t e = cudaEventSynchronize(s.done);
    s.pending = false;
    if (e != cudaSuccess) return cuda_fail(e, "slot synchronize");
    if (s.user_losses) memcpy(s.user_losses, s.pinned, 4 * sizeof(float));
    s.user_losses = nullptr;
    return WTPSE_OK;
}

void wtpse_host_plan_destroy(wtpse_host_plan* p) {
    if (!p) return;
    for (int i = 0; i < 2; ++i) {
        if (p->slot[i].pending && p->slot[i].done) cudaEventSynchronize(p->slot[i].done);
        free_slot(p->slot[i]);
    }
    delete p;
}

int wtpse_host_plan_create(int B, int64_t P, wtpse_host_plan** out) {
    if (!out) return fail(WTPSE_ERR_INVALID, "null plan pointer");
    if (B <= 0 || P <= 1) return fail(WTPSE_ERR_INVALID, "need B >= 1 and P >= 2");
    wtpse_host_plan* p = new wtpse_host_plan();
    memset(p, 0, sizeof(*p));
    p->B = B; p->P = P;
    const size_t nz = size_t(B) * WTPSE_CHANNELS * size_t(P) * sizeof(float);
    p->ws_bytes = wtpse_whitening_workspace_bytes(B, P);
    for (int i = 0; i < 2; ++i) {
        HostSlot& s = p->slot[i];
        cudaError_t e = cudaSuccess;
        if ((e = cudaMalloc(&s.z, nz)) != cudaSuccess || (e = cudaMalloc(&s.dz, nz)) != cudaSuccess ||
            (e = cudaMalloc(&s.gram, size_t(B) * 256 * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&s.rowstat, size_t(B) * 2 * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&s.domgrad, size_t(B) * 120 * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&s.losses, 4 * sizeof(float))) != cudaSuccess ||
            (e = cudaMalloc(&s.gvec, 4 * sizeof(float))) != cudaSuccess || (e = cudaMalloc(&s.ws, p->ws_bytes)) != cudaSuccess ||
            (e = cudaMemset(s.ws, 0, wtpse_whitening_ticket_bytes(B))) != cudaSuccess ||        // the ticket contract: zero once
            (e = cudaMallocHost(&s.pinned, 8 * sizeof(float))) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) {
            wtpse_host_plan_destroy(p);
            return cuda_fail(e, "host plan allocation");
        }
    }
    *out = p;
    return WTPSE_OK;
}

// Asynchronous: returns once the step is enqueued (all staging goes through the slot's page-locked buffer, so no copy
// blocks the host).  `losses_host` and `dz_host` are complete after wtpse_host_plan_wait() -- or after the second-next
// submit, which reuses the slot -- and must stay valid until then.
int wtpse_host_plan_submit(wtpse_host_plan* p, const float* z_host, int n_per_domain, int n_domains, float margin, float eps,
                           const float grad_w[3], float losses_host[4], float* dz_host) {
    if (!p || !z_host || !losses_host) return fail(WTPSE_ERR_INVALID, "null pointer");
    HostSlot& s = p->slot[p->submitted & 1];
    if (int rc = drain_slot(s)) return rc;             // the slot's previous step must have left the device
    const size_t nz = size_t(p->B) * WTPSE_CHANNELS * size_t(p->P) * sizeof(float);
    int rc = WTPSE_OK;
    cudaError_t e = cudaSuccess;
    bool enqueued = false;
    do {
        if ((e = cudaMemcpyAsync(s.z, z_host, nz, cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) { rc = cuda_fail(e, "H2D copy"); break; }
        enqueued = true;
        if ((rc = wtpse_whitening_forward(s.z, p->B, WTPSE_CHANNELS, p->P, n_per_domain, n_domains, margin, eps, s.losses, s.gram,
                                          s.rowstat, s.domgrad, s.ws, p->ws_bytes, s.stream))) break;
        if ((e = cudaMemcpyAsync(s.pinned, s.losses, 4 * sizeof(float), cudaMemcpyDeviceToHost, s.stream)) != cudaSuccess) { rc = cuda_fail(e, "D2H losses"); break; }
        if (dz_host) {
            s.pinned[4] = grad_w ? grad_w[0] : 1.f; s.pinned[5] = grad_w ? grad_w[1] : 1.f; s.pinned[6] = grad_w ? grad_w[2] : 1.f; s.pinned[7] = 0.f;
            if ((e = cudaMemcpyAsync(s.gvec, s.pinned + 4, 4 * sizeof(float), cudaMemcpyHostToDevice, s.stream)) != cudaSuccess) { rc = cuda_fail(e, "H2D grads"); break; }
            if ((rc = wtpse_whitening_backward(s.z, s.gram, s.rowstat, s.domgrad, s.gvec, s.gvec + 1, s.gvec + 2, p->B, WTPSE_CHANNELS,
                                               p->P, n_per_domain, n_domains, s.dz, s.stream))) break;
            if ((e = cudaMemcpyAsync(dz_host, s.dz, nz, cudaMemcpyDeviceToHost, s.stream)) != cudaSuccess) { rc = cuda_fail(e, "D2H dz"); break; }
        }
    } while (false);
    if (rc != WTPSE_OK) {
        // work may already be in flight on the slot's stream: let it finish before anyone reuses the slot's buffers
        if (enqueued) cudaStreamSynchronize(s.stream);
        return rc;
    }
    if ((e = cudaEventRecord(s.done, s.stream)) != cudaSuccess) {
        cudaStreamSynchronize(s.stream);
        return cuda_fail(e, "event record");
    }
    s.user_losses = losses_host;
    s.pending = true;
    ++p->submitted;
    return WTPSE_OK;
}

int wtpse_host_plan_wait(wtpse_host_plan* p) {
    if (!p) return fail(WTPSE_ERR_INVALID, "null pointer");
    int rc = WTPSE_OK;
    for (int i = 0; i < 2; ++i)
        if (int r = drain_slot(p->slot[i])) rc = r;
    return rc;
}

int wtpse_host_plan_run(wtpse_host_plan* p, const float* z_host, int n_per_domain, int n_domains, float margin, float eps,
                        const float grad_w[3], float losses_host[4], float* dz_host) {
    if (int rc = wtpse_host_plan_submit(p, z_host, n_per_domain, n_domains, margin, eps, grad_w, losses_host, dz_host)) return rc;
    return wtpse_host_plan_wait(p);
}

}  // extern "C"

/*
 * wtpse_b200.h -- C ABI of the B200-native shape-regularization hot path of WT-PSE.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch / C++ types) and
 * returns an int status (WTPSE_OK == 0).  Nothing throws across the boundary;
 * wtpse_last_error() returns the message of the last failing call on the calling thread.
 * Device-pointer entry points are asynchronous on `stream` (a cudaStream_t passed as void*),
 * never synchronise the host, never copy host<->device and never allocate.  The *_host entry
 * points take HOST buffers and do the copies themselves (the plugin-facing path bench.py times
 * as "e2e").
 *
 * The reference (tonyckc/WT-PSE-code) has no FFI: its boundary for this path is the Python
 * method surface that Trainer.py calls.  Each function below names the reference statement(s)
 * it replaces; INTEGRATION.md shows the ctypes binding and the method rebinding a maintainer
 * would add to the reference.
 *
 * Layouts: all tensors fp32, NCHW contiguous, C == 16 (self.dim, algorithms.py:1157).
 *   z       [B][16][P]      P = H*W
 *   gram    [B][16][16]     f_cor of algorithms.py:1283 (symmetric, eps on the diagonal)
 *   rowstat [B][2]          off_b (algorithms.py:1289) and diag_b (algorithms.py:1297), margin subtracted
 *   domgrad [B][120]        d L_dom / d v_b over the upper-triangle entries in torch.triu_indices(16,16,1) order
 *                           (algorithms.py:1305); rows of samples outside the MMD are not written and not read
 *   losses  [4]             L_off, L_diag, L_dom, L_off + L_diag
 * gram, rowstat and domgrad are what the forward SAVES for the backward.
 *
 * Diagnostic switches and the launch accounting bench.py reads live in wtpse_b200_debug.h, not here.
 */
#ifndef WTPSE_B200_H
#define WTPSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WTPSE_OK              0
#define WTPSE_ERR_INVALID     1   /* bad argument (shape, null pointer, C != 16, ...) */
#define WTPSE_ERR_CUDA        2   /* a CUDA runtime call or launch failed            */
#define WTPSE_ERR_WORKSPACE   3   /* workspace too small                             */

#define WTPSE_CHANNELS        16
#define WTPSE_ABI_VERSION     2

typedef void* wtpse_stream_t;     /* cudaStream_t */

int         wtpse_abi_version(void);
const char* wtpse_last_error(void);
/* Number of SMs the persistent kernels size their grids for on the current device. */
int         wtpse_sm_count(void);

/* ---- whitening (Gram) loss: algorithms.py:1277-1309, shape_networks.py:561-594 ------------- */

/* Bytes of scratch the forward needs for a [B][16][P] input.  The backward needs none. */
size_t wtpse_whitening_workspace_bytes(int B, int64_t P);
/*
 * The first wtpse_whitening_ticket_bytes(B) bytes of the forward's workspace are arrival tickets of the in-kernel
 * tail (the CTA that arrives last finishes the sample / the batch: no second launch, no float atomics).  CONTRACT:
 * they must be ZERO when a forward is enqueued, and every forward leaves them zero again when it completes -- so a
 * workspace that is zeroed once (cudaMemset at allocation) and then reused by stream-ordered calls needs nothing
 * more.  A caller that hands over fresh, uninitialised scratch every time clears that prefix with cudaMemsetAsync on
 * the same stream first.  (The rest of the workspace needs no initialisation.)
 */
size_t wtpse_whitening_ticket_bytes(int B);

/*
 * Forward.  Replaces the body of WT_PSE.compute_whitening_loss (algorithms.py:1280-1307) and of
 * ShapeVariationalDist_x.compute_whitening_loss (shape_networks.py:564-592) including
 * compute_MMD.forward (algorithms.py:102-121):
 *   gram    = bmm(z, z^T)/(P-1) + eps*I
 *   L_off   = mean_b clamp((sum_{i<j}|gram_ij| - margin)/120, 0)
 *   L_diag  = mean_b clamp((sum_i |gram_ii - 1| - margin)/16, 0)
 *   L_dom   = gaussian-kernel MMD over the n_domains chunks [k*n : (k+1)*n] of the 120-d
 *             upper-triangle vectors (python slice semantics: chunks are truncated at B; an empty
 *             chunk yields NaN exactly like the reference's mean over an empty tensor);
 *             0 when n_domains <= 1.
 * NaN/Inf in z propagate to the losses (the caller's NaN guard is Trainer.py:799-800).
 * ONE kernel launch for 16-byte-aligned z with P % 4 == 0 (anything else, and MMD batches beyond ~100 samples, take a
 * chain of small kernels with the same results).
 */
int wtpse_whitening_forward(const float* z, int B, int C, int64_t P,
                            int n_per_domain, int n_domains, float margin, float eps,
                            float* losses, float* gram, float* rowstat, float* domgrad,
                            void* workspace, size_t workspace_bytes, wtpse_stream_t stream);

/*
 * Backward (what autograd derives from the statements above):
 *   dz_b = (S_b + S_b^T) z_b / (P-1)
 * g_off / g_diag / g_dom are DEVICE pointers to the upstream gradients of L_off / L_diag / L_dom
 * (NULL == 0), so no host synchronisation is needed between forward and backward.  ONE kernel launch, no workspace:
 * the per-sample matrix is combined inside the kernel from gram, rowstat, domgrad and the three scalars.
 */
int wtpse_whitening_backward(const float* z, const float* gram, const float* rowstat, const float* domgrad,
                             const float* g_off, const float* g_diag, const float* g_dom,
                             int B, int C, int64_t P, int n_per_domain, int n_domains,
                             float* dz, wtpse_stream_t stream);

/*
 * Producer fusion with the DeepWT tail.  In DeepWT.forward (algorithms.py:1099-1113) each embedding
 * that feeds the loss is followed immediately by `F.relu(z_instance)` (:1105, :1112).  These two entry points do both
 * in the passes the loss makes anyway:
 *   forward : everything wtpse_whitening_forward does, and relu_out = relu(z)  (same shape as z; must not alias it)
 *   backward: dz = (S_b + S_b^T) z_b / (P-1) + [z > 0] * grad_relu              (grad_relu: upstream gradient of relu_out)
 * i.e. the ReLU forward, its backward (ATen threshold_backward: passes where z > 0 or z is NaN) and the sum autograd
 * forms of the two gradients reaching z.  HBM traffic for embedding + activation: 128 + 192 B/pixel instead of
 * (64 + 128) + (128 + 192 + 192).  relu_out / grad_relu need the same 16-byte alignment as z for the TMA path.
 */
int wtpse_whitening_relu_forward(const float* z, float* relu_out, int B, int C, int64_t P,
                                 int n_per_domain, int n_domains, float margin, float eps,
                                 float* losses, float* gram, float* rowstat, float* domgrad,
                                 void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
int wtpse_whitening_relu_backward(const float* z, const float* grad_relu, const float* gram, const float* rowstat,
                                  const float* domgrad, const float* g_off, const float* g_diag, const float* g_dom,
                                  int B, int C, int64_t P, int n_per_domain, int n_domains, float* dz,
                                  wtpse_stream_t stream);

/*
 * Decoder up-sampling of the U-Net stages: F.interpolate(x, scale_factor=2, mode='bilinear', align_corners=False)
 * (algorithms.py:947, ConvU.forward) on a CHANNELS-LAST tensor, and its adjoint (the backward pass).
 *   adjoint == 0: in [N][H][W][C] -> out [N][2H][2W][C]
 *   adjoint != 0: in [N][2H][2W][C] (gradient of the output) -> out [N][H][W][C] (gradient of the input)
 * H and W are the LOW-resolution sizes in both directions; C % 4 == 0; 16-byte aligned pointers.  Deterministic (the
 * adjoint gathers, no atomics).
 */
int wtpse_upsample2x_nhwc(const float* in, float* out, int64_t N, int H, int W, int C, int adjoint, wtpse_stream_t stream);

/*
 * Convolution bias (+ ReLU) pass of the backbone blocks that have no BatchNorm (DoubleConvWT algorithms.py:416-428, the
 * 1x1 heads :1019-1030, :1176-1183): y[p][c] = act(y[p][c] + bias[c]) IN PLACE on a channels-last activation of npix
 * pixels x C channels; relu != 0 applies max(., 0) with NaN propagating (ATen's clamp_min).  C % 4 == 0, 16-byte aligned.
 */
int wtpse_bias_act_nhwc(float* y, const float* bias, int64_t npix, int C, int relu, wtpse_stream_t stream);

/*
 * Bias gradient of a backbone convolution: out[c] = sum over the npix pixels of g[p][c] (`grad.sum((0, 2, 3))`, what
 * aten::convolution_backward reduces for Conv2d.bias) on a channels-last gradient.  C a power of two in [4, 1024];
 * two deterministic stages, no atomics.
 */
size_t wtpse_channel_sum_workspace_bytes(int64_t npix, int C);
/* Backward of wtpse_bias_act_nhwc(relu = 1) in one pass: gx = [out > 0] * g (ATen threshold_backward on the saved output; passes
 * where out is NaN) and bias_grad[c] = sum_p gx[p][c].  Workspace as for wtpse_channel_sum_nhwc; gx may alias g. */
int wtpse_relu_backward_channel_sum_nhwc(const float* g, const float* out, int64_t npix, int C, float* gx, float* bias_grad,
                                         void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
int wtpse_channel_sum_nhwc(const float* g, int64_t npix, int C, float* out,
                           void* workspace, size_t workspace_bytes, wtpse_stream_t stream);

/*
 * Training-mode nn.BatchNorm2d followed (relu != 0) by ReLU, on a channels-last activation of npix pixels x C channels
 * -- the conv -> bn -> relu stages of ConvD / ConvU / DoubleConv (algorithms.py:877-962, 398-413).
 *   forward : batch mean / biased variance per channel, y = relu?((x - mean) * gamma / sqrt(var + eps) + beta);
 *             running_mean = (1 - momentum) * running_mean + momentum * (mean + mean_shift)   (mean_shift NULL == 0: the bias of
 *             the preceding convolution when its add was folded away), running_var likewise with the UNBIASED variance
 *             (torch.nn.BatchNorm2d semantics; either may be NULL);  save_stats[3][C] = mean, invstd, gamma * invstd.
 *   backward: dx, dgamma, dbeta of that composite given dy (gradient of y); the ReLU mask is recomputed from x.
 * C a power of two in [4, 1024]; gamma and beta required; all pointers 16-byte aligned; deterministic (no atomics).
 */
size_t wtpse_batchnorm_workspace_bytes(int64_t npix, int C);
int wtpse_batchnorm_relu_forward(const float* x, int64_t npix, int C, const float* gamma, const float* beta,
                                 const float* mean_shift, float eps, float momentum, int relu,
                                 float* running_mean, float* running_var, float* y, float* save_stats,
                                 void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
int wtpse_batchnorm_relu_backward(const float* x, const float* dy, int64_t npix, int C, const float* gamma, const float* beta,
                                  const float* save_stats, int relu, float* dx, float* dgamma, float* dbeta,
                                  void* workspace, size_t workspace_bytes, wtpse_stream_t stream);

/*
 * Encoder down-sampling: F.max_pool2d(x, 2) (ConvD.forward, algorithms.py:897) on a channels-last tensor whose H and W are even.
 *   backward == 0: in [N][2Ho][2Wo][C] -> out [N][Ho][Wo][C], argmax[N*Ho*Wo*C] = position 0..3 inside the 2x2 window
 *                  (row-major; ATen's selection rule: first maximum, NaN wins)
 *   backward != 0: in = gradient of the pooled output, out = gradient of the input (every element written once)
 * C % 4 == 0; float pointers 16-byte aligned.
 */
int wtpse_maxpool2_nhwc(const float* in, float* out, unsigned char* argmax, int64_t N, int Ho, int Wo, int C, int backward,
                        wtpse_stream_t stream);

/*
 * Channels-last variants: z, relu_out, grad_relu and dz are [B][P][16] -- the memory of channels-last B x 16 x H x W
 * tensors (what a channels-last backbone hands over), so no layout conversion is needed around the loss.  Same
 * arithmetic and outputs as the NCHW entry points (summation order differs, results agree to fp32 rounding).
 * relu_out / grad_relu may be NULL: then these are plain wtpse_whitening_forward / _backward; non-NULL: the fused
 * DeepWT tail of wtpse_whitening_relu_forward / _backward.  All tensor pointers 16-byte aligned.
 */
int wtpse_whitening_forward_cl(const float* z, float* relu_out, int B, int C, int64_t P,
                               int n_per_domain, int n_domains, float margin, float eps,
                               float* losses, float* gram, float* rowstat, float* domgrad,
                               void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
int wtpse_whitening_backward_cl(const float* z, const float* grad_relu, const float* gram, const float* rowstat,
                                const float* domgrad, const float* g_off, const float* g_diag, const float* g_dom,
                                int B, int C, int64_t P, int n_per_domain, int n_domains, float* dz,
                                wtpse_stream_t stream);

/* ---- standalone MMD: compute_MMD.forward, algorithms.py:102-121 / shape_networks.py:283-309 ---- */

size_t wtpse_mmd_workspace_bytes(int B);
/* v is [B][D] with D == 120; loss[0] = mean over domain pairs of (Kxx + Kyy - 2Kxy), gamma = [1]. */
int wtpse_mmd_forward(const float* v, int B, int D, int n_per_domain, int n_domains, float* loss,
                      void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
/* dv[B][D] = gout * dloss/dv (rows beyond n_domains*n_per_domain get 0); gout DEVICE scalar, NULL == 1. */
int wtpse_mmd_backward(const float* v, const float* gout, int B, int D, int n_per_domain, int n_domains,
                       float* dv, void* workspace, size_t workspace_bytes, wtpse_stream_t stream);

/* ---- KD loss: ShapeVariationalDist_x.wasser_distance, shape_networks.py:596-597 ----------- */

size_t wtpse_mse_workspace_bytes(int64_t N);
/* loss[0] = mean((a-b)^2) over N elements. */
int wtpse_mse_forward(const float* a, const float* b, int64_t N, float* loss,
                      void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
/* da = 2*gout*(a-b)/N, db = -da; gout is a DEVICE scalar; da or db may be NULL. */
int wtpse_mse_backward(const float* a, const float* b, const float* gout, int64_t N,
                       float* da, float* db, wtpse_stream_t stream);

/* ---- integer label path + coarse-to-fine ROI ----------------------------- */

/*
 * custom_transforms.py:466-499 (Normalize_tf) + :581-599 (ToTensor), batched on the device:
 *   image_chw[b][c][h][w] = float(img_hwc[b][h][w][c]) / 127.5 - 1.0      (img_hwc may be NULL to skip)
 *   label_od = [raw_od <= 200], label_oc = [raw_od <= 50]  as {0,1} floats, [B][1][H][W]
 * raw_oc is accepted for signature parity; the reference overwrites it completely from the OD mask
 * (custom_transforms.py:493-494).  Bit-exact.
 */
int wtpse_prepare_batch(const unsigned char* img_hwc, const unsigned char* raw_od, const unsigned char* raw_oc,
                        int B, int H, int W, float* image_chw, float* label_od, float* label_oc,
                        wtpse_stream_t stream);

size_t wtpse_od_roi_workspace_bytes(void);
/*
 * Trainer.py:842-853 and :865-867:
 *   od_pred = (sigmoid(logits) > threshold) as float          [B][1][HW]
 *   image  += 1 (IN PLACE, as the reference does)              [B][C][HW]
 *   image_roi = image * od_pred - 1
 *   sums[0] = sum(od_pred), sums[1] = sum(od_pred * target_oc), sums[2] = sums[0]/sums[1] (1 if inf/nan)
 * target_oc and sums may be NULL.  Bit-exact against the same statements executed by ATen on the GPU.
 * Precondition: target_oc holds {0, 1} labels (custom_transforms.py:466-499); sums[1] counts the non-zero products, which is
 * exact and order-independent for binary targets and NOT the float sum for soft ones.
 */
int wtpse_od_roi(const float* logits, const float* target_oc, float* image, float* od_pred, float* image_roi,
                 int B, int C, int64_t HW, float threshold, float* sums,
                 void* workspace, size_t workspace_bytes, wtpse_stream_t stream);

/* ---- attention fuse: algorithms.py:1243-1249 (attention_layer :1120-1129) ---------------------- */

size_t wtpse_attention_fuse_workspace_bytes(int B, int64_t P);
/*
 * att = sigmoid(w * z_post + b)  (Conv2d(1,1,kernel_size=1) + Sigmoid), weight_bias = DEVICE {w, b}
 * att_mask = (att > threshold) as float, fuse = coef * emb + att * emb ; emb/fuse are [B][Ce][P].
 */
int wtpse_attention_fuse_forward(const float* emb, const float* z_post, const float* weight_bias, float coef,
                                 int B, int Ce, int64_t P, float threshold,
                                 float* fuse, float* att_mask, float* att, wtpse_stream_t stream);
/* d_emb / d_z_post / d_weight_bias ({dw, db}) may each be NULL. */
int wtpse_attention_fuse_backward(const float* grad_fuse, const float* emb, const float* z_post, const float* att,
                                  const float* weight_bias, float coef, int B, int Ce, int64_t P,
                                  float* d_emb, float* d_z_post, float* d_weight_bias,
                                  void* workspace, size_t workspace_bytes, wtpse_stream_t stream);

/* ---- Track W: wavelet transform + detail-coefficient shape loss -- PARITY UNPINNED --------------
 * The reference contains no wavelet code; these entry points replace nothing in it.
 * Conventions (oracle/wavelet_np.py): orthonormal Haar (wavelet = 0) / db2 (wavelet = 1), periodic
 * extension, Mallat-packed coefficients with the input's shape, J levels, H and W divisible by 2^J.
 * x / coef are [nmaps][H][W] fp32 (nmaps = batch * channels). */
size_t wtpse_wavelet_workspace_bytes(int nmaps, int H, int W, int J);
int wtpse_dwt2d_forward(const float* x, int nmaps, int H, int W, int wavelet, int J, float* coef,
                        void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
/* Synthesis == adjoint (orthonormal); scale is an optional DEVICE scalar multiplied into x. */
int wtpse_dwt2d_inverse(const float* coef, int nmaps, int H, int W, int wavelet, int J, float* x,
                        const float* scale, void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
/* loss = (1/nmaps) sum_maps sum_j w_j mean|detail_j| ; grad_coef receives dloss/dcoef (Mallat layout), so
 * the backward is wtpse_dwt2d_inverse(grad_coef, ..., scale = upstream gradient).  level_weights is a HOST
 * array of J floats (NULL = all ones). */
int wtpse_wavelet_loss_forward(const float* x, int nmaps, int H, int W, int wavelet, int J,
                               const float* level_weights, float* loss, float* grad_coef,
                               void* workspace, size_t workspace_bytes, wtpse_stream_t stream);

/*
 * Fused loss + gradient path: the loss AND dloss/dx from one call, no coefficient buffer in HBM (the L1 loss needs only
 * sign(d) of a detail coefficient: three 2-bit codes per site).
 *  - whole map resident (wavelet_resident.cu): every H x W map is held in the distributed shared memory of one
 *    thread-block cluster (row bands, 1-D TMA loads, filter-overlap rows exchanged through DSMEM); all J analysis
 *    levels, the L1 reduction and the whole synthesis of the gradient run shared-to-shared: 8 B per element of HBM
 *    traffic.  Chosen when the smallest cluster that holds a map has <= 2 CTAs.
 *  - streamed plan (wavelet_stream.cu, wavelet_tiles.cu): the first k levels go global-to-global through persistent
 *    TMA pipelines (detail bands reduced to one byte of packed signs per site), k growing until the k-th low-low band
 *    fits a cluster of <= 2 CTAs (512 x 512 maps: k = 1; 1024 x 1024: k = 2); that band is what the cluster-resident
 *    kernel then works on, in place.
 * wtpse_wavelet_resident_cluster returns the cluster size of the resident stage (1, 2, 4, 8; 1 also when every level
 * is streamed) or 0 when no fused plan exists for the shape (e.g. W not divisible by 2^(J+1) with level widths that
 * are not multiples of 32) -- use the per-level entry points above then.
 * grad_x (may be NULL: loss only) receives upstream * dloss/dx; upstream is an optional DEVICE scalar (NULL = 1).
 * Workspace: wtpse_wavelet_workspace_bytes.
 */
int wtpse_wavelet_resident_cluster(int H, int W, int wavelet, int J);
int wtpse_wavelet_loss_resident(const float* x, int nmaps, int H, int W, int wavelet, int J,
                                const float* level_weights, const float* upstream, float* loss, float* grad_x,
                                void* workspace, size_t workspace_bytes, wtpse_stream_t stream);
/* data[0..n) *= *scale unless *scale == 1 (decided on the device; no host synchronisation).  Used by the autograd
 * backward of the resident path: the gradient was already written for upstream = 1. */
int wtpse_scale_unless_one(float* data, int64_t n, const float* scale, wtpse_stream_t stream);

/* ---- host-buffer entry point (plugin-facing, used for the end-to-end number) --------------- */

typedef struct wtpse_host_plan wtpse_host_plan;

/* Allocates device buffers for one [B][16][P] batch on the current device. */
int  wtpse_host_plan_create(int B, int64_t P, wtpse_host_plan** plan);
void wtpse_host_plan_destroy(wtpse_host_plan* plan);
/*
 * z_host -> device, forward, backward with upstream weights grad_w[3] = (g_off, g_diag, g_dom),
 * dz -> dz_host (may be NULL to skip backward), losses_host[4] as in `losses`.
 * Synchronous: returns when the host buffers are complete.
 */
int  wtpse_host_plan_run(wtpse_host_plan* plan, const float* z_host,
                         int n_per_domain, int n_domains, float margin, float eps,
                         const float grad_w[3], float losses_host[4], float* dz_host);
/*
 * Asynchronous form: enqueue one step (same work as _run) and return; the host buffers of a step are complete
 * after wtpse_host_plan_wait(), or after the second-next submit (two device slots alternate, so the D2H of one
 * step overlaps the H2D of the next on the full-duplex link).  Each step needs its own host buffers until then.
 */
int  wtpse_host_plan_submit(wtpse_host_plan* plan, const float* z_host,
                            int n_per_domain, int n_domains, float margin, float eps,
                            const float grad_w[3], float losses_host[4], float* dz_host);
int  wtpse_host_plan_wait(wtpse_host_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* WTPSE_B200_H */

// Internal launch interface between the C-ABI layer (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "whitening_matrix.cuh"
#include "whitening_tail.cuh"

namespace wtpse {

struct GramPlan {
    bool tma;                    // 1-D TMA pipeline (P % 4 == 0, 16B-aligned base) or generic fallback
    long long tiles_per_sample;  // persistent path only
    long long T;                 // total tiles
    long long G;                 // grid (CTAs)
    int nslots;                  // partial slots per sample
    int group;                   // always 1: one contiguous tile range per CTA (round-robin groups measured slower)
    long long item_px;           // channels-last schedule only: pixels per work item
};

GramPlan plan_gram(const float* z, int B, long long P, int sm_count, const float* relu_out = nullptr);
size_t gram_partial_floats(int B, long long P, int sm_count);
// channels-last input ([B][P][16]); relu_out (nullable, same layout): also write relu(z)
GramPlan plan_gram_cl(int B, long long P, int sm_count);
cudaError_t launch_gram_cl(const float* z, float* relu_out, float* partial, int* slot_count, int B, long long P, const GramPlan& g,
                           cudaStream_t stream);
// channels-last backward: dz = M_b z (+ [z > 0] * grelu), all [B][P][16]
cudaError_t launch_apply_cl(const float* z, const float* grelu, const SeedArgs& seed, float* dz, int B, long long P,
                            int sm_count, cudaStream_t stream);
// relu_out: also write relu(z) (8(f).1); tail: run the rest of the forward inside the kernel (whitening_tail.cuh)
cudaError_t launch_gram(const float* z, float* partial, int* slot_count, int B, long long P, const GramPlan& g,
                        cudaStream_t stream, float* relu_out = nullptr, const TailParams* tail = nullptr);
bool gram_tail_fits(int B, int n_per_domain, int n_domains);

// channels-last kernels as tensor-map TMA pipelines (whitening_cl_tma.cu); cl_tma_ok: pointers (nullable) and size allow them
bool cl_tma_ok(const float* a, const float* b, const float* c, long long P);
GramPlan plan_gram_cl_tma(int B, long long P, int sm_count);
bool gram_cl_tail_fits(int B, int n_per_domain, int n_domains);
cudaError_t launch_gram_cl_tma(const float* z, float* relu_out, float* partial, int* slot_count, int B, long long P, const GramPlan& g,
                               cudaStream_t stream, const TailParams* tail);
cudaError_t launch_apply_cl_tma(const float* z, const float* grelu, const SeedArgs& seed, float* dz, int B, long long P, int sm_count,
                                cudaStream_t stream);

// stand-alone forward tail (whitening_epilogue.cu): the fallback of the in-kernel tail.  `scratch` is the global fallback
// for the single-CTA kernels' working set (epilogue_scratch_bytes).
size_t epilogue_scratch_bytes(int B, int K);
extern int g_apply_round_robin;
extern int g_l2_evict_first;
extern int g_cl_tma;

// stage 2a: one CTA per sample reduces the per-CTA partials -> gram, rowstat, vd [B][124], statd [B][2]
cudaError_t launch_gram_reduce(const float* partial, const int* slot_count, int nslots, int B, long long P, int n_per_domain,
                               int n_domains, float margin, float eps, float* gram, float* rowstat, float* vd, float* statd,
                               cudaStream_t stream);
// stage 2b: ONE CTA, vd / statd -> losses
cudaError_t launch_whiten_epilogue_fwd(int B, long long P, int n_per_domain, int n_domains, float* losses, void* scratch,
                                       cudaStream_t stream, const float* vd, const float* statd);

// standalone compute_MMD.forward / backward on v[B][stride]; dv == nullptr selects the forward
cudaError_t launch_mmd(const float* v, int stride, const float* gout, int B, int n_per_domain, int n_domains, float* loss,
                       float* dv, void* scratch, cudaStream_t stream);

// backward apply: dz_b = M_b z_b
// grelu != nullptr: dz_b = M_b z_b + [z_b > 0] * grelu_b (ReLU backward + gradient sum fused)
// M_b is derived in the kernel from the forward's saved tensors and the upstream scalars (whitening_matrix.cuh)
cudaError_t launch_apply(const float* z, const SeedArgs& seed, float* dz, int B, long long P, int sm_count,
                         cudaStream_t stream, const float* grelu = nullptr);

// whitening_apply_relu.cu: the TMA path of the grelu variant (z and grelu both staged by the producer warp)
bool apply_relu_tma_ok(const float* z, const float* grelu, const float* dz, long long P);
cudaError_t launch_apply_relu(const float* z, const float* grelu, const SeedArgs& seed, float* dz, int B, long long P,
                              int sm_count, cudaStream_t stream);

// KD MSE
size_t mse_partial_doubles(long long N, int sm_count);
cudaError_t launch_mse_fwd(const float* a, const float* b, long long N, float* loss, double* partial, int sm_count,
                           cudaStream_t stream);
cudaError_t launch_mse_bwd(const float* a, const float* b, const float* gout, long long N, float* da, float* db,
                           int sm_count, cudaStream_t stream);

// element-wise kernels (elementwise.cu)
cudaError_t launch_prepare_batch(const unsigned char* img, const unsigned char* raw_od, const unsigned char* raw_oc, int B,
                                 long long HW, float* image, float* label_od, float* label_oc, int sm_count,
                                 cudaStream_t stream);
size_t od_roi_workspace_bytes();
cudaError_t launch_od_roi(const float* logits, const float* target_oc, float* image, float* od_pred, float* image_roi, int B,
                          int C, long long HW, float thr, float* sums, void* workspace, int sm_count, cudaStream_t stream);
cudaError_t launch_fuse_fwd(const float* emb, const float* zp, const float* wb, float coef, int B, int Ce, long long P,
                            float thr, float* fuse, float* mask, float* att, int sm_count, cudaStream_t stream);
size_t fuse_bwd_partial_doubles(int B, long long P, int sm_count);
cudaError_t launch_fuse_bwd(const float* g, const float* emb, const float* zp, const float* att, const float* wb, float coef,
                            int B, int Ce, long long P, float* d_emb, float* d_zp, float* d_wb, double* partial, int sm_count,
                            cudaStream_t stream);

// bilinear x2 up-sampling, channels-last (backbone_elementwise.cu); adjoint = its backward.  H, W are the LOW-resolution sizes.
cudaError_t launch_upsample2x_nhwc(const float* in, float* out, long long N, int H, int W, int C, bool adjoint, int sm_count,
                                   cudaStream_t stream);

// y = act(y + bias[c]) in place on a channels-last tensor of npix pixels x C channels (backbone_elementwise.cu)
cudaError_t launch_bias_act_nhwc(float* y, const float* bias, long long npix, int C, bool relu, int sm_count, cudaStream_t stream);

// out[c] = sum_p g[p][c] on a channels-last tensor (a convolution's bias gradient); partial: channel_sum_blocks * C floats
int channel_sum_blocks(long long npix, int C, int sm_count);
bool channel_sum_supported(int C);
cudaError_t launch_channel_sum_nhwc(const float* g, long long npix, int C, float* out, float* partial, int sm_count,
                                    cudaStream_t stream);

// 2x2 stride-2 max pooling, channels-last; arg: one byte per pooled element (position inside the window)
cudaError_t launch_maxpool2_nhwc(const float* in, float* out, unsigned char* arg, long long N, int Ho, int Wo, int C, bool backward,
                                 int sm_count, cudaStream_t stream);

// training-mode BatchNorm2d (+ ReLU), channels-last (batchnorm.cu)
bool batchnorm_supported(int C);
size_t batchnorm_workspace_floats(long long npix, int C, int sm_count);
cudaError_t launch_batchnorm_fwd(const float* x, long long npix, int C, const float* gamma, const float* beta,
                                 const float* mean_shift, float eps, float momentum, bool relu, float* running_mean,
                                 float* running_var, float* y, float* save_mean, float* save_invstd, float* save_scale,
                                 float* workspace, int sm_count, cudaStream_t stream);
cudaError_t launch_batchnorm_bwd(const float* x, const float* dy, long long npix, int C, const float* gamma, const float* beta,
                                 const float* save_mean, const float* save_invstd, const float* save_scale, bool relu, float* dx,
                                 float* dgamma, float* dbeta, float* workspace, int sm_count, cudaStream_t stream);

// gx = [out > 0] * g and bias_grad[c] = sum_p gx[p][c] in one pass (backward of bias + ReLU); partial as for channel_sum
cudaError_t launch_relu_bwd_channel_sum_nhwc(const float* g, const float* out, float* gx, long long npix, int C, float* bias_grad,
                                             float* partial, int sm_count, cudaStream_t stream);

// Track W (wavelet.cu)
size_t wavelet_scratch_floats(long long nmaps, int H, int W);
size_t wavelet_partial_doubles(long long nmaps, int H, int W, int J);
cudaError_t launch_dwt(const float* x, int nmaps, int H, int W, int taps, int J, float* coef, float* scratch,
                       const float* weights_host, float* loss, double* partial, cudaStream_t stream);
cudaError_t launch_idwt(const float* coef, int nmaps, int H, int W, int taps, int J, float* x, float* scratch,
                        const float* scale, cudaStream_t stream);

cudaError_t launch_wavelet_loss_final(const double* partial, int n, float* loss, cudaStream_t stream);

// Track W, cluster-resident fused loss + gradient (wavelet_resident.cu)
extern int g_wavelet_resident;
extern int g_wavelet_cluster_max;
int wavelet_resident_cluster(int H, int W, int taps, int J, int nmaps = 0);   // cluster size, 0 = the map does not fit
cudaError_t launch_wavelet_resident(const float* x, int nmaps, int H, int W, int taps, int J, const float* weights_host,
                                    const float* upstream, float* loss, float* grad, double* partial, cudaStream_t stream,
                                    int* n_partials = nullptr);
// streaming level 1 + resident levels 2..J (wavelet_stream.cu)
extern int g_wavelet_split;
extern int g_wavelet_peel_max;
extern int g_wavelet_haar_min_log2px;
int wavelet_fused_plan(int H, int W, int taps, int J, int nmaps = 0);          // 0 none, 1 whole map resident, 2 streamed plan
int wavelet_stream_levels(int H, int W, int taps, int J, int nmaps, int* cs);  // streamed levels k (0: no such plan)
size_t wavelet_stream_partials(int nmaps, int H, int W);
cudaError_t launch_wavelet_loss_split(const float* x, int nmaps, int H, int W, int taps, int J, const float* weights_host,
                                      const float* upstream, float* loss, float* grad, float* scratch, double* partial,
                                      int sm_count, cudaStream_t stream);
// level 1 of the streamed plan as persistent TMA pipelines (wavelet_tiles.cu); R = 0: shape not taken
extern int g_wavelet_tiles;
void wavelet_tile_plan(int H, int W, int taps, bool has_ll, int* R_fwd, int* S_fwd, int* R_inv, int* S_inv);
cudaError_t launch_dwt1_tiles(const float* x, float* ll, unsigned char* sg, int nmaps, int H, int W, int taps, int R, int S,
                              float sc, bool grad, double* partial, int sm_count, cudaStream_t stream, int* n_partials);
cudaError_t launch_idwt1_tiles(const float* gll, const unsigned char* sg, float* out, int nmaps, int H, int W, int taps, int R,
                               int S, float sc, const float* upstream, bool has_ll, const double* partial, int n_partials,
                               float* loss, int sm_count, cudaStream_t stream);
// db2 levels in factored form, one or two levels per pass (wavelet_db2.cu)
extern int g_wavelet_db2;
extern int g_wavelet_db2_two;
extern int g_wavelet_db2_deep;
extern int g_wavelet_db2_rf, g_wavelet_db2_ri, g_wavelet_db2_nw2, g_wavelet_db2_rr;
extern int g_wavelet_haar_passes;
bool wavelet_db2_pass(int H, int W, bool two, bool has_ll, int* R_fwd, int* S_fwd, int* NC_fwd, int* R_inv, int* S_inv, bool haar = false);
cudaError_t launch_db2_analysis(const float* x, float* ll, unsigned char* sg1, unsigned char* sg2, int nmaps, int H, int W, bool two,
                                float sc1, float sc2, bool grad, bool pdl_wait, double* partial, int sm_count, cudaStream_t stream,
                                int* n_partials, bool haar = false);
cudaError_t launch_db2_synthesis(const float* g, const unsigned char* sg1, const unsigned char* sg2, float* out, int nmaps, int H, int W,
                                 bool two, bool has_ll, float sc1, float sc2, const float* upstream, const double* partial, int n_partials,
                                 float* loss, int sm_count, cudaStream_t stream, bool haar = false);
cudaError_t launch_scale_unless_one(float* data, long long n, const float* scale, int sm_count, cudaStream_t stream);

}  // namespace wtpse

/* sisr_b200 -- C ABI of the B200-native SRGAN training-step kernels.
 *
 * The reference (keyber/Single-Image-Super-Resolution) is pure Python: it has no FFI / plugin
 * interface of its own, every operator on its hot path is an ATen library call made from
 * nn.Module.forward.  This header is therefore the boundary a maintainer
 * would bind in place of those ATen calls; each entry cites the reference call site it replaces.
 * The reference-side binding (a ctypes stub) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless stated otherwise;
 *   - activations are NHWC bf16 (uint16_t storage), parameters / gradients / statistics fp32;
 *   - no allocation, no ownership: outputs and workspaces are caller-provided;
 *   - asynchronous on `stream` (a cudaStream_t passed as void*); CUDA-graph capturable;
 *   - return 0 on success, non-zero on error, message via sisr_last_error() (thread-local).
 */
#ifndef SISR_B200_H_
#define SISR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t sisr_bf16;

/* activation codes for fused epilogues */
enum { SISR_ACT_NONE = 0, SISR_ACT_RELU = 1, SISR_ACT_LEAKY = 2, SISR_ACT_PRELU = 3, SISR_ACT_TANH = 4 };

/* Convolution geometry: input [n,h,w,cin] -> output [n,oh,ow,cout], square kernel/stride/pad.
 * ps_r = 2 means the conv is followed by PixelShuffle(2): the output is written directly as
 * [n, 2*oh, 2*ow, cout/4] and the prepared weights carry the matching row permutation. */
typedef struct sisr_conv_desc {
  int n, h, w, cin;
  int oh, ow, cout;
  int k, stride, pad;
  int ps_r;
} sisr_conv_desc;

const char* sisr_last_error(void);
int sisr_abi_version(void);
/* number of partial-sum rows of a `stats` buffer (= SM count: one row per persistent CTA) */
int sisr_stats_rows(void);
/* debug / A-B timing: 0 = layers with cout <= 128 use the 128-pixel x cout tiles instead of the
 * 128-channel x 256-pixel (transposed) tiles */
int sisr_debug_transposed(int on);
/* debug / A-B timing: 0 = the 64 -> 64 channel stride-1 layers (generator trunk, model_generator.py:10,13,39;
 * VGG conv1_2) do not use the halo-fed kernel with the pixels on the UMMA M side (csrc/igemm_pm.cu); default 1 */
int sisr_debug_pm_mode(int on);
/* 1 if the tcgen05 implicit-GEMM engine takes this fprop/dgrad shape, 0 if the CUDA-core kernel does */
int sisr_conv_uses_tensor_cores(const sisr_conv_desc* d);

/* ---- layout at the module boundary (NCHW fp32 <-> NHWC bf16); replaces nothing in the reference,
 *      it is the price of keeping the reference's NCHW fp32 tensors at the nn.Module surface ---- */
int sisr_nchw_f32_to_nhwc_bf16(const float* x, sisr_bf16* y, int n, int c, int h, int w, void* stream);
int sisr_nhwc_bf16_to_nchw_f32(const sisr_bf16* x, float* y, int n, int c, int h, int w, void* stream);
/* x: [batch][rows][cols] -> y: [batch][cols][rows]  (discriminator flatten, model_discriminator.py:59) */
int sisr_transpose_bf16(const sisr_bf16* x, sisr_bf16* y, int batch, int rows, int cols, void* stream);
/* dpre[nhwc bf16] = dout[nchw f32] * (1 - y[nchw f32]^2)   (Tanh backward, model_generator.py:53,63) */
int sisr_tanh_bwd_nchw_to_nhwc(const float* dout, const float* y, sisr_bf16* dpre, int n, int c, int h,
                               int w, void* stream);

/* ---- spectral norm + weight preparation: torch.nn.utils.spectral_norm as applied at
 *      model_generator.py:10,13,33,39,45,52,123 and model_discriminator.py:10,39 ---- */
size_t sisr_sn_workspace_floats(int cout, int k);
int sisr_sn_power_iteration(const float* w_orig, float* u, float* v, float* sigma, int cout, int k,
                            int training, float eps, float* workspace, void* stream);
/* Batched variants: one launch chain for every conv of a network (37 layers in G, 8 in D).  `layers`
 * is a HOST array; the pointers inside are DEVICE pointers.  t [k] and s [cout] are scratch;
 * u_saved / v_saved receive the vectors this call's sigma was computed with (for the backward pass).
 * training = 0: no iteration, sigma = u^T W v with the stored vectors. */
typedef struct sisr_sn_layer {
  const float* w;
  float *u, *v, *t, *s, *sigma, *u_saved, *v_saved;
  int cout, k, training, reserved;
} sisr_sn_layer;
typedef struct sisr_prep_layer {
  const float* w;
  const float* sigma;      /* nullable */
  const float* bias;       /* nullable unless bias_perm */
  sisr_bf16* w_fprop;
  sisr_bf16* w_dgrad;      /* nullable */
  float* bias_perm;        /* nullable */
  int cout, cin, k, ps_r;
} sisr_prep_layer;
int sisr_sn_power_iteration_batched(const sisr_sn_layer* layers, int n_layers, float eps, void* stream);
int sisr_weight_prep_batched(const sisr_prep_layer* layers, int n_layers, void* stream);
/* w: [cout,cin,k,k] fp32 -> w_fprop: [cout',k,k,cin] bf16, w_dgrad: [cin,k,k,cout'] bf16 (nullable),
 * both scaled by 1/sigma (sigma nullable); bias_perm (nullable) = bias in the permuted row order. */
int sisr_weight_prep(const float* w, const float* sigma, const float* bias, sisr_bf16* w_fprop,
                     sisr_bf16* w_dgrad, float* bias_perm, int cout, int cin, int k, int ps_r,
                     void* stream);
/* g_prepared: fp32 [cout',k,k,cin] from sisr_conv_wgrad -> dw: [cout,cin,k,k] fp32 including the
 * gradient through sigma (u, v constant); sigma == NULL means no spectral norm. workspace: 4 floats */
int sisr_weight_grad_finish(const float* g_prepared, const float* w_orig, const float* u, const float* v,
                            const float* sigma, float* dw, const float* dbias_perm, float* dbias,
                            int cout, int cin, int k, int ps_r, int accumulate, float* workspace,
                            void* stream);

/* ---- convolutions: nn.Conv2d at model_generator.py:10,13,33,39,45,52,123,
 *      model_discriminator.py:10,39 and torchvision vgg19.features (model_content_extractor.py:43) ---- */
/* y (bf16 NHWC, or pixel-shuffled) and/or y_nchw_f32 (fp32 NCHW, edge layers only) = act(conv(x)+bias).
 * stats (nullable): fp32 [sisr_stats_rows()][2*cout], overwritten: row r holds the per-channel {sum, sum of
 * squares} of y over the output tiles of CTA r (BN batch statistics fused in the conv epilogue, no
 * atomics); sisr_bn_finalize adds the rows. */
int sisr_conv_fprop(const sisr_conv_desc* d, const sisr_bf16* x, const sisr_bf16* w_fprop,
                    const float* bias, int act, float slope, const float* slope_ptr, sisr_bf16* y,
                    float* y_nchw_f32, float* stats, void* stream);
/* dx = conv_transpose(dy): dy is the gradient w.r.t. the pre-activation conv output in the layout
 * fprop wrote it (pixel-shuffled when ps_r = 2). */
int sisr_conv_dgrad(const sisr_conv_desc* d, const sisr_bf16* dy, const sisr_bf16* w_fprop,
                    const sisr_bf16* w_dgrad, sisr_bf16* dx, void* stream);
/* dgrad with the activation backward of the conv's INPUT tensor fused into the store:
 * dx = conv_transpose(dy) * (x > 0 ? 1 : mask_slope), mask = x (the input of this conv, which is the
 * ReLU-family output of the previous layer; vgg19.features ReLU, model_content_extractor.py:43).
 * Only where sisr_conv_dgrad_fuses_mask(d) = 1 (tensor-core path). */
int sisr_conv_dgrad_fuses_mask(const sisr_conv_desc* d);
int sisr_conv_dgrad_masked(const sisr_conv_desc* d, const sisr_bf16* dy, const sisr_bf16* w_fprop,
                           const sisr_bf16* w_dgrad, sisr_bf16* dx, const sisr_bf16* mask, float mask_slope,
                           void* stream);
/* g_prepared: fp32 [cout',k,k,cin] (overwritten); dbias_perm: fp32 [cout'] nullable.
 * workspace: sisr_conv_wgrad_workspace_bytes(d) bytes (split-K partials), may be NULL if that is 0. */
size_t sisr_conv_wgrad_workspace_bytes(const sisr_conv_desc* d);
int sisr_conv_wgrad(const sisr_conv_desc* d, const sisr_bf16* x, const sisr_bf16* dy, float* g_prepared,
                    float* dbias_perm, void* workspace, void* stream);

/* sisr_conv_wgrad + sisr_weight_grad_finish in one call (what the autograd operators use): the split-K
 * partials are reduced, divided by sigma, corrected for the gradient through sigma and re-laid out as
 * dw [cout,cin,k,k] by one cooperative kernel.  dbias_in (nullable): bias gradient in prepared row order
 * already reduced by the caller (sisr_bn_bwd_apply / sisr_act_bwd colsum); NULL: computed here.
 * workspace: sisr_conv_wgrad_fused_workspace_bytes(d) bytes. */
size_t sisr_conv_wgrad_fused_workspace_bytes(const sisr_conv_desc* d);
int sisr_conv_wgrad_fused(const sisr_conv_desc* d, const sisr_bf16* x, const sisr_bf16* dy,
                          const float* w_orig, const float* u, const float* v, const float* sigma,
                          float* dw, const float* dbias_in, float* dbias, int accumulate, void* workspace,
                          void* stream);
/* debug: 1 = never use the cooperative launch (two ordinary kernels instead) */
int sisr_debug_disable_cooperative(int off);
/* tools only (tools/wgrad_phases.py): 8 int64 DEVICE counters that CTA 0 of every tensor-core weight-gradient launch
 * adds its phase cycles to; NULL (default) switches the instrumentation off */
int sisr_debug_wgrad_counters(long long* device_counters);

/* ---- BatchNorm2d (train / eval) fused with PReLU / LeakyReLU / residual add:
 *      model_generator.py:11-14,16-19,40,93 and model_discriminator.py:11-12 ---- */
int sisr_bn_stats(const sisr_bf16* y, long long rows, int c, float* stats, void* stream);
/* stats: [stats_rows][2c] partial sums (stats_rows = 1 for sisr_bn_stats output) */
int sisr_bn_finalize(const float* stats, int stats_rows, float count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, long long* num_batches_tracked,
                     float momentum, float eps, int training, float* scale, float* shift, float* mean,
                     float* invstd, int c, void* stream);
int sisr_bn_apply(const sisr_bf16* y, const float* scale, const float* shift, int act, float slope,
                  const float* slope_ptr, const sisr_bf16* residual, sisr_bf16* out, long long rows,
                  int c, void* stream);
/* sums: fp32 [2c+1] accumulated (caller zeroes): {sum g, sum g*xhat, sum dout*min(0,z)} */
int sisr_bn_bwd_reduce(const sisr_bf16* dout, const sisr_bf16* y, const float* mean, const float* invstd,
                       const float* scale, const float* shift, int act, float slope,
                       const float* slope_ptr, float* sums, long long rows, int c, void* stream);
/* colsum (nullable, fp32 [c], accumulated - caller zeroes): per-channel sum of the dy written, i.e. the
 * bias gradient of the conv that feeds this BatchNorm */
int sisr_bn_bwd_apply(const sisr_bf16* dout, const sisr_bf16* y, const float* mean, const float* invstd,
                      const float* scale, const float* shift, int act, float slope,
                      const float* slope_ptr, const float* sums, float count, sisr_bf16* dy,
                      float* colsum, long long rows, int c, void* stream);
/* backward of an activation fused in a conv epilogue, from its OUTPUT (needs slope > 0), on [rows, c]:
 * din = dout * f'(out); dslope (nullable, accumulated) += sum dout * min(0, pre);
 * colsum (nullable, fp32 [c], accumulated): per-channel sum of din (the conv's bias gradient) */
int sisr_act_bwd(const sisr_bf16* dout, const sisr_bf16* out, int act, float slope, const float* slope_ptr,
                 sisr_bf16* din, float* dslope, float* colsum, long long rows, int c, void* stream);

/* ---- MaxPool2d(2,2) of torchvision vgg19.features ---- */
int sisr_maxpool2_fwd(const sisr_bf16* x, sisr_bf16* y, int n, int h, int w, int c, void* stream);
int sisr_maxpool2_bwd(const sisr_bf16* x, const sisr_bf16* dy, sisr_bf16* dx, int n, int h, int w, int c,
                      void* stream);

/* ---- discriminator head: model_discriminator.py:47-53,59-60 ---- */
int sisr_dhead_forward(const sisr_bf16* x_flat, const float* w0, const float* b0, const float* w2,
                       const float* b2, float slope, float* h, float* p, int batch, int fc_in, int fc_mid,
                       void* stream);
int sisr_dhead_backward(const sisr_bf16* x_flat, const float* w0, const float* w2, const float* h,
                        const float* p, const float* dp, float slope, float* dh, float* dw0, float* db0,
                        float* dw2, float* db2, float* dx_flat, int batch, int fc_in, int fc_mid,
                        int need_wgrad, void* stream);

/* ---- stand-alone nn.PixelShuffle(2) on NHWC bf16 (model_generator_progressive.py:54, the 64->16->4
 *      channel stages whose convs are too narrow for the fused store): inverse = 0: x [n,h,w,4*c_out]
 *      -> y [n,2h,2w,c_out]; inverse = 1: x is the gradient [n,2h,2w,c_out], y receives [n,h,w,4*c_out] ---- */
int sisr_pixel_shuffle2(const sisr_bf16* x, sisr_bf16* y, int n, int h, int w, int c_out, int inverse,
                        void* stream);

/* ---- LR synthesis: utils.lr_from_hr (utils.py:16-31) = F.interpolate(bicubic, align_corners=True) +
 *      clamp to [-1,1]; NCHW fp32.  The backward (content_loss_on_lr mode, train.py:95-97) passes the
 *      gradient where the interpolated value stayed inside (-1, 1). ---- */
int sisr_lr_from_hr(const float* hr, float* lr, int n, int c, int h, int w, int oh, int ow, void* stream);
int sisr_lr_from_hr_bwd(const float* hr, const float* dlr, float* dhr, int n, int c, int h, int w, int oh,
                        int ow, void* stream);

/* ---- losses: nn.BCELoss (config.py:107; train.py:135,159,177), feature MSE (train.py:183-186) ---- */
int sisr_bce_fwd(const float* p, int n, float target, float* loss, float* mean_p, void* stream);
int sisr_bce_bwd(const float* p, int n, float target, const float* gout, float* dp, void* stream);
/* loss = coef * sum (a-b)^2 ; grads: gb = gout*2*coef*(b-a), ga = -gb (either nullable) */
int sisr_mse_fwd(const float* a, const float* b, long long n, float coef, float* loss, void* stream);
int sisr_mse_bwd(const float* a, const float* b, long long n, float coef, const float* gout, float* ga,
                 float* gb, void* stream);

/* ---- image-quality metrics: PSNR and SSIM (11 x 11 Gaussian window, sigma 1.5, K1 0.01, K2 0.03) per image
 *      of two NCHW fp32 batches with dynamic range `range` (2 for [-1, 1]).  The reference lists them as a todo
 *      (README.md:88) for its viewer flow (visualisation.py:46-52).  workspace: 2 * n floats. ---- */
int sisr_psnr_ssim(const float* a, const float* b, int n, int c, int h, int w, float range, float* workspace,
                   float* psnr, float* ssim, void* stream);

/* ---- data parallelism: replaces nn.DataParallel (config.py:114-118).  Gradient all-reduce is NCCL (host
 *      side, parallel.py); the per-layer SyncBN statistic exchange runs over NVLink peer memory:
 *      every rank allocates one workspace (sisr_peer_alloc), publishes its CUDA-IPC handle, maps the
 *      peers' workspaces (sisr_peer_open) and passes the HOST array bases[world] (own pointer at
 *      bases[rank]) to the calls below.  slot: a distinct index in [0, 512) per exchange of a step,
 *      the same on every rank.  Results are bit-identical on all ranks. ---- */
size_t sisr_peer_workspace_bytes(void);
int sisr_peer_handle_bytes(void);
int sisr_peer_alloc(void** ptr);
int sisr_peer_free(void* ptr);
int sisr_peer_get_handle(void* ptr, void* handle);
int sisr_peer_open(const void* handle, void** ptr);
int sisr_peer_close(void* ptr);
/* buf[i] <- sum over ranks of buf[i], i < n <= 1152 */
int sisr_peer_allreduce(void* const* bases, int rank, int world, int slot, float* buf, int n, void* stream);
/* sisr_bn_finalize (training) on the global batch: local partial rows are added, the [2c] sums are
 * exchanged, count = global number of elements per channel */
int sisr_bn_finalize_sync(void* const* bases, int rank, int world, int slot, const float* stats,
                          int stats_rows, float count, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches_tracked,
                          float momentum, float eps, float* scale, float* shift, float* mean,
                          float* invstd, int c, void* stream);

/* sisr_bn_bwd_reduce with the cross-GPU exchange fused in: sums (caller zeroes) receives the LOCAL
 * [2c+1] sums, sums_global the sums over all ranks; ticket: one zeroed 32-bit word of scratch */
int sisr_bn_bwd_reduce_sync(void* const* bases, int rank, int world, int slot, const sisr_bf16* dout,
                            const sisr_bf16* y, const float* mean, const float* invstd, const float* scale,
                            const float* shift, int act, float slope, const float* slope_ptr, float* sums,
                            float* sums_global, void* ticket, long long rows, int c, void* stream);

/* ---- optimiser: torch.optim.Adam + LambdaLR (config.py:170-180, 293-294; train.py:75,108,121-122) ---- */
int sisr_adam_tick(int* step, float lr0, float decay, float b1, float b2, float* hyper, void* stream);
/* p/g/m/v/numel are HOST arrays of n device pointers / element counts */
int sisr_adam_multi(int n, float* const* p, const float* const* g, float* const* m, float* const* v,
                    const long long* numel, const float* hyper, float b1, float b2, float eps,
                    float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SISR_B200_H_ */

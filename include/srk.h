/*
 * srk.h — C ABI of libsrk.so: the B200 (sm_100a) super-resolution hot-path kernels.
 *
 * The reference (Jaskieeeer/food101-super-resolution) has no FFI: its hot path is the set of
 * ATen ops dispatched from src/models.py, src/loss.py and src/metrics.py.  Every entry point
 * below replaces one such op family; the reference call sites are cited per function
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - the caller owns every buffer, including workspaces (sizes via srk_*_workspace_bytes);
 *   - `stream` is a cudaStream_t passed as void*; no entry point synchronises the host;
 *   - return value 0 = ok; non-zero = error, message via srk_last_error() (thread local);
 *   - re-entrant: may be called concurrently from the autograd thread and the main thread.
 *
 * Tensor layouts
 *   SRK_LAYOUT_IMAGE : NCHW fp32, contiguous — what the reference modules take and return.
 *   SRK_LAYOUT_ACT   : zero-bordered channels-last [N][H+2][W+2][C], dtype fp32 or bf16 — the
 *                      internal activation layout.  The 1-pixel border is always zero so a 3x3
 *                      tap is a constant offset in the flat pixel index.
 */
#ifndef SRK_H_
#define SRK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRK_F32 0
#define SRK_BF16 1

#define SRK_LAYOUT_IMAGE 0
#define SRK_LAYOUT_ACT 1

#define SRK_ACT_NONE 0
#define SRK_ACT_RELU 1
#define SRK_ACT_PRELU 2

#define SRK_IMPL_AUTO 0
#define SRK_IMPL_SIMT 1 /* fp32-accurate CUDA-core implicit GEMM */
#define SRK_IMPL_TC 2   /* tcgen05/TMEM/TMA implicit GEMM (bf16 in, fp32 accumulate) */

/* weight pack kinds (srk_weight_pack) */
#define SRK_PACK_FPROP_SIMT 0 /* fp32 [R][S][Cin][Cout] */
#define SRK_PACK_DGRAD_SIMT 1 /* fp32 [R][S][Cout][Cin], taps rotated by 180 degrees */
#define SRK_PACK_FPROP_TC 2   /* bf16 [R*S][Cout'][Cin]  (Cout' permuted when pixel_shuffle) */
#define SRK_PACK_DGRAD_TC 3   /* bf16 [R*S rot180][Cin][Cout'] */
#define SRK_PACK_FPROP_TC_N8 4 /* bf16 [R][NP][Cin], row n = s*3 + co, NP = 32 / 16: RGB-output convs on tcgen05 */
#define SRK_PACK_RGBIN_TC 5    /* bf16 [64][KP], k = r*SEG + s*3 + c (SEG = 3S rounded up to even), zero padded: 3 -> 64 conv forward */
#define SRK_PACK_RGBOUT_DGRAD_TC 6 /* bf16 [64 ci][KP], k = r'*SEG + s'*3 + co, taps rotated: 64 -> 3 conv dgrad */

typedef struct srk_tensor {
  void* data;
  int32_t layout; /* SRK_LAYOUT_* */
  int32_t dtype;  /* SRK_F32 / SRK_BF16 (IMAGE is always fp32) */
  int32_t n, c, h, w; /* logical (un-padded) sizes */
} srk_tensor;

const char* srk_last_error(void);
int srk_version(void);
/* 1 if the tcgen05 path can take this conv shape (used by the host to pick pack kinds). */
int srk_conv_tc_supported(int cin, int cout, int r, int s, int dtype, int pixel_shuffle);

/* ---- convolution family: F.conv2d / convolution_backward -------------------------------------
 * replaces nn.Conv2d forward at models.py:46,49,65,67,84-86,107,113,117,120,125,150,156,159,162,167
 * and its autograd backward (dgrad = fprop with SRK_PACK_DGRAD_* weights).
 * Fused epilogue: +bias, ReLU (models.py:99-100) / single-alpha PReLU (models.py:108,119,122,151,161,164),
 * +residual (models.py:185), PixelShuffle(2) store remap (models.py:118,121,160,163).
 * stride 1, "same" padding (pad = R/2), as every conv on the path.
 * `y` holds the output geometry; with pixel_shuffle=2, y is [N][Cout/4][2H][2W].
 * bn_sums (fp32 [2][Cout], WRITTEN, or NULL): per-channel sum and sum of squares of the conv output over interior
 * pixels - the statistics native_batch_norm needs (models.py:47,50,114) - produced by the conv epilogue from the
 * fp32 accumulators on the tcgen05 path, summed in a fixed order (needs reduce_ws, see srk_reduce_workspace_bytes).
 * bn_acc (or NULL; excludes bn_sums): the same statistics delivered into an ACCUMULATOR ("exact sums" below) that
 * srk_bn_apply_train consumes - the conv then ends without the serial tail of the ordered fold.  Returns 2, and
 * launches nothing, when the conv is not the single-pass 3x3 64 -> 64 bf16 ACT conv that path covers.
 * prelu_z (or NULL; act = PReLU, ACT outputs): a tensor of y's geometry / dtype that receives the PRE-activation
 * when - and only when - the slope is <= 0 (decided on the device).  nn.PReLU places no constraint on its slope
 * (models.py:48,66,108,119,122); for a slope <= 0 the backward cannot recover sign(z) / z from the output and reads
 * this copy instead (srk_act_bwd, srk_conv_rgbout_bwd_unshuffle).  Contents are undefined while the slope is > 0.
 */
int srk_conv_fprop(const srk_tensor* x, const srk_tensor* y, const void* w_packed, int pack_kind,
                   int cout, int r, int s, const float* bias, int act, const float* alpha,
                   const srk_tensor* residual, int pixel_shuffle, int impl, float* bn_sums,
                   void* reduce_ws, void* bn_acc, const srk_tensor* prelu_z, void* workspace, void* stream);

/* ---- deterministic reductions -------------------------------------------------------------------
 * No kernel of the training step accumulates floating-point values with atomics: every cross-block sum (BatchNorm
 * statistics and their backward reductions, PReLU-slope and SE gradients, weight / bias gradients, loss values) is
 * per-block partial rows + ONE fold in block-index order, so a step is bit-reproducible run to run.  The entry
 * points that reduce take `reduce_ws`: a caller-owned buffer of srk_reduce_workspace_bytes() bytes, zero-filled ONCE
 * when it is allocated (its tickets reset themselves), and not shared between streams whose kernels may run
 * concurrently (one per stream). */
int64_t srk_reduce_workspace_bytes(void);
/* Exact sums ("accumulators").  The ordered fold ends a kernel with a chain of dependent global-memory steps (5 us on
 * the 23 us trunk conv).  The trunk convs can instead add their per-CTA partial sums into an accumulator: each fp32
 * partial is split EXACTLY into radix-2^40 integer digits that are added with 64-bit integer reductions - associative,
 * so the total is bit-reproducible whatever the arrival order, and exact; non-finite partials (and partials of
 * 2^120 and more, which could overflow a counter) make the value read as NaN.  An accumulator is srk_acc_bytes() bytes of device memory, zero-filled ONCE by the caller; after that a
 * producer (srk_conv_fprop bn_acc, srk_conv_dgrad_bnred acc) and a consumer (srk_bn_apply_train acc,
 * srk_bn_bwd_apply_raw acc, srk_acc_read) must alternate on it, in stream order: the consumer converts the digits and
 * leaves the accumulator zero-filled again.  Values: [sum C | sum of squares C] or [sum g C | sum g*z C | dalpha], C = 64. */
int64_t srk_acc_bytes(void);
/* generic consumer: out[i] = value i for i < nv (fp32), accumulator reset */
int srk_acc_read(void* acc, int nv, float* out, void* stream);
/* bytes of `workspace` srk_conv_fprop needs for this input and pack kind (0 = may pass NULL): the tcgen05 path
 * carries the fp32 partial sums of a contraction over more than 64 input channels through it. */
int64_t srk_conv_fprop_workspace_bytes(const srk_tensor* x, int pack_kind);

/* dW (fp32, OIHW) and db (fp32 [Cout], may be NULL): added into dw / db when accumulate != 0, written otherwise.
 * x: conv input, dy: gradient w.r.t. the conv output (pre-activation, conv-output geometry).
 * workspace: srk_conv_wgrad_workspace_bytes() bytes (may be NULL when that returns 0).
 * perm_shuffle != 0 (tcgen05 path): dy's channels are sub-pixel-major (sub * Cout/4 + c, what
 * srk_conv_rgbout_bwd_unshuffle and srk_act_bwd(perm_tc = 1) produce for a PixelShuffle conv); dw / db come out in the
 * reference channel order 4c + sub all the same. */
int srk_conv_wgrad(const srk_tensor* x, const srk_tensor* dy, float* dw, float* db, int r, int s,
                   int impl, int accumulate, int perm_shuffle, void* workspace, void* stream);
int64_t srk_conv_wgrad_workspace_bytes(const srk_tensor* x, const srk_tensor* dy, int r, int s, int impl);

/* dgrad of a 3x3 64 -> 64 conv (bf16 ACT, w_packed_dgrad = SRK_PACK_DGRAD_TC weights) with the BatchNorm-backward
 * REDUCTION of the BN layer below it fused into the epilogue (native_batch_norm_backward's sums of
 * ResidualBlock.backward, models.py:56-57): z = that BN's saved input, alpha = the PReLU slope between the BN and
 * this conv (models.py:57) or NULL.  Writes sum_g[64], sum_gz[64] (raw sums; srk_bn_bwd_apply_raw turns them into
 * dgamma / the dy constants) and dalpha[1].  residual (or NULL): the skip connection's gradient, added to the
 * dgrad BEFORE the reduction - dx = dgrad(dz) + residual is then the whole gradient of a ResidualBlock's input
 * (models.py:60) and z the bn2 input of the block below it.  acc (or NULL): an accumulator that receives the three
 * sums instead of sum_g / sum_gz / dalpha (which may then be NULL, like reduce_ws); srk_bn_bwd_apply_raw consumes
 * it.  Returns 0 ok, 1 error, 2 = shape outside the fused
 * kernel (nothing launched; run srk_conv_fprop + srk_bn_bwd_reduce instead). */
int srk_conv_dgrad_bnred(const srk_tensor* dz, const srk_tensor* dx, const void* w_packed_dgrad, const srk_tensor* z,
                         const float* mean, const float* invstd, const float* gamma, const float* beta,
                         const float* alpha, float* sum_g, float* sum_gz, float* dalpha, const srk_tensor* residual,
                         void* reduce_ws, void* acc, void* stream);

/* ---- convolutions with an RGB side on tcgen05 (K = 9 or 5; im2col built in shared memory) ---------
 * input_conv / SRCNN conv1 (3 -> 64, models.py:84,107,150) and the backward of output_conv / SRCNN conv3
 * (64 -> 3, models.py:86,125,167).  img3: IMAGE fp32 [N,3,H,W]; y, t64, dx: bf16 ACT [N,64,H,W].
 *   fprop:  y = act(conv(img3, W) + bias), W packed SRK_PACK_RGBIN_TC; prelu_z as in srk_conv_fprop.
 *   bwd, rgb_out = 0 (3 -> 64 conv): img3 = the conv input, t64 = dZ; dw [64][3][K][K], db [64] WRITTEN.
 *   bwd, rgb_out = 1 (64 -> 3 conv): img3 = dY, t64 = the conv input; dw [3][64][K][K], db [3] WRITTEN;
 *        when dx != NULL also dx = dgrad(dY) with w_packed = SRK_PACK_RGBOUT_DGRAD_TC weights.
 * workspace: srk_conv_rgb_workspace_bytes(k) bytes (one partial gradient per CTA, summed in CTA order). */
int64_t srk_conv_rgb_workspace_bytes(int k);
int srk_conv_rgb_fprop(const srk_tensor* img3, const srk_tensor* y, const void* w_packed, int k,
                       const float* bias, int act, const float* alpha, const srk_tensor* prelu_z, void* stream);
int srk_conv_rgb_bwd(const srk_tensor* img3, const srk_tensor* t64, const void* w_packed,
                     const srk_tensor* dx, float* dw, float* db, int k, int rgb_out, void* workspace,
                     void* stream);
/* srk_conv_rgb_bwd (rgb_out = 1) fused with the PReLU + PixelShuffle(2) backward of the upsample stage below the output
 * conv (autograd of models.py:120-125): t64 = that stage's output, t64_z = its saved pre-activation (prelu_z of the
 * forward call, or NULL), alpha / dalpha = its PReLU slope and slope gradient (written).  Instead of dx it writes dz_ps = the gradient of the 64 -> 256 conv output, bf16 ACT [N,256,H/2,W/2],
 * channels SUB-PIXEL-MAJOR (sub * 64 + c).  Consumers: srk_conv_fprop with SRK_PACK_DGRAD_TC weights packed with
 * pixel_shuffle = 2, and srk_conv_wgrad with perm_shuffle = 1. */
int srk_conv_rgbout_bwd_unshuffle(const srk_tensor* dy_img, const srk_tensor* t64, const srk_tensor* t64_z,
                                  const void* w_packed, const srk_tensor* dz_ps, float* dw, float* db, const float* alpha,
                                  float* dalpha, int k, void* workspace, void* stream);

/* OIHW fp32 master weights -> kernel operand layouts (see SRK_PACK_*). */
int srk_weight_pack(const float* w_oihw, void* out, int cout, int cin, int r, int s, int kind,
                    int pixel_shuffle, void* stream);
int64_t srk_weight_pack_bytes(int cout, int cin, int r, int s, int kind);
/* `count` packs (square kernels, pixel_shuffle = 0) in ceil(count / 64) launches; host arrays of device pointers */
int srk_weight_pack_multi(int count, const float* const* w_oihw, void* const* out, const int32_t* cout,
                          const int32_t* cin, const int32_t* r, const int32_t* kind, const int32_t* pixel_shuffle,
                          void* stream);

/* ---- activation backward: _prelu_kernel_backward / threshold_backward (+ pixel_unshuffle) ----
 * out: saved post-activation tensor; dout: its gradient; dz: gradient of the pre-activation in
 * conv-output geometry ([N][4C][H/2][W/2] when pixel_unshuffle=2; channel order matches `perm_tc`:
 * 0 = reference order co=4c+2i+j, 1 = sub-pixel-major co'=(2i+j)*C+c used by the TC conv path).
 * dalpha (fp32[1], written; needs reduce_ws) only for PReLU.  For a slope > 0 the pre-activation follows from `out`
 * (sign(out) = sign(z), z = out / alpha on the negative side); for a slope <= 0 the kernel reads `zsave`, the
 * prelu_z copy the forward epilogue wrote (NULL: the slope must be > 0). */
int srk_act_bwd(const srk_tensor* dout, const srk_tensor* out, const srk_tensor* zsave, const srk_tensor* dz, int act,
                const float* alpha, float* dalpha, int pixel_unshuffle, int perm_tc, void* reduce_ws, void* stream);

/* ---- BatchNorm2d (models.py:47,50,56-57,114,140): native_batch_norm / _backward ----------------*/
/* per-channel sum and sum of squares over interior pixels: sums = fp32 [2][C], written */
int srk_bn_stats(const srk_tensor* y, float* sums, void* reduce_ws, void* stream);
/* training: batch mean / invstd from (sum,sumsq); updates running stats (momentum, unbiased var)
 * and num_batches_tracked (int64) when those pointers are non-NULL. */
int srk_bn_finalize(const float* sum, const float* sumsq, int c, int64_t count, float eps,
                    float momentum, float* running_mean, float* running_var,
                    int64_t* num_batches_tracked, float* mean, float* invstd, void* stream);
/* eval: mean = running_mean, invstd = rsqrt(running_var + eps) */
int srk_bn_eval_params(const float* running_mean, const float* running_var, int c, float eps,
                       float* mean, float* invstd, void* stream);
/* out = [PReLU_alpha](gamma*(y-mean)*invstd+beta) [+ residual]   (alpha/residual may be NULL) */
int srk_bn_apply(const srk_tensor* y, const float* mean, const float* invstd, const float* gamma,
                 const float* beta, const float* alpha, const srk_tensor* residual,
                 const srk_tensor* out, void* stream);
/* training-mode BatchNorm in one pass after the statistics: srk_bn_finalize folded into srk_bn_apply (every block
 * derives mean / invstd from (sum, sumsq) itself; block 0 publishes them to mean / invstd for the backward and updates
 * the running statistics and num_batches_tracked when those pointers are non-NULL).  F.batch_norm(training=True) of
 * models.py:56-57,140.  acc (or NULL): the statistics as an accumulator (consumed) instead of sum / sumsq. */
int srk_bn_apply_train(const srk_tensor* y, const float* sum, const float* sumsq, int64_t count, float eps,
                       float momentum, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                       float* mean, float* invstd, const float* gamma, const float* beta, const float* alpha,
                       const srk_tensor* residual, const srk_tensor* out, void* acc, void* stream);
/* backward pass 1: dgamma[C], dbeta[C], dalpha[1] (all fp32, written) */
int srk_bn_bwd_reduce(const srk_tensor* dout, const srk_tensor* y, const float* mean,
                      const float* invstd, const float* gamma, const float* beta,
                      const float* alpha, float* dgamma, float* dbeta, float* dalpha, void* reduce_ws,
                      void* stream);
/* backward pass 2: dy.  dgamma_b/dbeta_b are THIS batch's reductions (pass-1 outputs);
 * batch_stats=0 gives the eval-mode backward (dy = gamma*invstd*g). */
int srk_bn_bwd_apply(const srk_tensor* dout, const srk_tensor* y, const float* mean,
                     const float* invstd, const float* gamma, const float* beta, const float* alpha,
                     const float* dgamma_b, const float* dbeta_b, int batch_stats,
                     const srk_tensor* dy, void* stream);
/* as srk_bn_bwd_apply, fed with the raw sums of srk_conv_dgrad_bnred: sum_g = sum g (= dbeta), sum_gz = sum g * z;
 * dgamma_out[C] receives invstd * (sum_gz - mean * sum_g).  acc (or NULL): the raw sums as an accumulator (consumed)
 * instead of sum_g / sum_gz; dbeta_out[C] (= sum g) and dalpha_out[1] (may be NULL) are then written as well. */
int srk_bn_bwd_apply_raw(const srk_tensor* dout, const srk_tensor* y, const float* mean, const float* invstd,
                         const float* gamma, const float* beta, const float* alpha, const float* sum_g,
                         const float* sum_gz, int batch_stats, float* dgamma_out, const srk_tensor* dy, void* acc,
                         float* dbeta_out, float* dalpha_out, void* stream);

/* ---- squeeze-excite gate (models.py:26-41,76-78): mean / mm / sigmoid / mul / add ------------- */
/* pool[N][C] = mean over H,W of r */
int srk_se_pool(const srk_tensor* r, float* pool, void* reduce_ws, void* stream);
/* hidden[N][Cr] = relu(pool @ w1^T), gate[N][C] = sigmoid(hidden @ w2^T); w1 [Cr][C], w2 [C][Cr] */
int srk_se_fc(const float* pool, const float* w1, const float* w2, int n, int c, int cr,
              float* hidden, float* gate, void* stream);
/* out = (x ? x : 0) + scale * r * gate[n][c] */
int srk_se_apply(const srk_tensor* x, const srk_tensor* r, const float* gate, float scale,
                 const srk_tensor* out, void* stream);
/* dgate_raw[N][C] = sum_hw dout * r   (written) */
int srk_se_bwd_reduce(const srk_tensor* dout, const srk_tensor* r, float* dgate_raw, void* reduce_ws, void* stream);
/* tiny FC backward: dw1 [Cr][C], dw2 [C][Cr], dpool [N][C] (all written) */
int srk_se_fc_bwd(const float* dgate_raw, const float* gate, const float* hidden, const float* pool,
                  const float* w1, const float* w2, int n, int c, int cr, float scale, float* dw1,
                  float* dw2, float* dpool, void* stream);
/* dr = scale * gate * dout + dpool / (H*W) */
int srk_se_bwd_apply(const srk_tensor* dout, const float* gate, const float* dpool, float scale,
                     const srk_tensor* dr, void* stream);

/* ---- layout / elementwise helpers --------------------------------------------------------------*/
int srk_image_to_act(const srk_tensor* img, const srk_tensor* act, void* stream);
int srk_act_to_image(const srk_tensor* act, const srk_tensor* img, void* stream);
/* out = a + b on ACT tensors (residual sums models.py:60,141) ; out may alias a */
int srk_act_add(const srk_tensor* a, const srk_tensor* b, const srk_tensor* out, void* stream);

/* 2x2 / stride 2 max pooling, floor mode, on ACT tensors: nn.MaxPool2d(2, 2) of torchvision VGG19.features, which
 * PerceptualLoss runs both images through (reference loss.py:23-28).  out: [N, C, H/2, W/2].  bwd: dx (same geometry
 * as x, fully written) receives dout at the first maximum of each window in row-major order (ATen's index rule). */
int srk_maxpool2_fwd(const srk_tensor* x, const srk_tensor* out, void* stream);
int srk_maxpool2_bwd(const srk_tensor* x, const srk_tensor* dout, const srk_tensor* dx, void* stream);
/* F.interpolate(mode='bicubic', align_corners=False) models.py:98 (A=-0.75, clamped taps) */
int srk_bicubic_upsample(const srk_tensor* in, const srk_tensor* out, void* stream);

/* ---- losses (loss.py:81-86 nn.L1Loss / nn.MSELoss; loss.py:31-79 NLPDLoss) --------------------- */
/* mode 0 = L1, 1 = MSE.  loss[1] fp32 overwritten.  scratch: srk_pixel_loss_scratch_bytes() bytes (per-block fp64
 * partial sums, added in block order). */
int64_t srk_pixel_loss_scratch_bytes(void);
int srk_pixel_loss_fwd(const float* sr, const float* hr, int64_t numel, int mode, float* loss,
                       double* scratch, void* stream);
/* grad_sr = gout[0] * dLoss/dsr */
int srk_pixel_loss_bwd(const float* sr, const float* hr, int64_t numel, int mode, const float* gout,
                       float* grad_sr, void* stream);
int64_t srk_nlpd_workspace_bytes(int n, int c, int h, int w, int levels);
/* loss = alpha*L1 + (1-alpha)*sum_l mean|Lap_l(sr)-Lap_l(hr)|; kernel25 = the 5x5 blur taps (device).
 * clamp01 != 0 clamps both inputs to [0,1] first (metrics.py:16-17; forward only).
 * The workspace keeps the pyramid for the backward. */
int srk_nlpd_fwd(const float* sr, const float* hr, int n, int c, int h, int w, int levels,
                 float alpha, const float* kernel25, int clamp01, void* workspace, float* loss,
                 void* stream);
int srk_nlpd_bwd(int n, int c, int h, int w, int levels, float alpha, const float* kernel25,
                 void* workspace, const float* gout, float* grad_sr, void* stream);

/* ---- metrics (metrics.py:14-31 via torchmetrics 1.8.2 PSNR / SSIM) ----------------------------- */
/* per-image sum of squared error of clamp(sr,0,1) vs clamp(hr,0,1) (clamp01 != 0): sse[N] double */
int srk_psnr_sse(const float* sr, const float* hr, int n, int64_t per_image, int clamp01,
                 double* sse, void* stream);
/* per-image SSIM sums over the (H-10)x(W-10) valid windows x C: ssim_sum[N] double (overwritten) */
int srk_ssim(const float* sr, const float* hr, int n, int c, int h, int w, int clamp01,
             double* ssim_sum, void* stream);

/* ---- optimizer (train.py:55,120 optim.Adam(betas=(0.5,0.999))) --------------------------------- */
/* single flat fp32 buffer Adam step (torch.optim.Adam semantics, no amsgrad/weight decay);
 * step_count is the 1-based step number held on the device (int64[1]) so the call is graph-safe. */
int srk_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                  float lr, float beta1, float beta2, float eps, const int64_t* step_count,
                  float grad_scale, void* stream);

/* the same update for `count` tensors in ceil(count / 48) launches (host arrays of device pointers).
 * lr_dev (device float[1] or NULL): when set it REPLACES lr - a captured CUDA graph then follows a scheduler
 * (train.py:56,164 ReduceLROnPlateau) without being re-captured.  grad_scale_dev (device float[1] or NULL): extra
 * factor on the gradients, e.g. the clip coefficient of clip_grad_norm_ (train.py:113) computed on the device. */
int srk_adam_multi(int count, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* numel, float lr, float beta1, float beta2,
                   float eps, const int64_t* step_count, float grad_scale, const float* lr_dev,
                   const float* grad_scale_dev, void* stream);

/* ---- training-sample pipeline (dataset.py:14-41: RandomCrop / CenterCrop + RandomHorizontalFlip + ToTensor, then
 * transforms.Resize((crop/s, crop/s), BICUBIC) on the HR tensor = antialiased bicubic, ATen _upsample_bicubic2d_aa) ----
 * src_u8: decoded images, uint8, [N][Hs][Ws][3] (hwc != 0) or [N][3][Hs][Ws]; offsets int32 [N][2] = (top, left) of
 * each crop, flips uint8 [N] (drawn by the host with torch's generator, as torchvision draws them);
 * hr: fp32 [N][3][crop][crop] = crop (flipped) / 255; lr: fp32 [N][3][crop/scale][crop/scale], not clamped.
 * scale in {2, 3, 4}, crop % scale == 0. */
int srk_sr_make_batch(const void* src_u8, int hwc, int n, int hs, int ws, const int32_t* offsets, const uint8_t* flips,
                      int crop, int scale, float* hr, float* lr, void* stream);

/* ---- bring-up / self-test hooks (tests only) --------------------------------------------------- */
/* Runs the tcgen05 descriptor probe (see csrc/srk_probe.cu); results into out[] (host memory). */
int srk_tc_probe(int variant, float* out_host, int out_len);

#ifdef __cplusplus
}
#endif
#endif /* SRK_H_ */

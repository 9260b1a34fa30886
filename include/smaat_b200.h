/*
 * smaat_b200.h -- C ABI of libsmaat_b200.so: the B200 (sm_100a) kernels behind the
 * SmaAt-UNet hot path (depthwise-separable conv blocks + CBAM + their glue): forward for
 * inference, train-mode forward + backward, and the training step's loss/metric pass.
 *
 * The reference (HansBambel/SmaAt-UNet) has no FFI / plugin registry: its boundary is
 * the Python nn.Module interface.  Each entry point below
 * replaces the torch.nn call(s) cited next to it; `smaat_unet_b200/modules.py` is the
 * host-side mirror of the reference classes that binds them through ctypes, and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - all tensors are fp32, NCHW, dense in (H, W); device pointers owned by the caller
 *     (PyTorch caching allocator).  The library allocates nothing persistent.
 *   - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*):
 *     no device synchronisation, no allocation -> CUDA-graph capturable
 *     (the one exception is the debug hook smaat_debug_dsconv_timing).
 *   - return value: 0 on success, negative SMAAT_E_* otherwise; smaat_last_error()
 *     returns a thread-local description.  Nothing throws or exits across the ABI.
 *   - "bstride" arguments are batch strides in ELEMENTS (>= C*H*W) so a kernel can
 *     read from / write into a channel slice of a wider tensor without a copy.
 */
#ifndef SMAAT_B200_H_
#define SMAAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMAAT_ABI_VERSION 1

#define SMAAT_OK 0
#define SMAAT_E_BADARG (-1)   /* shape / pointer / alignment rejected by host-side validation */
#define SMAAT_E_CUDA (-2)     /* CUDA runtime or driver error at launch (see smaat_last_error) */
#define SMAAT_E_UNSUPPORTED (-3) /* valid request this build has no kernel for */

/* pointwise arithmetic modes (smaat_pw1x1_fwd) */
#define SMAAT_PW_FP32_SIMT 0  /* CUDA-core FFMA, exact fp32 products                      */
#define SMAAT_PW_TF32 1       /* tcgen05 kind::tf32, fp32 accumulate in TMEM (1 pass)     */
#define SMAAT_PW_TF32X3 2     /* tcgen05 3xTF32 split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo) */

int smaat_abi_version(void);
const char* smaat_last_error(void);
/* Number of kernel launches enqueued by this library in the calling process so far. */
uint64_t smaat_launch_count(void);

/* ---- depthwise 3x3, padding 1, groups=Cin, k outputs per input channel -------------
 * replaces DepthwiseSeparableConv.depthwise  (models/layers.py:38-44,48)
 * Input is the virtual channel-concat of x0 (C0 channels) and x1 (C1 channels, may be
 * NULL/0): this is torch.cat([x2, x1], dim=1) of UpDS.forward
 * (models/unet_parts_depthwise_separable.py:85) without materialising the concat.
 * y[b, o] = bias[o] + sum_{dy,dx} w[o,dy,dx] * in[b, o / k, i+dy-1, j+dx-1]
 *   w: (k*(C0+C1), 3, 3)   bias: (k*(C0+C1)) or NULL   y: (B, k*(C0+C1), H, W) dense
 * in_scale/in_shift (per INPUT channel, may be NULL): when given the kernel applies
 *   relu(in_scale[c] * x + in_shift[c]) on load -- the train-mode BatchNorm+ReLU of the
 *   producing layer (parts_ds.py:25-26), with zero padding applied AFTER the activation.
 * loader: 0 = auto, 1 = force the LDG loader, 2 = force TMA (error if ineligible). */
int smaat_dw3x3_fwd(const float* x0, int C0, int64_t x0_bstride,
                    const float* x1, int C1, int64_t x1_bstride,
                    const float* w, const float* bias,
                    const float* in_scale, const float* in_shift,
                    float* y, int B, int H, int W, int k, int loader, void* stream);

/* ---- pointwise 1x1 + per-channel affine (+ReLU) epilogue -----------------------------
 * replaces DepthwiseSeparableConv.pointwise (models/layers.py:45,49) fused with the
 * following nn.BatchNorm2d (eval) + nn.ReLU (parts_ds.py:25-26,34-35):
 *   y[b,o,p] = act( scale[o] * sum_c w[o,c] x[b,c,p] + shift[o] )
 *   x: (B, K, P) dense   w: (Cout, K)   scale/shift: (Cout) (scale NULL = 1, shift NULL = 0)
 *   y: (B, Cout, P) with batch stride y_bstride elements.
 * w_lo: (Cout, K) low parts for SMAAT_PW_TF32X3 (w must then hold the tf32-truncated
 *   high parts, see smaat_split_tf32); NULL otherwise.
 * stats: NULL, or (2*Cout) fp64 zero-initialised accumulators receiving per-channel
 *   sum and sum of squares of the PRE-activation value scale*acc+shift (train-mode BN). */
int smaat_pw1x1_fwd(const float* x, const float* w, const float* w_lo,
                    const float* scale, const float* shift,
                    float* y, int64_t y_bstride, double* stats,
                    int B, int K, int Cout, int P, int relu, int mode, void* stream);

/* ---- fused DepthwiseSeparableConv: depthwise 3x3 -> pointwise 1x1 -> affine (+ReLU) in ONE kernel ----
 * replaces DepthwiseSeparableConv.forward (models/layers.py:47-50) + eval BatchNorm2d + ReLU
 * (parts_ds.py:25-26,34-35); the k*Cin-channel depthwise result stays on chip (CUDA-core stencil writes
 * the tcgen05 A operand directly).  Arguments as smaat_dw3x3_fwd (input = virtual concat [x0, x1],
 * dw_w: (k*Cin,3,3), dw_b: (k*Cin) or NULL) and smaat_pw1x1_fwd (pw_w: (Cout, k*Cin) -- the tf32 hi
 * parts in TF32X3 mode, pw_w_lo the lo parts; scale/shift/stats/relu as there).
 * mode: SMAAT_PW_TF32 or SMAAT_PW_TF32X3.  Returns SMAAT_E_UNSUPPORTED for shapes the fused kernel
 * does not take (k not in {1,2}, Cout > 128, W % 4, patch waste > 35 %): callers then run
 * smaat_dw3x3_fwd + smaat_pw1x1_fwd.  smaat_dsconv_eligible returns 1/0 for the same test. */
int smaat_dsconv_eligible(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                          const float* pw_w, int H, int W, int k, int Cout);
/* Same test with the batch-statistics request made explicit: `with_stats` != 0 asks for the kernel that accumulates the
 * per-channel sum / sum of squares of its output (train-mode BatchNorm, parts_ds.py:25,34) -- the shared-memory-operand kernel. */
int smaat_dsconv_eligible2(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                           const float* pw_w, int H, int W, int k, int Cout, int with_stats);
int smaat_dsconv_fwd(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                     const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_w_lo,
                     const float* scale, const float* shift, float* y, int64_t y_bstride, double* stats,
                     int B, int H, int W, int k, int Cout, int relu, int mode, void* stream);
/* The network's last two modules in one kernel: the fused DS conv above followed by OutConv(Cout -> 1 class)
 * (models/SmaAt_UNet.py:55-56, unet_parts.py:67-73).  oc_w: (Cout), oc_b: (1) or NULL, logits: (B, 1, H, W); the
 * Cout-channel activation is reduced in the epilogue registers (TMEM lane = pixel) and never written.  Same eligibility
 * as smaat_dsconv_fwd. */
int smaat_dsconv_outconv_fwd(const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                             const float* dw_w, const float* dw_b, const float* pw_w, const float* pw_w_lo,
                             const float* scale, const float* shift, const float* oc_w, const float* oc_b, float* logits,
                             int B, int H, int W, int k, int Cout, int relu, int mode, void* stream);

/* Which kernel smaat_dsconv_fwd / smaat_dsconv_outconv_fwd run (same reference lines, models/layers.py:47-50): 0 = auto
 * (default: the TMEM-operand kernel, csrc/dsconv_tmem.cu, where it applies -- k = 2, Cout <= 128, no batch statistics --
 * else the shared-memory-operand kernel, csrc/dsconv_fused.cu), 1 = shared-memory-operand kernel only, 2 = TMEM-operand
 * kernel only.  Process-wide; the environment variable SMAAT_DS_IMPL presets it.  For A/B measurements and tests. */
int smaat_set_dsconv_impl(int impl);

/* Debug hook: stage timers of the fused kernel's CTA 0 (16 clock64 counters accumulated over launches; layout in
 * csrc/dsconv_fused.cu).  Copies them to the HOST array `out` and clears them; synchronises the device. */
int smaat_debug_dsconv_timing(unsigned long long* out);
/* Same for the TMEM-operand kernel (24 counters; layout in csrc/dsconv_tmem.cu). */
int smaat_debug_dsconv_tmem_timing(unsigned long long* out);
/* Per-CTA (start ns, end ns, SM id) of the last timed launch of the TMEM-operand kernel: 3 * n_ctas entries, n_ctas <= 256. */
int smaat_debug_dsconv_tmem_cta_timing(unsigned long long* out, int n_ctas);
/* Event trace of CTA 0 (launch made with SMAAT_DSCONV_TIMING=2): 16 clock64 stamps per unit, n_units <= 256 (layout in the .cu). */
int smaat_debug_dsconv_tmem_trace(long long* out, int n_units);

/* 1 if this (x, w, K, Cout, P) can take the tcgen05 path (P % 4 == 0, K % 4 == 0, 16-byte aligned
 * pointers, Cout >= 8), else 0: the caller then uses SMAAT_PW_FP32_SIMT. */
int smaat_pw1x1_tc_eligible(const float* x, const float* w, int K, int Cout, int P);

/* hi[i] = tf32_truncate(src[i]); lo[i] = src[i] - hi[i]   (weight preparation for TF32X3) */
int smaat_split_tf32(const float* src, float* hi, float* lo, int64_t n, void* stream);

/* ---- eval-mode BatchNorm folded to the affine the pw epilogue applies -----------------
 * nn.BatchNorm2d.eval (parts_ds.py:25,34):  scale = gamma / sqrt(rv + eps),
 *   shift = beta + (conv_bias - rm) * scale        (conv_bias may be NULL) */
int smaat_bn_fold(const float* gamma, const float* beta, const float* rm, const float* rv,
                  const float* conv_bias, float eps, float* scale, float* shift, int C, void* stream);

/* ---- train-mode BatchNorm2d (parts_ds.py:25,34; layers.py:120 in .train()) ----------------------
 * The producing kernel accumulates per-channel sum / sum of squares in fp64 (`stats`, 2*C doubles,
 * zero-initialised by the caller); smaat_channel_stats does the same for an existing tensor
 * x: (B, C, P) (SpatialAttention's 1-channel BN).
 * bn_finalize: mean = s1/n, var = s2/n - mean^2 (biased), scale = gamma/sqrt(var+eps),
 *   shift = beta - mean*scale; writes mean / inv-std for the backward pass (may be NULL) and updates
 *   running_mean <- (1-m) rm + m mean, running_var <- (1-m) rv + m var*n/(n-1)  (may be NULL).
 *   num_batches_tracked (nullable, device int64) is incremented by one, like nn.BatchNorm2d does in train mode.
 * affine_act: y[b,c,p] = act(scale[c]*x[b,c,p] + shift[c]), act 0 = none, 1 = ReLU, 2 = sigmoid
 *   (scale NULL = 1, shift NULL = 0). */
int smaat_channel_stats(const float* x, double* stats, int B, int C, int P, void* stream);
int smaat_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float eps, float momentum,
                      float* running_mean, float* running_var, float* scale, float* shift,
                      float* mean_out, float* invstd_out, long long* num_batches_tracked, int C, void* stream);
int smaat_affine_act_fwd(const float* x, const float* scale, const float* shift, float* y,
                         int B, int C, int P, int act, void* stream);

/* ---- nn.MaxPool2d(2) (parts_ds.py:48): stride 2, floor ---------------------------------
 * x: (N, H, W) planes -> y: (N, H/2, W/2) */
int smaat_maxpool2_fwd(const float* x, float* y, int64_t N, int H, int W, void* stream);

/* ---- nn.Upsample(x2, bilinear, align_corners=True) + F.pad to the skip size -----------
 * (parts_ds.py:64,78-81).  x: (B, C, H, W) -> y: (B, C, Ho, Wo) with Ho >= 2H, Wo >= 2W,
 * zero border of (Ho-2H)//2 rows on top, (Wo-2W)//2 columns on the left. */
int smaat_upsample2x_pad_fwd(const float* x, float* y, int64_t y_bstride,
                             int B, int C, int H, int W, int Ho, int Wo, void* stream);

/* ---- CBAM (models/layers.py:90-141) -----------------------------------------------------
 * pool:   avg[n] = mean_p x[n,p], mx[n] = max_p x[n,p] over N = B*C planes of P pixels
 *         (AdaptiveAvgPool2d(1)/AdaptiveMaxPool2d(1), layers.py:107-108)
 * mlp:    sc[b,c] = sigmoid( MLP(avg[b]) + MLP(mx[b]) ), MLP = W2 relu(W1 v + b1) + b2
 *         (layers.py:98-103,109)
 * reduce: pooled[b,0,p] = mean_c x[b,c,p]*sc[b,c]; pooled[b,1,p] = max_c ...  (layers.py:123-125)
 * gate:   a = conv_kxk(pooled, wsp)  (2->1 ch, pad k/2, no bias, layers.py:119,126);
 *         sa[b,p] = sigmoid(bn_affine[0] * a + bn_affine[1])  (BatchNorm2d(1) + sigmoid, :127-128);
 *         bn_affine = 2 floats ON THE DEVICE (smaat_bn_fold with C=1), NULL = identity;
 *         if raw != NULL the pre-BN conv output a is also written (train-mode statistics).
 * scale:  y[b,c,p] = x[b,c,p] * sc[b,c] * sa[b,p]   (layers.py:110,128) */
int smaat_cbam_pool_fwd(const float* x, float* avg, float* mx, int64_t N, int P, void* stream);
/* The same pools plus nn.MaxPool2d(2) of the same planes in ONE read of x (every encoder map of SmaAt-UNet feeds both
 * cbam_l and down_l: models/SmaAt_UNet.py:42-50).  x: (N, H, W) -> avg (N), mx (N), pooled (N, H/2, W/2).
 * SMAAT_E_UNSUPPORTED unless W % 4 == 0 and H % 2 == 0 (then run smaat_cbam_pool_fwd + smaat_maxpool2_fwd). */
int smaat_cbam_pool_maxpool_fwd(const float* x, float* avg, float* mx, float* pooled, int64_t N, int H, int W, void* stream);
int smaat_cbam_mlp_fwd(const float* avg, const float* mx, const float* w1, const float* b1,
                       const float* w2, const float* b2, float* sc, int B, int C, int hidden, void* stream);
int smaat_cbam_reduce_fwd(const float* x, const float* sc, float* pooled, int B, int C, int P, void* stream);
int smaat_cbam_gate_fwd(const float* pooled, const float* wsp, const float* bn_affine, float* sa, float* raw,
                        int B, int H, int W, int ks, void* stream);
int smaat_cbam_scale_fwd(const float* x, const float* sc, const float* sa, float* y, int64_t y_bstride,
                         int B, int C, int P, void* stream);

/* ---- OutConv (models/unet_parts.py:67-73): 1x1 conv Cin -> ncls (small), bias, no activation */
int smaat_outconv_fwd(const float* x, const float* w, const float* bias, float* y,
                      int B, int Cin, int ncls, int P, void* stream);

/* ======================================= backward (training step) =======================================
 * Gradients of the same path (BASELINE configs[2]): what torch autograd runs for these modules when the reference calls
 * loss.backward() (train_SmaAtUNet.py:55; Lightning automatic optimisation over UNetBase.training_step,
 * models/regression_lightning.py:67-78).  Each group names the forward lines it differentiates.  Notation: z = pre-BatchNorm activation,
 * a = act(scale*z+shift), dA = dL/da masked by the activation (act: 0 none, 1 ReLU recomputed from z).
 * Accumulating outputs (dW, db, dgamma, dbeta, d_sc, stats-like sums) are += : the caller zero-initialises.
 *
 * BatchNorm(+ReLU) backward -- nn.BatchNorm2d + nn.ReLU, parts_ds.py:25-26,34-35; BatchNorm2d(1), layers.py:120,127 --
 * = reduce -> coeffs -> apply:
 *   reduce: sums[c] += sum dA, sums[C+c] += sum dA*z                       (fp64)
 *   coeffs: dgamma += invstd*(S2 - mean*S1), dbeta += S1; per-channel a,b,cc with dz = a*dA + b*z + cc
 *           (train: batch-statistics terms; eval (train=0): dz = gamma*invstd*dA); dz_sum (nullable) += sum_{b,p} dz
 *           per channel, i.e. the bias gradient of the conv that produced z, from the sums alone (0 with batch statistics)
 *   apply : dz[b,c,p] = a[c]*dA + b[c]*z + cc[c] */
int smaat_bn_act_bwd_reduce(const float* dy, const float* z, const float* scale, const float* shift, double* sums,
                            int B, int C, int P, int act, void* stream);
int smaat_bn_bwd_coeffs(const double* sums, double count, const float* gamma, const float* mean, const float* invstd, int train,
                        float* a, float* b, float* cc, float* dgamma, float* dbeta, float* dz_sum, int C, void* stream);
int smaat_bn_act_bwd_apply(const float* dy, const float* z, const float* scale, const float* shift,
                           const float* a, const float* b, const float* cc, float* dz, int B, int C, int P, int act, void* stream);

/* depthwise 3x3 backward (DepthwiseSeparableConv.depthwise, models/layers.py:38-44,48): input gradient (split over the virtual concat x0|x1) and weight/bias gradient;
 * in_scale/in_shift: the forward's on-load BN+ReLU prologue (input of the conv was relu(in_scale*x+in_shift)). */
int smaat_dw3x3_bwd_input(const float* dd, const float* w, float* dx0, int C0, int64_t dx0_bstride,
                          float* dx1, int C1, int64_t dx1_bstride, int B, int H, int W, int k, void* stream);
int smaat_dw3x3_bwd_weight(const float* dd, const float* x0, int C0, int64_t x0_bstride, const float* x1, int C1, int64_t x1_bstride,
                           const float* in_scale, const float* in_shift, float* dw, float* db,
                           int B, int H, int W, int k, void* stream);

/* pointwise 1x1 backward (DepthwiseSeparableConv.pointwise, models/layers.py:45,49): dW[o][c] += sum_{b,p} dz[b,o,p]*d[b,c,p],
 * db[o] += sum dz.  (The input gradient is
 * smaat_pw1x1_fwd(dz, W^T): use smaat_transpose for W^T.) */
int smaat_pw1x1_bwd_weight(const float* dz, const float* d, float* dW, float* db, int B, int K, int Cout, int P, void* stream);
/* tensor-core (tcgen05, split over pixels + fp32 atomics) variant; mode SMAAT_PW_TF32 / SMAAT_PW_TF32X3;
 * SMAAT_E_UNSUPPORTED when P % 4 != 0 (use the CUDA-core smaat_pw1x1_bwd_weight). */
int smaat_pw1x1_bwd_weight_tc(const float* dz, const float* d, float* dW, float* db, int B, int K, int Cout, int P,
                              int mode, void* stream);
int smaat_transpose(const float* src, float* dst, int rows, int cols, void* stream);

/* glue backward: nn.MaxPool2d(2) (parts_ds.py:48), nn.Upsample(x2, bilinear, align_corners) + F.pad (parts_ds.py:64,78-81),
 * OutConv (unet_parts.py:67-73) */
int smaat_maxpool2_bwd(const float* x, const float* dy, float* dx, int64_t N, int H, int W, void* stream);
int smaat_upsample2x_pad_bwd(const float* dy, int64_t dy_bstride, float* dx, int B, int C, int H, int W, int Ho, int Wo, void* stream);
int smaat_outconv_bwd(const float* dy, const float* x, const float* w, float* dx, float* dW, float* db,
                      int B, int Cin, int ncls, int P, void* stream);

/* CBAM backward (models/layers.py:90-141 differentiated; see cbam_bwd.cu for the chain): gate_in -> [BN(1) backward] -> gate_bwd -> dsc -> mlp_bwd -> dx.
 * amax: (B, P) int32 channel argmax of x*sc written by gate_in; pkey: (B*C) uint64, zeroed by the caller, receives the
 * packed plane argmax of x in dsc and is consumed by dx; dsc (B, C) is accumulated into (caller zeroes it). */
int smaat_cbam_bwd_gate_in(const float* g, const float* x, const float* sc, const float* sa, float* dpre, int* amax,
                           int B, int C, int P, void* stream);
int smaat_cbam_gate_bwd(const float* draw, const float* pooled, const float* wsp, float* dpooled, float* dW,
                        int B, int H, int W, int ks, void* stream);
int smaat_cbam_bwd_dsc(const float* g, const float* x, const float* sa, const float* dpooled, const int* amax, float* dsc,
                       unsigned long long* pkey, int B, int C, int P, void* stream);
int smaat_cbam_mlp_bwd(const float* avg, const float* mx, const float* w1, const float* b1, const float* w2, const float* sc,
                       const float* dsc, float* dw1, float* db1, float* dw2, float* db2, float* davg, float* dmx,
                       int B, int C, int hidden, void* stream);
int smaat_cbam_bwd_dx(const float* g, const float* sc, const float* sa, const float* dpooled, const int* amax,
                      const float* davg, const float* dmx, const unsigned long long* pkey, float* dx,
                      int B, int C, int P, void* stream);

/* ---- loss + metric bookkeeping of one training/validation step, one pass, no host sync ------------
 * Replaces UNetBase.loss_func (models/regression_lightning.py:57-65) and PrecipitationMetrics.update
 * (metric/precipitation_metrics.py:37-95).  pred/target: n floats.  batch_acc: double[8], overwritten:
 *   [0] sum (p-y)^2  [1] sum (p*f-y*f)^2 (f = factor if denormalize else 1)  [2] number of NaNs in p or y
 *   [3] TN [4] FP [5] FN [6] TP of ((x*f)*12 > threshold)  [7] n
 * dpred (nullable): 2*(p-y)*grad_scale, the gradient of sum((p-y)^2)*grad_scale (grad_scale = 1/B).
 * smaat_metrics_commit adds one batch to totals (double[9]: total_loss, total_loss_denorm, total_samples,
 * total_pixels, TN, FP, FN, TP, skipped batches) on the device, skipping NaN batches like the reference (:46-48). */
int smaat_mse_metrics_fwd(const float* pred, const float* target, int64_t n, float factor, float threshold,
                          int denormalize, double* batch_acc, float* dpred, float grad_scale, void* stream);
int smaat_metrics_commit(const double* batch_acc, double* totals, int batch_size, int denormalize, void* stream);

/* CBAM in three launches (reference models/layers.py:90-141).
 * smaat_cbam_pool_mlp_fwd: ChannelAttention's global pools AND its shared MLP + sigmoid (layers.py:98-109): the last pooling
 *   CTA of each image finishes the MLP; pooled != NULL also emits MaxPool2d(2)(x) from the same read (parts_ds.py:48).
 *   counters: B ints, zero on entry and on exit.  C % 8 == 0, C <= 512, hidden <= 64 (else SMAAT_E_UNSUPPORTED).
 * smaat_cbam_reduce_fwd (above): per-pixel channel mean / max of x * sc (layers.py:123-125).
 * smaat_cbam_gate_scale_fwd: conv k x k (2 -> 1) + BatchNorm2d(1) affine + sigmoid AND y = (x * sc) * gate
 *   (layers.py:126-128, :110): the gate map never reaches HBM.  W % 4 == 0, 16-byte aligned x / y (else SMAAT_E_UNSUPPORTED). */
int smaat_cbam_pool_mlp_fwd(const float* x, float* avg, float* mx, float* pooled, const float* w1, const float* b1, const float* w2,
                            const float* b2, float* sc, int* counters, int B, int C, int H, int W, int hidden, void* stream);
int smaat_cbam_gate_scale_fwd(const float* pooled, const float* wsp, const float* bn_affine, const float* x, const float* sc, float* y,
                              int64_t y_bstride, int B, int C, int H, int W, int ks, void* stream);

/* ---- UpDS(bilinear=False): nn.ConvTranspose2d(in, in // 2, 2, stride=2) + F.pad (reference
 * models/unet_parts_depthwise_separable.py:72-73, 76-81).  Kernel = stride = 2: no overlapping taps, so the transposed conv is ONE
 * pointwise GEMM Cin -> 4 Cout packed channels (smaat_pw1x1_fwd on the repacked weight) followed by a 2x2 pixel shuffle.
 *   smaat_convt2x2_pack_weight:  W (Cin, Cout, 2, 2) -> Wp (4 Cout, Cin), row (2 dy + dx) Cout + o
 *   smaat_pixel_shuffle2_pad_fwd: t (B, 4 Cout, H, W) [+ bias (Cout) or NULL] -> y (B, Cout, Ho, Wo), zero pad frame
 *   smaat_pixel_shuffle2_pad_bwd: the gather transpose; smaat_convt2x2_unpack_wgrad ACCUMULATES dWp / dbp into dW / db. */
int smaat_convt2x2_pack_weight(const float* w, float* wp, int Cin, int Cout, void* stream);
int smaat_convt2x2_unpack_wgrad(const float* dwp, const float* dbp, float* dw, float* db, int Cin, int Cout, void* stream);
int smaat_pixel_shuffle2_pad_fwd(const float* t, const float* bias, float* y, int64_t y_bstride, int B, int Cout, int H, int W, int Ho,
                                 int Wo, void* stream);
int smaat_pixel_shuffle2_pad_bwd(const float* g, int64_t g_bstride, float* dt, int B, int Cout, int H, int W, int Ho, int Wo, void* stream);

/* ---- optimizer step (reference models/regression_lightning.py:47-48, train_SmaAtUNet.py:25: torch.optim.Adam with its
 * defaults) over flat fp32 buffers of n floats (n % 4 == 0, 16-byte aligned; parameters, gradients, first and second moment
 * share one layout; padding must carry zero gradients).  lr and step are DEVICE scalars (fp32; step = completed steps,
 * incremented by the call), so a CUDA graph holding this call follows learning-rate changes.  torch's arithmetic:
 *   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr / (1-b1^t) * m / (sqrt(v) / sqrt(1-b2^t) + eps). */
int smaat_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr,
                    float* step, double beta1, double beta2, double eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMAAT_B200_H_ */

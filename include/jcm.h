/* jcm.h - C ABI of libjcm.so: the B200 (sm_100a) kernels behind the joint-cnn-mrf hot path.
 *
 * The reference (max-andr/joint-cnn-mrf) has no FFI: its boundary is the Python function surface of main.py executing
 * TensorFlow-1.x ops.  Each entry point below replaces the TF op call sites named in its comment (file:line in the
 * reference).  The Python package `jcm` (joint-cnn-mrf_b200/jcm) binds these with ctypes and re-exposes the reference's
 * function names (model, spatial_model, conv_mrf, spatial_softmax, softmax_cross_entropy, ...); INTEGRATION.md shows
 * the binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (activations NHWC, fp32 unless stated); kernels never
 *     allocate; `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream.
 *   - "operand planes": bf16 NHWC tensors in pairs (hi, lo) with hi + lo == the fp32 value to ~2^-17.  Passing lo == NULL
 *     selects plain bf16 arithmetic (training config), passing lo selects the 3-term split product ("bf16x3", fp32 config).
 *   - return value: 0 = ok, < 0 = bad argument / unsupported shape (JCM_E*), > 0 = cudaError_t.  jcm_last_error()
 *     returns a thread-local message.  There is no CPU fallback anywhere.
 *   - devices: the design is one process per GPU (data parallelism = one replica per process, jcm.train.Trainer); calls act on the
 *     CUDA device that is current in the calling thread.  Per-device state (SM count, kernel attributes) is cached per device
 *     ordinal, so a process that switches devices stays correct; kernels of one call never span devices.
 *   - the library reads no environment variable (measurement switches exist only in builds with -DJCM_EXPERIMENTS).
 */
#ifndef JCM_H_
#define JCM_H_

#ifdef __cplusplus
extern "C" {
#endif

#define JCM_OK 0
#define JCM_EINVAL (-1)
#define JCM_ENOTSUP (-2)
#define JCM_EWORKSPACE (-3)

const char* jcm_last_error(void);
int jcm_version(void);
int jcm_sm_count(void);
long jcm_launch_count(void); /* kernels launched by this library since load */

/* ---- operand preparation --------------------------------------------------------------------------------------- */

/* x fp32 [B,H,W,3] -> space-to-depth bf16 planes [B,H/2,W/2,64], [B,H/4,W/4,64], [B,H/8,W/8,64] for the three banks: channel
 * dx*16 + (sy*2+sx)*3 + c holds x[2Y+sy, 2(X+dx-1)+sx, c] (the three horizontal taps of the 3x3 s2d kernel folded into channels).
 * Replaces tf.image.resize_images(x,[H/2,W/2]) / [H/4,W/4] (main.py:51,60) and prepares the stride-2 conv1_* (main.py:44,52,61),
 * which becomes a 3x1 stride-1 convolution (ksize 3, kw 1) over these planes. */
int jcm_prep_input(const float* x, int B, int H, int W, void* full_hi, void* full_lo, void* half_hi, void* half_lo,
                   void* quarter_hi, void* quarter_lo, void* stream);

/* conv kernel HWIO fp32 [k,k,Cin,Cout] (main.py:138-147 weight_variable layout) -> packed [k*k][Opad][Ipad] bf16 planes.
 * transpose = 0: forward operand (O = Cout, I = Cin); transpose = 1: data-gradient operand (taps flipped, O = Cin, I = Cout). */
int jcm_pack_weights(const float* w, int ksize, int Cin, int Cout, int Opad, int Ipad, int transpose, void* out_hi,
                     void* out_lo, void* stream);

/* Both layouts of up to 16 conv kernels in ONE launch (the re-pack after an optimizer step): layer i reads w [k,k,Cin,Cout] once and
 * writes fwd = [k*k][Cout][Cin] (as jcm_pack_weights transpose 0) and dgrad = [k*k][Cin][Cout] (as transpose 1).  Cin, Cout multiples
 * of 32 that need no padding (<= 256 or a multiple of 256).  lo pointers NULL: plain bf16.  layers: HOST array. */
typedef struct {
  const float* w;
  void *fwd_hi, *fwd_lo, *dgrad_hi, *dgrad_lo;
  int ksize, Cin, Cout;
} jcm_pack_desc;
int jcm_pack_weights_batch(const jcm_pack_desc* layers, int n, void* stream);

/* conv1_* kernels [5,5,3,Cout] -> [3][Cout][64] matching jcm_prep_input's channel order (5x5 s2 SAME == 3x1 s1 over x-folded s2d). */
int jcm_pack_weights_s2d(const float* w, int Cout, void* out_hi, void* out_lo, void* stream);

/* fp32 -> bf16 hi (+ lo) element-wise; n multiple of 4. */
int jcm_split_planes(const float* x, long n, void* hi, void* lo, void* stream);

/* ---- part detector --------------------------------------------------------------------------------------------- */

/* y = [relu](conv_SAME_stride1(x, w) + bias): tf.nn.conv2d + bias + tf.nn.relu, main.py:133-135,160-162.
 * x planes [B,H,W,Cin] (Cin a multiple of 16), w planes [ksize*kw][Cout_pad][Cin], y fp32 [B,H,W,Cout].
 * ksize = kernel height, kw = kernel width (0: square, kw = ksize).
 * TMA-fed implicit GEMM on tcgen05 tensor cores, fp32 accumulation in TMEM.
 * y: fp32 [B,H,W,Cout], or bf16 when y_bf16 != 0 (activations of the bf16 training configuration; Cout a multiple of 64). */
int jcm_conv2d_fwd(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, void* y, int y_bf16,
                   int B, int H, int W, int Cin, int Cout, int Cout_pad, int ksize, int kw, int relu, void* stream);

/* jcm_conv2d_fwd with the kernel variant forced - for tests and measurements; every variant computes the same values.
 * variant bit 0: single-CTA kernel instead of the CTA pair (cta_group::2) that N = 256 tiles use by default; bit 1: uniform tile grid
 * instead of the mixed-shape pixel-tile plan; bit 2: no N-split of the last partial wave. */
int jcm_conv2d_fwd_variant(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias, void* y,
                           int y_bf16, int B, int H, int W, int Cin, int Cout, int Cout_pad, int ksize, int kw, int relu, int variant,
                           void* stream);

/* tf.contrib.layers.batch_norm(decay=0.9, eps=1e-3, center, scale), main.py:128-130 and :112-113.
 * jcm_bn_stats: per-channel partial sums of x [M,C] into `partial` (jcm_bn_stats_blocks(M,C)*2*C floats).
 * jcm_bn_finalize: train != 0 -> batch statistics (biased variance), moving stats updated in place when update_moving
 * (unbiased variance); train == 0 -> moving statistics.  Emits scale = gamma*rstd, shift = beta - mean*scale. */
int jcm_bn_stats_blocks(long M, int C);
int jcm_bn_stats(const void* x, int x_bf16, long M, int C, float* partial, void* stream);   /* x: fp32, or bf16 when x_bf16 */
int jcm_bn_finalize(const float* partial, long M, int C, const float* gamma, const float* beta, float* moving_mean,
                    float* moving_var, float eps, float decay, int train, int update_moving, float* scale, float* shift,
                    float* save_mean, float* save_rstd, void* stream);

/* BN affine (+ tf.nn.max_pool 2x2 s2 SAME, main.py:172-174) -> operand planes and/or fp32. */
int jcm_bn_apply_pool(const void* a, int a_bf16, const float* scale, const float* shift, int B, int H, int W, int C, int pool,
                      void* out_hi, void* out_lo, float* out_f32, void* stream);

/* (bn(a1) + resize(bn(a2)) + resize(bn(a3))) / 3 with tf.image.resize_images legacy bilinear, main.py:58,67,69-70.
 * scale_shift = [6][C] = scale1, shift1, scale2, shift2, scale3, shift3. */
int jcm_upsample_avg3(const void* a1, const void* a2, const void* a3, int a_bf16, const float* scale_shift, int B, int H, int W,
                      int H2, int W2, int H3, int W3, int C, void* out_hi, void* out_lo, float* out_f32, void* stream);

/* ---- heads ----------------------------------------------------------------------------------------------------- */

/* spatial_softmax, main.py:212-217: softmax over S = H*W per (image, joint); logits/out [B,S,K]. */
int jcm_spatial_softmax(const float* logits, int B, int S, int K, float* out, void* stream);

/* softmax_cross_entropy, main.py:220-240: per_nk[B*K] = -sum_s labels*log_softmax(logits), loss[0] = mean.
 * labels [B,S,KL] (first K channels used); lse [B*K] optional (saved for the backward pass). */
int jcm_softmax_ce(const float* logits, const float* labels, int B, int S, int K, int KL, float* per_nk, float* lse,
                   float* loss, void* stream);

/* evaluation.get_joints_coords, evaluation.py:15-24 (= argmax_hm, main.py:389-397): out int32 [B,2,K] = (row, col) of the
 * first maximum in row-major order. */
int jcm_argmax_hw(const float* hm, int B, int H, int W, int K, int* out, void* stream);

/* ---- spatial model ---------------------------------------------------------------------------------------------- */

/* spatial_model + conv_mrf, main.py:77-125, after the bn_sm statistics (use jcm_bn_stats/finalize on heat_map [B*H*W, K+1]).
 * heat_map [B,H,W,K+1]; energies [P][2H][2W] (pairwise_energies, main.py:482-484), biases [P][H][W] (main.py:486-487);
 * pair_target / pair_cond int32 [P] on the device, sorted by (target, cond) = the reference's summation order. out [B,H,W,K]. */
long jcm_spatial_model_workspace(int B, int H, int W, int K, int P);
int jcm_spatial_model_fwd(const float* heat_map, const float* bn_scale, const float* bn_shift, const float* energies,
                          const float* biases, const int* pair_target, const int* pair_cond, float* out, void* workspace,
                          long workspace_bytes, int B, int H, int W, int K, int P, void* stream);

/* conv_mrf(A, B), main.py:77-91, on its own: A [2H,2W], Bmaps [b,H,W] (both used as given) -> out [b,H,W] =
 * legacy-bilinear resize of the 'valid' convolution.  Workspace: jcm_spatial_model_workspace(b, H, W, 0, 1) bytes. */
int jcm_conv_mrf_fwd(const float* A, const float* Bmaps, float* out, void* workspace, long workspace_bytes, int b, int H,
                     int W, void* stream);

/* ---- backward pass (the reference: TensorFlow autodiff via opt.compute_gradients(loss_tower), main.py:557-560) ------- */

/* d(mean softmax-CE)/d logits * scale = scale * (softmax * sum(labels) - labels); lse from jcm_softmax_ce (main.py:239). */
int jcm_softmax_ce_bwd(const float* logits, const float* labels, const float* lse, int B, int S, int K, int KL, float scale,
                       float* dlogits, void* stream);

/* backward of spatial_softmax (main.py:212-217): dx (+)= y * (dy - sum_s dy*y); dy has KD >= K channels. */
int jcm_spatial_softmax_bwd(const float* y, const float* dy, int B, int S, int K, int KD, int accumulate, float* dx, void* stream);

/* [2x2 SAME max-pool bwd] + training-mode batch-norm bwd + ReLU bwd of one conv_layer (main.py:156-169).
 * a = ReLU output [B,H,W,C]; dout = gradient w.r.t. the layer output ([B,ceil(H/2),ceil(W/2),C] when pool), times dy_scale.
 * dout: fp32, or bf16 when dout_bf16 (the bf16 configuration lets the data-gradient convolution write bf16).
 * Outputs: d_pre planes [B,H,W,C] (gradient w.r.t. the conv output), dgamma, dbeta, dbias [C].
 * workspace: (4 * jcm_bn_relu_bwd_blocks(M_out, C) + 2) * C floats. */
int jcm_bn_relu_bwd_blocks(long M_out, int C);
int jcm_bn_relu_bwd(const void* a, int a_bf16, const void* dout, int dout_bf16, const float* scale, const float* shift, const float* mean, const float* rstd,
                    float dy_scale, int B, int H, int W, int C, int pool, void* d_hi, void* d_lo, float* d_f32, float* dgamma,
                    float* dbeta, float* dbias, float* workspace, void* stream);

/* column sums of x [M,C] (bias gradient of the last layer); partial: jcm_bn_stats_blocks(M,C)*2*C floats. */
int jcm_colsum(const float* x, long M, int C, float* partial, float* out, void* stream);

/* transpose of the up-sampling + 3-way average (main.py:58,67,69-70): d2 [B,H2,W2,C], d3 [B,H3,W3,C] (the full-resolution
 * bank's gradient is dmerged / 3, folded into jcm_bn_relu_bwd's dy_scale). */
int jcm_upsample_avg3_bwd(const void* dmerged, int dm_bf16, int B, int H, int W, int H2, int W2, int H3, int W3, int C, float* d2, float* d3,
                          void* stream);      /* dmerged: fp32, or bf16 when dm_bf16 */

/* fp32 [M,C] -> bf16 planes [M,Cpad] with zero-padded channels. */
int jcm_pad_planes(const float* x, long M, int C, int Cpad, void* hi, void* lo, void* stream);

/* weight gradient of tf.nn.conv2d (main.py:135): dw [ksize*kw][Cin][dw_cout_stride] from input planes x [B,H,W,Cin] and output-gradient
 * planes g [B,H,W,Gc].  tcgen05 GEMM per tap with the contraction over pixels (MN-major operands), split-K partial sums in
 * `workspace` (jcm_conv2d_wgrad_workspace bytes), deterministic reduction.  The data gradient is jcm_conv2d_fwd on
 * jcm_pack_weights(..., transpose = 1). */
long jcm_conv2d_wgrad_workspace(int B, int H, int W, int Cin, int Gc, int ksize, int kw);
int jcm_conv2d_wgrad(const void* x_hi, const void* x_lo, const void* g_hi, const void* g_lo, float* dw, void* workspace,
                     long workspace_bytes, int B, int H, int W, int Cin, int Gc, int Cout, int dw_cout_stride, int ksize, int kw,
                     void* stream);

/* conv1 weight gradient in x-folded space-to-depth form [3][64][Cout] (as jcm_conv2d_wgrad writes it) -> [5,5,3,Cout]. */
int jcm_unpack_s2d_grad(const float* g9, int Cout, float* dw, void* stream);

/* backward of jcm_spatial_model_fwd.  fwd_workspace: the forward workspace of the same inputs. */
long jcm_spatial_model_bwd_workspace(int B, int H, int W, int K, int P);
int jcm_spatial_model_bwd(const float* g, const float* heat_map, const float* bn_scale, const float* bn_shift, const float* bn_mean,
                          const float* bn_rstd, int train, const float* energies, const float* biases, const int* pair_target,
                          const int* pair_cond, const void* fwd_workspace, void* workspace, long workspace_bytes, float* d_heat_map,
                          float* dE, float* db, float* dgamma, float* dbeta, int B, int H, int W, int K, int P, void* stream);

/* Tensor-core form of the spatial model (bf16 training configuration): same arguments, workspace protocol and results as
 * jcm_spatial_model_fwd / _bwd (main.py:94-125 and its autodiff), with the pairwise convolutions computed as grouped Toeplitz
 * GEMMs on tcgen05 with bf16 operands and fp32 accumulation.  The prior operand is sp(E) minus its per-pair mean; the common level
 * is added back in fp32, so on nearly flat priors (the reference's) the logits and every gradient that is linear in the prior
 * agree with the fp32 form to ~1e-6, the prior gradient dE to ~3e-3 of its maximum (K + 1 <= 32).  The two forms keep different
 * workspaces: pass the _tc_ forward workspace to the _tc_ backward. */
long jcm_spatial_model_tc_workspace(int B, int H, int W, int K, int P);
int jcm_spatial_model_tc_fwd(const float* heat_map, const float* bn_scale, const float* bn_shift, const float* energies,
                             const float* biases, const int* pair_target, const int* pair_cond, float* out, void* workspace,
                             long workspace_bytes, int B, int H, int W, int K, int P, void* stream);
long jcm_spatial_model_tc_bwd_workspace(int B, int H, int W, int K, int P);
int jcm_spatial_model_tc_bwd(const float* g, const float* heat_map, const float* bn_scale, const float* bn_shift, const float* bn_mean,
                             const float* bn_rstd, int train, const float* energies, const float* biases, const int* pair_target,
                             const int* pair_cond, const void* fwd_workspace, void* workspace, long workspace_bytes,
                             float* d_heat_map, float* dE, float* db, float* dgamma, float* dbeta, int B, int H, int W, int K, int P,
                             void* stream);

/* ---- tap-expanded form of a convolution with very few output channels (conv6: 9x9, 512 -> K, main.py:72) ------------------
 * y[p,co] = b[co] + sum_tap Z[p+tap-pad, tap*KP+co] with Z = 1x1 conv (jcm_conv2d_fwd, ksize 1) of the input with the weights packed
 * by jcm_pack_weights_taps (rows n = tap*KP+co; KP = Cout padded to a multiple of 4; Npad >= k*k*KP rows, zero filled).
 * Backward: jcm_tap_scatter_planes builds Gt[q, tap*KP+co] = G[q-(tap-pad), co] as bf16 operand planes; the weight gradient is
 * jcm_conv2d_wgrad(x, Gt, ksize 1) -> dwz [Cin][ZC] -> jcm_unpack_tap_grad -> [k,k,Cin,Cout]; the data gradient is jcm_conv2d_fwd
 * (ksize 1) of Gt with jcm_pack_weights_taps(transpose = 1). */
int jcm_pack_weights_taps(const float* w, int ksize, int Cin, int Cout, int KP, int Npad, int transpose, void* out_hi, void* out_lo,
                          void* stream);
int jcm_tap_gather(const float* z, const float* bias, int B, int H, int W, int ksize, int KP, int ZC, int Cout, float* y, void* stream);
int jcm_tap_scatter_planes(const float* g, int B, int H, int W, int ksize, int Cout, int KP, int Npad, void* hi, void* lo, void* stream);
int jcm_unpack_tap_grad(const float* dwz, int ksize, int Cin, int Cout, int KP, int ZC, float* dw, void* stream);

/* ---- optimizer on flat buffers (main.py:243-267 gradient mean, :195-205 weight decay, :302-309 clip, :501-506,577 Adam) --- */
int jcm_optim_blocks(long n);
/* g <- g * inv_world + lmbd * w (first n_decay elements); stats[0] = ||g||, stats[1] = sum w^2/2.  partial: 2*jcm_optim_blocks(n). */
int jcm_grad_prepare(float* g, const float* w, long n, long n_decay, float inv_world, float lmbd, float* partial, float* stats,
                     void* stream);
/* w <- Adam_TF1(w, g * clip / max(stats[0], clip)) (momentum != 0: MomentumOptimizer with b1 as the momentum). */
int jcm_clip_adam(float* w, const float* g, float* m, float* v, long n, const float* stats, float clip, float lr_t, float b1,
                  float b2, float eps, int momentum, void* stream);

/* ---- stand-alone pieces of the function surface (used by jcm.conv_layer / conv2d / weight_decay / average_gradients / grad_renorm;
 * inside the fused training step the same arithmetic lives in jcm_conv2d_fwd's epilogue, jcm_grad_prepare and jcm_clip_adam) ---- */
/* out[0] (+)= mul * sum x^2.  weight_decay(var_pattern), main.py:195-205 = sum of tf.nn.l2_loss: mul 0.5, accumulate from the second
 * variable on; squared global norm of tf.clip_by_global_norm, main.py:302-309: mul 1.  partial: jcm_optim_blocks(n) floats. */
int jcm_sumsq(const float* x, long n, float mul, int accumulate, float* partial, float* out, void* stream);
/* out = g * clip / max(sqrt(sumsq[0]), clip): grad_renorm, main.py:302-309 (sumsq[0] over the whole gradient list). */
int jcm_clip_scale(const float* g, long n, const float* sumsq, float clip, float* out, void* stream);
/* out = mean of n_towers (<= 16) gradient tensors on this device, summed in tower order: average_gradients, main.py:243-267.
 * towers: HOST array of device pointers. */
int jcm_tower_mean(const float* const* towers, int n_towers, long n, float* out, void* stream);
/* y = x[:, oy::2, ox::2, :]: turns the stride-1 SAME convolution into tf.nn.conv2d(strides=[1,2,2,1], 'SAME'), main.py:133-135
 * (oy/ox = 1 for an even input extent, 0 for an odd one). */
int jcm_subsample2(const float* x, int B, int H, int W, int C, int oy, int ox, float* y, void* stream);
/* out = [relu](x + bias) per channel of x [M,C]: `conv2d(...) + b`, tf.nn.relu of conv_layer, main.py:160-162. */
int jcm_bias_relu(const float* x, const float* bias, long M, int C, int relu, float* out, void* stream);

/* ---- training augmentation (augmentation.py:12-77, applied at main.py:495-499 before the tower forward) -------
 * NHWC fp32; prm [B][8] on the device = {flip 0/1, brightness delta, contrast factor, cos(angle), sin(angle), rh, rw, unused}:
 * the caller makes the random draws (tf.random_uniform in the reference).  src != out for the resampling kernels. */
int jcm_augment_color(const float* x, const float* prm, float* mean_ws /*[B*C]*/, int B, int H, int W, int C, float* out, void* stream);
int jcm_augment_flip_channels(const float* hm, const float* prm, const int* perm /*[C] device*/, int B, int H, int W, int C, float* out,
                              void* stream);
int jcm_augment_rotate(const float* src, const float* prm, int B, int H, int W, int C, float* out, void* stream);
int jcm_augment_crop_resize(const float* src, const float* prm, int B, int H, int W, int C, float crop_size, float* out, void* stream);
int jcm_augment_hm_renorm(const float* hm, int B, int H, int W, int C, float power, float eps, float* out, void* stream);

/* ---- measurement / test support --------------------------------------------------------------------------------- */

/* FP32 FMA peak loop (packed = 1: FFMA2); flops_out = FLOPs of one launch. scratch: blocks*512 floats. */
int jcm_fma_peak(float* scratch, int blocks, int iters, int packed, double* flops_out, void* stream);

/* The mixed-shape pixel-tile plan the convolution kernels use for an H x W map (csrc/tiling.cuh): cap = 128 (forward / data gradient
 * M tiles) or 64 (weight-gradient k-blocks).  out = [n_tiles, n_shapes, then x0, y0, box_w, box_h per tile]; returns the ints written,
 * 0 when no mixed plan beats `uniform_tiles`, < 0 (minus the size needed) when max_out is too small.  Host only, no GPU needed. */
int jcm_debug_tile_plan(int H, int W, int cap, int exact_px, int max_shapes, int uniform_tiles, int* out, int max_out);

/* Test / measurement switch of jcm_conv2d_wgrad (every variant computes the same values up to the summation order of the k-splits):
 * bit 0 = single-CTA kernel instead of the CTA pair, bit 1 = uniform patch grid instead of the mixed-shape plan, bit 2 = one tap per task
 * instead of tap groups on the narrow layers.  Returns the previous
 * value.  Process-wide; the product path never calls it. */
int jcm_debug_set_wgrad_variant(int variant);

/* Naive direct convolution on the same operand planes - used only by tests to cross-check the tcgen05 kernel. */
int jcm_debug_conv2d_naive(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                           float* y, int B, int H, int W, int Cin, int Cout, int Cout_pad, int ksize, int kw, int relu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JCM_H_ */

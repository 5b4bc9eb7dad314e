// HBM-bound glue kernels of the part detector and the loss heads (all NHWC, fp32 in, vectorised float4 access):
//   jcm_bn_stats / jcm_bn_finalize   per-channel batch statistics -> (scale, shift), moving-average update
//                                    (reference main.py:128-130,112-113: tf.contrib.layers.batch_norm, decay 0.9,
//                                     eps 1e-3, biased variance to normalise, unbiased into the moving average)
//   jcm_bn_apply_pool                BN affine (+ 2x2 s2 SAME max-pool, main.py:172-174) -> bf16 hi/lo operand planes
//   jcm_upsample_avg3                BN affine on the three bank outputs + legacy-bilinear up-sampling of the 1/2 and 1/4
//                                    banks + (x1+x2+x3)/3 (main.py:58,67,69-70) -> bf16 hi/lo operand planes
//   jcm_spatial_softmax              softmax over H*W per (image, joint)   (main.py:212-217)
//   jcm_softmax_ce                   soft-label cross entropy, mean over (image, joint) (main.py:220-240)
//   jcm_argmax_hw                    first-max (row, col) per (image, joint) (evaluation.py:15-24)
#include "common.cuh"

namespace {

constexpr int kStatThreads = 256;

// ------------------------------------------------------------------------------------------- BN statistics
// x [M, C]; each block reduces a contiguous slab of rows; thread t owns column group (t % G) and row lane (t / G).
template <int VEC>
__global__ void bn_stats_kernel(const void* __restrict__ x, int x_bf16, long M, int C, float* __restrict__ partial /*[grid][2][C]*/) {
  extern __shared__ float sh[];  // [kStatThreads][2*VEC]
  const int G = C / VEC;         // column groups
  const int lanes = kStatThreads / G;
  const int g = threadIdx.x % G, rl = threadIdx.x / G;
  float s[VEC], q[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { s[v] = 0.f; q[v] = 0.f; }
  const long rows_per_block = (M + gridDim.x - 1) / gridDim.x;
  const long r0 = blockIdx.x * rows_per_block;
  long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  if (rl < lanes) {
    for (long r = r0 + rl; r < r1; r += lanes) {
      if (VEC == 8) {
        float f[VEC];
        load_act_vec<VEC>(x, r * C + (long)g * VEC, x_bf16, f);
#pragma unroll
        for (int v = 0; v < VEC; ++v) { s[v] += f[v]; q[v] += f[v] * f[v]; }
      } else if (VEC == 4) {
        const float4 v = load_act4(x, (r * C + g * 4) >> 2, x_bf16);
        s[0] += v.x; q[0] += v.x * v.x;
        s[1 % VEC] += v.y; q[1 % VEC] += v.y * v.y;
        s[2 % VEC] += v.z; q[2 % VEC] += v.z * v.z;
        s[3 % VEC] += v.w; q[3 % VEC] += v.w * v.w;
      } else {
        const float v = load_act1(x, r * C + g, x_bf16);
        s[0] += v; q[0] += v * v;
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    sh[threadIdx.x * 2 * VEC + v] = s[v];
    sh[threadIdx.x * 2 * VEC + VEC + v] = q[v];
  }
  __syncthreads();
  // deterministic tree-free reduction: thread c sums its column over the row lanes in order
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int gg = c / VEC, v = c % VEC;
    float ts = 0.f, tq = 0.f;
    for (int l = 0; l < lanes; ++l) {
      ts += sh[(l * G + gg) * 2 * VEC + v];
      tq += sh[(l * G + gg) * 2 * VEC + VEC + v];
    }
    partial[((long)blockIdx.x * 2 + 0) * C + c] = ts;
    partial[((long)blockIdx.x * 2 + 1) * C + c] = tq;
  }
}

// train != 0: batch statistics from the partials, moving stats updated in place.  train == 0: moving stats.
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nblocks, long M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ moving_mean, float* __restrict__ moving_var,
                                   float eps, float decay, int train, int update_moving, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  __shared__ double sh[kPartY][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  double s = 0.0, q = 0.0;
  if (train) {
    s = partial_colsum(partial, nblocks, C, 0, c, sh);
    q = partial_colsum(partial, nblocks, C, 1, c, sh);
  }
  if (threadIdx.y != 0 || c >= C) return;
  double mean, var;
  if (train) {
    mean = s / (double)M;
    var = q / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    if (update_moving) {
      const double unb = M > 1 ? var * ((double)M / (double)(M - 1)) : var;
      moving_mean[c] = (float)((double)moving_mean[c] * decay + (1.0 - decay) * mean);
      moving_var[c] = (float)((double)moving_var[c] * decay + (1.0 - decay) * unb);
    }
  } else {
    mean = moving_mean[c];
    var = moving_var[c];
  }
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  if (save_mean) save_mean[c] = (float)mean;
  if (save_rstd) save_rstd[c] = rstd;
}

// ------------------------------------------------------------------------------------------- BN apply (+pool)
template <int VEC>
__global__ void bn_apply_pool_kernel(const void* __restrict__ a, int a_bf16, const float* __restrict__ scale, const float* __restrict__ shift,
                                     int B, int H, int W, int C, int pool, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo, float* __restrict__ out_f32) {
  const int Ho = pool ? (H + 1) / 2 : H, Wo = pool ? (W + 1) / 2 : W;
  const int CG = C / VEC;
  const long total = (long)B * Ho * Wo * CG;
  const long step = (long)gridDim.x * blockDim.x;
  const bool fixed_cg = (step % CG) == 0;      // every thread then keeps one channel group: scale / shift stay in registers
  float sc[VEC], sh[VEC];
  int cg_loaded = -1;
  // (measured: two items per thread with all loads first took this kernel from 54 to 116 registers and from 0.77 to 1.27 ms per
  // step - four resident blocks of one item each keep more bytes in flight than two blocks of two)
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += step) {
    int cg, xo, yo, n;
    split_index(i, CG, Wo, Ho, cg, xo, yo, n);
    float r[VEC];
    if (!fixed_cg || cg_loaded != cg) {
      load_f32_vec<VEC>(scale, (long)cg * VEC, sc);
      load_f32_vec<VEC>(shift, (long)cg * VEC, sh);
      cg_loaded = cg;
    }
    if (!pool) {
      load_act_vec<VEC>(a, i * VEC, a_bf16, r);
#pragma unroll
      for (int c = 0; c < VEC; ++c) r[c] = fmaf(r[c], sc[c], sh[c]);
    } else {
      const int y0 = 2 * yo, x0 = 2 * xo;
      const long base = (((long)n * H + y0) * W + x0) * C + (long)cg * VEC;
      const bool hx = x0 + 1 < W, hy = y0 + 1 < H;  // SAME pooling: the padded row/column never wins
      float v1[VEC], v2[VEC], v3[VEC];
      load_act_vec<VEC>(a, base, a_bf16, r);
      if (hx) load_act_vec<VEC>(a, base + C, a_bf16, v1);
      if (hy) load_act_vec<VEC>(a, base + (long)W * C, a_bf16, v2);
      if (hx && hy) load_act_vec<VEC>(a, base + (long)W * C + C, a_bf16, v3);
#pragma unroll
      for (int c = 0; c < VEC; ++c) r[c] = fmaf(r[c], sc[c], sh[c]);
      if (hx) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) r[c] = fmaxf(r[c], fmaf(v1[c], sc[c], sh[c]));
      }
      if (hy) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) r[c] = fmaxf(r[c], fmaf(v2[c], sc[c], sh[c]));
      }
      if (hx && hy) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) r[c] = fmaxf(r[c], fmaf(v3[c], sc[c], sh[c]));
      }
    }
    store_planes_vec<VEC>(hi, lo, out_f32, i * VEC, r);
  }
}

// bf16 configuration (bf16 activation in, bf16 hi plane out, grid stride a multiple of the channel-group count): the next item's
// loads (one 16-byte word, or the four of a pooling window) are in flight while the current item is processed.
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <int POOL>
__global__ void __launch_bounds__(256, 3)
bn_apply_pool_bf16_kernel(const uint4* __restrict__ a, const float* __restrict__ scale, const float* __restrict__ shift, int B, int H, int W,
                          int C, uint4* __restrict__ hi) {
  constexpr int VEC = 8;
  const int Ho = POOL ? (H + 1) / 2 : H, Wo = POOL ? (W + 1) / 2 : W;
  const int CG = C / VEC;
  const long total = (long)B * Ho * Wo * CG;
  const long step = (long)gridDim.x * blockDim.x;
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = (int)(i % CG);
  float sc[VEC], sh[VEC];
  load_f32_vec<VEC>(scale, (long)cg * VEC, sc);
  load_f32_vec<VEC>(shift, (long)cg * VEC, sh);
  uint4 nraw[POOL ? 4 : 1];
  unsigned nokm = 1u;
  const long row8 = (long)W * CG;
  auto fetch = [&](long it) {
    if (!POOL) {
      nraw[0] = a[it];
    } else {
      int cgi, xo, yo, n;
      split_index(it, CG, Wo, Ho, cgi, xo, yo, n);
      const int y0 = 2 * yo, x0 = 2 * xo;
      const long base = (((long)n * H + y0) * W + x0) * CG + cg;
      const bool hx = x0 + 1 < W, hy = y0 + 1 < H;      // SAME pooling: the padded row / column never wins
      nokm = 1u | (hx ? 2u : 0u) | (hy ? 4u : 0u) | ((hx && hy) ? 8u : 0u);
      nraw[0] = a[base];
      if (hx) nraw[POOL ? 1 : 0] = a[base + CG];
      if (hy) nraw[POOL ? 2 : 0] = a[base + row8];
      if (hx && hy) nraw[POOL ? 3 : 0] = a[base + row8 + CG];
    }
  };
  fetch(i);
  for (; i < total; i += step) {
    uint4 raw[POOL ? 4 : 1];
#pragma unroll
    for (int k = 0; k < (POOL ? 4 : 1); ++k) raw[k] = nraw[k];
    const unsigned okm = nokm;
    if (i + step < total) fetch(i + step);
    float r[VEC], v[VEC];
    unpack8(raw[0], v);
#pragma unroll
    for (int c = 0; c < VEC; ++c) r[c] = fmaf(v[c], sc[c], sh[c]);
    if (POOL) {
#pragma unroll
      for (int e = 1; e < 4; ++e) {
        if ((okm >> e) & 1u) {
          unpack8(raw[POOL ? e : 0], v);
#pragma unroll
          for (int c = 0; c < VEC; ++c) r[c] = fmaxf(r[c], fmaf(v[c], sc[c], sh[c]));
        }
      }
    }
    __align__(16) __nv_bfloat16 h[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) h[c] = __float2bfloat16_rn(r[c]);
    hi[i] = *reinterpret_cast<uint4*>(h);
  }
}

// ------------------------------------------------------------------------------------------- upsample + average
// legacy bilinear (tf.image.resize_images, align_corners=False, no half-pixel offset): src = dst * (in/out) in fp32
__device__ __forceinline__ void legacy_tap(int dst, int n_in, int n_out, int& lo, int& hi, float& w) {
  const float scale = (float)n_in / (float)n_out;
  const float src = (float)dst * scale;
  lo = (int)floorf(src);
  hi = min(lo + 1, n_in - 1);
  w = src - (float)lo;
}
// The taps of every destination row / column are tabulated once per block in shared memory (the per-item version recomputed four
// `(float)n_in / n_out` divisions and two floors per sample: 0.75 ms for 0.82 GB of traffic, six times its HBM time).
struct TapTab { int lo, hi; float w; };

// legacy-bilinear sample of the RAW activation (the batch-norm affine is applied by the caller AFTER the interpolation: the four
// weights sum to 1, so  interp(scale * a + shift) == scale * interp(a) + shift  - one FMA per channel instead of four)
template <int VEC>
__device__ __forceinline__ void resize_sample(const void* __restrict__ a, int a_bf16, long img_base, int Wi, int C, const TapTab ty,
                                              const TapTab tx, float (&out)[VEC]) {
  float tl[VEC], tr[VEC], bl[VEC], br[VEC];
  load_act_vec<VEC>(a, img_base + ((long)ty.lo * Wi + tx.lo) * C, a_bf16, tl);
  load_act_vec<VEC>(a, img_base + ((long)ty.lo * Wi + tx.hi) * C, a_bf16, tr);
  load_act_vec<VEC>(a, img_base + ((long)ty.hi * Wi + tx.lo) * C, a_bf16, bl);
  load_act_vec<VEC>(a, img_base + ((long)ty.hi * Wi + tx.hi) * C, a_bf16, br);
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    const float top = tl[c] + (tr[c] - tl[c]) * tx.w, bot = bl[c] + (br[c] - bl[c]) * tx.w;
    out[c] = top + (bot - top) * ty.w;
  }
}

// out = (bn1(a1) + resize(bn2(a2)) + resize(bn3(a3))) / 3 = s1 * a1 + s2 * resize(a2) + s3 * resize(a3) + t, with s_k = scale_k / 3 and
// t = (shift1 + shift2 + shift3) / 3: three FMAs per channel after the two interpolations.  (The first version applied the affine to
// all nine taps, divided by 3 in IEEE and converted the unused residual plane: 638 instructions per 8-channel item, instruction-bound
// at 0.82 ms for 0.82 GB of traffic.)
template <int VEC>
__global__ void __launch_bounds__(256, 2) upsample_avg3_kernel(const void* __restrict__ a1, const void* __restrict__ a2, const void* __restrict__ a3, int a_bf16,
                                     const float* __restrict__ ss /* [6][C]: scale1, shift1, scale2, shift2, scale3, shift3 */,
                                     int B, int H, int W, int H2, int W2, int H3, int W3, int C, __nv_bfloat16* __restrict__ hi,
                                     __nv_bfloat16* __restrict__ lo, float* __restrict__ out_f32) {
  extern __shared__ TapTab tabs[];          // [H] rows of bank 2, [W] columns of bank 2, [H] rows of bank 3, [W] columns of bank 3
  TapTab* ty2 = tabs;
  TapTab* tx2 = ty2 + H;
  TapTab* ty3 = tx2 + W;
  TapTab* tx3 = ty3 + H;
  for (int i = threadIdx.x; i < 2 * (H + W); i += blockDim.x) {
    const int bank3 = i >= H + W, j = bank3 ? i - (H + W) : i;
    const bool row = j < H;
    const int d = row ? j : j - H;
    TapTab t;
    legacy_tap(d, row ? (bank3 ? H3 : H2) : (bank3 ? W3 : W2), row ? H : W, t.lo, t.hi, t.w);
    tabs[i] = t;
  }
  __syncthreads();
  const int CG = C / VEC;
  const long total = (long)B * H * W * CG;
  const long step = (long)gridDim.x * blockDim.x;
  // when the grid stride is a multiple of the channel-group count every thread keeps ONE channel group: its constants stay in
  // registers for the whole loop
  const bool fixed_cg = (step % CG) == 0;
  float s1[VEC], s2[VEC], s3[VEC], t0[VEC];
  int cg_loaded = -1;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += step) {
    int cg, x, y, n;
    split_index(i, CG, W, H, cg, x, y, n);
    if (!fixed_cg || cg_loaded != cg) {
      float u[VEC];
      load_f32_vec<VEC>(ss + 0 * C, (long)cg * VEC, s1);
      load_f32_vec<VEC>(ss + 2 * C, (long)cg * VEC, s2);
      load_f32_vec<VEC>(ss + 4 * C, (long)cg * VEC, s3);
      load_f32_vec<VEC>(ss + 1 * C, (long)cg * VEC, t0);
      load_f32_vec<VEC>(ss + 3 * C, (long)cg * VEC, u);
#pragma unroll
      for (int c = 0; c < VEC; ++c) t0[c] += u[c];
      load_f32_vec<VEC>(ss + 5 * C, (long)cg * VEC, u);
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        t0[c] = (t0[c] + u[c]) * (1.0f / 3.0f);
        s1[c] *= (1.0f / 3.0f); s2[c] *= (1.0f / 3.0f); s3[c] *= (1.0f / 3.0f);
      }
      cg_loaded = cg;
    }
    float v1[VEC], v2[VEC], v3[VEC], r[VEC];
    load_act_vec<VEC>(a1, i * VEC, a_bf16, v1);
    resize_sample<VEC>(a2, a_bf16, (long)n * H2 * W2 * C + (long)cg * VEC, W2, C, ty2[y], tx2[x], v2);
    resize_sample<VEC>(a3, a_bf16, (long)n * H3 * W3 * C + (long)cg * VEC, W3, C, ty3[y], tx3[x], v3);
#pragma unroll
    for (int c = 0; c < VEC; ++c) r[c] = fmaf(s1[c], v1[c], fmaf(s2[c], v2[c], fmaf(s3[c], v3[c], t0[c])));
    store_planes_vec<VEC>(hi, lo, out_f32, i * VEC, r);
  }
}

// bf16 configuration (bf16 bank outputs, bf16 hi plane only, grid stride a multiple of the channel-group count): the same arithmetic
// with the NEXT item's nine 16-byte loads in flight while the current item is interpolated (the kernel above waits on the long
// scoreboard with 16 warps per SM: 0.78 ms for 0.82 GB).
__global__ void __launch_bounds__(256, 2)
upsample_avg3_bf16_kernel(const uint4* __restrict__ a1, const uint4* __restrict__ a2, const uint4* __restrict__ a3,
                          const float* __restrict__ ss, int B, int H, int W, int H2, int W2, int H3, int W3, int C, uint4* __restrict__ hi) {
  extern __shared__ TapTab tabs[];
  TapTab* ty2 = tabs;
  TapTab* tx2 = ty2 + H;
  TapTab* ty3 = tx2 + W;
  TapTab* tx3 = ty3 + H;
  for (int i = threadIdx.x; i < 2 * (H + W); i += blockDim.x) {
    const int bank3 = i >= H + W, j = bank3 ? i - (H + W) : i;
    const bool row = j < H;
    const int d = row ? j : j - H;
    TapTab t;
    legacy_tap(d, row ? (bank3 ? H3 : H2) : (bank3 ? W3 : W2), row ? H : W, t.lo, t.hi, t.w);
    tabs[i] = t;
  }
  __syncthreads();
  constexpr int VEC = 8;
  const int CG = C / VEC;
  const long total = (long)B * H * W * CG;
  const long step = (long)gridDim.x * blockDim.x;     // a multiple of CG (checked by the launcher): one channel group per thread
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cg = (int)(i % CG);
  float s1[VEC], s2[VEC], s3[VEC], t0[VEC];
  {
    float u[VEC];
    load_f32_vec<VEC>(ss + 0 * C, (long)cg * VEC, s1);
    load_f32_vec<VEC>(ss + 2 * C, (long)cg * VEC, s2);
    load_f32_vec<VEC>(ss + 4 * C, (long)cg * VEC, s3);
    load_f32_vec<VEC>(ss + 1 * C, (long)cg * VEC, t0);
    load_f32_vec<VEC>(ss + 3 * C, (long)cg * VEC, u);
#pragma unroll
    for (int c = 0; c < VEC; ++c) t0[c] += u[c];
    load_f32_vec<VEC>(ss + 5 * C, (long)cg * VEC, u);
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      t0[c] = (t0[c] + u[c]) * (1.0f / 3.0f);
      s1[c] *= (1.0f / 3.0f); s2[c] *= (1.0f / 3.0f); s3[c] *= (1.0f / 3.0f);
    }
  }
  uint4 nraw[9];
  float nw[4];      // wx2, wy2, wx3, wy3 of the prefetched item
  auto fetch = [&](long it) {
    int cgi, x, y, n;
    split_index(it, CG, W, H, cgi, x, y, n);
    const TapTab y2 = ty2[y], x2 = tx2[x], y3 = ty3[y], x3 = tx3[x];
    const long b2 = ((long)n * H2 * W2) * CG + cg, b3 = ((long)n * H3 * W3) * CG + cg;
    nraw[0] = a1[it];
    nraw[1] = a2[b2 + ((long)y2.lo * W2 + x2.lo) * CG];
    nraw[2] = a2[b2 + ((long)y2.lo * W2 + x2.hi) * CG];
    nraw[3] = a2[b2 + ((long)y2.hi * W2 + x2.lo) * CG];
    nraw[4] = a2[b2 + ((long)y2.hi * W2 + x2.hi) * CG];
    nraw[5] = a3[b3 + ((long)y3.lo * W3 + x3.lo) * CG];
    nraw[6] = a3[b3 + ((long)y3.lo * W3 + x3.hi) * CG];
    nraw[7] = a3[b3 + ((long)y3.hi * W3 + x3.lo) * CG];
    nraw[8] = a3[b3 + ((long)y3.hi * W3 + x3.hi) * CG];
    nw[0] = x2.w; nw[1] = y2.w; nw[2] = x3.w; nw[3] = y3.w;
  };
  fetch(i);
  for (; i < total; i += step) {
    uint4 raw[9];
    float w[4];
#pragma unroll
    for (int k = 0; k < 9; ++k) raw[k] = nraw[k];
#pragma unroll
    for (int k = 0; k < 4; ++k) w[k] = nw[k];
    if (i + step < total) fetch(i + step);
    float r[VEC];
    {
      float v[VEC];
      unpack8(raw[0], v);
#pragma unroll
      for (int c = 0; c < VEC; ++c) r[c] = fmaf(s1[c], v[c], t0[c]);
    }
#pragma unroll
    for (int bank = 0; bank < 2; ++bank) {
      float tl[VEC], tr[VEC], bl[VEC], br[VEC];
      unpack8(raw[1 + 4 * bank], tl);
      unpack8(raw[2 + 4 * bank], tr);
      unpack8(raw[3 + 4 * bank], bl);
      unpack8(raw[4 + 4 * bank], br);
      const float wx = w[2 * bank], wy = w[2 * bank + 1];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        const float top = tl[c] + (tr[c] - tl[c]) * wx, bot = bl[c] + (br[c] - bl[c]) * wx;
        r[c] = fmaf(bank ? s3[c] : s2[c], top + (bot - top) * wy, r[c]);
      }
    }
    __align__(16) __nv_bfloat16 h[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) h[c] = __float2bfloat16_rn(r[c]);
    hi[i] = *reinterpret_cast<uint4*>(h);
  }
}

// ------------------------------------------------------------------------------------------- softmax / CE / argmax
__device__ float block_reduce_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
  return r;
}
__device__ float block_reduce_max(float v, float* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = -INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r = fmaxf(r, sh[i]);
  return r;
}

// one block per (image, joint); logits [B, S, K]
__global__ void spatial_softmax_kernel(const float* __restrict__ logits, int S, int K, float* __restrict__ out) {
  __shared__ float sh[32];
  const int n = blockIdx.x / K, k = blockIdx.x % K;
  const float* src = logits + (long)n * S * K + k;
  float* dst = out + (long)n * S * K + k;
  float m = -INFINITY;
  for (int s = threadIdx.x; s < S; s += blockDim.x) m = fmaxf(m, src[(long)s * K]);
  m = block_reduce_max(m, sh);
  float sum = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) sum += expf(src[(long)s * K] - m);
  sum = block_reduce_sum(sum, sh);
  const float inv = 1.0f / sum;
  for (int s = threadIdx.x; s < S; s += blockDim.x) dst[(long)s * K] = expf(src[(long)s * K] - m) * inv;
}

// labels [B, S, KL] (first K channels used). per_nk[n*K+k] = lse * sum(y) - sum(y * x); lse_out optional (for backward)
__global__ void softmax_ce_kernel(const float* __restrict__ logits, const float* __restrict__ labels, int S, int K, int KL,
                                  float* __restrict__ per_nk, float* __restrict__ lse_out) {
  __shared__ float sh[32];
  const int n = blockIdx.x / K, k = blockIdx.x % K;
  const float* src = logits + (long)n * S * K + k;
  const float* lab = labels + (long)n * S * KL + k;
  float m = -INFINITY;
  for (int s = threadIdx.x; s < S; s += blockDim.x) m = fmaxf(m, src[(long)s * K]);
  m = block_reduce_max(m, sh);
  float sum = 0.f, sy = 0.f, syx = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float xv = src[(long)s * K];
    const float yv = lab[(long)s * KL];
    sum += expf(xv - m);
    sy += yv;
    syx += yv * xv;
  }
  sum = block_reduce_sum(sum, sh);
  sy = block_reduce_sum(sy, sh);
  syx = block_reduce_sum(syx, sh);
  if (threadIdx.x == 0) {
    const float lse = m + logf(sum);
    per_nk[blockIdx.x] = lse * sy - syx;
    if (lse_out) lse_out[blockIdx.x] = lse;
  }
}

__global__ void mean_kernel(const float* __restrict__ v, int n, float* __restrict__ out) {
  __shared__ float sh[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)v[i];
  float r = block_reduce_sum((float)acc, sh);
  if (threadIdx.x == 0) out[0] = r / (float)n;
}

// first maximal index in row-major order -> (row, col); out [B, 2, K] int32
__global__ void argmax_hw_kernel(const float* __restrict__ hm, int S, int W, int K, int* __restrict__ out) {
  __shared__ float shv[32];
  __shared__ int shi[32];
  const int n = blockIdx.x / K, k = blockIdx.x % K;
  const float* src = hm + (long)n * S * K + k;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float v = src[(long)s * K];
    if (v > bv || (v == bv && s < bi)) { bv = v; bi = s; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { shv[threadIdx.x >> 5] = bv; shi[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i)
      if (shv[i] > bv || (shv[i] == bv && shi[i] < bi)) { bv = shv[i]; bi = shi[i]; }
    const int row = bi / W;
    out[((long)n * 2 + 0) * K + k] = row;
    out[((long)n * 2 + 1) * K + k] = bi - row * W;
  }
}

inline int grid_for(long total, int threads) {
  long g = (total + threads - 1) / threads;
  long cap = (long)jcm_num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int jcm_bn_stats_blocks(long M, int C) {
  (void)C;
  long b = M / 64;
  if (b < 1) b = 1;
  long cap = (long)jcm_num_sms() * 4;
  return (int)(b < cap ? b : cap);
}

// partial: caller-owned workspace of jcm_bn_stats_blocks(M,C) * 2 * C floats
extern "C" int jcm_bn_stats(const void* x, int x_bf16, long M, int C, float* partial, void* stream) {
  JCM_CHECK_ARG(x && partial && M > 0 && C > 0, "jcm_bn_stats: bad arguments");
  const int blocks = jcm_bn_stats_blocks(M, C);
  if (x_bf16 && (C % 8) == 0 && C / 8 <= kStatThreads) {
    bn_stats_kernel<8><<<blocks, kStatThreads, kStatThreads * 16 * sizeof(float), (cudaStream_t)stream>>>(x, x_bf16, M, C, partial);
  } else if ((C % 4) == 0 && C / 4 <= kStatThreads) {
    bn_stats_kernel<4><<<blocks, kStatThreads, kStatThreads * 8 * sizeof(float), (cudaStream_t)stream>>>(x, x_bf16, M, C, partial);
  } else {
    JCM_CHECK_ARG(C <= kStatThreads, "jcm_bn_stats: C=%d not supported (must be a multiple of 4 <= 1024, or <= 256)", C);
    bn_stats_kernel<1><<<blocks, kStatThreads, kStatThreads * 2 * sizeof(float), (cudaStream_t)stream>>>(x, x_bf16, M, C, partial);
  }
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

namespace {
__global__ void colsum_from_partial_kernel(const float* __restrict__ partial, int nblocks, int C, float* __restrict__ out) {
  __shared__ double sh[kPartY][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const double s = partial_colsum(partial, nblocks, C, 0, c, sh);
  if (threadIdx.y == 0 && c < C) out[c] = (float)s;
}
}  // namespace

// out[c] = sum over rows of x [M,C] (bias gradient of the last conv layer); partial as for jcm_bn_stats
extern "C" int jcm_colsum(const float* x, long M, int C, float* partial, float* out, void* stream) {
  int rc = jcm_bn_stats(x, 0, M, C, partial, stream);
  if (rc) return rc;
  colsum_from_partial_kernel<<<jcm_cdiv(C, 32), dim3(32, kPartY), 0, (cudaStream_t)stream>>>(partial, jcm_bn_stats_blocks(M, C), C, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_bn_finalize(const float* partial, long M, int C, const float* gamma, const float* beta, float* moving_mean,
                               float* moving_var, float eps, float decay, int train, int update_moving, float* scale,
                               float* shift, float* save_mean, float* save_rstd, void* stream) {
  JCM_CHECK_ARG(gamma && beta && moving_mean && moving_var && scale && shift, "jcm_bn_finalize: null pointer");
  JCM_CHECK_ARG(!train || partial, "jcm_bn_finalize: training mode needs the partial sums");
  const int nblocks = train ? jcm_bn_stats_blocks(M, C) : 0;
  bn_finalize_kernel<<<jcm_cdiv(C, 32), dim3(32, kPartY), 0, (cudaStream_t)stream>>>(partial, nblocks, M, C, gamma, beta, moving_mean, moving_var, eps,
                                                                         decay, train, update_moving, scale, shift, save_mean, save_rstd);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_bn_apply_pool(const void* a, int a_bf16, const float* scale, const float* shift, int B, int H, int W, int C, int pool,
                                 void* out_hi, void* out_lo, float* out_f32, void* stream) {
  JCM_CHECK_ARG(a && scale && shift && (out_hi || out_f32), "jcm_bn_apply_pool: null pointer");
  JCM_CHECK_ARG((C % 4) == 0, "jcm_bn_apply_pool: C must be a multiple of 4, got %d", C);
  const int Ho = pool ? (H + 1) / 2 : H, Wo = pool ? (W + 1) / 2 : W;
  if (a_bf16 && (C % 8) == 0 && out_hi && !out_lo && !out_f32 && ((long)grid_for((long)B * Ho * Wo * (C / 8), 256) * 256) % (C / 8) == 0) {
    const long total = (long)B * Ho * Wo * (C / 8);
    if (pool)
      bn_apply_pool_bf16_kernel<1><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, scale, shift, B, H, W, C, (uint4*)out_hi);
    else
      bn_apply_pool_bf16_kernel<0><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)a, scale, shift, B, H, W, C, (uint4*)out_hi);
  } else if (a_bf16 && (C % 8) == 0) {
    const long total = (long)B * Ho * Wo * (C / 8);
    bn_apply_pool_kernel<8><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a, a_bf16, scale, shift, B, H, W, C, pool, (__nv_bfloat16*)out_hi,
                                                                                    (__nv_bfloat16*)out_lo, out_f32);
  } else {
    const long total = (long)B * Ho * Wo * (C / 4);
    bn_apply_pool_kernel<4><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(a, a_bf16, scale, shift, B, H, W, C, pool, (__nv_bfloat16*)out_hi,
                                                                                    (__nv_bfloat16*)out_lo, out_f32);
  }
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_upsample_avg3(const void* a1, const void* a2, const void* a3, int a_bf16, const float* scale_shift, int B, int H, int W,
                                 int H2, int W2, int H3, int W3, int C, void* out_hi, void* out_lo, float* out_f32, void* stream) {
  JCM_CHECK_ARG(a1 && a2 && a3 && scale_shift && (out_hi || out_f32), "jcm_upsample_avg3: null pointer");
  JCM_CHECK_ARG((C % 4) == 0, "jcm_upsample_avg3: C must be a multiple of 4, got %d", C);
  const size_t tab_bytes = (size_t)2 * (H + W) * sizeof(TapTab);
  JCM_CHECK_ARG(tab_bytes <= 40 * 1024, "jcm_upsample_avg3: map %dx%d too large for the tap tables", H, W);
  if (a_bf16 && (C % 8) == 0 && out_hi && !out_lo && !out_f32 && ((long)grid_for((long)B * H * W * (C / 8), 256) * 256) % (C / 8) == 0) {
    const long total = (long)B * H * W * (C / 8);
    upsample_avg3_bf16_kernel<<<grid_for(total, 256), 256, tab_bytes, (cudaStream_t)stream>>>(
        (const uint4*)a1, (const uint4*)a2, (const uint4*)a3, scale_shift, B, H, W, H2, W2, H3, W3, C, (uint4*)out_hi);
  } else if (a_bf16 && (C % 8) == 0) {
    const long total = (long)B * H * W * (C / 8);
    upsample_avg3_kernel<8><<<grid_for(total, 256), 256, tab_bytes, (cudaStream_t)stream>>>(a1, a2, a3, a_bf16, scale_shift, B, H, W, H2, W2, H3, W3, C,
                                                                                    (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_f32);
  } else {
    const long total = (long)B * H * W * (C / 4);
    upsample_avg3_kernel<4><<<grid_for(total, 256), 256, tab_bytes, (cudaStream_t)stream>>>(a1, a2, a3, a_bf16, scale_shift, B, H, W, H2, W2, H3, W3, C,
                                                                                    (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo, out_f32);
  }
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_spatial_softmax(const float* logits, int B, int S, int K, float* out, void* stream) {
  JCM_CHECK_ARG(logits && out && B > 0 && S > 0 && K > 0, "jcm_spatial_softmax: bad arguments");
  spatial_softmax_kernel<<<B * K, 256, 0, (cudaStream_t)stream>>>(logits, S, K, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_softmax_ce(const float* logits, const float* labels, int B, int S, int K, int KL, float* per_nk, float* lse,
                              float* loss, void* stream) {
  JCM_CHECK_ARG(logits && labels && per_nk && loss && KL >= K, "jcm_softmax_ce: bad arguments");
  softmax_ce_kernel<<<B * K, 256, 0, (cudaStream_t)stream>>>(logits, labels, S, K, KL, per_nk, lse);
  JCM_LAUNCH_CHECK();
  mean_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(per_nk, B * K, loss);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_argmax_hw(const float* hm, int B, int H, int W, int K, int* out, void* stream) {
  JCM_CHECK_ARG(hm && out && B > 0 && H > 0 && W > 0 && K > 0, "jcm_argmax_hw: bad arguments");
  argmax_hw_kernel<<<B * K, 256, 0, (cudaStream_t)stream>>>(hm, H * W, W, K, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

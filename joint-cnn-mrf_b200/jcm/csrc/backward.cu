// Backward (HBM-bound) kernels of the part detector and the loss heads.  The reference gets these from TensorFlow's
// autodiff of main.py:29-74,212-240 (`opt.compute_gradients(loss_tower)`, main.py:557-560); here they are written out:
//   jcm_softmax_ce_bwd        d(mean CE)/d logits = (softmax - y) / (B*K): the backprop output of TF1's
//                             SoftmaxCrossEntropyWithLogits op, whatever sum(y) is [TF1]                     (main.py:239)
//   jcm_spatial_softmax_bwd   dx = y * (dy - sum_s dy*y)   (joint training: spatial model input -> PD logits, main.py:523-530)
//   jcm_bn_relu_bwd           [2x2 SAME max-pool bwd] + training-mode batch-norm bwd + ReLU bwd in two passes
//                             (reduce: sum dy, sum dy*xhat per channel; apply: d_pre planes + per-block bias-grad partials)
//   jcm_colsum_finalize       deterministic reduction of per-block column partials (bias / gamma / beta gradients)
//   jcm_upsample_avg3_bwd     transpose of the legacy-bilinear up-sampling + 3-way average (gather form, no atomics)
//   jcm_pad_planes            fp32 [M,C] -> bf16 planes [M,Cpad] (zero padded channels; conv6's K-channel gradient)
//   jcm_unpack_s2d_grad       conv1 weight gradient [3][64][Cout] (x-folded space-to-depth form) -> [5,5,3,Cout]
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ float blk_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
  return r;
}

// one block per (image, joint)
__global__ void softmax_ce_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ labels, const float* __restrict__ lse,
                                      int S, int K, int KL, float scale, float* __restrict__ dlogits) {
  const int n = blockIdx.x / K, k = blockIdx.x % K;
  const float* src = logits + (long)n * S * K + k;
  const float* lab = labels + (long)n * S * KL + k;
  float* dst = dlogits + (long)n * S * K + k;
  // [TF1] tf.nn.softmax_cross_entropy_with_logits registers backprop = softmax - labels as the gradient (xent_op), NOT the
  // derivative of -sum(y * log_softmax) for unnormalised y (which would be softmax * sum(y) - y).  The two differ only for label
  // maps that do not sum to 1: the border-clipped blobs of data.py:180-186.
  const float l = lse[blockIdx.x];
  for (int s = threadIdx.x; s < S; s += blockDim.x)
    dst[(long)s * K] = scale * (expf(src[(long)s * K] - l) - lab[(long)s * KL]);
}

// y [B,S,K] softmax output, dy [B,S,KD] (first K channels), dx [B,S,K] (+= if accumulate)
__global__ void spatial_softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, int S, int K, int KD, int accumulate,
                                           float* __restrict__ dx) {
  __shared__ float sh[32];
  const int n = blockIdx.x / K, k = blockIdx.x % K;
  const float* yy = y + (long)n * S * K + k;
  const float* dd = dy + (long)n * S * KD + k;
  float* out = dx + (long)n * S * K + k;
  float dot = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) dot += yy[(long)s * K] * dd[(long)s * KD];
  dot = blk_sum(dot, sh);
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float v = yy[(long)s * K] * (dd[(long)s * KD] - dot);
    if (accumulate) out[(long)s * K] += v; else out[(long)s * K] = v;
  }
}

// ------------------------------------------------------------------------------------------------ BN + ReLU (+pool) backward
// a [B,H,W,C] = ReLU output (BN input); dout = gradient w.r.t. the BN output ([B,H,W,C], or the pooled [B,Ho,Wo,C] when pool).
// Work item = one output position of `dout` x 4 channels.  Row slabs per block as in bn_stats_kernel, so all per-channel
// partial sums are deterministic.  PASS 0: partial[blk][0][c] = sum dy, partial[blk][1][c] = sum dy*xhat.
// PASS 1: writes d_pre (gradient w.r.t. the conv output before ReLU) as bf16 planes (+fp32) and partial[blk][0][c] = sum d_pre.
struct BnBwdArgs {
  const void* a;           // ReLU output, fp32 or bf16 (a_bf16)
  int a_bf16;
  const void* dout;        // fp32, or bf16 (dout_bf16: the data-gradient convolution's bf16 output in the bf16 configuration)
  int dout_bf16;
  const float* scale;      // gamma * rstd
  const float* shift;      // beta - mean * scale
  const float* mean;
  const float* rstd;
  const float* sums;       // [2][C]: sum dy, sum dy*xhat (PASS 1)
  float dy_scale;          // constant factor folded into dout (1/3 of the bank average)
  int B, H, W, C, pool;
  float inv_count;         // 1 / (B*H*W)
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  float* d_f32;
  float* partial;          // [grid][2][C]
};

template <int PASS, int VEC, int POOL>
__global__ void __launch_bounds__(kThreads, 2) bn_relu_bwd_kernel(BnBwdArgs p) {
  extern __shared__ float sh[];  // [kThreads][2 * VEC]
  const int CG = p.C / VEC;      // channel groups of VEC channels (VEC = 8 for bf16-stored activations: 16-byte loads)
  const int lanes = kThreads / CG;
  const int g = threadIdx.x % CG, rl = threadIdx.x / CG;
  const int Ho = POOL ? (p.H + 1) / 2 : p.H, Wo = POOL ? (p.W + 1) / 2 : p.W;
  const long M = (long)p.B * Ho * Wo;
  const long per_blk = (M + gridDim.x - 1) / gridDim.x;
  const long r0 = blockIdx.x * per_blk;
  long r1 = r0 + per_blk;
  if (r1 > M) r1 = M;
  float s0[VEC], s1[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) { s0[c] = 0.f; s1[c] = 0.f; }
  if (rl < lanes) {
    float scv[VEC], sfv[VEC], muv[VEC], rsv[VEC], m_dy[VEC], m_dyx[VEC];
    load_f32_vec<VEC>(p.scale, (long)g * VEC, scv);
    load_f32_vec<VEC>(p.shift, (long)g * VEC, sfv);
    load_f32_vec<VEC>(p.mean, (long)g * VEC, muv);
    load_f32_vec<VEC>(p.rstd, (long)g * VEC, rsv);
#pragma unroll
    for (int c = 0; c < VEC; ++c) { m_dy[c] = 0.f; m_dyx[c] = 0.f; }
    if (PASS == 1) {
      load_f32_vec<VEC>(p.sums, (long)g * VEC, m_dy);
      load_f32_vec<VEC>(p.sums + p.C, (long)g * VEC, m_dyx);
#pragma unroll
      for (int c = 0; c < VEC; ++c) { m_dy[c] *= p.inv_count; m_dyx[c] *= p.inv_count; }
    }
    if (!POOL) {
      // The loop body needs three per-channel constants, not six: with xhat = (a - mean) * rstd
      //   PASS 0:  sum dy * xhat            = sum dy * (a - mean) * rstd                  -> (mean, rstd)
      //   PASS 1:  scale * (dy - m_dy - xhat * m_dyx) = ca * dy + cb + cc * a             -> (ca, cb, cc)
      // and four rows are in flight per thread (all loads first): at 126 registers the first version kept two blocks of 8 warps with
      // two 32-byte loads each per SM - a third of the bytes in flight that HBM latency x bandwidth asks for.
      float ca[VEC], cb[VEC], cc[VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        ca[c] = scv[c] * p.dy_scale;
        cc[c] = -scv[c] * m_dyx[c] * rsv[c];
        cb[c] = -scv[c] * m_dy[c] - cc[c] * muv[c];
        if (PASS == 0) { ca[c] = muv[c]; cb[c] = rsv[c] * p.dy_scale; }      // PASS 0 reuses the slots: ca = mean, cb = rstd * dy_scale
      }
      constexpr int R = (PASS == 1 && VEC == 8) ? 3 : 4;      // rows in flight (PASS 1 with 8 channels also holds the output row: 3 fit 128 registers)
      for (long r = r0 + rl; r < r1; r += (long)R * lanes) {
        float dv[R][VEC], av[R][VEC];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const long rr = r + (long)i * lanes;
          if (rr < r1) {
            load_act_vec<VEC>(p.dout, rr * p.C + (long)g * VEC, p.dout_bf16, dv[i]);
            load_act_vec<VEC>(p.a, rr * p.C + (long)g * VEC, p.a_bf16, av[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const long rr = r + (long)i * lanes;
          if (rr >= r1) break;
          float o[VEC];
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            if (PASS == 0) {
              s0[c] += dv[i][c];
              s1[c] = fmaf(dv[i][c], (av[i][c] - ca[c]) * cb[c], s1[c]);
            } else {
              const float da = fmaf(ca[c], dv[i][c], fmaf(cc[c], av[i][c], cb[c]));
              o[c] = av[i][c] > 0.f ? da : 0.f;
              s0[c] += o[c];
            }
          }
          if (PASS == 1) store_planes_vec<VEC>(p.hi, p.lo, p.d_f32, rr * p.C + (long)g * VEC, o);
        }
      }
      if (PASS == 0) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) s0[c] *= p.dy_scale;      // s1 already carries dy_scale through cb
      }
    } else {
      for (long r = r0 + rl; r < r1; r += lanes) {
        float dv[VEC];
        load_act_vec<VEC>(p.dout, r * p.C + (long)g * VEC, p.dout_bf16, dv);
#pragma unroll
        for (int c = 0; c < VEC; ++c) dv[c] *= p.dy_scale;
        int xo, yo, n;
        split_index3(r, Wo, Ho, xo, yo, n);
        const int y0 = 2 * yo, x0 = 2 * xo;
        float av[4][VEC];   // [window element][channel]
        bool ok[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int yy = y0 + (e >> 1), xx = x0 + (e & 1);
          ok[e] = yy < p.H && xx < p.W;
          if (ok[e]) {
            load_act_vec<VEC>(p.a, (((long)n * p.H + yy) * p.W + xx) * p.C + (long)g * VEC, p.a_bf16, av[e]);
          } else {
#pragma unroll
            for (int c = 0; c < VEC; ++c) av[e][c] = 0.f;
          }
        }
        float o[4][VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          int best = 0;
          float bv = fmaf(av[0][c], scv[c], sfv[c]);
#pragma unroll
          for (int e = 1; e < 4; ++e) {
            const float v = fmaf(av[e][c], scv[c], sfv[c]);
            if (ok[e] && v > bv) { bv = v; best = e; }
          }
          if (PASS == 0) {
            float ab = av[0][c];       // av[best][c] without dynamic register indexing
#pragma unroll
            for (int e = 1; e < 4; ++e) ab = (best == e) ? av[e][c] : ab;
            const float xh = (ab - muv[c]) * rsv[c];
            s0[c] += dv[c];
            s1[c] += dv[c] * xh;
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float xh = (av[e][c] - muv[c]) * rsv[c];
              const float dy = (e == best) ? dv[c] : 0.f;
              const float da = scv[c] * (dy - m_dy[c] - xh * m_dyx[c]);
              o[e][c] = (ok[e] && av[e][c] > 0.f) ? da : 0.f;
              s0[c] += o[e][c];
            }
          }
        }
        if (PASS == 1) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (!ok[e]) continue;
            const int yy = y0 + (e >> 1), xx = x0 + (e & 1);
            store_planes_vec<VEC>(p.hi, p.lo, p.d_f32, (((long)n * p.H + yy) * p.W + xx) * p.C + (long)g * VEC, o[e]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    sh[threadIdx.x * 2 * VEC + c] = s0[c];
    sh[threadIdx.x * 2 * VEC + VEC + c] = s1[c];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int gg = c / VEC, v = c % VEC;
    float t0 = 0.f, t1 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      t0 += sh[(l * CG + gg) * 2 * VEC + v];
      t1 += sh[(l * CG + gg) * 2 * VEC + VEC + v];
    }
    p.partial[((long)blockIdx.x * 2 + 0) * p.C + c] = t0;
    p.partial[((long)blockIdx.x * 2 + 1) * p.C + c] = t1;
  }
}

// Fast path of the bf16 configuration (no pooling, bf16 activation and bf16 data gradient, bf16 hi plane only): same work split,
// same partial sums and the same arithmetic as bn_relu_bwd_kernel<PASS, 8, 0>, but the two 16-byte loads of the NEXT group of rows are
// issued (as raw bf16 words) before the current group is converted and processed.  The generic kernel waits for its loads with
// nothing to do (ncu: 7.8 of 11.2 stall cycles per issue on the long scoreboard at 16 warps per SM, 3.4-3.8 TB/s).
__device__ __forceinline__ void unpack_bf16x8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <int PASS>
__global__ void __launch_bounds__(kThreads, 2) bn_relu_bwd_bf16_kernel(BnBwdArgs p) {
  extern __shared__ float sh[];  // [kThreads][16]
  constexpr int VEC = 8, R = 3;
  const int CG = p.C / VEC;
  const int lanes = kThreads / CG;
  const int g = threadIdx.x % CG, rl = threadIdx.x / CG;
  const long M = (long)p.B * p.H * p.W;
  const long per_blk = (M + gridDim.x - 1) / gridDim.x;
  const long r0 = blockIdx.x * per_blk;
  long r1 = r0 + per_blk;
  if (r1 > M) r1 = M;
  float s0[VEC], s1[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) { s0[c] = 0.f; s1[c] = 0.f; }
  if (rl < lanes) {
    float ca[VEC], cb[VEC], cc[VEC];
    {
      float scv[VEC], muv[VEC], rsv[VEC], m_dy[VEC], m_dyx[VEC];
      load_f32_vec<VEC>(p.scale, (long)g * VEC, scv);
      load_f32_vec<VEC>(p.mean, (long)g * VEC, muv);
      load_f32_vec<VEC>(p.rstd, (long)g * VEC, rsv);
#pragma unroll
      for (int c = 0; c < VEC; ++c) { m_dy[c] = 0.f; m_dyx[c] = 0.f; }
      if (PASS == 1) {
        load_f32_vec<VEC>(p.sums, (long)g * VEC, m_dy);
        load_f32_vec<VEC>(p.sums + p.C, (long)g * VEC, m_dyx);
      }
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        if (PASS == 0) {
          ca[c] = muv[c];
          cb[c] = rsv[c] * p.dy_scale;
          cc[c] = 0.f;
        } else {
          const float mdy = m_dy[c] * p.inv_count, mdyx = m_dyx[c] * p.inv_count;
          ca[c] = scv[c] * p.dy_scale;
          cc[c] = -scv[c] * mdyx * rsv[c];
          cb[c] = -scv[c] * mdy - cc[c] * muv[c];
        }
      }
    }
    const uint4* dptr = reinterpret_cast<const uint4*>(p.dout);
    const uint4* aptr = reinterpret_cast<const uint4*>(p.a);
    uint4* optr = reinterpret_cast<uint4*>(p.hi);
    const long stride = (long)R * lanes;
    uint4 nd[R], na[R];
    long r = r0 + rl;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const long rr = r + (long)i * lanes;
      if (rr < r1) {
        nd[i] = dptr[(rr * p.C >> 3) + g];
        na[i] = aptr[(rr * p.C >> 3) + g];
      }
    }
    for (; r < r1; r += stride) {
      uint4 cd[R], cu[R];
#pragma unroll
      for (int i = 0; i < R; ++i) { cd[i] = nd[i]; cu[i] = na[i]; }
      const long rn = r + stride;
#pragma unroll
      for (int i = 0; i < R; ++i) {      // next group's loads fly while this group is processed
        const long rr = rn + (long)i * lanes;
        if (rr < r1) {
          nd[i] = dptr[(rr * p.C >> 3) + g];
          na[i] = aptr[(rr * p.C >> 3) + g];
        }
      }
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const long rr = r + (long)i * lanes;
        if (rr >= r1) break;
        float dv[VEC], av[VEC];
        unpack_bf16x8(cd[i], dv);
        unpack_bf16x8(cu[i], av);
        if (PASS == 0) {
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            s0[c] += dv[c];
            s1[c] = fmaf(dv[c], (av[c] - ca[c]) * cb[c], s1[c]);
          }
        } else {
          __align__(16) __nv_bfloat16 h[VEC];
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            const float da = fmaf(ca[c], dv[c], fmaf(cc[c], av[c], cb[c]));
            const float o = av[c] > 0.f ? da : 0.f;
            s0[c] += o;
            h[c] = __float2bfloat16_rn(o);
          }
          optr[(rr * p.C >> 3) + g] = *reinterpret_cast<uint4*>(h);
        }
      }
    }
    if (PASS == 0) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) s0[c] *= p.dy_scale;
    }
  }
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    sh[threadIdx.x * 2 * VEC + c] = s0[c];
    sh[threadIdx.x * 2 * VEC + VEC + c] = s1[c];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int gg = c / VEC, v = c % VEC;
    float t0 = 0.f, t1 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      t0 += sh[(l * CG + gg) * 2 * VEC + v];
      t1 += sh[(l * CG + gg) * 2 * VEC + VEC + v];
    }
    p.partial[((long)blockIdx.x * 2 + 0) * p.C + c] = t0;
    p.partial[((long)blockIdx.x * 2 + 1) * p.C + c] = t1;
  }
}

// The pooled layers (conv1_*, conv2_*: the largest activations) of the bf16 configuration, with the same prefetching scheme: the work
// item is one POOLED position x 8 channels = one 16-byte load of the data gradient + the four 16-byte loads of its 2x2 window; the
// next item's five loads are in flight while the current one is processed.  Arithmetic and partial sums as bn_relu_bwd_kernel<PASS, 8, 1>.
template <int PASS>
__global__ void __launch_bounds__(kThreads, 2) bn_relu_bwd_pool_bf16_kernel(BnBwdArgs p) {
  extern __shared__ float sh[];  // [kThreads][16]
  constexpr int VEC = 8;
  const int CG = p.C / VEC;
  const int lanes = kThreads / CG;
  const int g = threadIdx.x % CG, rl = threadIdx.x / CG;
  const int Ho = (p.H + 1) / 2, Wo = (p.W + 1) / 2;
  const long M = (long)p.B * Ho * Wo;
  const long per_blk = (M + gridDim.x - 1) / gridDim.x;
  const long r0 = blockIdx.x * per_blk;
  long r1 = r0 + per_blk;
  if (r1 > M) r1 = M;
  float s0[VEC], s1[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) { s0[c] = 0.f; s1[c] = 0.f; }
  if (rl < lanes) {
    float scv[VEC], sfv[VEC], ca[VEC], cb[VEC], cc[VEC];
    {
      float muv[VEC], rsv[VEC], m_dy[VEC], m_dyx[VEC];
      load_f32_vec<VEC>(p.scale, (long)g * VEC, scv);
      load_f32_vec<VEC>(p.shift, (long)g * VEC, sfv);
      load_f32_vec<VEC>(p.mean, (long)g * VEC, muv);
      load_f32_vec<VEC>(p.rstd, (long)g * VEC, rsv);
#pragma unroll
      for (int c = 0; c < VEC; ++c) { m_dy[c] = 0.f; m_dyx[c] = 0.f; }
      if (PASS == 1) {
        load_f32_vec<VEC>(p.sums, (long)g * VEC, m_dy);
        load_f32_vec<VEC>(p.sums + p.C, (long)g * VEC, m_dyx);
      }
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        if (PASS == 0) {
          ca[c] = muv[c];                 // sum dy * xhat = sum dy * (a - mean) * rstd
          cb[c] = rsv[c] * p.dy_scale;
          cc[c] = 0.f;
        } else {
          // d_pre(e) = scale * (dy(e) - m_dy - xhat(e) * m_dyx) = ca * dy(e) + cb + cc * a(e),  dy(e) = dout for the arg-max element, else 0
          const float mdy = m_dy[c] * p.inv_count, mdyx = m_dyx[c] * p.inv_count;
          ca[c] = scv[c] * p.dy_scale;
          cc[c] = -scv[c] * mdyx * rsv[c];
          cb[c] = -scv[c] * mdy - cc[c] * muv[c];
        }
      }
    }
    const uint4* dptr = reinterpret_cast<const uint4*>(p.dout);
    const uint4* aptr = reinterpret_cast<const uint4*>(p.a);
    uint4* optr = reinterpret_cast<uint4*>(p.hi);
    const int c8 = p.C >> 3;
    // window of pooled position r: index (in 16-byte units) of its top-left element and a 4-bit mask of the elements inside the map
    // (32-bit indices: the launcher routes tensors of 2^31 or more 16-byte units to the generic kernel)
    const int row8 = p.W * c8;
    auto locate = [&](long r, int& base, unsigned& okm) {
      int xo, yo, n;
      split_index3(r, Wo, Ho, xo, yo, n);
      const int y0 = 2 * yo, x0 = 2 * xo;
      base = ((n * p.H + y0) * p.W + x0) * c8 + g;
      okm = 1u | (x0 + 1 < p.W ? 2u : 0u) | (y0 + 1 < p.H ? 4u : 0u) | ((x0 + 1 < p.W && y0 + 1 < p.H) ? 8u : 0u);
    };
    uint4 nd, na[4];
    int nbase = 0;
    unsigned nokm = 0;
    long r = r0 + rl;
    if (r < r1) {
      locate(r, nbase, nokm);
      nd = dptr[r * c8 + g];
#pragma unroll
      for (int e = 0; e < 4; ++e) na[e] = ((nokm >> e) & 1u) ? aptr[nbase + (e & 1) * c8 + (e >> 1) * row8] : make_uint4(0u, 0u, 0u, 0u);
    }
    for (; r < r1; r += lanes) {
      const uint4 cd = nd;
      uint4 cu[4];
      const int base = nbase;
      const unsigned okm = nokm;
#pragma unroll
      for (int e = 0; e < 4; ++e) cu[e] = na[e];
      const long rn = r + lanes;
      if (rn < r1) {                      // the next item's loads fly while this one is processed
        locate(rn, nbase, nokm);
        nd = dptr[rn * c8 + g];
#pragma unroll
        for (int e = 0; e < 4; ++e) na[e] = ((nokm >> e) & 1u) ? aptr[nbase + (e & 1) * c8 + (e >> 1) * row8] : make_uint4(0u, 0u, 0u, 0u);
      }
      bool ok[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) ok[e] = (okm >> e) & 1u;
      float dv[VEC], av[4][VEC];
      unpack_bf16x8(cd, dv);
#pragma unroll
      for (int e = 0; e < 4; ++e) unpack_bf16x8(cu[e], av[e]);
      __align__(16) __nv_bfloat16 h[4][VEC];
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        int best = 0;
        float bv = fmaf(av[0][c], scv[c], sfv[c]);
#pragma unroll
        for (int e = 1; e < 4; ++e) {
          const float v = fmaf(av[e][c], scv[c], sfv[c]);
          if (ok[e] && v > bv) { bv = v; best = e; }
        }
        if (PASS == 0) {
          float ab = av[0][c];
#pragma unroll
          for (int e = 1; e < 4; ++e) ab = (best == e) ? av[e][c] : ab;
          s0[c] += dv[c];
          s1[c] = fmaf(dv[c], (ab - ca[c]) * cb[c], s1[c]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float da = fmaf(cc[c], av[e][c], cb[c]) + ((e == best) ? ca[c] * dv[c] : 0.f);
            const float o = (ok[e] && av[e][c] > 0.f) ? da : 0.f;
            s0[c] += o;
            h[e][c] = __float2bfloat16_rn(o);
          }
        }
      }
      if (PASS == 1) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (ok[e]) optr[base + (e & 1) * c8 + (e >> 1) * row8] = *reinterpret_cast<uint4*>(h[e]);
      }
    }
    if (PASS == 0) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) s0[c] *= p.dy_scale;
    }
  }
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    sh[threadIdx.x * 2 * VEC + c] = s0[c];
    sh[threadIdx.x * 2 * VEC + VEC + c] = s1[c];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
    const int gg = c / VEC, v = c % VEC;
    float t0 = 0.f, t1 = 0.f;
    for (int l = 0; l < lanes; ++l) {
      t0 += sh[(l * CG + gg) * 2 * VEC + v];
      t1 += sh[(l * CG + gg) * 2 * VEC + VEC + v];
    }
    p.partial[((long)blockIdx.x * 2 + 0) * p.C + c] = t0;
    p.partial[((long)blockIdx.x * 2 + 1) * p.C + c] = t1;
  }
}

// sums[j][c] = sum over blocks of partial[blk][j][c] (double accumulation); optionally scaled by mul[c]
__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int nblocks, int C, int rows, float* __restrict__ sums,
                                       float* __restrict__ copy0, float* __restrict__ copy1) {
  __shared__ double sh[kPartY][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = 0; j < rows; ++j) {
    const double s = partial_colsum(partial, nblocks, C, j, c, sh);
    if (threadIdx.y == 0 && c < C) {
      sums[j * C + c] = (float)s;
      float* cp = j == 0 ? copy0 : copy1;
      if (cp) cp[c] = (float)s;
    }
  }
}

// ------------------------------------------------------------------------------------------------ upsample/average backward
__device__ __forceinline__ void legacy_tap(int dst, int n_in, int n_out, int& lo, int& hi, float& w) {
  const float scale = (float)n_in / (float)n_out;
  const float src = (float)dst * scale;
  lo = (int)floorf(src);
  hi = min(lo + 1, n_in - 1);
  w = src - (float)lo;
}

// dlow[n,p,q,c] = mul * sum over (y,x) of dm[n,y,x,c] * wy(y->p) * wx(x->q)   (transpose of resize_images [Hi,Wi] -> [H,W])
// Gather form, no atomics.  Per block, shared-memory tables hold the (lo, hi, w) taps of every destination row / column and, for
// every source row / column, the range of destination indices that touch it - so an item visits exactly its contributing
// destinations (3 x 3 for the 1/2 bank, 7 x 7 for the 1/4 bank) with table look-ups instead of recomputing `(float)n_in / n_out`
// and a floor per candidate (the first version: 0.72 ms for the two launches).  dm: fp32 (VEC 4) or bf16 (VEC 8).
struct TapTabB { int lo, hi; float w; };
template <int VEC>
__global__ void upsample_bwd_kernel(const void* __restrict__ dm, int dm_bf16, int B, int H, int W, int Hi, int Wi, int C, float mul,
                                    float* __restrict__ dlow) {
  extern __shared__ int sh_raw[];
  TapTabB* ty = reinterpret_cast<TapTabB*>(sh_raw);     // [H]
  TapTabB* tx = ty + H;                                 // [W]
  int* yfirst = reinterpret_cast<int*>(tx + W);         // [Hi] first / last destination row whose lo or hi tap is this source row
  int* ylast = yfirst + Hi;
  int* xfirst = ylast + Hi;                             // [Wi]
  int* xlast = xfirst + Wi;
  for (int i = threadIdx.x; i < H + W; i += blockDim.x) {
    TapTabB t;
    if (i < H) legacy_tap(i, Hi, H, t.lo, t.hi, t.w); else legacy_tap(i - H, Wi, W, t.lo, t.hi, t.w);
    ty[i] = t;      // ty and tx are contiguous
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Hi + Wi; i += blockDim.x) {
    const bool row = i < Hi;
    const int src = row ? i : i - Hi, n_dst = row ? H : W;
    const TapTabB* t = row ? ty : tx;
    int f = n_dst, l = -1;
    for (int d = 0; d < n_dst; ++d)
      if (t[d].lo == src || t[d].hi == src) { if (d < f) f = d; l = d; }
    if (row) { yfirst[src] = f; ylast[src] = l; } else { xfirst[src] = f; xlast[src] = l; }
  }
  __syncthreads();
  const int CG = C / VEC;
  const long total = (long)B * Hi * Wi * CG;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int cg, q, pp, n;
    split_index(i, CG, Wi, Hi, cg, q, pp, n);
    float acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
    for (int y = yfirst[pp]; y <= ylast[pp]; ++y) {
      const TapTabB t = ty[y];
      const float wyv = (t.lo == pp ? 1.f - t.w : 0.f) + (t.hi == pp ? t.w : 0.f);
      if (wyv == 0.f) continue;
      for (int x = xfirst[q]; x <= xlast[q]; ++x) {
        const TapTabB u = tx[x];
        const float wxv = (u.lo == q ? 1.f - u.w : 0.f) + (u.hi == q ? u.w : 0.f);
        if (wxv == 0.f) continue;
        float d[VEC];
        load_act_vec<VEC>(dm, (((long)n * H + y) * W + x) * C + (long)cg * VEC, dm_bf16, d);
        const float ww = wyv * wxv;
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] += d[c] * ww;
      }
    }
#pragma unroll
    for (int c = 0; c < VEC / 4; ++c)
      reinterpret_cast<float4*>(dlow)[i * (VEC / 4) + c] = make_float4(acc[4 * c] * mul, acc[4 * c + 1] * mul, acc[4 * c + 2] * mul, acc[4 * c + 3] * mul);
  }
}

__global__ void pad_planes_kernel(const float* __restrict__ x, long M, int C, int Cpad, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo) {
  const long total = M * Cpad;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % Cpad);
    const long r = i / Cpad;
    const float v = c < C ? x[r * C + c] : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// g9 [3][64][Cout] (fp32, x-folded s2d form as jcm_conv2d_wgrad writes it: vertical tap, channel, cout) -> dw [5][5][3][Cout]
__global__ void unpack_s2d_grad_kernel(const float* __restrict__ g9, int Cout, float* __restrict__ dw) {
  const int total = 75 * Cout;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int co = idx % Cout;
    const int r = idx / Cout;
    const int ci = r % 3, tx = (r / 3) % 5, ty = r / 15;
    const int by = (ty + 1) >> 1, sy = (ty + 1) & 1, bx = (tx + 1) >> 1, sx = (tx + 1) & 1;
    dw[idx] = g9[((by * 64 + bx * 16) + (sy * 2 + sx) * 3 + ci) * Cout + co];
  }
}

inline int grid_for(long total, int threads) {
  long g = (total + threads - 1) / threads;
  long cap = (long)jcm_num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int jcm_softmax_ce_bwd(const float* logits, const float* labels, const float* lse, int B, int S, int K, int KL, float scale,
                                  float* dlogits, void* stream) {
  JCM_CHECK_ARG(logits && labels && lse && dlogits && KL >= K, "jcm_softmax_ce_bwd: bad arguments");
  softmax_ce_bwd_kernel<<<B * K, 256, 0, (cudaStream_t)stream>>>(logits, labels, lse, S, K, KL, scale, dlogits);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_spatial_softmax_bwd(const float* y, const float* dy, int B, int S, int K, int KD, int accumulate, float* dx,
                                       void* stream) {
  JCM_CHECK_ARG(y && dy && dx && KD >= K, "jcm_spatial_softmax_bwd: bad arguments");
  spatial_softmax_bwd_kernel<<<B * K, 256, 0, (cudaStream_t)stream>>>(y, dy, S, K, KD, accumulate, dx);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_bn_relu_bwd_blocks(long M_out, int C) {
  (void)C;
  // two blocks per SM: that is what the kernels' register budgets keep resident, so the grid is one wave - and the deterministic
  // finalize kernels that follow read 296 partial rows per channel instead of 1184 (they cost 13 us each, 26 per step)
  long b = M_out / 32;
  if (b < 1) b = 1;
  long cap = (long)jcm_num_sms() * 2;
  return (int)(b < cap ? b : cap);
}

// Two passes.  workspace: 2 * blocks * 2 * C + 2 * C floats (partials of both passes + the reduced sums).
// Outputs: d_pre planes [B,H,W,C] (hi[, lo]) and optionally fp32; dgamma[C], dbeta[C], dbias[C] (= column sums of d_pre).
extern "C" int jcm_bn_relu_bwd(const void* a, int a_bf16, const void* dout, int dout_bf16, const float* scale, const float* shift, const float* mean,
                               const float* rstd, float dy_scale, int B, int H, int W, int C, int pool, void* d_hi, void* d_lo,
                               float* d_f32, float* dgamma, float* dbeta, float* dbias, float* workspace, void* stream) {
  JCM_CHECK_ARG(a && dout && scale && shift && mean && rstd && d_hi && dgamma && dbeta && dbias && workspace, "jcm_bn_relu_bwd: null pointer");
  JCM_CHECK_ARG((C % 4) == 0 && C / 4 <= kThreads, "jcm_bn_relu_bwd: C must be a multiple of 4 and <= 1024 (got %d)", C);
  const int Ho = pool ? (H + 1) / 2 : H, Wo = pool ? (W + 1) / 2 : W;
  const long M_out = (long)B * Ho * Wo;
  const int blocks = jcm_bn_relu_bwd_blocks(M_out, C);
  cudaStream_t st = (cudaStream_t)stream;
  float* part0 = workspace;
  float* part1 = workspace + (long)blocks * 2 * C;
  float* sums = part1 + (long)blocks * 2 * C;
  BnBwdArgs p;
  p.a = a; p.a_bf16 = a_bf16; p.dout = dout; p.dout_bf16 = dout_bf16; p.scale = scale; p.shift = shift; p.mean = mean; p.rstd = rstd; p.sums = sums; p.dy_scale = dy_scale;
  p.B = B; p.H = H; p.W = W; p.C = C; p.pool = pool; p.inv_count = 1.0f / (float)((long)B * H * W);
  p.hi = (__nv_bfloat16*)d_hi; p.lo = (__nv_bfloat16*)d_lo; p.d_f32 = d_f32; p.partial = part0;
  // bf16-stored activations with C a multiple of 8: 8 channels per thread (16-byte loads); else 4
  const bool v8 = a_bf16 && (C % 8) == 0 && C / 8 <= kThreads;
  const size_t shb = kThreads * (v8 ? 16 : 8) * sizeof(float);
  const bool fast = v8 && dout_bf16 && !d_lo && !d_f32 && (long)B * H * W * (C / 8) < (1L << 31);      // bf16 configuration: the prefetching kernels
  auto launch = [&](int pass) {
    if (fast && !pool) {
      if (pass == 0) bn_relu_bwd_bf16_kernel<0><<<blocks, kThreads, shb, st>>>(p);
      else bn_relu_bwd_bf16_kernel<1><<<blocks, kThreads, shb, st>>>(p);
    } else if (fast) {
      if (pass == 0) bn_relu_bwd_pool_bf16_kernel<0><<<blocks, kThreads, shb, st>>>(p);
      else bn_relu_bwd_pool_bf16_kernel<1><<<blocks, kThreads, shb, st>>>(p);
    } else if (pass == 0) {
      if (v8) { if (pool) bn_relu_bwd_kernel<0, 8, 1><<<blocks, kThreads, shb, st>>>(p); else bn_relu_bwd_kernel<0, 8, 0><<<blocks, kThreads, shb, st>>>(p); }
      else { if (pool) bn_relu_bwd_kernel<0, 4, 1><<<blocks, kThreads, shb, st>>>(p); else bn_relu_bwd_kernel<0, 4, 0><<<blocks, kThreads, shb, st>>>(p); }
    } else {
      if (v8) { if (pool) bn_relu_bwd_kernel<1, 8, 1><<<blocks, kThreads, shb, st>>>(p); else bn_relu_bwd_kernel<1, 8, 0><<<blocks, kThreads, shb, st>>>(p); }
      else { if (pool) bn_relu_bwd_kernel<1, 4, 1><<<blocks, kThreads, shb, st>>>(p); else bn_relu_bwd_kernel<1, 4, 0><<<blocks, kThreads, shb, st>>>(p); }
    }
  };
  launch(0);
  JCM_LAUNCH_CHECK();
  colsum_finalize_kernel<<<jcm_cdiv(C, 32), dim3(32, kPartY), 0, st>>>(part0, blocks, C, 2, sums, dbeta, dgamma);   // dbeta = sum dy, dgamma = sum dy * xhat
  JCM_LAUNCH_CHECK();
  p.partial = part1;
  launch(1);
  JCM_LAUNCH_CHECK();
  colsum_finalize_kernel<<<jcm_cdiv(C, 32), dim3(32, kPartY), 0, st>>>(part1, blocks, C, 1, dbias, nullptr, nullptr);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// dmerged: fp32, or bf16 when dm_bf16 (C a multiple of 8)
extern "C" int jcm_upsample_avg3_bwd(const void* dmerged, int dm_bf16, int B, int H, int W, int H2, int W2, int H3, int W3, int C, float* d2,
                                     float* d3, void* stream) {
  JCM_CHECK_ARG(dmerged && d2 && d3 && (C % 4) == 0 && (!dm_bf16 || (C % 8) == 0), "jcm_upsample_avg3_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int Hs[2] = {H2, H3}, Ws[2] = {W2, W3};
  float* outs[2] = {d2, d3};
  for (int k = 0; k < 2; ++k) {
    const size_t shb = (size_t)(H + W) * sizeof(TapTabB) + (size_t)2 * (Hs[k] + Ws[k]) * sizeof(int);
    JCM_CHECK_ARG(shb <= 40 * 1024, "jcm_upsample_avg3_bwd: map too large for the tap tables");
    if (dm_bf16)
      upsample_bwd_kernel<8><<<grid_for((long)B * Hs[k] * Ws[k] * (C / 8), 256), 256, shb, st>>>(dmerged, 1, B, H, W, Hs[k], Ws[k], C, 1.0f / 3.0f, outs[k]);
    else
      upsample_bwd_kernel<4><<<grid_for((long)B * Hs[k] * Ws[k] * (C / 4), 256), 256, shb, st>>>(dmerged, 0, B, H, W, Hs[k], Ws[k], C, 1.0f / 3.0f, outs[k]);
    JCM_LAUNCH_CHECK();
  }
  return JCM_OK;
}

extern "C" int jcm_pad_planes(const float* x, long M, int C, int Cpad, void* hi, void* lo, void* stream) {
  JCM_CHECK_ARG(x && hi && Cpad >= C && M > 0, "jcm_pad_planes: bad arguments");
  pad_planes_kernel<<<grid_for(M * Cpad, 256), 256, 0, (cudaStream_t)stream>>>(x, M, C, Cpad, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_unpack_s2d_grad(const float* g9, int Cout, float* dw, void* stream) {
  JCM_CHECK_ARG(g9 && dw && Cout > 0, "jcm_unpack_s2d_grad: bad arguments");
  unpack_s2d_grad_kernel<<<grid_for(75L * Cout, 256), 256, 0, (cudaStream_t)stream>>>(g9, Cout, dw);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

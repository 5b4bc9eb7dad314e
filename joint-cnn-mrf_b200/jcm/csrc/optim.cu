// Optimizer step on flat fp32 parameter / gradient buffers (all trainable variables are views into one buffer, so the
// data-parallel gradient exchange is ONE NCCL all-reduce and the update is two kernels):
//   jcm_grad_prepare   g <- g * inv_world (mean over replicas, main.py:243-267) + lmbd * w on the weight-decayed prefix
//                      (d/dw of lmbd * sum tf.nn.l2_loss(w), main.py:195-205,541); block partials of sum g^2 and sum w^2/2
//   jcm_grad_finish    global norm (tf.clip_by_global_norm, main.py:302-309) and the weight-decay loss term
//   jcm_clip_adam      g * clip / max(norm, clip), then TF1 Adam (lr_t = lr sqrt(1-b2^t)/(1-b1^t), eps outside the sqrt)
//                      or TF MomentumOptimizer(0.9)  (main.py:501-506,577)
#include "common.cuh"

namespace {
constexpr int kT = 256;

__global__ void grad_prepare_kernel(float* __restrict__ g, const float* __restrict__ w, long n, long n_decay, float inv_world, float lmbd,
                                    float* __restrict__ partial /*[grid][2]*/) {
  __shared__ float sh[2][kT / 32];
  double sg = 0.0, sw = 0.0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gv = g[i] * inv_world;
    if (i < n_decay) {
      const float wv = w[i];
      gv = fmaf(lmbd, wv, gv);
      sw += 0.5 * (double)wv * (double)wv;
    }
    g[i] = gv;
    sg += (double)gv * (double)gv;
  }
  float a = warp_sum((float)sg), b = warp_sum((float)sw);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = a; sh[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
    for (int i = 0; i < kT / 32; ++i) { ta += sh[0][i]; tb += sh[1][i]; }
    partial[blockIdx.x * 2 + 0] = ta;
    partial[blockIdx.x * 2 + 1] = tb;
  }
}

// stats[0] = global norm, stats[1] = sum w^2/2 over the decayed prefix
__global__ void grad_finish_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ stats) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < nblocks; ++i) { a += (double)partial[2 * i]; b += (double)partial[2 * i + 1]; }
    stats[0] = (float)sqrt(a);
    stats[1] = (float)b;
  }
}

__global__ void clip_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                                 const float* __restrict__ stats, float clip, float lr_t, float b1, float b2, float eps, int momentum) {
  const float norm = stats[0];
  const float scale = clip > 0.f ? clip / fmaxf(norm, clip) : 1.f;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float gv = g[i] * scale;
    if (momentum) {
      const float acc = fmaf(b1, m[i], gv);   // accum = momentum * accum + grad ; var -= lr * accum
      m[i] = acc;
      w[i] -= lr_t * acc;
    } else {
      const float mv = fmaf(b1, m[i], (1.f - b1) * gv);
      const float vv = fmaf(b2, v[i], (1.f - b2) * gv * gv);
      m[i] = mv;
      v[i] = vv;
      w[i] -= lr_t * mv / (sqrtf(vv) + eps);
    }
  }
}

// partial[blk] = sum of x^2 over the block's grid-stride share (double accumulation per thread, fixed order: deterministic)
__global__ void sumsq_kernel(const float* __restrict__ x, long n, float* __restrict__ partial) {
  __shared__ float sh[kT / 32];
  double s = 0.0;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) s += (double)x[i] * (double)x[i];
  const float a = warp_sum((float)s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kT / 32; ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}
__global__ void sumsq_finish_kernel(const float* __restrict__ partial, int nblocks, float mul, int accumulate, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < nblocks; ++i) a += (double)partial[i];
    out[0] = (float)(a * (double)mul) + (accumulate ? out[0] : 0.f);
  }
}
// out = g * clip / max(sqrt(sumsq[0]), clip)      (tf.clip_by_global_norm, main.py:302-309)
__global__ void clip_scale_kernel(const float* __restrict__ g, long n, const float* __restrict__ sumsq, float clip, float* __restrict__ out) {
  const float scale = clip / fmaxf(sqrtf(sumsq[0]), clip);
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) out[i] = g[i] * scale;
}
constexpr int kMaxTowers = 16;
struct TowerPtrs { const float* t[kMaxTowers]; };
// out = mean over towers, summed in tower order (expand_dims / concat / reduce_mean of main.py:253-259)
__global__ void tower_mean_kernel(TowerPtrs tp, int n_towers, long n, float* __restrict__ out) {
  const float inv = 1.0f / (float)n_towers;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float s = tp.t[0][i];
    for (int k = 1; k < n_towers; ++k) s += tp.t[k][i];
    out[i] = s * inv;
  }
}
// y[b, i, j, :] = x[b, 2i + oy, 2j + ox, :]
__global__ void subsample2_kernel(const float* __restrict__ x, int B, int H, int W, int C, int oy, int ox, int Ho, int Wo,
                                  float* __restrict__ y) {
  const long total = (long)B * Ho * Wo * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long t = i / C;
    const int j = (int)(t % Wo);
    t /= Wo;
    const int r = (int)(t % Ho);
    const int b = (int)(t / Ho);
    y[i] = x[(((long)b * H + 2 * r + oy) * W + 2 * j + ox) * C + c];
  }
}

// out[r, c] = [relu](x[r, c] + bias[c])
__global__ void bias_relu_kernel(const float* __restrict__ x, const float* __restrict__ bias, long M, int C, int relu, float* __restrict__ out) {
  const long total = M * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i] + bias[(int)(i % C)];
    out[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

inline int blocks_for(long n) {
  long b = (n + kT * 4 - 1) / (kT * 4);
  long cap = (long)jcm_num_sms() * 8;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}
}  // namespace

extern "C" int jcm_optim_blocks(long n) { return blocks_for(n); }

// partial: jcm_optim_blocks(n)*2 floats; stats: 2 floats (norm, weight-decay loss term)
extern "C" int jcm_grad_prepare(float* g, const float* w, long n, long n_decay, float inv_world, float lmbd, float* partial,
                                float* stats, void* stream) {
  JCM_CHECK_ARG(g && w && partial && stats && n > 0 && n_decay >= 0 && n_decay <= n, "jcm_grad_prepare: bad arguments");
  const int nb = blocks_for(n);
  grad_prepare_kernel<<<nb, kT, 0, (cudaStream_t)stream>>>(g, w, n, n_decay, inv_world, lmbd, partial);
  JCM_LAUNCH_CHECK();
  grad_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partial, nb, stats);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_clip_adam(float* w, const float* g, float* m, float* v, long n, const float* stats, float clip, float lr_t,
                             float b1, float b2, float eps, int momentum, void* stream) {
  JCM_CHECK_ARG(w && g && m && (v || momentum) && stats && n > 0, "jcm_clip_adam: bad arguments");
  clip_adam_kernel<<<blocks_for(n), kT, 0, (cudaStream_t)stream>>>(w, g, m, v, n, stats, clip, lr_t, b1, b2, eps, momentum);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// out[0] (+)= mul * sum x^2.  weight_decay(var_pattern) of main.py:195-205 is the sum of tf.nn.l2_loss = sum(w^2)/2 over the matching
// variables: mul = 0.5, accumulate != 0 from the second variable on; the squared global norm of main.py:302-309: mul = 1.
// partial: jcm_optim_blocks(n) floats.
extern "C" int jcm_sumsq(const float* x, long n, float mul, int accumulate, float* partial, float* out, void* stream) {
  JCM_CHECK_ARG(x && partial && out && n > 0, "jcm_sumsq: bad arguments");
  const int nb = blocks_for(n);
  sumsq_kernel<<<nb, kT, 0, (cudaStream_t)stream>>>(x, n, partial);
  JCM_LAUNCH_CHECK();
  sumsq_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partial, nb, mul, accumulate, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// out = g * clip / max(sqrt(sumsq[0]), clip): grad_renorm / tf.clip_by_global_norm (main.py:302-309) for one tensor of the list,
// sumsq[0] = squared global norm over the WHOLE list (accumulated with jcm_sumsq).  out may alias g.
extern "C" int jcm_clip_scale(const float* g, long n, const float* sumsq, float clip, float* out, void* stream) {
  JCM_CHECK_ARG(g && sumsq && out && n > 0 && clip > 0.f, "jcm_clip_scale: bad arguments");
  clip_scale_kernel<<<blocks_for(n), kT, 0, (cudaStream_t)stream>>>(g, n, sumsq, clip, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// out = mean of n_towers same-sized gradient tensors that live on THIS device (average_gradients of main.py:243-267 for towers of
// one process; across processes the mean is the NCCL all-reduce in jcm.train.Trainer).  towers: HOST array of device pointers.
extern "C" int jcm_tower_mean(const float* const* towers, int n_towers, long n, float* out, void* stream) {
  JCM_CHECK_ARG(towers && out && n > 0 && n_towers >= 1 && n_towers <= kMaxTowers, "jcm_tower_mean: 1..%d towers supported, got %d",
                kMaxTowers, n_towers);
  TowerPtrs tp;
  for (int k = 0; k < kMaxTowers; ++k) tp.t[k] = k < n_towers ? towers[k] : nullptr;
  for (int k = 0; k < n_towers; ++k) JCM_CHECK_ARG(tp.t[k] != nullptr, "jcm_tower_mean: null tower pointer");
  tower_mean_kernel<<<blocks_for(n), kT, 0, (cudaStream_t)stream>>>(tp, n_towers, n, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// y [B, (H - oy + 1) / 2, (W - ox + 1) / 2, C] = x[:, oy::2, ox::2, :].  A stride-2 SAME convolution (tf.nn.conv2d strides [1,2,2,1],
// main.py:133-135) equals the stride-1 SAME convolution sampled at rows/columns 2i + o with o = 1 for an even input extent (TF pads
// (k-3)/2 before) and o = 0 for an odd one; the part detector's own stride-2 layers (conv1_*) do not use this - they run on the
// space-to-depth planes of jcm_prep_input - it serves the stand-alone conv2d(x, W, 2) of the function surface.
extern "C" int jcm_subsample2(const float* x, int B, int H, int W, int C, int oy, int ox, float* y, void* stream) {
  JCM_CHECK_ARG(x && y && B > 0 && H > 0 && W > 0 && C > 0 && (oy == 0 || oy == 1) && (ox == 0 || ox == 1) && oy < H && ox < W,
                "jcm_subsample2: bad arguments");
  const int Ho = (H - oy + 1) / 2, Wo = (W - ox + 1) / 2;
  subsample2_kernel<<<blocks_for((long)B * Ho * Wo * C), kT, 0, (cudaStream_t)stream>>>(x, B, H, W, C, oy, ox, Ho, Wo, y);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// out = [relu](x + bias) per channel of x [M, C]: the `conv2d(x, w, stride) + b` / tf.nn.relu of conv_layer (main.py:160-162) for the
// stand-alone function surface (inside model() bias and ReLU are the convolution kernel's epilogue).  out may alias x.
extern "C" int jcm_bias_relu(const float* x, const float* bias, long M, int C, int relu, float* out, void* stream) {
  JCM_CHECK_ARG(x && bias && out && M > 0 && C > 0, "jcm_bias_relu: bad arguments");
  bias_relu_kernel<<<blocks_for(M * C), kT, 0, (cudaStream_t)stream>>>(x, bias, M, C, relu, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

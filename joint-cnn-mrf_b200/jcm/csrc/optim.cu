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

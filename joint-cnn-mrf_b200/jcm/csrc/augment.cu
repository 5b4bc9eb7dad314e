// In-graph training augmentation of the reference (augmentation.py:12-77, called at main.py:495-499 right before the tower
// forward) as HBM-bound sm_100a kernels on NHWC fp32 tensors.  SURVEY 8(f2).  The random draws are made by the caller and
// passed per image as prm[B][8] = {flip (0/1), brightness delta, contrast factor, cos(angle), sin(angle), rh, rw, unused}.
//   jcm_augment_color        flip_left_right + random_brightness (x + delta) + random_contrast ((x - mean_hw) f + mean_hw) + clip [0,1]
//   jcm_augment_flip_channels flip_left_right of the heat maps + the left/right joint channel permutation (augmentation.py:20)
//   jcm_augment_rotate       tf.contrib.image.rotate(BILINEAR): out(x,y) = in(cos x - sin y + x_off, sin x + cos y + y_off), 0 outside
//   jcm_augment_crop_resize  tf.image.crop_and_resize(box [rh, rw, rh+s, rw+s]) to the input size, bilinear, extrapolation 0
//   jcm_augment_hm_renorm    hm ** 1.6 + 1e-5, normalised over H*W per (image, channel)   (augmentation.py:52-54)
#include "common.cuh"

namespace {

inline int grid_for(long total, int threads) {
  long g = (total + threads - 1) / threads;
  long cap = (long)jcm_num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

__device__ float blk_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
  return r;
}

// one block per (image, channel): mean over H*W
__global__ void aug_mean_kernel(const float* __restrict__ x, int S, int C, float* __restrict__ mean) {
  __shared__ float sh[32];
  const int n = blockIdx.x / C, c = blockIdx.x % C;
  const float* src = x + (long)n * S * C + c;
  float s = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s += src[(long)i * C];
  s = blk_sum(s, sh);
  if (threadIdx.x == 0) mean[blockIdx.x] = s / (float)S;
}

__global__ void aug_color_kernel(const float* __restrict__ x, const float* __restrict__ prm, const float* __restrict__ mean, int B, int H,
                                 int W, int C, float* __restrict__ out) {
  const long total = (long)B * H * W * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long t = i / C;
    const int xo = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const float* q = prm + n * 8;
    const int xs = q[0] != 0.f ? W - 1 - xo : xo;
    const float m = mean[n * C + c] + q[1];
    const float v = x[(((long)n * H + y) * W + xs) * C + c] + q[1];
    out[i] = fminf(fmaxf((v - m) * q[2] + m, 0.f), 1.f);
  }
}

__global__ void aug_flip_channels_kernel(const float* __restrict__ hm, const float* __restrict__ prm, const int* __restrict__ perm, int B,
                                         int H, int W, int C, float* __restrict__ out) {
  const long total = (long)B * H * W * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long t = i / C;
    const int xo = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int n = (int)(t / H);
    const bool f = prm[n * 8] != 0.f;
    out[i] = hm[(((long)n * H + y) * W + (f ? W - 1 - xo : xo)) * C + (f ? perm[c] : c)];
  }
}

__device__ __forceinline__ float tap_or_zero(const float* __restrict__ img, int H, int W, int C, int c, int yy, int xx) {
  return (yy >= 0 && yy < H && xx >= 0 && xx < W) ? img[((long)yy * W + xx) * C + c] : 0.f;
}

__global__ void aug_rotate_kernel(const float* __restrict__ src, const float* __restrict__ prm, int B, int H, int W, int C,
                                  float* __restrict__ out) {
  const long total = (long)B * H * W * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long t = i / C;
    const int xo = (int)(t % W);
    t /= W;
    const int yo = (int)(t % H);
    const int n = (int)(t / H);
    const float cs = prm[n * 8 + 3], sn = prm[n * 8 + 4];
    const float x_off = ((float)(W - 1) - (cs * (float)(W - 1) - sn * (float)(H - 1))) * 0.5f;
    const float y_off = ((float)(H - 1) - (sn * (float)(W - 1) + cs * (float)(H - 1))) * 0.5f;
    const float sx = cs * (float)xo - sn * (float)yo + x_off;
    const float sy = sn * (float)xo + cs * (float)yo + y_off;
    const float fx = floorf(sx), fy = floorf(sy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx = sx - fx, wy = sy - fy;
    const float* img = src + (long)n * H * W * C;
    const float v = (tap_or_zero(img, H, W, C, c, y0, x0) * (1.f - wx) + tap_or_zero(img, H, W, C, c, y0, x0 + 1) * wx) * (1.f - wy) +
                    (tap_or_zero(img, H, W, C, c, y0 + 1, x0) * (1.f - wx) + tap_or_zero(img, H, W, C, c, y0 + 1, x0 + 1) * wx) * wy;
    out[i] = v;
  }
}

__global__ void aug_crop_kernel(const float* __restrict__ src, const float* __restrict__ prm, int B, int H, int W, int C, float crop,
                                float* __restrict__ out) {
  const long total = (long)B * H * W * C;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long t = i / C;
    const int xo = (int)(t % W);
    t /= W;
    const int yo = (int)(t % H);
    const int n = (int)(t / H);
    const float y1 = prm[n * 8 + 5], x1 = prm[n * 8 + 6];
    const float sy = y1 * (float)(H - 1) + (float)yo * crop;     // (y2 - y1)(H-1)/(crop_h - 1) with crop_h == H
    const float sx = x1 * (float)(W - 1) + (float)xo * crop;
    float v = 0.f;
    if (sy >= 0.f && sy <= (float)(H - 1) && sx >= 0.f && sx <= (float)(W - 1)) {
      const int top = (int)floorf(sy), left = (int)floorf(sx);
      const int bot = min((int)ceilf(sy), H - 1), right = min((int)ceilf(sx), W - 1);
      const float wy = sy - (float)top, wx = sx - (float)left;
      const float* img = src + (long)n * H * W * C + c;
      const float tl = img[((long)top * W + left) * C], tr = img[((long)top * W + right) * C];
      const float bl = img[((long)bot * W + left) * C], br = img[((long)bot * W + right) * C];
      const float tp = tl + (tr - tl) * wx, bt = bl + (br - bl) * wx;
      v = tp + (bt - tp) * wy;
    }
    out[i] = v;
  }
}

// one block per (image, channel)
__global__ void aug_hm_renorm_kernel(const float* __restrict__ hm, int S, int C, float power, float eps, float* __restrict__ out) {
  __shared__ float sh[32];
  const int n = blockIdx.x / C, c = blockIdx.x % C;
  const float* src = hm + (long)n * S * C + c;
  float* dst = out + (long)n * S * C + c;
  float s = 0.f;
  for (int i = threadIdx.x; i < S; i += blockDim.x) s += powf(fmaxf(src[(long)i * C], 0.f), power) + eps;
  s = blk_sum(s, sh);
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < S; i += blockDim.x) dst[(long)i * C] = (powf(fmaxf(src[(long)i * C], 0.f), power) + eps) * inv;
}

}  // namespace

extern "C" int jcm_augment_color(const float* x, const float* prm, float* mean_ws, int B, int H, int W, int C, float* out, void* stream) {
  JCM_CHECK_ARG(x && prm && mean_ws && out && B > 0 && H > 0 && W > 0 && C > 0, "jcm_augment_color: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  aug_mean_kernel<<<B * C, 256, 0, st>>>(x, H * W, C, mean_ws);
  JCM_LAUNCH_CHECK();
  aug_color_kernel<<<grid_for((long)B * H * W * C, 256), 256, 0, st>>>(x, prm, mean_ws, B, H, W, C, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_augment_flip_channels(const float* hm, const float* prm, const int* perm, int B, int H, int W, int C, float* out,
                                         void* stream) {
  JCM_CHECK_ARG(hm && prm && perm && out && B > 0 && H > 0 && W > 0 && C > 0, "jcm_augment_flip_channels: bad arguments");
  aug_flip_channels_kernel<<<grid_for((long)B * H * W * C, 256), 256, 0, (cudaStream_t)stream>>>(hm, prm, perm, B, H, W, C, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_augment_rotate(const float* src, const float* prm, int B, int H, int W, int C, float* out, void* stream) {
  JCM_CHECK_ARG(src && prm && out && src != out && B > 0 && H > 0 && W > 0 && C > 0, "jcm_augment_rotate: bad arguments");
  aug_rotate_kernel<<<grid_for((long)B * H * W * C, 256), 256, 0, (cudaStream_t)stream>>>(src, prm, B, H, W, C, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_augment_crop_resize(const float* src, const float* prm, int B, int H, int W, int C, float crop_size, float* out,
                                       void* stream) {
  JCM_CHECK_ARG(src && prm && out && src != out && B > 0 && H > 1 && W > 1 && C > 0 && crop_size > 0.f, "jcm_augment_crop_resize: bad arguments");
  aug_crop_kernel<<<grid_for((long)B * H * W * C, 256), 256, 0, (cudaStream_t)stream>>>(src, prm, B, H, W, C, crop_size, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_augment_hm_renorm(const float* hm, int B, int H, int W, int C, float power, float eps, float* out, void* stream) {
  JCM_CHECK_ARG(hm && out && B > 0 && H > 0 && W > 0 && C > 0, "jcm_augment_hm_renorm: bad arguments");
  aug_hm_renorm_kernel<<<B * C, 256, 0, (cudaStream_t)stream>>>(hm, H * W, C, power, eps, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

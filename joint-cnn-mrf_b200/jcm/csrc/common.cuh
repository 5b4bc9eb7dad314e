// Shared helpers for the jcm sm_100a kernels (error reporting, PTX wrappers).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define JCM_OK 0
#define JCM_EINVAL (-1)
#define JCM_ENOTSUP (-2)
#define JCM_EWORKSPACE (-3)

void jcm_set_error(const char* fmt, ...);

#define JCM_CHECK_ARG(cond, ...)                  \
  do {                                            \
    if (!(cond)) {                                \
      jcm_set_error(__VA_ARGS__);                 \
      return JCM_EINVAL;                          \
    }                                             \
  } while (0)

#define JCM_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      jcm_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

void jcm_count_launch();

#define JCM_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    jcm_count_launch();                                                             \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess) {                                                       \
      jcm_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                              \
    }                                                                               \
  } while (0)

static inline int jcm_cdiv(int a, int b) { return (a + b - 1) / b; }

int jcm_num_sms();

// Internal launcher of the tcgen05 implicit-GEMM kernel (conv_tcgen05.cu).  jcm_conv2d_fwd is the grp == 0 case; the grouped
// forms serve the tensor-core spatial model (see ConvParams::grp for their meaning).
struct ConvExArgs {
  const void *x_hi, *x_lo, *w_hi, *w_lo;
  const float* bias;
  void* y;
  int y_bf16;
  int B, H, W, Cin, Cout, Cout_pad, ksize, kw, relu;
  int pad_y;                                   // rows of padding above; < 0: (ksize - 1) / 2
  int grp, a_div, w_cin, k_rows, sm_pad, sm_rows, sm_rows_in;
  const int* img_map;                          // device array: grp 1: A image of conv image i; grp 2: weight plane of batch q
  int map_images;                              // number of distinct images / planes img_map points into
  int map_on_a;                                // grp 2 only: img_map selects the A image of batch q instead of its weight plane
  int variant;                                 // 0: automatic; bit 0 single-CTA kernel, bit 1 uniform tiles, bit 2 no N-split tail
  void* stream;
};
int jcm_conv_igemm_ex(const ConvExArgs& a);

// ---------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 1/alpha * softplus(alpha x), alpha = 5 (reference main.py:106-108), stable form max(z,0)+log1p(exp(-|z|))
__device__ __forceinline__ float softplus5(float x) {
  float z = 5.0f * x;
  return 0.2f * (fmaxf(z, 0.0f) + log1pf(expf(-fabsf(z))));
}
// d/dx softplus5 = sigmoid(5x)
__device__ __forceinline__ float sigmoid5(float x) { return 1.0f / (1.0f + expf(-5.0f * x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Flat work-item index -> (r0, r1, r2, q) with i = ((q * d2 + r2) * d1 + r1) * d0 + r0.  Work items of the element-wise kernels
// are counted in 64 bits, but 64-bit division by a run-time divisor costs ~100 instructions and four of them per item made these
// kernels instruction-bound (upsample_avg3: 0.80 ms for 0.82 GB); every launch of the supported configurations fits 32 bits.
__device__ __forceinline__ void split_index(long i, int d0, int d1, int d2, int& r0, int& r1, int& r2, int& q) {
  if (i <= 0xffffffffL) {
    unsigned u = (unsigned)i, v = u / (unsigned)d0;
    r0 = (int)(u - v * (unsigned)d0);
    u = v / (unsigned)d1;
    r1 = (int)(v - u * (unsigned)d1);
    v = u / (unsigned)d2;
    r2 = (int)(u - v * (unsigned)d2);
    q = (int)v;
  } else {
    long tt = i / d0;
    r0 = (int)(i - tt * d0);
    long uu = tt / d1;
    r1 = (int)(tt - uu * d1);
    tt = uu / d2;
    r2 = (int)(uu - tt * d2);
    q = (int)tt;
  }
}

// three-way form: i = (q * d1 + r1) * d0 + r0
__device__ __forceinline__ void split_index3(long i, int d0, int d1, int& r0, int& r1, int& q) {
  if (i <= 0xffffffffL) {
    unsigned u = (unsigned)i, v = u / (unsigned)d0;
    r0 = (int)(u - v * (unsigned)d0);
    u = v / (unsigned)d1;
    r1 = (int)(v - u * (unsigned)d1);
    q = (int)u;
  } else {
    long tt = i / d0;
    r0 = (int)(i - tt * d0);
    long uu = tt / d1;
    r1 = (int)(tt - uu * d1);
    q = (int)uu;
  }
}

// split an fp32 value into bf16 hi + bf16 lo (hi + lo == x to ~2^-17 relative)
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// Deterministic column sum of per-block partials laid out [nblocks][2][C]: call from a (32, kPartY) thread block; thread
// (tx, ty) adds blocks ty, ty + kPartY, ... of column c, the kPartY lanes are then added in order.  Valid in ty == 0.
constexpr int kPartY = 16;
__device__ __forceinline__ double partial_colsum(const float* __restrict__ partial, int nblocks, int C, int j, int c,
                                                 double (*sh)[33]) {
  double s = 0.0;
  if (c < C) {
    // the loads are independent of the running sum: unrolled so that 8 of them are in flight (the rolled loop paid one L2 round
    // trip per block: 18-23 us per finalize launch, 40 launches per step); the additions keep their order
#pragma unroll 8
    for (int b = threadIdx.y; b < nblocks; b += kPartY) s += (double)partial[((long)b * 2 + j) * C + c];
  }
  __syncthreads();
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.y == 0)
    for (int l = 0; l < kPartY; ++l) t += sh[l][threadIdx.x];
  return t;
}

// Activations (post-ReLU conv outputs kept for the BN / backward passes) are fp32, or bf16 in the bf16 training configuration.
// 4 consecutive channels starting at element index i4*4 of `base`:
__device__ __forceinline__ float4 load_act4(const void* __restrict__ base, long i4, int is_bf16) {
  if (is_bf16) {
    const uint2 r = reinterpret_cast<const uint2*>(base)[i4];
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&r.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&r.y);
    const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return reinterpret_cast<const float4*>(base)[i4];
}
__device__ __forceinline__ float load_act1(const void* __restrict__ base, long i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]) : reinterpret_cast<const float*>(base)[i];
}

// VEC (4 or 8) consecutive channels starting at element index `e` (a multiple of VEC) of an activation tensor.
// VEC == 8 is used for bf16-stored activations (one 16-byte load).
template <int VEC>
__device__ __forceinline__ void load_act_vec(const void* __restrict__ base, long e, int is_bf16, float (&f)[VEC]) {
  if (VEC == 8 && is_bf16) {
    const uint4 r = reinterpret_cast<const uint4*>(base)[e >> 3];
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[i]));
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC / 4; ++i) {
      const float4 t = load_act4(base, (e >> 2) + i, is_bf16);
      f[4 * i] = t.x; f[4 * i + 1] = t.y; f[4 * i + 2] = t.z; f[4 * i + 3] = t.w;
    }
  }
}
template <int VEC>
__device__ __forceinline__ void load_f32_vec(const float* __restrict__ base, long e, float (&f)[VEC]) {
#pragma unroll
  for (int i = 0; i < VEC / 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(base)[(e >> 2) + i];
    f[4 * i] = t.x; f[4 * i + 1] = t.y; f[4 * i + 2] = t.z; f[4 * i + 3] = t.w;
  }
}
// fp32 values -> bf16 hi (+ lo) planes / fp32 copy, VEC consecutive channels at element index e
template <int VEC>
__device__ __forceinline__ void store_planes_vec(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, float* __restrict__ f32,
                                                 long e, const float (&v)[VEC]) {
  __align__(16) __nv_bfloat16 h[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) h[i] = __float2bfloat16_rn(v[i]);
  if (VEC == 8) {
    if (hi) reinterpret_cast<uint4*>(hi)[e >> 3] = *reinterpret_cast<uint4*>(h);
  } else {
    if (hi) reinterpret_cast<uint2*>(hi)[e >> 2] = *reinterpret_cast<uint2*>(h);
  }
  if (lo) {      // the residual plane exists in the fp32 configuration only: its conversions are skipped otherwise (uniform branch)
    __align__(16) __nv_bfloat16 l[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) l[i] = __float2bfloat16_rn(v[i] - __bfloat162float(h[i]));
    if (VEC == 8) reinterpret_cast<uint4*>(lo)[e >> 3] = *reinterpret_cast<uint4*>(l);
    else reinterpret_cast<uint2*>(lo)[e >> 2] = *reinterpret_cast<uint2*>(l);
  }
  if (f32) {
#pragma unroll
    for (int i = 0; i < VEC / 4; ++i)
      reinterpret_cast<float4*>(f32)[(e >> 2) + i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

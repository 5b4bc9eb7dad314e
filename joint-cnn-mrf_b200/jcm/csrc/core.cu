// Library core: thread-local error string, device queries, the FP32-FMA peak micro-benchmark used as the roofline
// denominator of the spatial-model kernel, and a deliberately naive direct convolution used ONLY by the GPU tests to
// cross-check the tcgen05 kernel on identical bf16 operands (it is not on any product path).
#include "common.cuh"
#include <stdarg.h>

static thread_local char g_err[512] = "";

void jcm_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static long g_launches = 0;
void jcm_count_launch() { __atomic_add_fetch(&g_launches, 1, __ATOMIC_RELAXED); }
// number of kernels this library has launched since it was loaded (bench.py reports the per-step delta)
extern "C" long jcm_launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int jcm_num_sms() {
  static int sms[64] = {0};     // per device ordinal
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!sms[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

extern "C" const char* jcm_last_error() { return g_err; }
extern "C" int jcm_version() { return 100; }
extern "C" int jcm_sm_count() { return jcm_num_sms(); }

// ------------------------------------------------------------------------------------------- FMA peak
namespace {
template <int PACKED>
__global__ void __launch_bounds__(512, 1) fma_peak_kernel(float* out, int iters, float a, float b) {
  if (PACKED) {
    unsigned long long acc[16];
    unsigned long long av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float s = (float)(threadIdx.x + i);
      asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(s));
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float x, y;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i]));
      s += x + y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  } else {
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = (float)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], a, b);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
}

// Operand-pattern probes of the spatial-model inner loop (no memory traffic in the loop): 7 x 4 accumulators per lane,
// a 7-slot rotating window operand and a per-step 4-image operand.  MODE 2: FFMA2 with duplicated (broadcast) window
// values, MODE 3: FFMA2 with two different values per window pair (forces the 64-bit operand form), MODE 4: scalar FFMA.
template <int MODE>
__global__ void __launch_bounds__(640, 1) fma_pattern_kernel(float* io, int iters) {
  const float* src = io + (threadIdx.x & 31);
  if (MODE == 4 || MODE == 7) {
    float acc[7][4], win[7], l[7][4];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      win[k] = src[k * 32];
#pragma unroll
      for (int e = 0; e < 4; ++e) { acc[k][e] = 0.f; l[k][e] = src[(7 + k * 4 + e) * 32]; }
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        if (MODE == 4) {
#pragma unroll
          for (int k = 0; k < 7; ++k)
#pragma unroll
            for (int e = 0; e < 4; ++e) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k][e]) : "f"(win[(s + k) % 7]), "f"(l[s][e]));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int k = 0; k < 7; ++k) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[k][e]) : "f"(win[(s + k) % 7]), "f"(l[s][e]));
        }
      }
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int e = 0; e < 4; ++e) r += acc[k][e];
    io[4096 + blockIdx.x * blockDim.x + threadIdx.x] = r;
  } else {
    unsigned long long acc[7][2], win[7], l[7][2];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float a = src[k * 32], b = (MODE == 3) ? src[(40 + k) * 32] : a;
      asm("mov.b64 %0, {%1, %2};" : "=l"(win[k]) : "f"(a), "f"(b));
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        acc[k][e] = 0ull;
        asm("mov.b64 %0, {%1, %2};" : "=l"(l[k][e]) : "f"(src[(7 + k * 4 + 2 * e) * 32]), "f"(src[(8 + k * 4 + 2 * e) * 32]));
      }
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        if (MODE == 5) {          // likelihood-major: 7 window values per likelihood pair
#pragma unroll
          for (int e = 0; e < 2; ++e)
#pragma unroll
            for (int k = 0; k < 7; ++k) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k][e]) : "l"(win[(s + k) % 7]), "l"(l[s][e]));
        } else if (MODE == 6) {   // snake: every instruction changes exactly one of the two shared operands
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            const int e0 = (k & 1), e1 = e0 ^ 1;
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k][e0]) : "l"(win[(s + k) % 7]), "l"(l[s][e0]));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k][e1]) : "l"(win[(s + k) % 7]), "l"(l[s][e1]));
          }
        } else {
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k][0]) : "l"(win[(s + k) % 7]), "l"(l[s][0]));
            asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[k][1]) : "l"(win[(s + k) % 7]), "l"(l[s][1]));
          }
        }
      }
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 7; ++k)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float x, y;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[k][e]));
        r += x + y;
      }
    io[4096 + blockIdx.x * blockDim.x + threadIdx.x] = r;
  }
}
}  // namespace

// Runs the FMA loop once on `stream`; flops_out = FLOPs executed (2 per FMA).  scratch: blocks*512 floats.
// packed = 0: scalar FFMA, packed = 1: FFMA2 (fma.rn.f32x2).  Time it from the caller with CUDA events.
extern "C" int jcm_fma_peak(float* scratch, int blocks, int iters, int packed, double* flops_out, void* stream) {
  JCM_CHECK_ARG(scratch && blocks > 0 && iters > 0, "jcm_fma_peak: bad arguments");
  if (packed >= 2) {  // operand-pattern probes: scratch needs 4096 + blocks*640 floats
    if (packed == 2) fma_pattern_kernel<2><<<blocks, 640, 0, (cudaStream_t)stream>>>(scratch, iters);
    else if (packed == 3) fma_pattern_kernel<3><<<blocks, 640, 0, (cudaStream_t)stream>>>(scratch, iters);
    else if (packed == 4) fma_pattern_kernel<4><<<blocks, 640, 0, (cudaStream_t)stream>>>(scratch, iters);
    else if (packed == 5) fma_pattern_kernel<5><<<blocks, 640, 0, (cudaStream_t)stream>>>(scratch, iters);
    else if (packed == 6) fma_pattern_kernel<6><<<blocks, 640, 0, (cudaStream_t)stream>>>(scratch, iters);
    else fma_pattern_kernel<7><<<blocks, 640, 0, (cudaStream_t)stream>>>(scratch, iters);
    JCM_LAUNCH_CHECK();
    if (flops_out) *flops_out = 2.0 * 7.0 * 7.0 * 4.0 * (double)iters * 640.0 * (double)blocks;
    return JCM_OK;
  }
  if (packed)
    fma_peak_kernel<1><<<blocks, 512, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999f, 0.001f);
  else
    fma_peak_kernel<0><<<blocks, 512, 0, (cudaStream_t)stream>>>(scratch, iters, 0.999f, 0.001f);
  JCM_LAUNCH_CHECK();
  if (flops_out) *flops_out = 2.0 * 32.0 * 8.0 * (double)iters * 512.0 * (double)blocks;
  return JCM_OK;
}

// ------------------------------------------------------------------------------------------- test-only naive conv
namespace {
__global__ void conv_naive_kernel(const __nv_bfloat16* __restrict__ x_hi, const __nv_bfloat16* __restrict__ x_lo,
                                  const __nv_bfloat16* __restrict__ w_hi, const __nv_bfloat16* __restrict__ w_lo,
                                  const float* __restrict__ bias, float* __restrict__ y, int B, int H, int W, int Cin, int Cout,
                                  int Cout_pad, int ksize, int kw, int relu) {
  const long total = (long)B * H * W * Cout;
  const int pad = (ksize - 1) / 2, pad_x = (kw - 1) / 2;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % Cout);
    long t = idx / Cout;
    const int ox = (int)(t % W);
    t /= W;
    const int oy = (int)(t % H);
    const int n = (int)(t / H);
    float acc = 0.f;
    for (int dy = 0; dy < ksize; ++dy) {
      const int iy = oy + dy - pad;
      if (iy < 0 || iy >= H) continue;
      for (int dx = 0; dx < kw; ++dx) {
        const int ix = ox + dx - pad_x;
        if (ix < 0 || ix >= W) continue;
        const long xo = (((long)n * H + iy) * W + ix) * Cin;
        const long wo = ((long)(dy * kw + dx) * Cout_pad + co) * Cin;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xh = __bfloat162float(x_hi[xo + ci]), wh = __bfloat162float(w_hi[wo + ci]);
          acc = fmaf(xh, wh, acc);
          if (x_lo) {
            acc = fmaf(__bfloat162float(x_lo[xo + ci]), wh, acc);
            acc = fmaf(xh, __bfloat162float(w_lo[wo + ci]), acc);
          }
        }
      }
    }
    acc += bias ? bias[co] : 0.f;
    if (relu) acc = fmaxf(acc, 0.f);
    y[idx] = acc;
  }
}
}  // namespace

extern "C" int jcm_debug_conv2d_naive(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                                      float* y, int B, int H, int W, int Cin, int Cout, int Cout_pad, int ksize, int kw, int relu,
                                      void* stream) {
  JCM_CHECK_ARG(x_hi && w_hi && y, "jcm_debug_conv2d_naive: null pointer");
  if (kw <= 0) kw = ksize;
  const long total = (long)B * H * W * Cout;
  long grid = (total + 127) / 128;
  if (grid > 148L * 32) grid = 148L * 32;
  conv_naive_kernel<<<(int)grid, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo,
                                                                 (const __nv_bfloat16*)w_hi, (const __nv_bfloat16*)w_lo, bias, y, B, H, W,
                                                                 Cin, Cout, Cout_pad, ksize, kw, relu);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// "Tap-expanded" form of a convolution with very few output channels (conv6: 9x9, 512 -> K joints, reference main.py:72).
//
// As an implicit GEMM conv6 has N = K (7) output channels: the tensor core runs at N = 16 and, worse, the 128-pixel A tile
// (16 KB) is re-read from L2 for each of the 81 taps with nothing to amortise it over - measured 4.5 ms forward, 7.6 ms
// weight gradient at batch 64 for 0.8 % of the step's FLOPs.  Re-associating the sums turns all three passes into plain
// GEMMs over pixels with N or K = 81 * KP (KP = K padded to a multiple of 4), which the 1x1 path of the tcgen05 kernels
// runs at full rate, plus one HBM-bound gather/scatter over the tiny K-channel tensor:
//
//   forward   Z[q, tap*KP+co] = sum_ci X[q, ci] W[tap, ci, co]              (1x1 conv, N = 81*KP)
//             y[p, co]        = b[co] + sum_tap Z[p + tap - pad, tap*KP+co]  (jcm_tap_gather)
//   backward  Gt[q, tap*KP+co] = G[q - (tap - pad), co]                      (jcm_tap_scatter_planes, bf16 operand planes)
//             dW[tap, ci, co] = sum_q X[q, ci] Gt[q, tap*KP+co]              (1x1 weight gradient)
//             dX[q, ci]       = sum_n Gt[q, n] W[tap(n), ci, co(n)]          (1x1 conv with the transposed packing)
#include "common.cuh"

namespace {

inline int grid_for(long total, int threads) {
  long g = (total + threads - 1) / threads;
  long cap = (long)jcm_num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// transpose = 0: out[n][ci] (n = tap*KP+co < Npad rows, Cin columns)   - forward operand [1][Npad][Cin]
// transpose = 1: out[ci][n] (Cin rows, Npad columns)                    - data-gradient operand [1][Cin][Npad]
__global__ void pack_weights_taps_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int KP, int Npad, int transpose,
                                         __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long total = (long)Npad * Cin;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int n, ci;
    if (transpose) { ci = (int)(i / Npad); n = (int)(i % Npad); } else { n = (int)(i / Cin); ci = (int)(i % Cin); }
    const int tap = n / KP, co = n - tap * KP;
    const float v = (tap < taps && co < Cout) ? w[((long)tap * Cin + ci) * Cout + co] : 0.f;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// y[n,py,px,co] = bias[co] + sum_tap z[n, py+dy-pad, px+dx-pad, tap*KP + co]; one thread per (pixel, group of 4 channels)
__global__ void tap_gather_kernel(const float* __restrict__ z, const float* __restrict__ bias, int B, int H, int W, int ksize, int KP,
                                  int ZC, int Cout, float* __restrict__ y) {
  const int Q = KP / 4, pad = (ksize - 1) / 2;
  const long total = (long)B * H * W * Q;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int q, px, py, n;
    split_index(i, Q, W, H, q, px, py, n);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dy = 0; dy < ksize; ++dy) {
      const int yy = py + dy - pad;
      if (yy < 0 || yy >= H) continue;
      const float* zrow = z + ((long)n * H + yy) * W * ZC + (long)(dy * ksize) * KP + q * 4;
#pragma unroll 3
      for (int dx = 0; dx < ksize; ++dx) {
        const int xx = px + dx - pad;
        if (xx < 0 || xx >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(zrow + (long)xx * ZC + dx * KP);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    const float r[4] = {acc.x, acc.y, acc.z, acc.w};
    float* o = y + (((long)n * H + py) * W + px) * Cout;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int co = q * 4 + c;
      if (co < Cout) o[co] = r[c] + (bias ? bias[co] : 0.f);
    }
  }
}

// planes[n,qy,qx, tap*KP+co] = g[n, qy-(dy-pad), qx-(dx-pad), co]   (0 outside the map, for co >= Cout and in the channel padding)
// one thread per (pixel, group of 8 plane channels): 16-byte stores
__global__ void tap_scatter_planes_kernel(const float* __restrict__ g, int B, int H, int W, int ksize, int Cout, int KP, int Npad,
                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int G8 = Npad / 8, pad = (ksize - 1) / 2, taps = ksize * ksize;
  const long total = (long)B * H * W * G8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int g8, qx, qy, n;
    split_index(i, G8, W, H, g8, qx, qy, n);
    __align__(16) __nv_bfloat16 h[8];
    __align__(16) __nv_bfloat16 l[8];
    // walk the 8 channels with running (tap, co): one division per thread instead of three per element
    int tap = (g8 * 8) / KP, co = (g8 * 8) - tap * KP;
    const float* src = nullptr;
    auto locate = [&]() {
      src = nullptr;
      if (tap < taps) {
        const int dy = tap / ksize, dx = tap - dy * ksize;
        const int sy = qy - (dy - pad), sx = qx - (dx - pad);
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) src = g + (((long)n * H + sy) * W + sx) * Cout;
      }
    };
    locate();
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float v = (src != nullptr && co < Cout) ? src[co] : 0.f;
      h[e] = __float2bfloat16_rn(v);
      if (lo) l[e] = __float2bfloat16_rn(v - __bfloat162float(h[e]));      // residual plane: fp32 configuration only
      if (++co == KP) { co = 0; ++tap; locate(); }
    }
    reinterpret_cast<uint4*>(hi)[i] = *reinterpret_cast<uint4*>(h);
    if (lo) reinterpret_cast<uint4*>(lo)[i] = *reinterpret_cast<uint4*>(l);
  }
}

// The same for KP a multiple of 8 (K = 5..8, 13..16 joints): a group of 8 plane channels then lies inside ONE tap, so a thread locates
// its source pixel once and converts 8 values straight away (the generic kernel above walks (tap, co) element by element with a
// branch per channel: 277 instructions per thread, 0.33 ms for 531 MB written).  bf16 hi plane only.
__global__ void tap_scatter_planes8_kernel(const float* __restrict__ g, int B, int H, int W, int ksize, int Cout, int KP, int Npad,
                                           uint4* __restrict__ hi) {
  const int G8 = Npad / 8, pad = (ksize - 1) / 2, taps = ksize * ksize, per_tap = KP / 8;
  const long total = (long)B * H * W * G8;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    int g8, qx, qy, n;
    split_index(i, G8, W, H, g8, qx, qy, n);
    const int tap = g8 / per_tap, co0 = (g8 - tap * per_tap) * 8;
    __align__(16) __nv_bfloat16 h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = __float2bfloat16_rn(0.f);
    if (tap < taps) {
      const int dy = tap / ksize, dx = tap - dy * ksize;
      const int sy = qy - (dy - pad), sx = qx - (dx - pad);
      if (sy >= 0 && sy < H && sx >= 0 && sx < W) {
        const float* src = g + (((long)n * H + sy) * W + sx) * Cout + co0;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (co0 + e < Cout) h[e] = __float2bfloat16_rn(src[e]);
      }
    }
    hi[i] = *reinterpret_cast<uint4*>(h);
  }
}

// dwz [Cin][ZC] (column n = tap*KP+co) -> dw [taps][Cin][Cout]
__global__ void unpack_tap_grad_kernel(const float* __restrict__ dwz, int taps, int Cin, int Cout, int KP, int ZC, float* __restrict__ dw) {
  const long total = (long)taps * Cin * Cout;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    const long t = i / Cout;
    const int ci = (int)(t % Cin);
    const int tap = (int)(t / Cin);
    dw[i] = dwz[(long)ci * ZC + tap * KP + co];
  }
}

}  // namespace

extern "C" int jcm_pack_weights_taps(const float* w, int ksize, int Cin, int Cout, int KP, int Npad, int transpose, void* out_hi,
                                     void* out_lo, void* stream) {
  JCM_CHECK_ARG(w && out_hi && ksize > 0 && Cin > 0 && Cout > 0, "jcm_pack_weights_taps: bad arguments");
  JCM_CHECK_ARG(KP >= Cout && (KP % 4) == 0 && Npad >= ksize * ksize * KP && (Npad % 16) == 0,
                "jcm_pack_weights_taps: need KP >= Cout, KP %% 4 == 0, Npad >= k*k*KP, Npad %% 16 == 0 (KP=%d Npad=%d)", KP, Npad);
  pack_weights_taps_kernel<<<grid_for((long)Npad * Cin, 256), 256, 0, (cudaStream_t)stream>>>(
      w, ksize * ksize, Cin, Cout, KP, Npad, transpose, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_tap_gather(const float* z, const float* bias, int B, int H, int W, int ksize, int KP, int ZC, int Cout, float* y,
                              void* stream) {
  JCM_CHECK_ARG(z && y && B > 0 && H > 0 && W > 0 && (ksize & 1), "jcm_tap_gather: bad arguments");
  JCM_CHECK_ARG((KP % 4) == 0 && KP >= Cout && ZC >= ksize * ksize * KP && (ZC % 4) == 0, "jcm_tap_gather: bad channel layout (KP=%d ZC=%d)", KP, ZC);
  tap_gather_kernel<<<grid_for((long)B * H * W * (KP / 4), 128), 128, 0, (cudaStream_t)stream>>>(z, bias, B, H, W, ksize, KP, ZC, Cout, y);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_tap_scatter_planes(const float* g, int B, int H, int W, int ksize, int Cout, int KP, int Npad, void* hi, void* lo,
                                      void* stream) {
  JCM_CHECK_ARG(g && hi && B > 0 && H > 0 && W > 0 && (ksize & 1), "jcm_tap_scatter_planes: bad arguments");
  JCM_CHECK_ARG(KP >= Cout && (Npad % 8) == 0 && Npad >= ksize * ksize * KP, "jcm_tap_scatter_planes: bad channel layout (KP=%d Npad=%d)", KP, Npad);
  if (!lo && (KP % 8) == 0)
    tap_scatter_planes8_kernel<<<grid_for((long)B * H * W * (Npad / 8), 256), 256, 0, (cudaStream_t)stream>>>(g, B, H, W, ksize, Cout, KP, Npad,
                                                                                                          (uint4*)hi);
  else
    tap_scatter_planes_kernel<<<grid_for((long)B * H * W * (Npad / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        g, B, H, W, ksize, Cout, KP, Npad, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_unpack_tap_grad(const float* dwz, int ksize, int Cin, int Cout, int KP, int ZC, float* dw, void* stream) {
  JCM_CHECK_ARG(dwz && dw && ksize > 0 && Cin > 0 && Cout > 0 && KP >= Cout && ZC >= ksize * ksize * KP, "jcm_unpack_tap_grad: bad arguments");
  unpack_tap_grad_kernel<<<grid_for((long)ksize * ksize * Cin * Cout, 256), 256, 0, (cudaStream_t)stream>>>(dwz, ksize * ksize, Cin, Cout, KP,
                                                                                                         ZC, dw);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

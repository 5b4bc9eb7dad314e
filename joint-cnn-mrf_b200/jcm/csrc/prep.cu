// Operand preparation for the tcgen05 convolution kernel (HBM-bound, coalesced/vectorised):
//   * jcm_prep_input      x fp32 NHWC [B,H,W,3] -> three space-to-depth bf16 operand tensors (full, 1/2, 1/4 bank).
//                         Folds the reference's `tf.image.resize_images(x, [H/2,W/2])` / `[H/4,W/4]`
//                         (main.py:51,60; legacy bilinear at integer factors == exact strided sub-sampling, SURVEY
//                         Appendix B) into the gather, and turns the 5x5 stride-2 SAME conv1 (main.py:44,52,61;
//                         TF pads 1 before / 2 after) into a 3x1 stride-1 conv over 64 channels (3 horizontal taps x 16, 36 used).
//   * jcm_pack_weights    HWIO fp32 -> [tap][Cout_pad][Cin_pad] bf16 hi/lo (K-major B operand); optional
//                         flip+transpose for the data-gradient convolution.
//   * jcm_pack_weights_s2d  the conv1 weights [5,5,3,C] -> [3][C][64] matching jcm_prep_input's channel order.
//   * jcm_split_planes    fp32 -> bf16 hi (+ lo) element-wise.
#include "common.cuh"

namespace {

// One thread per output pixel (n, Y, X) and x-tap dxi: writes the 16-channel group dxi of the 64-channel pixel,
//   out[n,Y,X, dxi*16 + (sy*2+sx)*3 + c] = x[n, step*(2Y+sy), step*(2(X+dxi-1)+sx), c]     (0 outside the image, channels 12..15 of
// each group and the whole group 3 are padding).  Folding the three horizontal taps of the 3x3 space-to-depth kernel into the
// channel axis makes conv1 a 3x1 convolution with a 64-channel (128-byte) contraction per tap instead of 9 taps of 16 channels.
__global__ void prep_input_kernel(const float* __restrict__ x, int B, int H, int W, int step,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int Ho = H / (2 * step), Wo = W / (2 * step);
  const long total = (long)B * Ho * Wo * 4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int dxi = (int)(i & 3);
    const long pix = i >> 2;
    int X, Y, n;
    split_index3(pix, Wo, Ho, X, Y, n);
    __align__(16) __nv_bfloat16 vh[16];
    __align__(16) __nv_bfloat16 vl[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { vh[c] = __float2bfloat16_rn(0.f); vl[c] = vh[c]; }
    const int Xs = X + dxi - 1;
    if (dxi < 3 && Xs >= 0 && Xs < Wo) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int sy = s >> 1, sx = s & 1;
        const float* px = x + (((long)n * H + (long)step * (2 * Y + sy)) * W + (long)step * (2 * Xs + sx)) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) split_bf16(px[c], vh[s * 3 + c], vl[s * 3 + c]);
      }
    }
    uint4* dh = reinterpret_cast<uint4*>(hi + i * 16);
    dh[0] = reinterpret_cast<uint4*>(vh)[0];
    dh[1] = reinterpret_cast<uint4*>(vh)[1];
    if (lo) {
      uint4* dl = reinterpret_cast<uint4*>(lo + i * 16);
      dl[0] = reinterpret_cast<uint4*>(vl)[0];
      dl[1] = reinterpret_cast<uint4*>(vl)[1];
    }
  }
}

// out[tap][o][i] (o < Opad, i < Ipad), zero padded.
//   transpose == 0 (forward):  o = cout, i = cin,  value = W[tap][cin][cout]
//   transpose == 1 (dgrad):    o = cin,  i = cout, value = W[k*k-1-tap][cin][cout]
__global__ void pack_weights_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int Opad, int Ipad,
                                    int transpose, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long total = (long)taps * Opad * Ipad;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int i, o, tap;
    split_index3(idx, Ipad, Opad, i, o, tap);
    float v = 0.f;
    if (!transpose) {
      if (o < Cout && i < Cin) v = w[((long)tap * Cin + i) * Cout + o];
    } else {
      if (o < Cin && i < Cout) v = w[((long)(taps - 1 - tap) * Cin + o) * Cout + i];
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[idx] = h;
    if (lo) lo[idx] = l;
  }
}

// conv1: w [5][5][3][Cout] -> out [3 vertical taps][Cout][64], channel = bx*16 + (sy*2+sx)*3 + ci as prep_input_kernel writes it;
// original tap t in 0..4 -> block tap (t+1)/2, sub position (t+1)&1 (input row = 2*o - 1 + t under TF SAME padding (1 before, 2 after)).
__global__ void pack_weights_s2d_kernel(const float* __restrict__ w, int Cout, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo) {
  const int total = 3 * Cout * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int ch64 = idx & 63;
    const int co = (idx >> 6) % Cout;
    const int by = (idx >> 6) / Cout;
    const int bx = ch64 >> 4, ch = ch64 & 15;
    float v = 0.f;
    if (bx < 3 && ch < 12) {
      const int s = ch / 3, ci = ch % 3;
      const int sy = s >> 1, sx = s & 1;
      const int ty = 2 * by + sy - 1, tx = 2 * bx + sx - 1;  // original 5x5 tap
      if (ty >= 0 && ty < 5 && tx >= 0 && tx < 5) v = w[((ty * 5 + tx) * 3 + ci) * Cout + co];
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[idx] = h;
    if (lo) lo[idx] = l;
  }
}

__global__ void split_planes_kernel(const float* __restrict__ x, long n4, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    __align__(8) __nv_bfloat16 h[4];
    __align__(8) __nv_bfloat16 l[4];
    split_bf16(v.x, h[0], l[0]);
    split_bf16(v.y, h[1], l[1]);
    split_bf16(v.z, h[2], l[2]);
    split_bf16(v.w, h[3], l[3]);
    reinterpret_cast<uint2*>(hi)[i] = *reinterpret_cast<uint2*>(h);
    if (lo) reinterpret_cast<uint2*>(lo)[i] = *reinterpret_cast<uint2*>(l);
  }
}

inline int grid_for(long total, int threads) {
  long g = (total + threads - 1) / threads;
  long cap = (long)jcm_num_sms() * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace

extern "C" int jcm_prep_input(const float* x, int B, int H, int W, void* full_hi, void* full_lo, void* half_hi,
                              void* half_lo, void* quarter_hi, void* quarter_lo, void* stream) {
  JCM_CHECK_ARG(x && full_hi && half_hi && quarter_hi, "jcm_prep_input: null pointer");
  JCM_CHECK_ARG(B > 0 && H > 0 && W > 0 && (H % 8) == 0 && (W % 8) == 0, "jcm_prep_input: H and W must be multiples of 8 (got %d x %d)", H, W);
  void* hi[3] = {full_hi, half_hi, quarter_hi};
  void* lo[3] = {full_lo, half_lo, quarter_lo};
  for (int b = 0; b < 3; ++b) {
    const int step = 1 << b;
    const long total = (long)B * (H / (2 * step)) * (W / (2 * step)) * 4;
    prep_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, step, (__nv_bfloat16*)hi[b],
                                                                              (__nv_bfloat16*)lo[b]);
    JCM_LAUNCH_CHECK();
  }
  return JCM_OK;
}

extern "C" int jcm_pack_weights(const float* w, int ksize, int Cin, int Cout, int Opad, int Ipad, int transpose,
                                void* out_hi, void* out_lo, void* stream) {
  JCM_CHECK_ARG(w && out_hi, "jcm_pack_weights: null pointer");
  const int O = transpose ? Cin : Cout, I = transpose ? Cout : Cin;
  JCM_CHECK_ARG(Opad >= O && Ipad >= I, "jcm_pack_weights: padded dims (%d,%d) smaller than (%d,%d)", Opad, Ipad, O, I);
  const long total = (long)ksize * ksize * Opad * Ipad;
  pack_weights_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, ksize * ksize, Cin, Cout, Opad, Ipad, transpose,
                                                                              (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// ------------------------------------------------------------------------------------------------ batched re-pack after an update
// All regular conv kernels (Cin, Cout multiples of 32) of the model in ONE launch, both operand layouts from one read of the fp32
// master copy: a block owns a [tap][32 ci][32 co] tile, reads it coalesced along co, writes the data-gradient layout
// [taps-1-tap][ci][co] directly (co fastest) and the forward layout [tap][co][ci] through a shared-memory transpose (ci fastest).
// The per-layer kernel above gathers the forward layout with a stride of Cout floats between neighbouring threads (0.49 ms per step
// for the 20 launches of a training step; this one moves the same 340 MB at HBM speed).
namespace {
constexpr int kPackMax = 16;
struct PackLayer {
  const float* w;
  __nv_bfloat16 *f_hi, *f_lo, *d_hi, *d_lo;
  int taps, Cin, Cout, tile0;
};
struct PackBatch {
  PackLayer l[kPackMax];
  int n;
};
__global__ void __launch_bounds__(256) pack_batch_kernel(const __grid_constant__ PackBatch pb) {
  __shared__ float tile[32][33];
  int li = 0;
  while (li + 1 < pb.n && (int)blockIdx.x >= pb.l[li + 1].tile0) ++li;
  const PackLayer& L = pb.l[li];
  const int t = blockIdx.x - L.tile0;
  const int tco = L.Cout >> 5, per_tap = (L.Cin >> 5) * tco;
  const int tap = t / per_tap, r0 = t - tap * per_tap;
  const int ci0 = (r0 / tco) << 5, co0 = (r0 - (r0 / tco) * tco) << 5;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long dtap = (long)(L.taps - 1 - tap) * L.Cin;
  for (int r = ty; r < 32; r += 8) {
    const float v = L.w[((long)tap * L.Cin + ci0 + r) * L.Cout + co0 + tx];
    tile[r][tx] = v;
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    const long di = (dtap + ci0 + r) * L.Cout + co0 + tx;
    L.d_hi[di] = h;
    if (L.d_lo) L.d_lo[di] = l;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    __nv_bfloat16 h, l;
    split_bf16(tile[tx][r], h, l);
    const long fi = ((long)tap * L.Cout + co0 + r) * L.Cin + ci0 + tx;
    L.f_hi[fi] = h;
    if (L.f_lo) L.f_lo[fi] = l;
  }
}
}  // namespace

// layers: HOST array of n (<= 16) descriptors {w, fwd_hi, fwd_lo, dgrad_hi, dgrad_lo, ksize, Cin, Cout} (jcm_pack_desc in jcm.h);
// Cin and Cout multiples of 32 with Cout <= 256 or a multiple of 256 (no padding in either layout): fwd = [k*k][Cout][Cin] as
// jcm_pack_weights(transpose 0) writes it, dgrad = [k*k][Cin][Cout] as transpose 1 does.  lo pointers NULL: plain bf16.
struct jcm_pack_desc_c {
  const float* w;
  void *fwd_hi, *fwd_lo, *dgrad_hi, *dgrad_lo;
  int ksize, Cin, Cout;
};
extern "C" int jcm_pack_weights_batch(const jcm_pack_desc_c* layers, int n, void* stream) {
  JCM_CHECK_ARG(layers && n > 0 && n <= kPackMax, "jcm_pack_weights_batch: 1..%d layers, got %d", kPackMax, n);
  PackBatch pb;
  memset(&pb, 0, sizeof(pb));
  pb.n = n;
  long tiles = 0;
  for (int i = 0; i < n; ++i) {
    const jcm_pack_desc_c& d = layers[i];
    JCM_CHECK_ARG(d.w && d.fwd_hi && d.dgrad_hi && d.ksize > 0, "jcm_pack_weights_batch: null pointer in layer %d", i);
    JCM_CHECK_ARG((d.Cin % 32) == 0 && (d.Cout % 32) == 0 && (d.Cout <= 256 || (d.Cout % 256) == 0) && (d.Cin <= 256 || (d.Cin % 256) == 0),
                  "jcm_pack_weights_batch: layer %d (%d -> %d channels) needs padding, use jcm_pack_weights", i, d.Cin, d.Cout);
    PackLayer& L = pb.l[i];
    L.w = d.w;
    L.f_hi = (__nv_bfloat16*)d.fwd_hi; L.f_lo = (__nv_bfloat16*)d.fwd_lo;
    L.d_hi = (__nv_bfloat16*)d.dgrad_hi; L.d_lo = (__nv_bfloat16*)d.dgrad_lo;
    L.taps = d.ksize * d.ksize; L.Cin = d.Cin; L.Cout = d.Cout;
    L.tile0 = (int)tiles;
    tiles += (long)L.taps * (d.Cin / 32) * (d.Cout / 32);
  }
  JCM_CHECK_ARG(tiles < (1L << 31), "jcm_pack_weights_batch: too many tiles");
  pack_batch_kernel<<<(int)tiles, 256, 0, (cudaStream_t)stream>>>(pb);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_pack_weights_s2d(const float* w, int Cout, void* out_hi, void* out_lo, void* stream) {
  JCM_CHECK_ARG(w && out_hi && Cout > 0, "jcm_pack_weights_s2d: bad arguments");
  pack_weights_s2d_kernel<<<grid_for(3L * Cout * 64, 256), 256, 0, (cudaStream_t)stream>>>(w, Cout, (__nv_bfloat16*)out_hi,
                                                                                            (__nv_bfloat16*)out_lo);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

extern "C" int jcm_split_planes(const float* x, long n, void* hi, void* lo, void* stream) {
  JCM_CHECK_ARG(x && hi && n > 0 && (n % 4) == 0, "jcm_split_planes: n must be a positive multiple of 4");
  split_planes_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>(x, n / 4, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

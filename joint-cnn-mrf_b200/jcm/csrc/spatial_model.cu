// Fully-connected MRF "spatial model" forward (reference main.py:77-125), FP32 on the CUDA cores (FFMA2), sm_100a.
//
//   out[n,:,:,i] = log(sp(h_i)+d) + sum_{j != i, ascending} log( resize_61x91->60x90( conv_valid(sp(E_ij), sp(h_j)) ) + sp(b_ij) + d )
//   h = BN_bn_sm(heat_map),  sp(x) = softplus(5x)/5,  d = 1e-6
//
// `conv_mrf` (main.py:77-91: transpose + reverse + conv2d VALID + resize_images) is, per (pair, image), the true 2-D
// 'valid' convolution  C[y,x] = sum_{u<H, v<W} P[y+u, x+v] * Lf[u,v],  P = sp(E_ij) [2H,2W],  Lf[u,v] = sp(h_j)[H-1-u, W-1-v],
// y in [0,H], x in [0,W]  ->  (H+1)(W+1)HW = 29.98 M MAC for 60x90.  The reference runs K^2 of these as separate
// conv2d nodes; here all pairs x images run in ONE persistent kernel (sm_conv_kernel):
//
//   * task = (pair, group of 4 images, slice of 32 output "strips"); a strip = 7 consecutive x of one output row.
//     One warp owns one task per pass: each lane keeps 7 x 4 fp32 accumulators as 14 packed f32x2 registers and
//     slides a 7-wide register window of P along v (one scalar LDS per step), multiplying by the 4 images'
//     likelihood values fetched with one broadcast LDS.128:  14 FFMA2 (= 28 FMA/lane) per 2 shared-memory loads,
//     so the FMA pipe, not issue or shared memory, is the limiter.
//   * the whole prior P (softplus applied while staging) lives in shared memory (row stride chosen so a warp's
//     32 strips hit 32 different banks); likelihood rows stream through a cp.async double buffer.
//   * the task list is cut into equal contiguous ranges, one per SM (persistent CTAs, 20 warps = 5 per SM
//     sub-partition), so quantisation loss is at the granularity of 4 warp-tasks, not of whole pairs.
//   * sm_prep_kernel applies BN + softplus to the K+1 heat maps once and writes them flipped and image-interleaved
//     ([cond][group][u][v][4]); sm_finish_kernel does resize + softplus(bias) + delta + log + the ordered sum over j
//     + the unary term (deterministic: no atomics anywhere).
#include "common.cuh"

namespace {

constexpr int TX = 7;          // strip width (x positions per lane)
constexpr int NI = 4;          // images per task
constexpr int NW = 20;         // warps per CTA
constexpr int URC = 4;         // likelihood rows per staged chunk (12 was measured 5 % slower: longer exposed first-chunk load)
constexpr float kDelta = 1e-6f;

struct SmDims {
  int B, H, W, K;      // K predicted joints, K+1 heat-map channels
  int P;               // number of (target, cond) pairs
  int G;               // image groups = ceil(B / 4)
  int OH, OW;          // output extent of the sliding-window kernel ((H+1)x(W+1) forward, HxW for d/d likelihood)
  int Hp, Wp;          // kernel (streamed operand) extent padded to URC rows / TX columns
  int KH;              // real row count of the streamed operand (rows >= KH are zero padding and are skipped)
  int XG, tiles, NS;   // strips per output row, strips per (image, band), slices (32 strips) per (image, band)
  int NBD, TB;         // output-row bands per image and rows per band: only TB - 1 + Hp prior rows are resident at a time
  int pstride, prows;  // shared-memory layout of the prior
  int raw;             // 1: operands are used as given (conv_mrf entry point), 0: BN + softplus applied while staging
  // addressing of the raw convolution results C[p][n][y][x] ((H+1)x(W+1)) and of the likelihood gradients dL[p][n][y][x] (HxW) as
  // the glue kernels read them: element = base + p*sp + n*sn + y*sy + x.  The FFMA path stores [p][n][y][x] (dL flipped in y and
  // x), the tensor-core path [p][y][n][x] with padded rows (dL not flipped).
  long cb_sp, dl_sp;
  int cb_sn, cb_sy, dl_sn, dl_sy, dl_flip;
  // tensor-core path only (NULL otherwise): the GEMMs run on sp(E) - c0[pair] (bf16 operands resolve the prior's variation, not its
  // common level), so c0[p] * hsum[n][cond(p)] is added to every convolution result and c0[p] * dtsum[p][n] to every dL value
  const float* cm_c;      // [P][B]:   c0[p] * (sum over the map of sp(bn(heat map[n, :, :, cond(p)])))
  const float* cm_dl;     // [B][K+1]: sum over the pairs conditioned on channel j of c0[p] * (sum over the map of dT[p][n])  (the sum
                          //           of dT equals the sum of dC: the resize matrices have unit row sums)
};

__device__ __forceinline__ unsigned long long pack2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
// acc += (w, w) * l with the window value as a broadcast scalar operand (SASS: FFMA2 Rd, Rw.F32, Rl, Rd).  volatile: the
// issue order below is chosen so that consecutive FFMA2 share either the window or the likelihood operand - the register
// file delivers one new 64-bit source per FFMA2 slot at full rate, the other source must come from the operand-reuse
// cache (measured: 58 TFLOP/s with a new pair on both sources vs 73 TFLOP/s peak, profiles/r01).
__device__ __forceinline__ void ffma2s(unsigned long long& d, float w, unsigned long long l) {
  asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %1};\n\tfma.rn.f32x2 %0, t, %2, %0;\n\t}" : "+l"(d) : "f"(w), "l"(l));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------- prep
// Lt[j][g][u][v][e] = sp(scale_j * hm[4g+e, H-1-u, W-1-v, j] + shift_j)   (0 in the padding and for images >= B)
__global__ void sm_prep_kernel(const float* __restrict__ hm, const float* __restrict__ scale, const float* __restrict__ shift,
                               SmDims d, float* __restrict__ Lt) {
  const int KC = d.K + 1;
  const long total = (long)KC * d.G * d.Hp * d.Wp;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int v = (int)(i % d.Wp);
    long t = i / d.Wp;
    const int u = (int)(t % d.Hp);
    t /= d.Hp;
    const int g = (int)(t % d.G);
    const int j = (int)(t / d.G);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < d.H && v < d.W) {
      const int p = d.H - 1 - u, q = d.W - 1 - v;
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int n = 4 * g + e;
        if (n < d.B) {
          const float hv = hm[(((long)n * d.H + p) * d.W + q) * KC + j];
          r[e] = d.raw ? hv : softplus5(fmaf(hv, scale[j], shift[j]));
        } else {
          r[e] = 0.f;
        }
      }
      o = make_float4(r[0], r[1], r[2], r[3]);
    }
    reinterpret_cast<float4*>(Lt)[i] = o;
  }
}

// ---------------------------------------------------------------------------------------------- main kernel
__global__ void __launch_bounds__(NW * 32, 1)
sm_conv_kernel(const float* __restrict__ energies /*[P][2H][2W]*/, const float* __restrict__ Lt, const int* __restrict__ pair_cond,
               SmDims d, float* __restrict__ Cb /*[P][4G][H+1][W+1]*/) {
  extern __shared__ __align__(16) float smem[];
  float* Ps = smem;                                   // [prows][pstride]
  float* Ls = smem + (((size_t)d.prows * d.pstride + 3) & ~(size_t)3);  // [2 buffers][2 groups][URC][Wp][4]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lgroup = URC * d.Wp * 4;                  // floats per (group, chunk)
  const int lbuf = 2 * lgroup;

  // segment = (pair, band of output rows): the tasks of a segment share the staged prior rows [band*TB, band*TB + prows)
  const long tasks_per_seg = (long)d.G * d.NS;
  const long T = (long)d.P * d.NBD * tasks_per_seg;
  long t = (long)blockIdx.x * T / gridDim.x;
  const long t_end = (long)(blockIdx.x + 1) * T / gridDim.x;

  while (t < t_end) {
    const int seg = (int)(t / tasks_per_seg);
    const int pair = seg / d.NBD, band = seg - pair * d.NBD;
    const int yb = band * d.TB;
    const long pair_base = (long)seg * tasks_per_seg;
    const long seg_end = min(t_end, pair_base + tasks_per_seg);
    const int j = pair_cond ? pair_cond[pair] : pair;   // which streamed operand this pair reads

    // ---- stage the prior: Ps = softplus(E_pair) rows yb .. yb+prows-1, zero padding (rows >= 2H, columns >= 2W)
    __syncthreads();
    {
      const float* E = energies + (long)pair * (2 * d.H) * (2 * d.W);
      const int n = d.prows * d.pstride;
      for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int rl = idx / d.pstride, c = idx - rl * d.pstride;
        const int r = rl + yb;
        float pv = 0.f;
        if (r < 2 * d.H && c < 2 * d.W) {
          pv = E[(long)r * (2 * d.W) + c];
          if (!d.raw) pv = softplus5(pv);
        }
        Ps[idx] = pv;
      }
    }
    __syncthreads();

    while (t < seg_end) {
      const int g_lo = (int)((t - pair_base) / d.NS);
      // a pass covers at most NW tasks and at most two image groups (that is what the likelihood buffers hold)
      const long pass_end = min(min(seg_end, t + NW), pair_base + (long)(g_lo + 2) * d.NS);
      const long mytask = t + warp;
      const bool active = mytask < pass_end;
      const int g_hi = (int)((pass_end - 1 - pair_base) / d.NS);
      const int ng = g_hi - g_lo + 1;  // 1 or 2
      const long rel = (active ? mytask : t) - pair_base;
      const int g = (int)(rel / d.NS);
      const int slice = (int)(rel - (long)g * d.NS);
      int tile = slice * 32 + lane;
      bool lane_valid = active && tile < d.tiles;
      if (tile >= d.tiles) tile = 0;
      const int y = tile / d.XG, x0 = (tile - y * d.XG) * TX;   // y: row inside the band
      lane_valid = lane_valid && (yb + y) < d.OH;
      const int gsel = g - g_lo;

      unsigned long long acc[TX][2];
#pragma unroll
      for (int k = 0; k < TX; ++k) { acc[k][0] = 0ull; acc[k][1] = 0ull; }

      const int nchunks = d.Hp / URC;
      auto stage = [&](int c, int buf) {
        // copy ng x (URC x Wp x 4 floats) contiguous runs
        const int per_group16 = lgroup / 4;  // 16-byte packets per group
        for (int idx = threadIdx.x; idx < ng * per_group16; idx += blockDim.x) {
          const int gg = idx / per_group16, o = idx - gg * per_group16;
          const float* src = Lt + ((((long)j * d.G + (g_lo + gg)) * d.Hp + (long)c * URC) * d.Wp) * 4 + (long)o * 4;
          cp_async16(smem_u32(Ls + buf * lbuf + gg * lgroup + o * 4), src);
        }
        cp_async_commit();
      };

      stage(0, 0);
      for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
          stage(c + 1, (c + 1) & 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();
        if (active) {
          const float* lbase = Ls + (c & 1) * lbuf + gsel * lgroup;
          const int nrows = min(URC, d.KH - c * URC);
#pragma unroll 1
          for (int ul = 0; ul < nrows; ++ul) {
            const int u = c * URC + ul;
            const float* prow = Ps + (y + u) * d.pstride + x0;
            const ulonglong2* lrow = reinterpret_cast<const ulonglong2*>(lbase + ul * d.Wp * 4);
            float win[TX];
#pragma unroll
            for (int k = 0; k < TX - 1; ++k) win[k] = prow[k];
#pragma unroll 1
            for (int vb = 0; vb < d.Wp; vb += TX) {
#pragma unroll
              for (int s = 0; s < TX; ++s) {
                win[(s + TX - 1) % TX] = prow[vb + s + TX - 1];
                const ulonglong2 l = lrow[vb + s];
#pragma unroll
                for (int k = 0; k < TX; ++k) {   // snake order: each FFMA2 changes only one of (window, likelihood)
                  if ((k & 1) == 0) {
                    ffma2s(acc[k][0], win[(s + k) % TX], l.x);
                    ffma2s(acc[k][1], win[(s + k) % TX], l.y);
                  } else {
                    ffma2s(acc[k][1], win[(s + k) % TX], l.y);
                    ffma2s(acc[k][0], win[(s + k) % TX], l.x);
                  }
                }
              }
            }
          }
        }
        __syncthreads();
      }

      if (lane_valid) {
        const int OW = d.OW, OH = d.OH;
#pragma unroll
        for (int k = 0; k < TX; ++k) {
          const int x = x0 + k;
          if (x < OW) {
            float a0, a1, a2, a3;
            unpack2(acc[k][0], a0, a1);
            unpack2(acc[k][1], a2, a3);
            float* o = Cb + (((long)pair * (4 * d.G) + 4 * g) * OH + yb + y) * OW + x;
            const long istr = (long)OH * OW;
            o[0] = a0;
            o[istr] = a1;
            o[2 * istr] = a2;
            o[3 * istr] = a3;
          }
        }
      }
      t = pass_end;
    }
  }
}

// ---------------------------------------------------------------------------------------------- finish
__device__ __forceinline__ void legacy_tap(int dst, int n_in, int n_out, int& lo, int& hi, float& w) {
  const float scale = (float)n_in / (float)n_out;
  const float src = (float)dst * scale;
  lo = (int)floorf(src);
  hi = min(lo + 1, n_in - 1);
  w = src - (float)lo;
}

// out[n,y,x,i] = log(sp(bn(hm[n,y,x,i])) + d) + sum over pairs with target i (in list order) log(resize(C) + sp(b) + d)
// One CTA per (image, output row).  Phase 1: the P x W pairwise terms log(...) are independent - thread t -> (pair = t / W, x = t % W),
// reads of C and of the biases coalesced along x, many loads in flight per thread - and go to shared memory.  Phase 2: thread
// (i, x) adds the unary term and its pairs' terms IN LIST ORDER (the reference's summation order, main.py:114-123; the result does
// not depend on how phase 1 was scheduled); the [W][K] output row is transposed through shared memory and written contiguously.
__global__ void sm_finish_kernel(const float* __restrict__ hm, const float* __restrict__ scale, const float* __restrict__ shift,
                                 const float* __restrict__ Cb, const float* __restrict__ biases /*[P][H][W]*/,
                                 const int* __restrict__ pair_target, SmDims d, int pc, float* __restrict__ out) {
  extern __shared__ float fsm[];           // [W*K] output row, [K+2] first-pair table (as ints), [pc*W] pairwise terms of one chunk
  int* first = reinterpret_cast<int*>(fsm + d.W * d.K);
  float* terms = fsm + d.W * d.K + d.K + 2;
  const int n = blockIdx.x / d.H, y = blockIdx.x % d.H;
  const int KC = d.K + 1, OH = d.H + 1, OW = d.W + 1;
  if (threadIdx.x <= d.K) {                // pairs are sorted by target: first[i] = index of the first pair of target i
    int f = 0;
    while (f < d.P && pair_target[f] < (int)threadIdx.x) ++f;
    first[threadIdx.x] = f;
  }
  __syncthreads();
  int ylo, yhi;
  float wy;
  legacy_tap(y, OH, d.H, ylo, yhi, wy);
  // chunks of whole targets whose pairs fit the term buffer (pc >= the pairs of any one target; one chunk for K = 7 at 60x90)
  for (int i0 = 0; i0 < d.K;) {
    int i1 = i0 + 1;
    while (i1 < d.K && first[i1 + 1] - first[i0] <= pc) ++i1;
    const int p0 = first[i0], np = first[i1] - p0;
    for (int t = threadIdx.x; t < np * d.W; t += blockDim.x) {
      const int pl = t / d.W, x = t - pl * d.W, p = p0 + pl;
      int xlo, xhi;
      float wx;
      legacy_tap(x, OW, d.W, xlo, xhi, wx);
      const float* C = Cb + (long)p * d.cb_sp + (long)n * d.cb_sn;
      const float tl = C[ylo * d.cb_sy + xlo], tr = C[ylo * d.cb_sy + xhi];
      const float bl = C[yhi * d.cb_sy + xlo], br = C[yhi * d.cb_sy + xhi];
      const float top = tl + (tr - tl) * wx;
      const float bot = bl + (br - bl) * wx;
      float val = top + (bot - top) * wy;
      if (d.cm_c) val += d.cm_c[p * d.B + n];
      terms[t] = logf(val + softplus5(biases[((long)p * d.H + y) * d.W + x]) + kDelta);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < (i1 - i0) * d.W; t += blockDim.x) {
      const int il = t / d.W, x = t - il * d.W, i = i0 + il;
      const float h = fmaf(hm[(((long)n * d.H + y) * d.W + x) * KC + i], scale[i], shift[i]);
      float m = logf(softplus5(h) + kDelta);
      for (int p = first[i]; p < first[i + 1]; ++p) m += terms[(p - p0) * d.W + x];
      fsm[x * d.K + i] = m;
    }
    __syncthreads();                       // the term buffer is reused by the next chunk; the last pass also orders fsm[] before the copy-out
    i0 = i1;
  }
  float* orow = out + ((long)n * d.H + y) * d.W * d.K;
  for (int t = threadIdx.x; t < d.W * d.K; t += blockDim.x) orow[t] = fsm[t];
}

int launch_sm_finish(const float* hm, const float* scale, const float* shift, const float* Cb, const float* biases, const int* pair_target,
                     const SmDims& d, float* out, cudaStream_t st) {
  int threads = (((d.P > d.K ? d.P : d.K) * d.W + 31) / 32) * 32;
  if (threads > 512) threads = 512;          // two CTAs per SM: one's loads overlap the other's logarithms and its ordered sum
  // pairwise terms of as many whole targets as fit 96 KB next to the output row (two CTAs per SM); a target has at most K + 1 pairs
  const size_t fixed = ((size_t)d.W * d.K + d.K + 2) * sizeof(float);
  long pc = ((long)96 * 1024 - (long)fixed) / ((long)d.W * (long)sizeof(float));
  if (pc > d.P) pc = d.P;
  if (pc < d.K + 1) pc = d.K + 1;
  const size_t fsmem = fixed + (size_t)pc * d.W * sizeof(float);
  if (fsmem > 227 * 1024) {
    jcm_set_error("spatial model: %d joints on %d-wide heat maps need %zu B of shared memory in the finishing kernel", d.K, d.W, fsmem);
    return JCM_ENOTSUP;
  }
  if (fsmem > 48 * 1024) JCM_CUDA(cudaFuncSetAttribute(sm_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
  sm_finish_kernel<<<d.B * d.H, threads, fsmem, st>>>(hm, scale, shift, Cb, biases, pair_target, d, (int)pc, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// out[n,y,x] = legacy-bilinear resize of C[n] (H+1 x W+1) to H x W   (conv_mrf's tf.image.resize_images, main.py:89)
__global__ void sm_resize_kernel(const float* __restrict__ Cb, SmDims d, float* __restrict__ out) {
  const long total = (long)d.B * d.H * d.W;
  const int OH = d.H + 1, OW = d.W + 1;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % d.W);
    long t = idx / d.W;
    const int y = (int)(t % d.H);
    const int n = (int)(t / d.H);
    int ylo, yhi, xlo, xhi;
    float wy, wx;
    legacy_tap(y, OH, d.H, ylo, yhi, wy);
    legacy_tap(x, OW, d.W, xlo, xhi, wx);
    const float* C = Cb + (long)n * OH * OW;
    const float tl = C[ylo * OW + xlo], tr = C[ylo * OW + xhi];
    const float bl = C[yhi * OW + xlo], br = C[yhi * OW + xhi];
    const float top = tl + (tr - tl) * wx;
    const float bot = bl + (br - bl) * wx;
    out[idx] = top + (bot - top) * wy;
  }
}

size_t sm_smem_bytes(const SmDims& d) {
  return ((((size_t)d.prows * d.pstride + 3) & ~(size_t)3) + (size_t)2 * 2 * URC * d.Wp * 4) * sizeof(float);
}

// mode 0: forward  (output (H+1)x(W+1), streamed kernel HxW);  mode 1: d/d likelihood (output HxW, streamed kernel (H+1)x(W+1))
int fill_dims(SmDims& d, int B, int H, int W, int K, int P, int mode = 0) {
  d.raw = 0;
  d.cm_c = nullptr; d.cm_dl = nullptr;
  d.B = B; d.H = H; d.W = W; d.K = K; d.P = P;
  d.G = jcm_cdiv(B, NI);
  const int KH = mode ? H + 1 : H, KW = mode ? W + 1 : W;
  d.OH = mode ? H : H + 1;
  d.OW = mode ? W : W + 1;
  d.KH = KH;
  d.Hp = jcm_cdiv(KH, URC) * URC;
  d.Wp = jcm_cdiv(KW, TX) * TX;
  d.XG = jcm_cdiv(d.OW, TX);
  int need = d.XG * TX + d.Wp - 1;
  if (need < 2 * W) need = 2 * W;
  // bank-conflict-free: a warp's 32 consecutive strips (y*XG + xg) must map to addresses == 7*strip (mod 32)
  const int want = (TX * d.XG) % 32;
  int ps = need;
  while (ps % 32 != want) ++ps;
  d.pstride = ps;
  // output-row bands: the fewest such that the resident prior rows (TB - 1 + Hp) + the likelihood buffers fit in shared memory
  // (one band for 60x90 maps; two for the 96x128 maps of the K=14 configuration)
  for (d.NBD = 1;; ++d.NBD) {
    d.TB = jcm_cdiv(d.OH, d.NBD);
    d.prows = d.TB - 1 + d.KH;     // prior rows read by a band: y_local + u <= TB - 1 + KH - 1 (padding rows are skipped)
    d.tiles = d.TB * d.XG;
    d.NS = jcm_cdiv(d.tiles, 32);
    if (sm_smem_bytes(d) <= (size_t)227 * 1024 - 256 || d.TB <= 1) break;
  }
  d.cb_sp = (long)4 * d.G * (H + 1) * (W + 1); d.cb_sn = (H + 1) * (W + 1); d.cb_sy = W + 1;
  d.dl_sp = (long)4 * d.G * H * W; d.dl_sn = H * W; d.dl_sy = W; d.dl_flip = 1;
  return 0;
}

}  // namespace

// Workspace (floats): Lt = (K+1)*G*Hp*Wp*4, Cb = P*4G*(H+1)*(W+1)
extern "C" long jcm_spatial_model_workspace(int B, int H, int W, int K, int P) {
  SmDims d;
  fill_dims(d, B, H, W, K, P);
  const long lt = (long)(K + 1) * d.G * d.Hp * d.Wp * 4;
  const long cb = (long)P * 4 * d.G * (H + 1) * (W + 1);
  return (lt + cb) * (long)sizeof(float) + 16;
}

// heat_map [B,H,W,K+1] fp32 (already concatenated: K part-detector maps + conditioning channel), bn_scale/bn_shift [K+1]
// (from jcm_bn_finalize), energies [P][2H][2W], biases [P][H][W], pair_target/pair_cond [P] int32 (device), sorted by
// (target, cond) = the reference's summation order.  out [B,H,W,K].  cbuf_out (optional) receives the address of the raw
// (H+1)x(W+1) convolution results inside the workspace (kept for the backward pass).
extern "C" int jcm_spatial_model_fwd(const float* heat_map, const float* bn_scale, const float* bn_shift, const float* energies,
                                     const float* biases, const int* pair_target, const int* pair_cond, float* out,
                                     void* workspace, long workspace_bytes, int B, int H, int W, int K, int P, void* stream) {
  JCM_CHECK_ARG(heat_map && bn_scale && bn_shift && energies && biases && pair_target && pair_cond && out && workspace,
                "jcm_spatial_model_fwd: null pointer");
  JCM_CHECK_ARG(B > 0 && H > 1 && W > 1 && K > 0 && P > 0, "jcm_spatial_model_fwd: bad shape");
  SmDims d;
  fill_dims(d, B, H, W, K, P);
  if (workspace_bytes < jcm_spatial_model_workspace(B, H, W, K, P)) {
    jcm_set_error("jcm_spatial_model_fwd: workspace too small (%ld < %ld bytes)", workspace_bytes,
                  jcm_spatial_model_workspace(B, H, W, K, P));
    return JCM_EWORKSPACE;
  }
  const size_t smem = sm_smem_bytes(d);
  if (smem > 227 * 1024) {
    jcm_set_error("jcm_spatial_model_fwd: heat-map size %dx%d not supported by this build (needs %zu B shared memory)", H, W, smem);
    return JCM_ENOTSUP;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* Lt = (float*)workspace;
  float* Cb = Lt + (long)(K + 1) * d.G * d.Hp * d.Wp * 4;

  {
    const long total = (long)(K + 1) * d.G * d.Hp * d.Wp;
    int grid = (int)((total + 255) / 256);
    sm_prep_kernel<<<grid, 256, 0, st>>>(heat_map, bn_scale, bn_shift, d, Lt);
    JCM_LAUNCH_CHECK();
  }
  {
    JCM_CUDA(cudaFuncSetAttribute(sm_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    const long T = (long)P * d.NBD * d.G * d.NS;
    long grid = jcm_num_sms();
    if (grid > (T + NW - 1) / NW) grid = (T + NW - 1) / NW;
    sm_conv_kernel<<<(int)grid, NW * 32, smem, st>>>(energies, Lt, pair_cond, d, Cb);
    JCM_LAUNCH_CHECK();
  }
  {
    const int rc_f = launch_sm_finish(heat_map, bn_scale, bn_shift, Cb, biases, pair_target, d, out, st);
    if (rc_f) return rc_f;
  }
  return JCM_OK;
}

// conv_mrf(A, B), main.py:77-91, on its own: A [2H,2W] prior, Bmaps [b,H,W] likelihoods (both used as given, the reference
// applies softplus before the call) -> out [b,H,W] = resize(conv_valid(A, B_n)).  Workspace: jcm_spatial_model_workspace(b,H,W,0,1).
extern "C" int jcm_conv_mrf_fwd(const float* A, const float* Bmaps, float* out, void* workspace, long workspace_bytes, int b, int H,
                                int W, void* stream) {
  JCM_CHECK_ARG(A && Bmaps && out && workspace && b > 0 && H > 1 && W > 1, "jcm_conv_mrf_fwd: bad arguments");
  SmDims d;
  fill_dims(d, b, H, W, 0, 1);
  d.raw = 1;
  if (workspace_bytes < jcm_spatial_model_workspace(b, H, W, 0, 1)) {
    jcm_set_error("jcm_conv_mrf_fwd: workspace too small");
    return JCM_EWORKSPACE;
  }
  const size_t smem = sm_smem_bytes(d);
  if (smem > 227 * 1024) {
    jcm_set_error("jcm_conv_mrf_fwd: heat-map size %dx%d not supported by this build (needs %zu B shared memory)", H, W, smem);
    return JCM_ENOTSUP;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* Lt = (float*)workspace;
  float* Cb = Lt + (long)d.G * d.Hp * d.Wp * 4;
  int* zero = (int*)(Cb + (long)4 * d.G * (H + 1) * (W + 1));  // one int: cond index 0 (space reserved by the workspace query)
  JCM_CUDA(cudaMemsetAsync(zero, 0, sizeof(int), st));
  const long total = (long)d.G * d.Hp * d.Wp;
  sm_prep_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(Bmaps, nullptr, nullptr, d, Lt);
  JCM_LAUNCH_CHECK();
  JCM_CUDA(cudaFuncSetAttribute(sm_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
  const long T = (long)d.NBD * d.G * d.NS;
  long grid = jcm_num_sms();
  if (grid > (T + NW - 1) / NW) grid = (T + NW - 1) / NW;
  sm_conv_kernel<<<(int)grid, NW * 32, smem, st>>>(A, Lt, zero, d, Cb);
  JCM_LAUNCH_CHECK();
  const long tot2 = (long)b * H * W;
  sm_resize_kernel<<<(int)((tot2 + 255) / 256), 256, 0, st>>>(Cb, d, out);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// =====================================================================================================================
// Backward of the spatial model (the reference gets it from TensorFlow autodiff of main.py:94-125).
//   T_p = resize(C_p) + sp(b_p) + d,  dT = g_i / T,  db_p = sigmoid(5 b_p) * sum_n dT,  dC = R_y^T dT R_x
//   dL_j[n] = sum_{p: cond(p)=j} corr(dC_p[n], P_p)       -> the forward kernel again (mode 1: output HxW, streamed (H+1)x(W+1))
//   dP_p    = sum_n  dC_p[n] (*) flip(L_j[n])  (full convolution)                     -> sm_bwd_dp_kernel
//   dE_p = dP_p * sigmoid(5 E_p);  dh = (dL * sigmoid(5 h) + unary term) -> batch-norm backward over the K+1 channels
// =====================================================================================================================
namespace {

// thread per (pair, y, x): dT for every image (zero for the padding images) and the bias gradient
__global__ void sm_bwd_dt_kernel(const float* __restrict__ g, const float* __restrict__ Cb, const float* __restrict__ biases,
                                 const int* __restrict__ pair_target, SmDims d, float* __restrict__ dT, float* __restrict__ db) {
  const int OH = d.H + 1, OW = d.W + 1;
  const long total = (long)d.P * d.H * d.W;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % d.W);
    long t = idx / d.W;
    const int y = (int)(t % d.H);
    const int p = (int)(t / d.H);
    const int i = pair_target[p];
    int ylo, yhi, xlo, xhi;
    float wy, wx;
    legacy_tap(y, OH, d.H, ylo, yhi, wy);
    legacy_tap(x, OW, d.W, xlo, xhi, wx);
    const float bv = biases[idx];
    const float sb = softplus5(bv) + kDelta;
    float acc = 0.f;
    for (int n = 0; n < 4 * d.G; ++n) {
      float tv = 0.f;
      if (n < d.B) {
        const float* C = Cb + (long)p * d.cb_sp + (long)n * d.cb_sn;
        const float tl = C[ylo * d.cb_sy + xlo], tr = C[ylo * d.cb_sy + xhi];
        const float bl = C[yhi * d.cb_sy + xlo], br = C[yhi * d.cb_sy + xhi];
        const float top = tl + (tr - tl) * wx;
        const float bot = bl + (br - bl) * wx;
        float cv = top + (bot - top) * wy;
        if (d.cm_c) cv += d.cm_c[p * d.B + n];
        tv = g[(((long)n * d.H + y) * d.W + x) * d.K + i] / (cv + sb);
        acc += tv;
      }
      dT[(((long)p * (4 * d.G) + n) * d.H + y) * d.W + x] = tv;
    }
    db[idx] = acc * sigmoid5(bv);
  }
}

// dCs[p][g][u][v][e] = (R_y^T dT R_x)[4g+e][u][v] on the (H+1)x(W+1) grid, zero in the padding.  The legacy resize
// (H+1 -> H) has lo(y') = y', hi(y') = y'+1 (checked on the host), so row u receives (1-w(u)) from y'=u and w(u-1) from y'=u-1.
__global__ void sm_bwd_dc_kernel(const float* __restrict__ dT, SmDims dm, float* __restrict__ dCs) {
  const long total = (long)dm.P * dm.G * dm.Hp * dm.Wp;
  const int H = dm.H, W = dm.W;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int v = (int)(idx % dm.Wp);
    long t = idx / dm.Wp;
    const int u = (int)(t % dm.Hp);
    t /= dm.Hp;
    const int gi = (int)(t % dm.G);
    const int p = (int)(t / dm.G);
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    if (u <= H && v <= W) {
      float wyv[2], wxv[2];
      int ys[2], xs[2];
      int lo, hi;
      float w;
      ys[0] = u; wyv[0] = 0.f;
      if (u < H) { legacy_tap(u, H + 1, H, lo, hi, w); wyv[0] = 1.f - w; }
      ys[1] = u - 1; wyv[1] = 0.f;
      if (u >= 1) { legacy_tap(u - 1, H + 1, H, lo, hi, w); wyv[1] = w; }
      xs[0] = v; wxv[0] = 0.f;
      if (v < W) { legacy_tap(v, W + 1, W, lo, hi, w); wxv[0] = 1.f - w; }
      xs[1] = v - 1; wxv[1] = 0.f;
      if (v >= 1) { legacy_tap(v - 1, W + 1, W, lo, hi, w); wxv[1] = w; }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float* src = dT + ((long)p * (4 * dm.G) + 4 * gi + e) * H * W;
        float s = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int b = 0; b < 2; ++b)
            if (wyv[a] != 0.f && wxv[b] != 0.f) s += wyv[a] * wxv[b] * src[ys[a] * W + xs[b]];
        o[e] = s;
      }
    }
    reinterpret_cast<float4*>(dCs)[idx] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ---------------------------------------------------------------------------------------------- dP (prior gradient)
// dP[a][b] = sum_n sum_{y<=H, x<=W} dC[n][y][x] * Ltf[n][a-y][b-x]      a in [0,2H), b in [0,2W)      (full correlation, reduced over n)
// with Ltf = the flipped likelihood of the forward pass.  Every one of the (H+1)(W+1)HW products per image is needed exactly
// once; a gather over zero-padded operands would do 4x that.  The mapping below does the exact work (+ padding to 32 / 12):
//
//   * rows:    lane <-> output rows (a, a + Hm), Hm = 32*ceil(H/32).  For a fixed dC row y (warp-uniform, broadcast loads) the lane
//              reads likelihood row u = (a - y) mod Hm: y <= a contributes to row a, y > a to row a + Hm - so ONE accumulator set
//              serves both rows; it is flushed to the partial output once when y passes a (and once at the end).
//   * columns: warp <-> strip pair (b0 .. b0+11, b0+Wm .. b0+Wm+11), Wm = 12*ceil((W+1)/12), b0 = 12*strip.  x runs over Wm/12
//              blocks of 12; column b0+k at step x reads likelihood column (b0+k-x) mod Wm: x <= b0+k goes to the low strip, else to
//              the high one.  Because b0 is a multiple of the block size this is decided per BLOCK (all-low / mixed / all-high),
//              so the inner loop is branch free: one per-lane LDS.64 (new window element) + one broadcast LDS.64 (dC) per 12 FFMA2.
//   * the two halves of every FFMA2 are two images (their sum is taken at the flush); a CTA = (pair, chunk of image pairs) runs all
//     strips x row groups as warps, streams its image pairs through a cp.async double buffer and adds into its own partial
//     output [2Hm][2Wm] (each element is owned by exactly one lane, which adds to it in program order: deterministic).
//   * sm_bwd_dp_finish_kernel: dE = sigmoid(5E) * sum over chunks.
constexpr int TXD = 12;        // strip width
constexpr int DP_MAXW = 16;    // warps per CTA (strips x row groups are looped over when there are more)

struct DpDims {
  int B, H, W, P, G;
  int Hp_l, Wp_l, Hp_c, Wp_c;     // padded extents of Lt ([Hp_l][Wp_l][4]) and dCs ([Hp_c][Wp_c][4])
  int Hm, Wm, NB, RG;             // row / column moduli, column blocks (= strips), row groups
  int LS;                         // likelihood row stride in float2 units (odd; TXD-1 wrap-around halo in front)
  int npairs, nchunks, nbuf;      // image pairs (2 images each), chunks of image pairs (grid.y), shared-memory buffers (1 or 2)
  int lbuf, cbuf;                 // float2 units per likelihood / dC buffer
};

__device__ __forceinline__ void ffma2p(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}

// one block of TXD x-steps.  MODE 0: every column accumulates into lo; 2: into hi; 1: column k into lo iff step <= k.
// lcol points at likelihood column cbase of the lane's row (physical index, halo included); crow at dC[y][xb*TXD].
template <int MODE>
__device__ __forceinline__ void dp_block(unsigned long long (&lo)[TXD], unsigned long long (&hi)[TXD], unsigned long long (&win)[TXD],
                                         const unsigned long long* __restrict__ lcol, const unsigned long long* __restrict__ crow) {
#pragma unroll
  for (int s = 0; s < TXD; ++s) {
    win[(TXD - s) % TXD] = lcol[-s];
    const unsigned long long cv = crow[s];
#pragma unroll
    for (int k = 0; k < TXD; ++k) {
      const bool to_lo = MODE == 0 || (MODE == 1 && s <= k);
      if (to_lo) ffma2p(lo[k], win[(k - s + TXD) % TXD], cv);
      else ffma2p(hi[k], win[(k - s + TXD) % TXD], cv);
    }
  }
}

__global__ void __launch_bounds__(DP_MAXW * 32, 1)
sm_bwd_dp_kernel(const float* __restrict__ Lt, const float* __restrict__ dCs, const int* __restrict__ pair_cond, DpDims d,
                 float* __restrict__ partial /*[P][nchunks][2Hm][2Wm]*/) {
  extern __shared__ __align__(16) unsigned long long dsm[];
  unsigned long long* Lsm = dsm;                               // [nbuf][Hm][LS]
  unsigned long long* Csm = dsm + (size_t)d.nbuf * d.lbuf;      // [nbuf][H+1][Wm]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int p = blockIdx.x, chunk = blockIdx.y;
  const int j = pair_cond[p];
  const int ip0 = (int)((long)chunk * d.npairs / d.nchunks), ip1 = (int)((long)(chunk + 1) * d.npairs / d.nchunks);
  float* out = partial + ((long)p * d.nchunks + chunk) * (2 * d.Hm) * (2 * d.Wm);
  const int HALO = TXD - 1;

  // zero everything once: the padding (rows >= H, columns >= W of the likelihood, columns > W of dC) is never written again
  for (int i = threadIdx.x; i < d.nbuf * (d.lbuf + d.cbuf); i += blockDim.x) dsm[i] = 0ull;
  __syncthreads();

  auto stage = [&](int ip, int buf) {
    const int g = ip >> 1, h = ip & 1;
    const float* lsrc = Lt + (((long)j * d.G + g) * d.Hp_l * d.Wp_l) * 4 + 2 * h;
    unsigned long long* ldst = Lsm + (size_t)buf * d.lbuf;
    for (int i = threadIdx.x; i < d.H * d.W; i += blockDim.x) {
      const int u = i / d.W, v = i - u * d.W;
      cp_async8(smem_u32(ldst + u * d.LS + HALO + v), lsrc + ((long)u * d.Wp_l + v) * 4);
      if (v >= d.Wm - HALO) cp_async8(smem_u32(ldst + u * d.LS + v - (d.Wm - HALO)), lsrc + ((long)u * d.Wp_l + v) * 4);   // wrap-around halo
    }
    const float* csrc = dCs + (((long)p * d.G + g) * d.Hp_c * d.Wp_c) * 4 + 2 * h;
    unsigned long long* cdst = Csm + (size_t)buf * d.cbuf;
    for (int i = threadIdx.x; i < (d.H + 1) * (d.W + 1); i += blockDim.x) {
      const int y = i / (d.W + 1), x = i - y * (d.W + 1);
      cp_async8(smem_u32(cdst + y * d.Wm + x), csrc + ((long)y * d.Wp_c + x) * 4);
    }
    cp_async_commit();
  };

  if (ip0 < ip1) stage(ip0, 0);
  for (int ip = ip0; ip < ip1; ++ip) {
    const int buf = d.nbuf == 2 ? ((ip - ip0) & 1) : 0;
    if (d.nbuf == 2 && ip + 1 < ip1) {
      stage(ip + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const unsigned long long* Lb = Lsm + (size_t)buf * d.lbuf + HALO;
    const unsigned long long* Cb = Csm + (size_t)buf * d.cbuf;
    for (int wt = warp; wt < d.NB * d.RG; wt += nwarps) {
      const int strip = wt % d.NB, rg = wt / d.NB;
      const int a = rg * 32 + lane;               // low output row; high = a + Hm
      const int b0 = strip * TXD;
      unsigned long long lo[TXD], hi[TXD], win[TXD];
#pragma unroll
      for (int k = 0; k < TXD; ++k) { lo[k] = 0ull; hi[k] = 0ull; }
      auto flush = [&](int row) {
        float* orow = out + (long)row * (2 * d.Wm) + b0;
#pragma unroll
        for (int k = 0; k < TXD; ++k) {
          // fire-and-forget reductions (RED.ADD.F32): no load latency inside the y loop.  Every address is owned by exactly one
          // lane, which adds to it in program order, so the sum is still evaluated in a fixed order (deterministic).
          float s0, s1;
          unpack2(lo[k], s0, s1);
          atomicAdd(orow + k, s0 + s1);
          unpack2(hi[k], s0, s1);
          atomicAdd(orow + d.Wm + k, s0 + s1);
          lo[k] = 0ull;
          hi[k] = 0ull;
        }
      };
#pragma unroll 1
      for (int y = 0; y <= d.H; ++y) {
        if (y == a + 1) flush(a);                 // rows y <= a went to output row a; from here on to row a + Hm
        int u = a - y;
        if (u < 0) u += d.Hm;
        const unsigned long long* lrow = Lb + u * d.LS;
        const unsigned long long* crow = Cb + y * d.Wm;
#pragma unroll
        for (int k = 1; k < TXD; ++k) win[k] = lrow[b0 + k];
#pragma unroll 1
        for (int xb = 0; xb < strip; ++xb) dp_block<0>(lo, hi, win, lrow + TXD * (strip - xb), crow + xb * TXD);
        dp_block<1>(lo, hi, win, lrow, crow + strip * TXD);
#pragma unroll 1
        for (int xb = strip + 1; xb < d.NB; ++xb) dp_block<2>(lo, hi, win, lrow + TXD * (strip - xb + d.NB), crow + xb * TXD);
      }
      flush(a < d.H ? a + d.Hm : a);              // a >= H never switched rows: everything it summed belongs to its low row
    }
    __syncthreads();                              // everyone is done with `buf` before it is refilled
    if (d.nbuf == 1 && ip + 1 < ip1) stage(ip + 1, 0);
  }
}

// dE[p][a][b] = sigmoid(5 E[p][a][b]) * sum_chunks partial[p][chunk][a][b]
__global__ void sm_bwd_dp_finish_kernel(const float* __restrict__ partial, const float* __restrict__ energies, DpDims d, float* __restrict__ dE) {
  const long total = (long)d.P * 2 * d.H * 2 * d.W;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const int b = (int)(idx % (2 * d.W));
    long t = idx / (2 * d.W);
    const int a = (int)(t % (2 * d.H));
    const int p = (int)(t / (2 * d.H));
    float s = 0.f;
    for (int c = 0; c < d.nchunks; ++c) s += partial[(((long)p * d.nchunks + c) * (2 * d.Hm) + a) * (2 * d.Wm) + b];
    dE[idx] = s * sigmoid5(energies[idx]);
  }
}

void fill_dp_dims(DpDims& dp, const SmDims& df, const SmDims& dm, int B, int H, int W, int P) {
  dp.B = B; dp.H = H; dp.W = W; dp.P = P; dp.G = df.G;
  dp.Hp_l = df.Hp; dp.Wp_l = df.Wp; dp.Hp_c = dm.Hp; dp.Wp_c = dm.Wp;
  dp.Hm = jcm_cdiv(H, 32) * 32;
  dp.RG = dp.Hm / 32;
  dp.NB = jcm_cdiv(W + 1, TXD);
  dp.Wm = dp.NB * TXD;
  dp.LS = dp.Wm + TXD - 1;
  if ((dp.LS & 1) == 0) ++dp.LS;
  dp.npairs = 2 * df.G;
  // chunks of image pairs per pair (grid.y): the count (<= 8) whose P * n CTAs fill whole waves of the SMs best
  // (K=7: 49 pairs x 3 = 147 CTAs on 148 SMs; K=14: 196 x 3 = 588 = 3.97 waves instead of 196 = 1.32 waves)
  int nch = 1;
  double best = -1.0;
  for (int n = 1; n <= 8 && n <= dp.npairs; ++n) {
    const long ctas = (long)P * n;
    const double eff = (double)ctas / (double)(jcm_cdiv((int)ctas, jcm_num_sms()) * (long)jcm_num_sms());
    if (eff > best + 0.02) { best = eff; nch = n; }
  }
  dp.nchunks = nch;
  dp.lbuf = dp.Hm * dp.LS;
  dp.cbuf = (H + 1) * dp.Wm;
  dp.nbuf = ((size_t)2 * (dp.lbuf + dp.cbuf) * 8 <= (size_t)227 * 1024 - 256) ? 2 : 1;
}

// dhbn[n,y,x,j] = sigmoid(5 hbn) * ( sum_{p: cond(p)=j} dLf[p][n][H-1-y][W-1-x]  +  [j<K] g[n,y,x,j] / (sp(hbn) + d) )
__global__ void sm_bwd_dh_kernel(const float* __restrict__ hm, const float* __restrict__ scale, const float* __restrict__ shift,
                                 const float* __restrict__ g, const float* __restrict__ dLf, const int* __restrict__ pair_cond, SmDims d,
                                 float* __restrict__ dhbn) {
  // thread <-> (n, y, j, x) with x fastest: a warp reads 32 consecutive x of one pair's dL row (the big operand) coalesced and the
  // cond(p) == j test is warp-uniform; the small [.., KC] tensors are accessed with a KC-float stride
  const int KC = d.K + 1;
  const long total = (long)d.B * d.H * d.W * KC;
  for (long idx = blockIdx.x * (long)blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    int x, j, y, n;
    split_index(idx, d.W, KC, d.H, x, j, y, n);
    const long e = (((long)n * d.H + y) * d.W + x) * KC + j;
    const float hb = fmaf(hm[e], scale[j], shift[j]);
    const float* src = dLf + (long)n * d.dl_sn + (long)(d.dl_flip ? d.H - 1 - y : y) * d.dl_sy + (d.dl_flip ? d.W - 1 - x : x);
    float s = 0.f;
    for (int p = 0; p < d.P; ++p)
      if (pair_cond[p] == j) s += src[(long)p * d.dl_sp];
    if (d.cm_dl) s += d.cm_dl[n * KC + j];
    if (j < d.K) s += g[(((long)n * d.H + y) * d.W + x) * d.K + j] / (softplus5(hb) + kDelta);
    dhbn[e] = s * sigmoid5(hb);
  }
}

// batch-norm backward over [M, KC] (KC = K+1 <= 32 channels).  Pass 1: per-block partial sums of dy and dy*xhat
// (block = 32 rows x KC... threads (c, r)); pass 2: totals (fixed order) + d_in = scale * (dy - mean(dy) - xhat * mean(dy*xhat)).
constexpr int BNB_THREADS = 256;
__global__ void sm_bn_bwd_partial_kernel(const float* __restrict__ hm, const float* __restrict__ dy, const float* __restrict__ mean,
                                         const float* __restrict__ rstd, long M, int KC, float* __restrict__ partial /*[grid][2][KC]*/) {
  __shared__ float sh[2][BNB_THREADS];
  const int lanes = BNB_THREADS / KC;            // rows handled in parallel
  const int c = threadIdx.x % KC, rl = threadIdx.x / KC;
  const long per_blk = (M + gridDim.x - 1) / gridDim.x;
  const long r0 = blockIdx.x * per_blk;
  long r1 = r0 + per_blk;
  if (r1 > M) r1 = M;
  float s0 = 0.f, s1 = 0.f;
  if (rl < lanes) {
    const float mu = mean ? mean[c] : 0.f, rs = rstd ? rstd[c] : 1.f;
    for (long r = r0 + rl; r < r1; r += lanes) {
      const float dv = dy[r * KC + c];
      s0 += dv;
      s1 += dv * ((hm[r * KC + c] - mu) * rs);
    }
  }
  sh[0][threadIdx.x] = s0;
  sh[1][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.x < KC) {
    float t0 = 0.f, t1 = 0.f;
    for (int l = 0; l < lanes; ++l) { t0 += sh[0][l * KC + threadIdx.x]; t1 += sh[1][l * KC + threadIdx.x]; }
    partial[((long)blockIdx.x * 2 + 0) * KC + threadIdx.x] = t0;
    partial[((long)blockIdx.x * 2 + 1) * KC + threadIdx.x] = t1;
  }
}

__global__ void sm_bn_bwd_apply_kernel(const float* __restrict__ hm, const float* __restrict__ dy, const float* __restrict__ scale,
                                       const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ partial,
                                       int nblocks, long M, int KC, int train, float* __restrict__ d_in, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta) {
  __shared__ float tot[2][32];
  if (threadIdx.x < 2 * KC) {
    const int j = threadIdx.x / KC, c = threadIdx.x % KC;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += (double)partial[((long)b * 2 + j) * KC + c];
    tot[j][c] = (float)s;
    if (blockIdx.x == 0) { if (j == 0) dbeta[c] = (float)s; else dgamma[c] = (float)s; }
  }
  __syncthreads();
  const long total = M * KC;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int c = (int)(i % KC);
    const float mu = mean ? mean[c] : 0.f, rs = rstd ? rstd[c] : 1.f;
    const float m0 = train ? tot[0][c] / (float)M : 0.f, m1 = train ? tot[1][c] / (float)M : 0.f;
    const float xh = (hm[i] - mu) * rs;
    d_in[i] = scale[c] * (dy[i] - m0 - xh * m1);
  }
}

}  // namespace

// extra workspace (bytes) the backward pass needs on top of the forward workspace (which must be kept: it holds Lt and C)
extern "C" long jcm_spatial_model_bwd_workspace(int B, int H, int W, int K, int P) {
  SmDims df, dm;
  fill_dims(df, B, H, W, K, P, 0);
  fill_dims(dm, B, H, W, K, P, 1);
  const long dT = (long)P * 4 * df.G * H * W;
  const long dCs = (long)P * dm.G * dm.Hp * dm.Wp * 4;
  const long dLf = (long)P * 4 * dm.G * H * W;
  const long dh = (long)B * H * W * (K + 1);
  DpDims dp;
  fill_dp_dims(dp, df, dm, B, H, W, P);
  const long part = (long)P * dp.nchunks * (2 * dp.Hm) * (2 * dp.Wm);
  return (dT + dCs + dLf + dh + part) * (long)sizeof(float) + 64;
}

// g = d loss / d out [B,H,W,K].  fwd_workspace = the workspace jcm_spatial_model_fwd filled for the SAME inputs.
// bn_mean / bn_rstd: saved batch statistics of bn_sm (train != 0) or NULL.  Outputs: d_heat_map [B,H,W,K+1], dE [P][2H][2W],
// db [P][H][W], dgamma / dbeta [K+1].
extern "C" int jcm_spatial_model_bwd(const float* g, const float* heat_map, const float* bn_scale, const float* bn_shift,
                                     const float* bn_mean, const float* bn_rstd, int train, const float* energies, const float* biases,
                                     const int* pair_target, const int* pair_cond, const void* fwd_workspace, void* workspace,
                                     long workspace_bytes, float* d_heat_map, float* dE, float* db, float* dgamma, float* dbeta, int B,
                                     int H, int W, int K, int P, void* stream) {
  JCM_CHECK_ARG(g && heat_map && bn_scale && bn_shift && energies && biases && pair_target && pair_cond && fwd_workspace && workspace &&
                    d_heat_map && dE && db && dgamma && dbeta, "jcm_spatial_model_bwd: null pointer");
  JCM_CHECK_ARG(!train || (bn_mean && bn_rstd), "jcm_spatial_model_bwd: training mode needs the saved batch statistics");
  JCM_CHECK_ARG(K + 1 <= 32, "jcm_spatial_model_bwd: at most 31 joints are supported (got %d)", K);
  if (workspace_bytes < jcm_spatial_model_bwd_workspace(B, H, W, K, P)) {
    jcm_set_error("jcm_spatial_model_bwd: workspace too small");
    return JCM_EWORKSPACE;
  }
  // the gather form of the resize transpose assumes lo(y') = y' for the (H+1) -> H legacy resize
  for (int yy = 0; yy < H; ++yy) {
    const float src = (float)yy * ((float)(H + 1) / (float)H);
    JCM_CHECK_ARG((int)floorf(src) == yy, "jcm_spatial_model_bwd: unsupported heat-map height %d", H);
  }
  for (int xx = 0; xx < W; ++xx) {
    const float src = (float)xx * ((float)(W + 1) / (float)W);
    JCM_CHECK_ARG((int)floorf(src) == xx, "jcm_spatial_model_bwd: unsupported heat-map width %d", W);
  }
  SmDims df, dm;
  fill_dims(df, B, H, W, K, P, 0);
  fill_dims(dm, B, H, W, K, P, 1);
  const size_t smem_conv = sm_smem_bytes(dm);
  DpDims dp;
  fill_dp_dims(dp, df, dm, B, H, W, P);
  const size_t smem_dp = (size_t)dp.nbuf * (dp.lbuf + dp.cbuf) * sizeof(unsigned long long);
  if (smem_conv > 227 * 1024 || smem_dp > 227 * 1024 - 256) {
    jcm_set_error("jcm_spatial_model_bwd: heat-map size %dx%d not supported by this build (shared memory %zu / %zu B)", H, W, smem_conv, smem_dp);
    return JCM_ENOTSUP;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float* Lt = (const float*)fwd_workspace;
  const float* Cb = Lt + (long)(K + 1) * df.G * df.Hp * df.Wp * 4;
  float* dT = (float*)workspace;
  float* dCs = dT + (long)P * 4 * df.G * H * W;
  float* dLf = dCs + (long)P * dm.G * dm.Hp * dm.Wp * 4;
  float* dh = dLf + (long)P * 4 * dm.G * H * W;
  float* part = dh + (long)B * H * W * (K + 1);

  {
    const long total = (long)P * H * W;
    sm_bwd_dt_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(g, Cb, biases, pair_target, df, dT, db);
    JCM_LAUNCH_CHECK();
  }
  {
    const long total = (long)P * dm.G * dm.Hp * dm.Wp;
    sm_bwd_dc_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(dT, dm, dCs);
    JCM_LAUNCH_CHECK();
  }
  {
    JCM_CUDA(cudaFuncSetAttribute(sm_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    const long T = (long)P * dm.NBD * dm.G * dm.NS;
    long grid = jcm_num_sms();
    if (grid > (T + NW - 1) / NW) grid = (T + NW - 1) / NW;
    sm_conv_kernel<<<(int)grid, NW * 32, smem_conv, st>>>(energies, dCs, nullptr, dm, dLf);
    JCM_LAUNCH_CHECK();
  }
  {
    JCM_CUDA(cudaFuncSetAttribute(sm_bwd_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256));
    JCM_CUDA(cudaMemsetAsync(part, 0, (size_t)P * dp.nchunks * (2 * dp.Hm) * (2 * dp.Wm) * sizeof(float), st));
    int nw = dp.NB * dp.RG;
    if (nw > DP_MAXW) nw = DP_MAXW;
    sm_bwd_dp_kernel<<<dim3(P, dp.nchunks), nw * 32, smem_dp, st>>>(Lt, dCs, pair_cond, dp, part);
    JCM_LAUNCH_CHECK();
    const long tot = (long)P * 4 * H * W;
    sm_bwd_dp_finish_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(part, energies, dp, dE);
    JCM_LAUNCH_CHECK();
  }
  {
    const long total = (long)B * H * W * (K + 1);
    sm_bwd_dh_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(heat_map, bn_scale, bn_shift, g, dLf, pair_cond, df, dh);
    JCM_LAUNCH_CHECK();
    const long Mr = (long)B * H * W;
    int nb = (int)((Mr + 255) / 256);
    if (nb > 2 * jcm_num_sms()) nb = 2 * jcm_num_sms();
    float* bnpart = part;   // the dP partial buffer is free again (its finish kernel ran above): reuse its first 2*nb*(K+1) floats
    sm_bn_bwd_partial_kernel<<<nb, BNB_THREADS, 0, st>>>(heat_map, dh, bn_mean, bn_rstd, Mr, K + 1, bnpart);
    JCM_LAUNCH_CHECK();
    sm_bn_bwd_apply_kernel<<<2 * jcm_num_sms(), 256, 0, st>>>(heat_map, dh, bn_scale, bn_mean, bn_rstd, bnpart, nb, Mr, K + 1, train,
                                                             d_heat_map, dgamma, dbeta);
    JCM_LAUNCH_CHECK();
  }
  return JCM_OK;
}

// =====================================================================================================================
// Tensor-core form of the spatial model (bf16 training configuration, opt-in for fp32 inference; same interface).  The prior
// operand is centred per pair (P = sp(E) - c0[pair]) and the common level is added back in fp32 by the glue kernels, so everything
// that is linear in the prior comes out at fp32 accuracy on nearly flat priors; the prior gradient (a product of two bf16-rounded
// activations) at the plain bf16 level.
//
// conv_mrf is a 2-D convolution of an HxW map with a 2Hx2W kernel: along x it is a Toeplitz matrix product, along y a 2H-tap
// 1-D convolution.  With h = sp(bn(heat map)) (not flipped), P = sp(E):
//
//   forward   C[n][y][x]  = sum_dy sum_v  h[n][y+dy-H][v]      * P[2H-1-dy][x-v+W-1]          y <= H, x <= W
//   dL        dL[n][u][v] = sum_dy sum_x dC[n][u+dy-(H-1)][x]  * P[dy][x-v+W-1]
//                         = sum_dy sum_x' T[n][u+dy-(H-1)][x'] * ((1-w(x')) P[dy][x'-v+W-1] + w(x') P[dy][x'-v+W])     T = R_y^T dT, x' < W
//   dP        dP[r][c] = sum_{x-v+W-1=c} ( sum_{n,y'} h[n][y'][v] * dC[n][y'+r-(H-1)][x] )                              r = 2H-1-dy
//
// The first two are convolutions with 2H x 1 taps over "images" [rows y][columns = the batch][channels = x or v] whose tap
// weights are the Toeplitz blocks T_dy[x][v] of one prior row, a different set for every (target, cond) pair: the grouped form
// grp 1 of the tcgen05 implicit-GEMM kernel (conv_tcgen05.cu).  The third is, per (pair, r), a GEMM over (n, y') of h^T (M = v:
// exactly W rows) with a row-shifted dC^T (N = x): form grp 2, followed by a sum along the diagonals of each block.  The Toeplitz
// blocks are materialised in HBM as bf16 (2 x P x 2H x NP x CPf: 289 MB for K = 7, written once per step at HBM speed - cheaper than
// any on-the-fly form the UMMA descriptors could express).  Work: 3 x ~0.3 PFLOP-equivalent tiles on the tensor pipe instead of
// 3 x 94 GMAC of FFMA.
// =====================================================================================================================
namespace {

constexpr int kHsumChunks = 16;   // pixel chunks per image of the heat-map sums (smt_hsum_partial_kernel)

struct SmtDims {
  int B, H, W, K, P;
  int Hc;   // H + 1: rows of the convolution output and of the zero-padded operands
  int NP;   // GEMM N extent: W + 1 rounded up to 16 (at least 32)
  int CPf;  // GEMM K extent of the forward pass (v < W) and of dL (x' < W: the x half of the resize is folded into the dL
            // weights, see smt_pack_prior_kernel): W rounded up to 64 - for W = 128 one k-block less than W + 1 would need
  int Bp;   // images rounded up to 16; the dP GEMM contracts over k = row * Bp + image
};

int fill_smt(SmtDims& t, int B, int H, int W, int K, int P) {
  t.B = B; t.H = H; t.W = W; t.K = K; t.P = P;
  t.Hc = H + 1;
  t.NP = jcm_cdiv(W + 1, 16) * 16;          // UMMA N granularity at M = 128 (W = 128: 144 columns, not 160)
  if (t.NP < 32) t.NP = 32;
  t.CPf = jcm_cdiv(W, 64) * 64;
  t.Bp = jcm_cdiv(B, 16) * 16;
  JCM_CHECK_ARG(t.NP <= 256, "jcm_spatial_model_tc: heat-map width %d not supported (at most 255)", W);
  return JCM_OK;
}

inline size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

struct SmtFwdWs { size_t spE, c0, hsum, cmc, Xh, Wf, Cb, total; };
SmtFwdWs smt_fwd_layout(const SmtDims& t) {
  SmtFwdWs w;
  size_t o = 0;
  w.spE = o; o += al256((size_t)t.P * 2 * t.H * 2 * t.W * 4);
  w.c0 = o;  o += al256((size_t)t.P * 4);
  w.hsum = o; o += al256((size_t)t.B * kHsumChunks * (t.K + 1) * 4);
  w.cmc = o;  o += al256((size_t)t.P * t.B * 4);
  w.Xh = o;  o += al256((size_t)(t.K + 1) * t.Hc * t.B * t.CPf * 2);
  w.Wf = o;  o += al256((size_t)t.P * 2 * t.H * t.NP * t.CPf * 2);
  w.Cb = o;  o += al256((size_t)t.P * t.Hc * t.B * t.NP * 4);
  w.total = o + 256;
  return w;
}
struct SmtBwdWs { size_t dT, dtsum, cmdl, Xc, XcT, Ht, Wd, dL, blk, dh, part, total; };
SmtBwdWs smt_bwd_layout(const SmtDims& t) {
  SmtBwdWs w;
  const int G4 = 4 * jcm_cdiv(t.B, NI);
  size_t o = 0;
  w.dT = o;   o += al256((size_t)t.P * G4 * t.H * t.W * 4);
  w.dtsum = o; o += al256((size_t)t.P * G4 * 4);
  w.cmdl = o; o += al256((size_t)t.B * (t.K + 1) * 4);
  w.Xc = o;   o += al256((size_t)t.P * t.Hc * t.B * t.CPf * 2);
  w.XcT = o;  o += al256((size_t)t.P * t.NP * t.Hc * t.Bp * 2);
  w.Ht = o;   o += al256((size_t)(t.K + 1) * t.W * t.H * t.Bp * 2);
  w.Wd = o;   o += al256((size_t)t.P * 2 * t.H * t.NP * t.CPf * 2);
  w.dL = o;   o += al256((size_t)t.P * t.Hc * t.B * t.NP * 4);
  w.blk = o;  o += al256((size_t)t.P * 2 * t.H * t.W * t.NP * 4);
  w.dh = o;   o += al256((size_t)t.B * t.H * t.W * (t.K + 1) * 4);
  w.part = o; o += al256((size_t)4 * jcm_num_sms() * (t.K + 1) * 4 + 1024);
  w.total = o + 256;
  return w;
}

// deterministic block sum (fixed shuffle tree, then the warp sums in order); every thread gets the result.  blockDim.x multiple of 32
__device__ __forceinline__ float smt_block_sum(float v, float* red /*[33]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// one CTA per pair: c0[p] = mean of sp(E_p), spE[p] = sp(E_p) - c0[p].  The prior is nearly flat (sp(0) = ln2/5 = 0.1386 plus a
// variation of a few 1e-3), so rounding sp(E) itself to bf16 (steps of 1e-3 at this magnitude) would lose most of its structure;
// the residual is resolved to 2^-9 of ITS magnitude and the common level is added back exactly, in fp32, by the glue kernels.
__global__ void __launch_bounds__(1024)
smt_softplus_center_kernel(const float* __restrict__ E, int n, float* __restrict__ spE, float* __restrict__ c0) {
  __shared__ float red[33];
  const float* src = E + (long)blockIdx.x * n;
  float* dst = spE + (long)blockIdx.x * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = softplus5(src[i]);
    dst[i] = v;
    s += v;
  }
  const float mean = smt_block_sum(s, red) / (float)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] -= mean;   // re-reads this thread's own writes
  if (threadIdx.x == 0) c0[blockIdx.x] = mean;
}

// hpart[n][chunk][j] = sum over the chunk's pixels of sp(bn(hm[n, pixel, j])): grid (chunks, images); a thread walks its pixels with
// K+1 <= 32 running sums (one 4(K+1)-byte read per pixel), then the block adds them up channel by channel in a fixed order
__global__ void __launch_bounds__(256)
smt_hsum_partial_kernel(const float* __restrict__ hm, const float* __restrict__ scale, const float* __restrict__ shift, int HW, int KC,
                        float* __restrict__ hpart) {
  __shared__ float red[33];
  __shared__ float sc[32], sh[32];
  if (threadIdx.x < KC) { sc[threadIdx.x] = scale[threadIdx.x]; sh[threadIdx.x] = shift[threadIdx.x]; }
  __syncthreads();
  const int per = (HW + kHsumChunks - 1) / kHsumChunks;
  const int lo = blockIdx.x * per, hi = min(HW, lo + per);
  const float* src = hm + (long)blockIdx.y * HW * KC;
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < KC) acc[j] += softplus5(fmaf(src[(long)i * KC + j], sc[j], sh[j]));
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < KC) {
      const float t = smt_block_sum(acc[j], red);
      if (threadIdx.x == 0) hpart[((long)blockIdx.y * kHsumChunks + blockIdx.x) * KC + j] = t;
    }
  }
}

// cm_c[p][n] = c0[p] * sum_chunks hpart[n][chunk][cond(p)]      (thread per (p, n))
__global__ void smt_cmc_kernel(const float* __restrict__ hpart, const float* __restrict__ c0, const int* __restrict__ pair_cond, int P, int B,
                               int KC, float* __restrict__ cmc) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P * B) return;
  const int p = idx / B, n = idx - p * B;
  const int j = pair_cond[p];
  float h = 0.f;
  for (int c = 0; c < kHsumChunks; ++c) h += hpart[((long)n * kHsumChunks + c) * KC + j];
  cmc[idx] = c0[p] * h;
}

// cm_dl[n][j] = sum_{p: cond(p) = j, in list order} c0[p] * dtsum[p][n]      (thread per (n, j))
__global__ void smt_cmdl_kernel(const float* __restrict__ dtsum, const float* __restrict__ c0, const int* __restrict__ pair_cond, int P, int B,
                                int G4, int KC, float* __restrict__ cmdl) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * KC) return;
  const int n = idx / KC, j = idx - n * KC;
  float s = 0.f;
  for (int p = 0; p < P; ++p)
    if (pair_cond[p] == j) s += c0[p] * dtsum[(long)p * G4 + n];
  cmdl[idx] = s;
}

// dtsum[q] = sum of the HW values of dT[q], q = pair * 4G + image: one CTA per q
__global__ void __launch_bounds__(256)
smt_dtsum_kernel(const float* __restrict__ dT, int HW, float* __restrict__ dtsum) {
  __shared__ float red[33];
  const float* src = dT + (long)blockIdx.x * HW;
  float s = 0.f;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) s += src[i];
  const float t = smt_block_sum(s, red);
  if (threadIdx.x == 0) dtsum[blockIdx.x] = t;
}

// Toeplitz blocks of the prior, 8 consecutive K elements (16 bytes) per thread:
//   Wf[p][dy][x (NP rows)][v (CPf)]  = spE[p][2H-1-dy][x-v+W-1]    forward  (x <= W, v < W; 0 elsewhere)
//   Wd[p][dy][v (NP rows)][x' (CPf)] = (1-w(x')) spE[p][dy][x'-v+W-1] + w(x') spE[p][dy][x'-v+W]     dL  (v < W, x' < W)
// The dL weights carry the x half of the transposed resize: dC[y][x] = (1-w(x)) T[y][x] + w(x-1) T[y][x-1] with T = R_y^T dT, so
// sum_{x<=W} dC[y][x] P[x-v+W-1] = sum_{x'<W} T[y][x'] Wd[v][x'] - a contraction over W instead of W + 1 columns.
// One CTA per (dy, pair): the prior row (2W floats) and the interpolation weights w(x') go through shared memory; a thread owns one
// 16-byte column group c8 and walks down the NP rows.
__global__ void __launch_bounds__(256)
smt_pack_prior_kernel(const float* __restrict__ spE, SmtDims t, int dl, __nv_bfloat16* __restrict__ Wt) {
  extern __shared__ float psm[];            // [2W] prior row, [W] w(x')
  float* srow = psm;
  float* wtab = psm + 2 * t.W;
  const int dy = blockIdx.x, p = blockIdx.y;
  const int C8 = t.CPf / 8, H2 = 2 * t.H, W2 = 2 * t.W, W = t.W;
  const float* src = spE + ((long)p * H2 + (dl ? dy : H2 - 1 - dy)) * W2;
  for (int i = threadIdx.x; i < W2; i += blockDim.x) srow[i] = src[i];
  if (dl)
    for (int i = threadIdx.x; i < W; i += blockDim.x) {
      int lo, hi;
      float w;
      legacy_tap(i, W + 1, W, lo, hi, w);
      wtab[i] = w;
    }
  __syncthreads();
  const int rows_per_pass = blockDim.x / C8;
  const int c8 = threadIdx.x % C8, r0 = threadIdx.x / C8;
  if (r0 >= rows_per_pass) return;
  uint4* dst = reinterpret_cast<uint4*>(Wt) + ((long)p * H2 + dy) * t.NP * C8;
  const int cbase = c8 * 8;
  float wcol[8];                            // w(x') of this thread's eight columns (dL form)
#pragma unroll
  for (int e = 0; e < 8; ++e) wcol[e] = (dl && cbase + e < W) ? wtab[cbase + e] : 0.f;
  for (int row = r0; row < t.NP; row += rows_per_pass) {
    uint32_t o[4] = {0u, 0u, 0u, 0u};
    if (cbase < W && row < W + (dl ? 0 : 1)) {
      float v[8];
      if (dl) {
        const float* q = srow + (cbase - row + W - 1);      // q[e] = a, q[e + 1] = b of column cbase + e
        float prev = q[0];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool ok = cbase + e < W;
          const float nxt = ok ? q[e + 1] : 0.f;
          v[e] = ok ? prev + (nxt - prev) * wcol[e] : 0.f;
          prev = nxt;
        }
      } else {
        const float* q = srow + (row - cbase + W - 1);      // column cbase + e reads q[-e]
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (cbase + e < W) ? q[-e] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
        o[e] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    }
    dst[(long)row * C8 + c8] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Xh[j][y (Hc)][n (B)][v (CPf)] = h_j[n][y][v] (0 in the padding), one "image" per heat-map channel j (the GEMM kernels pick the
// pair's conditioning channel through pair_cond);  Ht[j][v (W)][y*Bp + n] = the same values, transposed (the dP GEMM's A
// operand: pixel rows v, rows y < H only).  One CTA per (32 columns, row, channel); either output may be NULL.
__global__ void __launch_bounds__(256)
smt_prep_kernel(const float* __restrict__ hm, const float* __restrict__ scale, const float* __restrict__ shift, SmtDims t,
                __nv_bfloat16* __restrict__ Xh, __nv_bfloat16* __restrict__ Ht) {
  __shared__ float tile[64][33];
  const int v0 = blockIdx.x * 32, y = blockIdx.y, p = blockIdx.z;
  const int j = p, KC = t.K + 1;
  const float sc = scale[j], sh = shift[j];
  const long KH = (long)t.H * t.Bp;
  for (int nc = 0; nc < t.Bp; nc += 64) {
    {
      const int tx = threadIdx.x & 31, tn = threadIdx.x >> 5;
      const int v = v0 + tx;
      for (int pass = 0; pass < 8; ++pass) {
        const int nn = pass * 8 + tn, n = nc + nn;
        float val = 0.f;
        if (n < t.B && y < t.H && v < t.W) val = softplus5(fmaf(hm[(((long)n * t.H + y) * t.W + v) * KC + j], sc, sh));
        tile[nn][tx] = val;
        if (Xh && n < t.B && v < t.CPf) Xh[(((long)p * t.Hc + y) * t.B + n) * t.CPf + v] = __float2bfloat16_rn(val);
      }
    }
    __syncthreads();
    if (Ht && y < t.H) {
      const int nn = threadIdx.x & 63, x0 = threadIdx.x >> 6;
      const int n = nc + nn;
      for (int pass = 0; pass < 8; ++pass) {
        const int xx = pass * 4 + x0, v = v0 + xx;
        if (v < t.W && n < t.Bp) Ht[((long)p * t.W + v) * KH + (long)y * t.Bp + n] = __float2bfloat16_rn(tile[nn][xx]);
      }
    }
    __syncthreads();
  }
}

// The two bf16 operands of the backward GEMMs from dT (P x images x H x W):
//   Xc[p][u (Hc)][n (B)][x' (CPf)] = T = R_y^T dT  (rows resized back to H + 1, columns still W)     dL: A operand
//   XcT[p][x (NP)][u*Bp + n]       = dC = T R_x  on the (H+1) x (W+1) grid (as sm_bwd_dc_kernel)     dP: weight rows x, K-major
__global__ void __launch_bounds__(256)
smt_dc_kernel(const float* __restrict__ dT, SmtDims t, int G4, __nv_bfloat16* __restrict__ Xc, __nv_bfloat16* __restrict__ XcT) {
  __shared__ float tile[64][33];
  const int x0 = blockIdx.x * 32, u = blockIdx.y, p = blockIdx.z;
  const int H = t.H, W = t.W;
  const long KA = (long)t.Hc * t.Bp;
  int lo, hi;
  float w;
  float wy0 = 0.f, wy1 = 0.f;
  if (u < H) { legacy_tap(u, H + 1, H, lo, hi, w); wy0 = 1.f - w; }
  if (u >= 1) { legacy_tap(u - 1, H + 1, H, lo, hi, w); wy1 = w; }
  const int tx = threadIdx.x & 31, tn = threadIdx.x >> 5;
  const int x = x0 + tx;
  float wx0 = 0.f, wx1 = 0.f;
  if (x < W) { legacy_tap(x, W + 1, W, lo, hi, w); wx0 = 1.f - w; }
  if (x >= 1 && x <= W) { legacy_tap(x - 1, W + 1, W, lo, hi, w); wx1 = w; }
  const int rows = t.Bp < 64 ? t.Bp : 64;           // images per tile pass (a multiple of 16): no idle half tile for batches of 16 / 32 / 48
  for (int nc = 0; nc < t.Bp; nc += rows) {
    for (int pass = 0; pass * 8 < rows; ++pass) {
      const int nn = pass * 8 + tn, n = nc + nn;
      float t0 = 0.f, t1 = 0.f;            // T[u][x], T[u][x-1]
      if (n < t.B && x <= W) {
        const float* src = dT + ((long)p * G4 + n) * H * W;
        if (x < W) {
          if (wy0 != 0.f) t0 += wy0 * src[u * W + x];
          if (wy1 != 0.f) t0 += wy1 * src[(u - 1) * W + x];
        }
        if (x >= 1) {
          if (wy0 != 0.f) t1 += wy0 * src[u * W + x - 1];
          if (wy1 != 0.f) t1 += wy1 * src[(u - 1) * W + x - 1];
        }
      }
      tile[nn][tx] = wx0 * t0 + wx1 * t1;
      if (n < t.B && x < t.CPf) Xc[(((long)p * t.Hc + u) * t.B + n) * t.CPf + x] = __float2bfloat16_rn(t0);
    }
    __syncthreads();
    {
      const int nn = threadIdx.x % rows, xb = threadIdx.x / rows, xper = 256 / rows;   // rows in {16, 32, 48, 64}: 48 leaves 16 threads idle
      const int n = nc + nn;
      if (xb < xper)
        for (int xx = xb; xx < 32; xx += xper)
          if (n < t.Bp && x0 + xx < t.NP) XcT[((long)p * t.NP + x0 + xx) * KA + (long)u * t.Bp + n] = __float2bfloat16_rn(tile[nn][xx]);
    }
    __syncthreads();
  }
}

// dE[p][r][c] = sigmoid(5 E[p][r][c]) * sum over the diagonal x - v + W - 1 = c of blk[p][r][v][x]   (blk rows v, NP columns x)
// One CTA per (r, pair), one thread per c.  All lanes of a warp walk the SAME rows v (from the first row any of its diagonals touches
// to the last), each reading its own column x = c + v - (W - 1) when that lies in [0, W]: every warp load is one contiguous run
// of a row, eight rows in flight per thread.
__global__ void __launch_bounds__(256, 4)
smt_dp_reduce_kernel(const float* __restrict__ blk, const float* __restrict__ E, SmtDims t, float* __restrict__ dE) {
  const int r = blockIdx.x, p = blockIdx.y;
  const int W = t.W, NP = t.NP;
  const float* src = blk + ((long)p * 2 * t.H + r) * W * NP;
  for (int cb = threadIdx.x & ~31; cb < 2 * W; cb += blockDim.x) {          // warp-uniform loop: cb = the warp's first diagonal
    const int c = cb + (threadIdx.x & 31);
    const int v_lo = max(0, W - 1 - (cb + 31)), v_hi = min(W - 1, 2 * W - 1 - cb);
    const int x0 = c - (W - 1);                                              // x = x0 + v
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int v = v_lo;
#pragma unroll 1
    for (; v + 8 <= v_hi + 1; v += 8) {
      float a[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int x = x0 + v + k;
        a[k] = ((unsigned)x <= (unsigned)W) ? src[(long)(v + k) * NP + x] : 0.f;
      }
      s0 += a[0]; s1 += a[1]; s2 += a[2]; s3 += a[3];
      s0 += a[4]; s1 += a[5]; s2 += a[6]; s3 += a[7];
    }
#pragma unroll 1
    for (; v <= v_hi; ++v) {
      const int x = x0 + v;
      if ((unsigned)x <= (unsigned)W) s0 += src[(long)v * NP + x];
    }
    if (c < 2 * W) {
      const long o = ((long)p * 2 * t.H + r) * 2 * W + c;
      dE[o] = ((s0 + s1) + (s2 + s3)) * sigmoid5(E[o]);
    }
  }
}

int smt_conv(const void* x, const void* w, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int pad_y, int grp, int a_div,
             int w_cin, int k_rows, int sm_pad, int sm_rows, int sm_rows_in, const int* img_map, int map_images, void* stream,
             int map_on_a = 0) {
  ConvExArgs a;
  memset(&a, 0, sizeof(a));
  a.x_hi = x; a.w_hi = w; a.y = y;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.Cout_pad = Cout; a.ksize = ksize; a.kw = 1;
  a.pad_y = pad_y; a.grp = grp; a.a_div = a_div; a.w_cin = w_cin; a.k_rows = k_rows;
  a.sm_pad = sm_pad; a.sm_rows = sm_rows; a.sm_rows_in = sm_rows_in;
  a.img_map = img_map; a.map_images = map_images; a.map_on_a = map_on_a;
  a.stream = stream;
  return jcm_conv_igemm_ex(a);
}

}  // namespace

extern "C" long jcm_spatial_model_tc_workspace(int B, int H, int W, int K, int P) {
  SmtDims t;
  if (fill_smt(t, B, H, W, K, P)) return -1;
  return (long)smt_fwd_layout(t).total;
}

// Same arguments and result as jcm_spatial_model_fwd; the convolutions run on the tensor cores with bf16 operands.
extern "C" int jcm_spatial_model_tc_fwd(const float* heat_map, const float* bn_scale, const float* bn_shift, const float* energies,
                                        const float* biases, const int* pair_target, const int* pair_cond, float* out,
                                        void* workspace, long workspace_bytes, int B, int H, int W, int K, int P, void* stream) {
  JCM_CHECK_ARG(heat_map && bn_scale && bn_shift && energies && biases && pair_target && pair_cond && out && workspace,
                "jcm_spatial_model_tc_fwd: null pointer");
  JCM_CHECK_ARG(B > 0 && H > 1 && W > 1 && K > 0 && P > 0, "jcm_spatial_model_tc_fwd: bad shape");
  JCM_CHECK_ARG(K + 1 <= 32, "jcm_spatial_model_tc_fwd: at most 31 joints are supported (got %d)", K);
  SmtDims t;
  int rc = fill_smt(t, B, H, W, K, P);
  if (rc) return rc;
  const SmtFwdWs L = smt_fwd_layout(t);
  if (workspace_bytes < (long)L.total) {
    jcm_set_error("jcm_spatial_model_tc_fwd: workspace too small (%ld < %zu bytes)", workspace_bytes, L.total);
    return JCM_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float* spE = (float*)(ws + L.spE);
  float* c0 = (float*)(ws + L.c0);
  float* hsum = (float*)(ws + L.hsum);
  float* cmc = (float*)(ws + L.cmc);
  __nv_bfloat16* Xh = (__nv_bfloat16*)(ws + L.Xh);
  __nv_bfloat16* Wf = (__nv_bfloat16*)(ws + L.Wf);
  float* Cb = (float*)(ws + L.Cb);
  {
    smt_softplus_center_kernel<<<P, 1024, 0, st>>>(energies, 4 * H * W, spE, c0);
    JCM_LAUNCH_CHECK();
    smt_hsum_partial_kernel<<<dim3(kHsumChunks, B), 256, 0, st>>>(heat_map, bn_scale, bn_shift, H * W, K + 1, hsum);
    JCM_LAUNCH_CHECK();
    smt_cmc_kernel<<<(P * B + 255) / 256, 256, 0, st>>>(hsum, c0, pair_cond, P, B, K + 1, cmc);
    JCM_LAUNCH_CHECK();
    smt_pack_prior_kernel<<<dim3(2 * H, P), 256, (size_t)3 * W * sizeof(float), st>>>(spE, t, 0, Wf);
    JCM_LAUNCH_CHECK();
    smt_prep_kernel<<<dim3(t.CPf / 32, t.Hc, K + 1), 256, 0, st>>>(heat_map, bn_scale, bn_shift, t, Xh, nullptr);
    JCM_LAUNCH_CHECK();
  }
  // C[p][y][n][x] = sum_dy Xh[cond(p)][y+dy-H][n][:] . Wf[p][dy][x][:]
  rc = smt_conv(Xh, Wf, Cb, P, t.Hc, B, t.CPf, t.NP, 2 * H, H, 1, 1, 0, 0, 0, 0, 0, pair_cond, K + 1, stream);
  if (rc) return rc;
  SmDims d;
  fill_dims(d, B, H, W, K, P);
  d.cb_sp = (long)t.Hc * B * t.NP; d.cb_sy = B * t.NP; d.cb_sn = t.NP;
  d.cm_c = cmc;
  {
    const int rc_f = launch_sm_finish(heat_map, bn_scale, bn_shift, Cb, biases, pair_target, d, out, st);
    if (rc_f) return rc_f;
  }
  return JCM_OK;
}

extern "C" long jcm_spatial_model_tc_bwd_workspace(int B, int H, int W, int K, int P) {
  SmtDims t;
  if (fill_smt(t, B, H, W, K, P)) return -1;
  return (long)smt_bwd_layout(t).total;
}

// Same arguments and results as jcm_spatial_model_bwd; fwd_workspace = the workspace jcm_spatial_model_tc_fwd filled.
extern "C" int jcm_spatial_model_tc_bwd(const float* g, const float* heat_map, const float* bn_scale, const float* bn_shift,
                                        const float* bn_mean, const float* bn_rstd, int train, const float* energies,
                                        const float* biases, const int* pair_target, const int* pair_cond, const void* fwd_workspace,
                                        void* workspace, long workspace_bytes, float* d_heat_map, float* dE, float* db, float* dgamma,
                                        float* dbeta, int B, int H, int W, int K, int P, void* stream) {
  JCM_CHECK_ARG(g && heat_map && bn_scale && bn_shift && energies && biases && pair_target && pair_cond && fwd_workspace && workspace &&
                    d_heat_map && dE && db && dgamma && dbeta, "jcm_spatial_model_tc_bwd: null pointer");
  JCM_CHECK_ARG(!train || (bn_mean && bn_rstd), "jcm_spatial_model_tc_bwd: training mode needs the saved batch statistics");
  JCM_CHECK_ARG(K + 1 <= 32, "jcm_spatial_model_tc_bwd: at most 31 joints are supported (got %d)", K);
  SmtDims t;
  int rc = fill_smt(t, B, H, W, K, P);
  if (rc) return rc;
  const SmtFwdWs LF = smt_fwd_layout(t);
  const SmtBwdWs L = smt_bwd_layout(t);
  if (workspace_bytes < (long)L.total) {
    jcm_set_error("jcm_spatial_model_tc_bwd: workspace too small (%ld < %zu bytes)", workspace_bytes, L.total);
    return JCM_EWORKSPACE;
  }
  for (int yy = 0; yy < H; ++yy) {
    const float src = (float)yy * ((float)(H + 1) / (float)H);
    JCM_CHECK_ARG((int)floorf(src) == yy, "jcm_spatial_model_tc_bwd: unsupported heat-map height %d", H);
  }
  for (int xx = 0; xx < W; ++xx) {
    const float src = (float)xx * ((float)(W + 1) / (float)W);
    JCM_CHECK_ARG((int)floorf(src) == xx, "jcm_spatial_model_tc_bwd: unsupported heat-map width %d", W);
  }
  cudaStream_t st = (cudaStream_t)stream;
  const uint8_t* fws = (const uint8_t*)(((uintptr_t)fwd_workspace + 255) & ~(uintptr_t)255);
  const float* spE = (const float*)(fws + LF.spE);
  const float* c0 = (const float*)(fws + LF.c0);
  const float* cmc = (const float*)(fws + LF.cmc);
  const float* Cb = (const float*)(fws + LF.Cb);
  uint8_t* ws = (uint8_t*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  float* dT = (float*)(ws + L.dT);
  float* dtsum = (float*)(ws + L.dtsum);
  float* cmdl = (float*)(ws + L.cmdl);
  __nv_bfloat16* Xc = (__nv_bfloat16*)(ws + L.Xc);
  __nv_bfloat16* XcT = (__nv_bfloat16*)(ws + L.XcT);
  __nv_bfloat16* Ht = (__nv_bfloat16*)(ws + L.Ht);
  __nv_bfloat16* Wd = (__nv_bfloat16*)(ws + L.Wd);
  float* dL = (float*)(ws + L.dL);
  float* blk = (float*)(ws + L.blk);
  float* dh = (float*)(ws + L.dh);
  float* bnpart = (float*)(ws + L.part);

  SmDims d;
  fill_dims(d, B, H, W, K, P);
  d.cb_sp = (long)t.Hc * B * t.NP; d.cb_sy = B * t.NP; d.cb_sn = t.NP;
  d.dl_sp = (long)t.Hc * B * t.NP; d.dl_sy = B * t.NP; d.dl_sn = t.NP; d.dl_flip = 0;
  d.cm_c = cmc; d.cm_dl = cmdl;
  const int G4 = 4 * d.G;
  {
    const long total = (long)P * H * W;
    sm_bwd_dt_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(g, Cb, biases, pair_target, d, dT, db);
    JCM_LAUNCH_CHECK();
    smt_dtsum_kernel<<<P * G4, 256, 0, st>>>(dT, H * W, dtsum);
    JCM_LAUNCH_CHECK();
    smt_cmdl_kernel<<<(B * (K + 1) + 127) / 128, 128, 0, st>>>(dtsum, c0, pair_cond, P, B, G4, K + 1, cmdl);
    JCM_LAUNCH_CHECK();
    smt_dc_kernel<<<dim3(jcm_cdiv(t.NP > t.CPf ? t.NP : t.CPf, 32), t.Hc, P), 256, 0, st>>>(dT, t, G4, Xc, XcT);
    JCM_LAUNCH_CHECK();
    smt_prep_kernel<<<dim3(jcm_cdiv(W, 32), t.Hc, K + 1), 256, 0, st>>>(heat_map, bn_scale, bn_shift, t, nullptr, Ht);
    JCM_LAUNCH_CHECK();
    smt_pack_prior_kernel<<<dim3(2 * H, P), 256, (size_t)3 * W * sizeof(float), st>>>(spE, t, 1, Wd);
    JCM_LAUNCH_CHECK();
  }
  // dL[p][u][n][v] = sum_dy Xc[p][u+dy-(H-1)][n][:] . Wd[p][dy][v][:]
  rc = smt_conv(Xc, Wd, dL, P, t.Hc, B, t.CPf, t.NP, 2 * H, H - 1, 1, 1, 0, 0, 0, 0, 0, nullptr, 0, stream);
  if (rc) return rc;
  // blk[p*2H+r][v][x] = sum_{y',n} Ht[cond(p)][v][y'*Bp+n] * XcT[p][x][(y'+r-(H-1))*Bp+n]      (r = 2H-1-dy: the prior row)
  // the M side is v: exactly W pixel rows (ONE 128-row tile for W <= 128, where x = W + 1 rows would need two), N = x
  rc = smt_conv(Ht, XcT, blk, P * 2 * H, 1, W, H * t.Bp, t.NP, 1, 0, 2, 2 * H, t.Hc * t.Bp, t.Bp, H - 1, H, t.Hc, pair_cond, K + 1, stream, 1);
  if (rc) return rc;
  {
    int threads = jcm_cdiv(2 * W, 32) * 32;
    if (threads > 256) threads = 256;
    smt_dp_reduce_kernel<<<dim3(2 * H, P), threads, 0, st>>>(blk, energies, t, dE);
    JCM_LAUNCH_CHECK();
  }
  {
    const long total = (long)B * H * W * (K + 1);
    sm_bwd_dh_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(heat_map, bn_scale, bn_shift, g, dL, pair_cond, d, dh);
    JCM_LAUNCH_CHECK();
    const long Mr = (long)B * H * W;
    int nb = (int)((Mr + 255) / 256);
    if (nb > 2 * jcm_num_sms()) nb = 2 * jcm_num_sms();
    sm_bn_bwd_partial_kernel<<<nb, BNB_THREADS, 0, st>>>(heat_map, dh, bn_mean, bn_rstd, Mr, K + 1, bnpart);
    JCM_LAUNCH_CHECK();
    sm_bn_bwd_apply_kernel<<<2 * jcm_num_sms(), 256, 0, st>>>(heat_map, dh, bn_scale, bn_mean, bn_rstd, bnpart, nb, Mr, K + 1, train,
                                                             d_heat_map, dgamma, dbeta);
    JCM_LAUNCH_CHECK();
  }
  return JCM_OK;
}

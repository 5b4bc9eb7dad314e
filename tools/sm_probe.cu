// Development harness for the FFMA2 spatial-model kernel (measurement tool, not part of libjcm.so): the kernel under test on
// synthetic operands at the bench shape, checked against a naive kernel, timed with CUDA events and with clock64 per CTA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/sm_probe tools/sm_probe.cu ; ./tools/sm_probe [B]
// Compile-time switches: -DNIMG=4|8 (images per task), -DNWARPS=n, -DST=n (ring stages), -DVARIANT=3|0 (14-register window ring /
// the round-1 window), -DEXP=1|2|3 (timing only: likelihood / window / both loads hoisted out of the inner loop).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

#ifndef NWARPS
#define NWARPS 20
#endif
#ifndef NIMG
#define NIMG 4
#endif
constexpr int TX = 7, NI = NIMG, NW = NWARPS;   // NIMG = 8: two LDS.128 per step for 28 FFMA2 (0.107 loads per FFMA2 instead of 0.143)
#ifndef ST
#define ST 4
#endif
#ifndef ORDER
#define ORDER 1
#endif

#ifndef EXP
#define EXP 0
#endif
#ifndef VARIANT
#define VARIANT 3
#endif
// the 14 FFMA2 of one step: window slots (sft + k) % RING, k = 0..6, against the two image pairs of l
#define STEP_FMAS(win, sft, RING, l)                                                                     \
  _Pragma("unroll") for (int q4 = 0; q4 < NI / 4; ++q4) {                                                \
    _Pragma("unroll") for (int k = 0; k < TX; ++k) ffma2s(acc[k][2 * q4], win[((sft) + k) % (RING)], l[q4].x); \
    _Pragma("unroll") for (int k = 0; k < TX; ++k) ffma2s(acc[TX - 1 - k][2 * q4 + 1], win[((sft) + TX - 1 - k) % (RING)], l[q4].y); \
  }

struct Dims {
  int B, H, W, P, G, OH, OW, KH, Hp, Wp, XG, tiles, NS, NBD, TB, pstride, prows;
};

__device__ long long g_cyc[1024];

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ffma2s(unsigned long long& d, float w, unsigned long long l) {
  asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %1};\n\tfma.rn.f32x2 %0, t, %2, %0;\n\t}" : "+l"(d) : "f"(w), "l"(l));
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// v2: warps are independent.  The SM's contiguous task range is dealt to its warps round-robin; every warp streams the
// likelihood rows of its own tasks through a private ST-deep ring of one-row buffers (one cp.async.bulk per row issued by
// lane 0, completion on a per-stage mbarrier), so the only block-wide barriers are around the staging of a new prior.
__global__ void __launch_bounds__(NW * 32, 1)
sm_conv2_kernel(const float* __restrict__ energies, const float* __restrict__ Lt, const int* __restrict__ pair_cond, Dims d, float* __restrict__ Cb) {
  extern __shared__ __align__(16) float smem[];
  float* Ps = smem;
  const int ps_floats = (d.prows * d.pstride + 3) & ~3;
  const int rowf = d.Wp * NI;                                  // floats per likelihood row (NI images interleaved)
  float* Lw = smem + ps_floats;                                // [NW][ST][rowf]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(Lw + (size_t)NW * ST * rowf);  // [NW][ST]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long cyc0 = clock64();

  if (threadIdx.x == 0) {
    for (int i = 0; i < NW * ST; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t my_bars = smem_u32(bars + warp * ST);
  float* my_L = Lw + (size_t)warp * ST * rowf;
  const uint32_t my_L_u32 = smem_u32(my_L);
  const uint32_t rowbytes = (uint32_t)rowf * 4;
  uint32_t q = 0;                                              // rows consumed by this warp so far (stage = q % ST, parity = (q / ST) & 1)

  const long tasks_per_seg = (long)d.G * d.NS;
  const long T = (long)d.P * d.NBD * tasks_per_seg;
  const long t_begin = (long)blockIdx.x * T / gridDim.x;
  const long t_end = (long)(blockIdx.x + 1) * T / gridDim.x;

  long t = t_begin;
  while (t < t_end) {
    const int seg = (int)(t / tasks_per_seg);
    const int pair = seg / d.NBD, band = seg - pair * d.NBD;
    const int yb = band * d.TB;
    const long seg_base = (long)seg * tasks_per_seg;
    const long seg_end = min(t_end, seg_base + tasks_per_seg);
    const int j = pair_cond ? pair_cond[pair] : pair;

    __syncthreads();
    {
      const float* E = energies + (long)pair * (2 * d.H) * (2 * d.W);
      const int n = d.prows * d.pstride;
      for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int rl = idx / d.pstride, c = idx - rl * d.pstride;
        const int r = rl + yb;
        float pv = 0.f;
        if (r < 2 * d.H && c < 2 * d.W) pv = E[(long)r * (2 * d.W) + c];
        Ps[idx] = pv;
      }
    }
    __syncthreads();

    // this warp's tasks of the segment: t + warp, t + warp + NW, ...
    const long first = t + warp;
    const int ntask = first < seg_end ? (int)((seg_end - first + NW - 1) / NW) : 0;
    const int R = ntask * d.KH;                                // row items of this warp in this segment
    const float* Lj = Lt + (long)j * d.G * d.Hp * d.Wp * NI;
    auto row_src = [&](int ti, int u) {
      const long rel = first + (long)ti * NW - seg_base;
      const int g = (int)(rel / d.NS);
      return Lj + (((long)g * d.Hp + u) * d.Wp) * NI;
    };
    // producer state (lane 0): next row item to issue
    int p_ti = 0, p_u = 0, p_issued = 0;
    uint32_t q_issue = q;                                      // rows issued so far (global count)
    if (lane == 0) {
      for (int k = 0; k < ST - 1 && p_issued < R; ++k) {
        const uint32_t s = q_issue % ST;
        mbar_expect_tx(my_bars + s * 8, rowbytes);
        bulk_g2s(my_L_u32 + s * rowbytes, row_src(p_ti, p_u), rowbytes, my_bars + s * 8);
        ++q_issue; ++p_issued;
        if (++p_u == d.KH) { p_u = 0; ++p_ti; }
      }
    }

    for (int ti = 0; ti < ntask; ++ti) {
      const long mytask = first + (long)ti * NW;
      const long rel = mytask - seg_base;
      const int g = (int)(rel / d.NS);
      const int slice = (int)(rel - (long)g * d.NS);
      int tile = slice * 32 + lane;
      bool lane_valid = tile < d.tiles;
      if (tile >= d.tiles) tile = 0;
      const int y = tile / d.XG, x0 = (tile - y * d.XG) * TX;
      lane_valid = lane_valid && (yb + y) < d.OH;

      unsigned long long acc[TX][NI / 2];
#pragma unroll
      for (int k = 0; k < TX; ++k)
#pragma unroll
        for (int e = 0; e < NI / 2; ++e) acc[k][e] = 0ull;

#pragma unroll 1
      for (int u = 0; u < d.KH; ++u) {
        __syncwarp();                                          // every lane is done with the row consumed one iteration ago
        if (lane == 0 && p_issued < R) {
          const uint32_t s = q_issue % ST;
          mbar_expect_tx(my_bars + s * 8, rowbytes);
          bulk_g2s(my_L_u32 + s * rowbytes, row_src(p_ti, p_u), rowbytes, my_bars + s * 8);
          ++q_issue; ++p_issued;
          if (++p_u == d.KH) { p_u = 0; ++p_ti; }
        }
        const uint32_t s = q % ST;
        mbar_wait(my_bars + s * 8, (q / ST) & 1);
        ++q;
        const float* prow = Ps + (y + u) * d.pstride + x0;
        const ulonglong2* lrow = reinterpret_cast<const ulonglong2*>(my_L + (size_t)s * rowf);
#if VARIANT == 3
        // 14-register window ring: slot i % 14 holds P[x0 + i]; step i uses slots i .. i+6 and then reloads its own (dead) slot
        // with the value 14 columns ahead, first needed 8 steps later - no load sits next to the last use of its register.
        float win[2 * TX];
#pragma unroll
        for (int k = 0; k < 2 * TX; ++k) win[k] = prow[k];
        const int npair = (d.Wp / TX) >> 1;
        ulonglong2 l0[NI / 4];
#pragma unroll
        for (int q4 = 0; q4 < NI / 4; ++q4) l0[q4] = lrow[(lane & 1) * (NI / 4) + q4];
        (void)l0;
#pragma unroll 1
        for (int it = 0; it < npair; ++it) {
          const int vb = it * 2 * TX;
#pragma unroll
          for (int sft = 0; sft < 2 * TX; ++sft) {
#if (EXP & 1)
            ulonglong2 l[NI / 4];
#pragma unroll
            for (int q4 = 0; q4 < NI / 4; ++q4) l[q4] = l0[q4];
#else
            ulonglong2 l[NI / 4];
#pragma unroll
            for (int q4 = 0; q4 < NI / 4; ++q4) l[q4] = lrow[(vb + sft) * (NI / 4) + q4];
#endif
            STEP_FMAS(win, sft, 2 * TX, l)
#if !(EXP & 2)
            win[sft] = prow[vb + sft + 2 * TX];
#endif
          }
        }
        if ((d.Wp / TX) & 1) {
          const int vb = npair * 2 * TX;
#pragma unroll
          for (int sft = 0; sft < TX; ++sft) {
            ulonglong2 l[NI / 4];
#pragma unroll
            for (int q4 = 0; q4 < NI / 4; ++q4) l[q4] = lrow[(vb + sft) * (NI / 4) + q4];
            STEP_FMAS(win, sft, 2 * TX, l)
          }
        }
#else
        float win[TX];
#pragma unroll
        for (int k = 0; k < TX - 1; ++k) win[k] = prow[k];
#pragma unroll 1
        for (int vb = 0; vb < d.Wp; vb += TX) {
#pragma unroll
          for (int sft = 0; sft < TX; ++sft) {
            win[(sft + TX - 1) % TX] = prow[vb + sft + TX - 1];
            ulonglong2 l[NI / 4];
#pragma unroll
            for (int q4 = 0; q4 < NI / 4; ++q4) l[q4] = lrow[(vb + sft) * (NI / 4) + q4];
            STEP_FMAS(win, sft, TX, l)
          }
        }
#endif
      }

      if (lane_valid) {
        const int OW = d.OW, OH = d.OH;
#pragma unroll
        for (int k = 0; k < TX; ++k) {
          const int x = x0 + k;
          if (x < OW) {
            float* o = Cb + (((long)pair * (NI * d.G) + NI * g) * OH + yb + y) * OW + x;
            const long istr = (long)OH * OW;
#pragma unroll
            for (int e = 0; e < NI / 2; ++e) {
              float a0, a1;
              unpack2(acc[k][e], a0, a1);
              o[(2 * e) * istr] = a0; o[(2 * e + 1) * istr] = a1;
            }
          }
        }
      }
    }
    t = seg_end;
  }
  __syncthreads();
  if (threadIdx.x == 0) g_cyc[blockIdx.x] = clock64() - cyc0;
}

// naive check: one thread per output element of selected (pair, image)
__global__ void naive_kernel(const float* E, const float* Lt, const int* pair_cond, Dims d, int pair, int n, float* out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= d.OH * d.OW) return;
  const int y = idx / d.OW, x = idx % d.OW;
  const int j = pair_cond[pair], g = n / NI, e = n % NI;
  const float* P = E + (long)pair * (2 * d.H) * (2 * d.W);
  double s = 0;
  for (int u = 0; u < d.H; ++u)
    for (int v = 0; v < d.W; ++v)
      s += (double)P[(y + u) * (2 * d.W) + x + v] * (double)Lt[((((long)j * d.G + g) * d.Hp + u) * d.Wp + v) * NI + e];
  out[idx] = (float)s;
}

static int cdiv(int a, int b) { return (a + b - 1) / b; }

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64;
  const int H = 60, W = 90, K = 7, P = 49;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  Dims d; d.B = B; d.H = H; d.W = W; d.P = P; d.G = cdiv(B, NI);
  d.OH = H + 1; d.OW = W + 1; d.KH = H; d.Hp = cdiv(H, 4) * 4; d.Wp = cdiv(W, TX) * TX; d.XG = cdiv(d.OW, TX);
  int need = d.XG * TX + d.Wp + TX; if (need < 2 * W) need = 2 * W;
  const int want = (TX * d.XG) % 32; int ps = need; while (ps % 32 != want) ++ps; d.pstride = ps;
  d.NBD = 1; d.TB = d.OH; d.prows = d.TB - 1 + d.KH; d.tiles = d.TB * d.XG; d.NS = cdiv(d.tiles, 32);
  const size_t smem = ((size_t)((d.prows * d.pstride + 3) & ~3) + (size_t)NW * ST * d.Wp * NI) * 4 + NW * ST * 8;
  printf("B=%d G=%d tiles=%d NS=%d pstride=%d prows=%d smem=%zu B  ST=%d NI=%d NW=%d VARIANT=%d\n", B, d.G, d.tiles, d.NS, d.pstride, d.prows, smem, ST, NI, NW, VARIANT);

  const long nE = (long)P * 2 * H * 2 * W, nL = (long)(K + 1) * d.G * d.Hp * d.Wp * NI, nC = (long)P * NI * d.G * d.OH * d.OW;
  std::vector<float> hE(nE), hL(nL, 0.f);
  std::vector<int> hcond(P);
  uint32_t s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.f; };
  for (auto& v : hE) v = 0.1f + 0.05f * rnd();
  for (int j = 0; j < K + 1; ++j)
    for (int g = 0; g < d.G; ++g)
      for (int u = 0; u < H; ++u)
        for (int v = 0; v < W; ++v)
          for (int e = 0; e < NI; ++e)
            if (NI * g + e < B) hL[((((long)j * d.G + g) * d.Hp + u) * d.Wp + v) * NI + e] = rnd() * 1e-3f;
  for (int p = 0; p < P; ++p) { int i = p / K, c = p % K; hcond[p] = c >= i ? c + 1 : c; }
  float *E, *L, *C, *ref; int* cond;
  CK(cudaMalloc(&E, nE * 4)); CK(cudaMalloc(&L, nL * 4)); CK(cudaMalloc(&C, nC * 4)); CK(cudaMalloc(&ref, d.OH * d.OW * 4)); CK(cudaMalloc(&cond, P * 4));
  CK(cudaMemcpy(E, hE.data(), nE * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(L, hL.data(), nL * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(cond, hcond.data(), P * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(C, 0xff, nC * 4));
  CK(cudaFuncSetAttribute(sm_conv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaEventRecord(e0));
    sm_conv2_kernel<<<sms, NW * 32, smem>>>(E, L, cond, d, C);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep >= 2 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  static long long hc[1024];
  CK(cudaMemcpyFromSymbol(hc, g_cyc, sizeof(long long) * sms));
  double cavg = 0, cmax = 0; for (int i = 0; i < sms; ++i) { cavg += (double)hc[i]; if ((double)hc[i] > cmax) cmax = (double)hc[i]; } cavg /= sms;
  const double mac = (double)P * B * (double)d.OH * d.OW * H * W;
  const double T = (double)P * d.G * d.NS;
  const double ideal_cyc = T / (sms * 4.0) * (double)d.KH * d.Wp * (TX * NI / 2) * 2;   // FFMA2 pipe cycles per SM sub-partition, executed work
  const double alg_cyc = mac / (sms * 128.0);                                // algorithmic MACs at 128 FMA/clk/SM
  printf("v2 kernel: %.3f ms  %.2f TFLOP/s algorithmic   clock %.0f MHz   cycles avg %.0f max %.0f   executed-FFMA2 pipe eff %.4f   algorithmic eff %.4f\n",
         best, 2 * mac / best / 1e9, cmax / best / 1e3, cavg, cmax, ideal_cyc / cmax, alg_cyc / cmax);

  // check a few (pair, image)
  std::vector<float> hC(d.OH * d.OW), hR(d.OH * d.OW);
  double worst = 0;
  const int checks[][2] = {{0, 0}, {5, 3}, {17, B - 1}, {48, B / 2}, {31, 6 % B}};
  for (auto& c : checks) {
    naive_kernel<<<cdiv(d.OH * d.OW, 128), 128>>>(E, L, cond, d, c[0], c[1], ref);
    CK(cudaMemcpy(hR.data(), ref, d.OH * d.OW * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hC.data(), C + ((long)c[0] * NI * d.G + c[1]) * d.OH * d.OW, d.OH * d.OW * 4, cudaMemcpyDeviceToHost));
    double mx = 0, mr = 0;
    for (int i = 0; i < d.OH * d.OW; ++i) { mx = fmax(mx, fabs((double)hC[i] - hR[i])); mr = fmax(mr, fabs((double)hR[i])); }
    if (!(mx / mr < 1e30)) mx = 1e30;
    worst = fmax(worst, mx / mr);
  }
  printf("check vs naive: worst relative error %.3e  %s\n", worst, worst < 1e-5 ? "OK" : "MISMATCH");
  return 0;
}

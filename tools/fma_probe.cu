// FFMA2 operand-pattern probes for the spatial-model inner loop (measurement tool, not part of libjcm.so).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/fma_probe tools/fma_probe.cu ; ./tools/fma_probe
// Every kernel is register-only inside its loop; FLOPs = 2 per FMA lane-op.  Prints TFLOP/s per variant.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef unsigned long long u64;
__device__ long long g_cyc[1024];
#define CYC_BEGIN __syncthreads(); const long long cyc0 = clock64();
#define CYC_END __syncthreads(); if (threadIdx.x == 0) g_cyc[blockIdx.x] = clock64() - cyc0;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float sum2(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }
// d += (w,w) * l    (SASS: FFMA2 Rd, Rw.F32, Rl, Rd)
__device__ __forceinline__ void f2s(u64& d, float w, u64 l) {
  asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %1};\n\tfma.rn.f32x2 %0, t, %2, %0;\n\t}" : "+l"(d) : "f"(w), "l"(l));
}
// d += a * l with a 64-bit a
__device__ __forceinline__ void f2p(u64& d, u64 a, u64 l) { asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(l)); }
__device__ __forceinline__ void f1(float& d, float a, float b) { asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(d) : "f"(a), "f"(b)); }

// ---- outer-product probes: NA window values x NB packed pairs, one "step" = NA*NB FFMA2, window rotates by one per step
// ORDER 0: a-major  1: b-major  2: a-major snake  3: b-major snake
// AFORM 0: scalar broadcast a (.F32)   1: 64-bit duplicated a     2: scalar FFMA (NB pairs = 2*NB scalars)
template <int NA, int NB, int ORDER, int AFORM, int NT>
__global__ void __launch_bounds__(NT, 1) outer_kernel(float* io, int iters) {
  CYC_BEGIN
  const float* src = io + (threadIdx.x & 31);
  float a[NA]; u64 ap[NA]; u64 b[2][NB]; u64 acc[NA][NB];
  float bs[2][NB * 2], accs[NA][NB * 2];
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    a[k] = src[k * 32]; ap[k] = pk(a[k], a[k]);
#pragma unroll
    for (int e = 0; e < NB; ++e) { acc[k][e] = 0ull; accs[k][2 * e] = 0.f; accs[k][2 * e + 1] = 0.f; }
  }
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int e = 0; e < NB; ++e) {
      bs[s][2 * e] = src[(64 + s * 16 + 2 * e) * 32]; bs[s][2 * e + 1] = src[(65 + s * 16 + 2 * e) * 32];
      b[s][e] = pk(bs[s][2 * e], bs[s][2 * e + 1]);
    }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s = 0; s < 2 * NA; ++s) {
#pragma unroll
      for (int i = 0; i < NA * NB; ++i) {
        int k, e;
        if (ORDER == 0) { k = i / NB; e = i % NB; }
        else if (ORDER == 1) { e = i / NA; k = i % NA; }
        else if (ORDER == 2) { k = i / NB; e = i % NB; if (k & 1) e = NB - 1 - e; }
        else { e = i / NA; k = i % NA; if (e & 1) k = NA - 1 - k; }
        if (AFORM == 0) f2s(acc[k][e], a[(s + k) % NA], b[s & 1][e]);
        else if (AFORM == 1) f2p(acc[k][e], ap[(s + k) % NA], b[s & 1][e]);
        else { f1(accs[k][2 * e], a[(s + k) % NA], bs[s & 1][2 * e]); f1(accs[k][2 * e + 1], a[(s + k) % NA], bs[s & 1][2 * e + 1]); }
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int k = 0; k < NA; ++k)
#pragma unroll
    for (int e = 0; e < NB; ++e) r += sum2(acc[k][e]) + accs[k][2 * e] + accs[k][2 * e + 1];
  io[8192 + blockIdx.x * NT + threadIdx.x] = r;
  CYC_END
}

// ---- slot probes: 16 accumulators in slot C; which of a (32-bit scalar) / b (64-bit) change per instruction
// MODE 0: a, b constant   1: a rotates over 8 registers, b constant   2: b rotates over 8 pairs, a constant   3: both rotate
// MODE 4: as 0 but accumulator in slot A (d = d * b + c: the classic peak loop)   5: 64-bit a rotating, b const  6: 64-bit a and b rotate
template <int MODE, int NT>
__global__ void __launch_bounds__(NT, 1) slot_kernel(float* io, int iters) {
  CYC_BEGIN
  const float* src = io + (threadIdx.x & 31);
  float a[8]; u64 ap[8], b[8], acc[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) { a[k] = src[k * 32]; ap[k] = pk(a[k], a[k]); b[k] = pk(src[(64 + 2 * k) * 32], src[(65 + 2 * k) * 32]); }
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = pk((float)i, (float)threadIdx.x);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int ia = (MODE == 1 || MODE == 3 || MODE == 5 || MODE == 6) ? (i + r) % 8 : 0;
        const int ib = (MODE == 2 || MODE == 3 || MODE == 6) ? (i + 3 * r) % 8 : 0;
        if (MODE == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(ap[0]), "l"(b[0]));
        else if (MODE >= 5) f2p(acc[i], ap[ia], b[ib]);
        else f2s(acc[i], a[ia], b[ib]);
      }
    }
  }
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) r += sum2(acc[i]);
  io[8192 + blockIdx.x * NT + threadIdx.x] = r;
  CYC_END
}


// ---- issue-slot probe: the b-major 7 x 2 FFMA2 step (14 FFMA2) with NX extra independent instructions of another pipe per step
// XT 0: IADD3 (alu pipe)   1: LDS.32 (conflict-free, result unused by the FFMA2)   2: LDS.128 broadcast   3: LDS.32 + LDS.128 feeding the FFMA2
template <int NX, int XT, int NT>
__global__ void __launch_bounds__(NT, 1) mix_kernel(float* io, int iters) {
  CYC_BEGIN
  __shared__ __align__(16) float sh[4096];
  for (int i = threadIdx.x; i < 4096; i += NT) sh[i] = io[i & 2047];
  __syncthreads();
  const float* src = io + (threadIdx.x & 31);
  float a[7]; u64 b[2], acc[7][2];
#pragma unroll
  for (int k = 0; k < 7; ++k) { a[k] = src[k * 32]; acc[k][0] = 0ull; acc[k][1] = 0ull; }
  b[0] = pk(src[64 * 32], src[65 * 32]); b[1] = pk(src[66 * 32], src[67 * 32]);
  int cnt = threadIdx.x; float fx = 0.f; float4 f4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* lp = sh + (threadIdx.x & 31);
  const float4* bp = reinterpret_cast<const float4*>(sh + 1024);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      if (XT == 3) {
        a[(s + 6) % 7] = lp[((it * 7 + s) & 63) * 32];
        const float4 t = bp[(it * 7 + s) & 127];
        b[0] = pk(t.x, t.y); b[1] = pk(t.z, t.w);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int k = 0; k < 7; ++k) f2s(acc[k][e], a[(s + k) % 7], b[e]);
#pragma unroll
      for (int x = 0; x < NX; ++x) {
        if (XT == 0) asm volatile("add.s32 %0, %0, %1;" : "+r"(cnt) : "r"(it));
        else if (XT == 1) { float t; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"((unsigned)__cvta_generic_to_shared(lp + ((s + x) & 7) * 32))); fx = t; }
        else if (XT == 2) { float4 t; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"((unsigned)__cvta_generic_to_shared(bp + ((s + x) & 7)))); f4 = t; }
      }
    }
  }
  float r = (float)cnt + fx + f4.x + f4.y + f4.z + f4.w;
#pragma unroll
  for (int k = 0; k < 7; ++k) r += sum2(acc[k][0]) + sum2(acc[k][1]);
  io[8192 + blockIdx.x * NT + threadIdx.x] = r;
  CYC_END
}

static float* g_io;
static int g_sms;
template <typename F>
static void run(const char* name, F launch, double flops_per_thread_iter, int nt, int iters) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    launch(iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  CK(cudaGetLastError());
  const double fl = flops_per_thread_iter * iters * (double)nt * g_sms;
  static long long hc[1024];
  CK(cudaMemcpyFromSymbol(hc, g_cyc, sizeof(long long) * g_sms));
  double cyc = 0; for (int i = 0; i < g_sms; ++i) cyc += (double)hc[i]; cyc /= g_sms;
  // FMA-pipe cycles needed per SM sub-partition: FFMA2 per thread (flops / 4) * warps per sub-partition * 2 (scalar FFMA: flops / 2 * 1)
  const double need = flops_per_thread_iter * iters / 4.0 * (nt / 128.0) * 2.0;
  printf("%-46s nt=%4d  %7.2f TFLOP/s  (%.3f ms)  pipe eff %.4f  clock %.0f MHz\n", name, nt, fl / best / 1e9, best, need / cyc, cyc / best / 1e3);
  fflush(stdout);
}

#define OUTER(NA, NB, ORDER, AFORM, NT) \
  run("outer NA=" #NA " NB=" #NB " order=" #ORDER " aform=" #AFORM, [&](int it) { outer_kernel<NA, NB, ORDER, AFORM, NT><<<g_sms, NT>>>(g_io, it); }, \
      2.0 * NA * 2.0 * NA * NB * 2.0, NT, 40000 / NA)
#define MIX(NX, XT, NT) \
  run("mix 14 FFMA2 + NX=" #NX " extra, type=" #XT, [&](int it) { mix_kernel<NX, XT, NT><<<g_sms, NT>>>(g_io, it); }, 7.0 * 14 * 4.0, NT, 6000)
#define SLOT(MODE, NT) \
  run("slot mode=" #MODE, [&](int it) { slot_kernel<MODE, NT><<<g_sms, NT>>>(g_io, it); }, 2.0 * 8 * 16 * 2.0, NT, 10000)

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  g_sms = p.multiProcessorCount;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %d kHz, FMA peak at max clock %.1f TFLOP/s\n", p.name, g_sms, clk, g_sms * 128 * 2.0 * clk * 1e3 / 1e12);
  CK(cudaMalloc(&g_io, (8192 + 1024 * 256) * sizeof(float)));
  {
    float* h = (float*)malloc(8192 * sizeof(float));
    for (int i = 0; i < 8192; ++i) h[i] = 1e-3f * (float)((i * 37) % 101) / 101.f;
    CK(cudaMemcpy(g_io, h, 8192 * sizeof(float), cudaMemcpyHostToDevice));
    free(h);
  }
  for (int w = 0; w < 30; ++w) slot_kernel<4, 512><<<g_sms, 512>>>(g_io, 10000);  // warm the clocks up
  CK(cudaDeviceSynchronize());
  SLOT(4, 640); SLOT(1, 640);
  MIX(0, 0, 640); MIX(1, 0, 640); MIX(2, 0, 640); MIX(4, 0, 640);
  MIX(1, 1, 640); MIX(2, 1, 640); MIX(4, 1, 640);
  MIX(1, 2, 640); MIX(2, 2, 640); MIX(4, 2, 640);
  MIX(0, 3, 640); MIX(0, 3, 512); MIX(0, 3, 768);
  return 0;
}

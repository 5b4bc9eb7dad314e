// Part-detector convolutions as a TMA-fed implicit GEMM on the 5th-gen tensor cores (tcgen05 / TMEM), sm_100a.
//
// Replaces the reference's `tf.nn.conv2d(x, W, strides, padding='SAME') + b` (+ `tf.nn.relu`)
// (reference main.py:133-135, 156-162) for every stride-1 layer; the three stride-2 `conv1_*` layers are
// mapped onto the same kernel by a space-to-depth transform done in prep.cu (5x5 s2 over 3 ch == 3x3 s1 over 16 ch).
//
// GEMM view:  D[m, co] = sum_{tap, ci} X[pix(m) + tap - pad, ci] * Wp[tap, co, ci]
//   * M tile = 128 output pixels = a TW x TH spatial patch of ONE image (TW*TH == 128). The A operand of tap
//     (dy,dx) is the same patch shifted by (dy-pad, dx-pad): one 4-D TMA box {kc, TW, TH, 1} per k-block, the
//     SAME zero padding comes for free from TMA out-of-bounds fill (coordinates may be negative).
//   * N tile = block_n output channels (16..256), B operand = packed weights [tap][Cout_pad][Cin] (K-major),
//     3-D TMA box {kc, block_n, 1}.
//   * K loop = taps x (Cin / kc) x terms.  kc = 64 (SWIZZLE_128B), 32 (SWIZZLE_64B) or 16 (SWIZZLE_32B, the s2d conv1 layers).
//   * terms = 1: plain bf16 operands (training config).  terms = 3: "bf16x3" split accumulation
//     (A_hi*B_hi + A_lo*B_hi + A_hi*B_lo) which reproduces fp32 products to ~2^-16 - this is the fp32 config.
//   * accumulators: fp32 in TMEM, 2 stages x 256 columns, so the epilogue of tile t overlaps the MMAs of t+1.
//   * warp roles: w0 TMA producer, w1 MMA issuer (one elected lane), w2 TMEM allocator, w4-7 epilogue
//     (tcgen05.ld -> +bias -> ReLU -> fp32 NHWC store).  Persistent grid, one CTA per SM.
#include "common.cuh"
#include "tiling.cuh"
#include <cuda.h>
#include <stdlib.h>

// Measurement switches (environment variables that change tiling / skip work, and the never-adopted CTA-pair kernel) exist only in
// builds with -DJCM_EXPERIMENTS (`python joint-cnn-mrf_b200/build.py --experiments` -> libjcm_exp.so, used by tests/gpu_diag.py).
// The product library compiles them out: it reads no environment variable and carries no skip-the-work path.
#ifdef JCM_EXPERIMENTS
#define JCM_ENV_INT(name, dflt) (getenv(name) ? atoi(getenv(name)) : (dflt))
#define JCM_DBG(p) ((p).dbg)
#define JCM_SPLITN(p) ((p).mma_split_n)
#else
#define JCM_ENV_INT(name, dflt) (dflt)
#define JCM_DBG(p) 0
#define JCM_SPLITN(p) 0
#endif

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kTileM = 128;

struct ConvParams {
  int B, H, W, Cout;
  int TW, TH, tiles_x, tiles_y;
  int n_tiles, block_n;
  int kc, cblocks, ksize, kw, pad, pad_x, terms;   // ksize x kw taps (kernel height x width), SAME padding pad / pad_x
  int stages, a_bytes, b_bytes, stage_bytes;
  int relu;
  int y_bf16;           // 1: the output activation is stored as bf16 (bf16 training configuration), TMA-store epilogue only
  int tma_store;        // 1: epilogue stages 32-column chunks in swizzled shared memory and writes them with TMA stores
  // halo mode (layers whose A operand traffic is the limiter: N <= 128): the A tile is loaded ONCE per (dx, channel block) as a
  // (TH + kh - 1)-row halo patch and the kh vertical taps address it at row offsets dy*TW (a multiple of the 8-row swizzle
  // period, so the UMMA descriptor just moves its start address); weights stream through their own ring, one tap at a time.
  int halo, a_stages, b_stages, a_halo_bytes, a_stride, b_stride;
  int mgroup;           // halo mode: M tiles that share every weight stage (2 where shared memory and TMEM allow: the weights are then
                        // streamed from L2 once per PAIR of tiles - for the 5x5 layers they are 3/4 of a tile's L2->SM traffic)
  int b_group, b_slot;  // halo mode: vertical taps per weight stage (one full/empty handshake and one tcgen05.commit per group: the
                        // handshake costs ~350 clk, more than the MMAs of one tap when N <= 128) and bytes per tap slot
  int mma_split_n;      // measurement switch (JCM_MMA_SPLITN): issue every MMA as two independent half-N MMAs
  int nacc;             // independent accumulators per tile (K steps are dealt round-robin to them and summed in the epilogue):
                        // back-to-back tcgen05.mma into the SAME TMEM tile serialise at ~170 clk each whatever N is, so for
                        // N <= 128 the tensor pipe idles unless consecutive MMAs target different accumulators
  // grouped forms used by the tensor-core spatial model (spatial_model.cu, "smt" section); plain mode only:
  //   grp 1: every image has its OWN weights (tap plane = img * grp_taps + tap) and taps whose input rows all lie outside the
  //          map are skipped (120-tap vertical kernels over 61-row maps: half of the taps of every tile)
  //   grp 2: batched GEMM with a shifted B operand: image = (batch a_div * q + s), A image q, weight plane q, the weight's K
  //          coordinate is offset by (s - sm_pad) * k_rows and only the k-blocks that can be non-zero are visited; an image map
  //          redirects either the weight plane (b_map) or the A image (a_map) of batch q
  int grp, grp_taps, a_div, k_rows, sm_pad, sm_rows, sm_rows_in;
  const int* a_map;     // grp 1: A image = a_map[img] (several pairs share one conditioning map); grp 2: a_map[img / a_div]; NULL: img
  const int* b_map;     // grp 2: weight plane = b_map[img / a_div]; NULL: img / a_div
  int dbg;              // measurement switches (JCM_CONV_DBG bit 0: epilogue releases TMEM without storing, bit 1: no MMAs issued,
                        // bit 2: producer skips the B loads) - results are garbage, timing only
  // mixed-shape pixel-tile plan (tiling.cuh; plain mode, one term, TMA-store epilogue): plan_n > 0 tiles per image, tile i is the box
  // psw[s] x psh[s] (s = pshape[i]) at (px0[i], py0[i]); its A box / output box come from ShapeMaps::a[s] / ::y[s]
  int plan_n;
  uint8_t px0[kPlanMaxTiles], py0[kPlanMaxTiles], pshape[kPlanMaxTiles];
  uint8_t psw[4], psh[4];
  // N-split tail: the last `tail_tiles` (M, N) tiles - the partial last wave of the persistent grid - run as tail_split sub-tiles of
  // block_n / tail_split output channels each on twice as many SMs (set only when the halves fit ONE sub-round, see the launcher)
  int tail_tiles, tail_split;
  const float* bias;
  float* y;
};

constexpr int kIgemmShapes = 4;
struct ShapeMaps {
  CUtensorMap a[kIgemmShapes];     // A-operand boxes {kc, sw, sh, 1} per plan shape
  CUtensorMap a_lo[kIgemmShapes];  // the same over the lo planes (fp32 configuration: bf16x3 split products)
  CUtensorMap y[kIgemmShapes];     // output boxes {32 | 64 channels, sw, sh, 1} per plan shape
  CUtensorMap b_tail;              // weight box {kc, block_n / tail_split, 1}
};

// (M, N) tile of the persistent loop.  Virtual index v < full: tile v, all block_n channels.  v >= full: tail sub-tile.
struct TileId { int nt, mt, n_off, n_cols; };
__device__ __forceinline__ TileId decode_tile(const ConvParams& p, int v, int m_tiles, int total_tiles) {
  TileId t;
  int tile = v;
  t.n_off = 0;
  t.n_cols = p.block_n;
  const int full = total_tiles - p.tail_tiles;
  if (v >= full) {
    const int u = v - full;
    tile = full + u / p.tail_split;
    t.n_cols = p.block_n / p.tail_split;
    t.n_off = (u - (u / p.tail_split) * p.tail_split) * t.n_cols;
  }
  t.nt = tile / m_tiles;
  t.mt = tile - t.nt * m_tiles;
  return t;
}
// pixel patch of M tile mt: image, origin, plan shape (-1: the uniform TW x TH grid)
struct Patch { int img, x0, y0, ty, shape; };
__device__ __forceinline__ Patch decode_patch(const ConvParams& p, int mt) {
  Patch q;
  if (p.plan_n) {
    q.img = mt / p.plan_n;
    const int i = mt - q.img * p.plan_n;
    q.x0 = p.px0[i];
    q.y0 = p.py0[i];
    q.shape = p.pshape[i];
    q.ty = 0;
  } else {
    const int per = p.tiles_y * p.tiles_x;
    q.img = mt / per;
    const int r = mt - q.img * per;
    q.ty = r / p.tiles_x;
    q.x0 = (r - q.ty * p.tiles_x) * p.TW;
    q.y0 = q.ty * p.TH;
    q.shape = -1;
  }
  return q;
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// one lane of a converged warp (elect.sync): unlike `lane == 0` the compiler knows the region is executed by a single thread
// and emits the uniform-datapath instructions (UTCHMMA / UTMALDG / UTCBAR) directly instead of wrapping each in an election loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread <-> TMEM lane)
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, rows of `row_bytes` (= swizzle span: 128 or 32 bytes),
// 8-row core groups `sbo` bytes apart.  Field layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);           // start address  [0,14)
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major) [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32; // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version = 1 (Blackwell) [46,48)
  d |= (uint64_t)(layout_type & 7) << 61;           // 2 = SWIZZLE_128B, 6 = SWIZZLE_32B
  return d;
}

// k-loop extent of one tile: taps [tap_lo, tap_hi) x channel blocks [cb_lo, cb_hi).  Whole extent for ordinary convolutions.
struct KRange { int tap_lo, tap_hi, cb_lo, cb_hi; };
__device__ __forceinline__ KRange tile_krange(const ConvParams& p, int img, int ty) {
  KRange r;
  r.tap_lo = 0; r.tap_hi = p.ksize * p.kw; r.cb_lo = 0; r.cb_hi = p.cblocks;
  if (p.grp == 1) {             // kw == 1: tap dy reads input rows y + dy - pad, y in [y0, y0 + TH); keep those that touch [0, H)
    const int y0 = ty * p.TH;
    r.tap_lo = max(0, p.pad - (y0 + p.TH - 1));
    r.tap_hi = min(p.ksize, p.H + p.pad - y0);
    if (r.tap_hi <= r.tap_lo) r.tap_hi = r.tap_lo + 1;
  } else if (p.grp == 2) {      // shift s: rows y of the A operand with 0 <= y + s - sm_pad < sm_rows_in, as a range of k-blocks
    const int s = img % p.a_div;
    const int ylo = max(0, p.sm_pad - s), yhi = min(p.sm_rows, p.sm_rows_in + p.sm_pad - s);
    r.cb_lo = (ylo * p.k_rows) / p.kc;
    r.cb_hi = min(p.cblocks, (yhi * p.k_rows + p.kc - 1) / p.kc);
    if (r.cb_hi <= r.cb_lo) { r.cb_lo = 0; r.cb_hi = 1; }
  }
  return r;
}

// ------------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(kThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                  const __grid_constant__ CUtensorMap map_y, const __grid_constant__ ShapeMaps smaps, const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);
  const uint32_t bar_afull = smem_u32(&bars[2 * kMaxStages + 4]);          // halo mode: ring of A halo tiles
  const uint32_t bar_aempty = smem_u32(&bars[3 * kMaxStages + 4]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
      mbar_init(bar_afull + 8 * s, 1);
      mbar_init(bar_aempty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int m_tiles = p.B * (p.plan_n ? p.plan_n : p.tiles_y * p.tiles_x);
  const int total_tiles = m_tiles * p.n_tiles;
  const int virt_tiles = total_tiles + p.tail_tiles * (p.tail_split - 1);   // loop extent of the plain-mode roles (tail sub-tiles)
  const int taps = p.ksize * p.kw;
  const int num_kb = taps * p.cblocks * p.terms;
  const uint32_t smem_b0 = smem_base + p.a_stages * p.a_stride;     // halo mode: start of the weight ring

  if (warp == 0 && p.halo) {
    // ===================== TMA producer, halo mode (map_a_lo carries the halo-box tensor map) =====================
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const int m_groups = (m_tiles + p.mgroup - 1) / p.mgroup;
      for (int tg = blockIdx.x; tg < m_groups * p.n_tiles; tg += gridDim.x) {
        const int nt = tg / m_groups, mt0 = (tg - nt * m_groups) * p.mgroup;
        const int n_in = min(p.mgroup, m_tiles - mt0);       // tiles of this group (the last group may hold one)
        const Patch pt0 = decode_patch(p, mt0), pt1 = decode_patch(p, mt0 + n_in - 1);
        for (int dx = 0; dx < p.kw; ++dx) {
          for (int cb = 0; cb < p.cblocks; ++cb) {
            for (int i = 0; i < n_in; ++i) {
              const Patch& pt = i ? pt1 : pt0;
              mbar_wait(bar_aempty + 8 * sa, pa ^ 1);
              mbar_expect_tx(bar_afull + 8 * sa, p.a_halo_bytes);
              tma_load_4d(smem_base + sa * p.a_stride, &map_a_lo, bar_afull + 8 * sa, cb * p.kc, pt.x0 - p.pad_x + dx, pt.y0 - p.pad, pt.img);
              if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            }
            if (p.b_group == 1) {
              for (int dy = 0; dy < p.ksize; ++dy) {
                mbar_wait(bar_empty + 8 * sb, pb ^ 1);
                mbar_expect_tx(bar_full + 8 * sb, p.b_bytes);
                tma_load_3d(smem_b0 + sb * p.b_stride, &map_b_hi, bar_full + 8 * sb, cb * p.kc, nt * p.block_n, dy * p.kw + dx);
                if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
              }
            } else
            for (int dy0 = 0; dy0 < p.ksize; dy0 += p.b_group) {
              const int gn = min(p.b_group, p.ksize - dy0);
              mbar_wait(bar_empty + 8 * sb, pb ^ 1);
              if (JCM_DBG(p) & 4) {
                mbar_arrive(bar_full + 8 * sb);
              } else {
                mbar_expect_tx(bar_full + 8 * sb, gn * p.b_bytes);
                for (int i = 0; i < gn; ++i)
                  tma_load_3d(smem_b0 + sb * p.b_stride + i * p.b_slot, &map_b_hi, bar_full + 8 * sb, cb * p.kc, nt * p.block_n,
                              (dy0 + i) * p.kw + dx);
              }
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1 && p.halo) {
    // ===================== MMA issuer, halo mode =====================
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t row_bytes = p.kc * 2;
      const uint32_t layout = (p.kc == 64) ? 2u : (p.kc == 32 ? 4u : 6u);
      const uint32_t sbo = 8 * row_bytes;
      const uint32_t dy_bytes = (uint32_t)p.TW * row_bytes;     // one patch row of the halo tile; TW % 8 == 0 keeps the swizzle phase
      const int kk = p.kc / 16;
      const uint64_t desc_hi = make_smem_desc(0, sbo, layout);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      if (p.mgroup == 1) {
      // one M tile per weight stage: the issue loops below are kept exactly as tuned (the single issuing thread sets the pace of the
      // small-K layers; restructuring them for the two-tile case cost 20 % on those layers)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accumulate = 0, mcount = 0;
        for (int it = 0; it < p.kw * p.cblocks; ++it) {
          mbar_wait(bar_afull + 8 * sa, pa);
          tc_fence_after();
          const uint32_t a_addr = smem_base + sa * p.a_stride;
          if (p.b_group == 1) {
            // one tap per weight stage (N = 128 layers: grouping does not pay there; this flat loop compiles to the tighter issue loop)
            for (int dy = 0; dy < p.ksize; ++dy) {
              mbar_wait(bar_full + 8 * sb, pb);
              tc_fence_after();
              const uint64_t adesc = desc_hi | (uint64_t)(((a_addr + dy * dy_bytes) >> 4) & 0x3FFF);
              const uint64_t bdesc = desc_hi | (uint64_t)(((smem_b0 + sb * p.b_stride) >> 4) & 0x3FFF);
              if (!(JCM_DBG(p) & 2))
              for (int k = 0; k < kk; ++k) {
                const int j = (mcount++) & (p.nacc - 1);
                tc_mma_bf16(d_tmem + j * p.block_n, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (accumulate >> j) & 1u);
                accumulate |= 1u << j;
              }
              tc_commit(bar_empty + 8 * sb);
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
            }
          } else
          for (int dy0 = 0; dy0 < p.ksize; dy0 += p.b_group) {
            const int gn = min(p.b_group, p.ksize - dy0);
            mbar_wait(bar_full + 8 * sb, pb);
            tc_fence_after();
            for (int i = 0; i < gn; ++i) {
              const uint64_t adesc = desc_hi | (uint64_t)(((a_addr + (dy0 + i) * dy_bytes) >> 4) & 0x3FFF);
              const uint64_t bdesc = desc_hi | (uint64_t)(((smem_b0 + sb * p.b_stride + i * p.b_slot) >> 4) & 0x3FFF);
              if (!(JCM_DBG(p) & 2))
              for (int k = 0; k < kk; ++k) {
                const int j = (mcount++) & (p.nacc - 1);
                tc_mma_bf16(d_tmem + j * p.block_n, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (accumulate >> j) & 1u);
                accumulate |= 1u << j;
              }
            }
            tc_commit(bar_empty + 8 * sb);
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
          tc_commit(bar_aempty + 8 * sa);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
        tc_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      } else {
      // two M tiles per weight stage (one tap per stage): tile 0 accumulates in columns [0, block_n), tile 1 in [block_n, 2 block_n).
      // The MMAs here are short (N <= 128: 32-64 clocks) and one thread issues eight per tap, so its address arithmetic is kept
      // minimal: the descriptors' high word is constant, the low word is a base (per halo patch / weight stage) plus small offsets.
      const int m_groups = (m_tiles + 1) / 2;
      const uint32_t d_hi = (uint32_t)(desc_hi >> 32), d_lo0 = (uint32_t)desc_hi;
      const uint32_t dy4 = dy_bytes >> 4, bstr4 = (uint32_t)p.b_stride >> 4, b04 = smem_b0 >> 4;
      for (int tg = blockIdx.x; tg < m_groups * p.n_tiles; tg += gridDim.x) {
        const bool two = 2 * (tg % m_groups) + 1 < m_tiles;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * 256, d1 = d0 + p.block_n;
        uint32_t accf = 0;
        for (int it = 0; it < p.kw * p.cblocks; ++it) {
          mbar_wait(bar_afull + 8 * sa, pa);
          uint32_t a0l = d_lo0 | ((smem_base + sa * p.a_stride) >> 4);
          int sa1 = sa + 1;
          uint32_t pa1 = pa;
          if (sa1 == p.a_stages) { sa1 = 0; pa1 ^= 1; }
          if (two) mbar_wait(bar_afull + 8 * sa1, pa1);
          uint32_t a1l = d_lo0 | ((smem_base + sa1 * p.a_stride) >> 4);
          tc_fence_after();
          for (int dy = 0; dy < p.ksize; ++dy) {
            mbar_wait(bar_full + 8 * sb, pb);
            tc_fence_after();
            const uint32_t bl = d_lo0 | (b04 + sb * bstr4);
#pragma unroll 4
            for (int k = 0; k < kk; ++k) {
              uint64_t ad, bd;
              asm("mov.b64 %0, {%1, %2};" : "=l"(ad) : "r"(a0l + 2 * k), "r"(d_hi));
              asm("mov.b64 %0, {%1, %2};" : "=l"(bd) : "r"(bl + 2 * k), "r"(d_hi));
              tc_mma_bf16(d0, ad, bd, idesc, k ? 1u : accf);
            }
            if (two) {
#pragma unroll 4
              for (int k = 0; k < kk; ++k) {
                uint64_t ad, bd;
                asm("mov.b64 %0, {%1, %2};" : "=l"(ad) : "r"(a1l + 2 * k), "r"(d_hi));
                asm("mov.b64 %0, {%1, %2};" : "=l"(bd) : "r"(bl + 2 * k), "r"(d_hi));
                tc_mma_bf16(d1, ad, bd, idesc, k ? 1u : accf);
              }
            }
            accf = 1u;
            a0l += dy4;
            a1l += dy4;
            tc_commit(bar_empty + 8 * sb);
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
          tc_commit(bar_aempty + 8 * sa);
          sa = sa1; pa = pa1;
          if (two) {
            tc_commit(bar_aempty + 8 * sa);
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
          }
        }
        tc_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      }
    }
  } else if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int vt = blockIdx.x; vt < virt_tiles; vt += gridDim.x) {
        const TileId tid = decode_tile(p, vt, m_tiles, total_tiles);
        const int nt = tid.nt;
        const Patch pt = decode_patch(p, tid.mt);
        const int img = pt.img, ty = pt.ty;
        const int x0 = pt.x0 - p.pad_x, y0 = pt.y0 - p.pad;
        if (p.grp) {
          const KRange kr = tile_krange(p, img, ty);
          const int q = img / p.a_div;
          const int a_img = p.grp == 2 ? (p.a_map ? p.a_map[q] : q) : (p.a_map ? p.a_map[img] : img);
          const int b_plane = p.grp == 2 ? (p.b_map ? p.b_map[q] : q) : img * p.grp_taps;
          const int b_koff = p.grp == 2 ? (img % p.a_div - p.sm_pad) * p.k_rows : 0;
          for (int tap = kr.tap_lo; tap < kr.tap_hi; ++tap) {
            for (int cb = kr.cb_lo; cb < kr.cb_hi; ++cb) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              const uint32_t sa = smem_base + stage * p.stage_bytes;
              const uint32_t fb = bar_full + 8 * stage;
              mbar_expect_tx(fb, p.a_bytes + p.b_bytes);
              tma_load_4d(sa, &map_a_hi, fb, cb * p.kc, x0, y0 + tap, a_img);
              tma_load_3d(sa + p.a_bytes, &map_b_hi, fb, cb * p.kc + b_koff, nt * p.block_n, b_plane + (p.grp == 2 ? 0 : tap));
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
          continue;
        }
        // bytes one k-block brings in: the A box of this patch shape (a clipped plan shape has fewer rows) + this tile's weight rows
        const uint32_t a_tx = pt.shape >= 0 ? (uint32_t)p.psw[pt.shape] * p.psh[pt.shape] * p.kc * 2 : (uint32_t)p.a_bytes;
        const uint32_t b_tx = (uint32_t)tid.n_cols * p.kc * 2;
        const void* mb_hi = tid.n_cols == p.block_n ? (const void*)&map_b_hi : (const void*)&smaps.b_tail;
        const void* ma_plan = pt.shape >= 0 ? (const void*)&smaps.a[pt.shape] : (const void*)&map_a_hi;
        const void* ma_plan_lo = pt.shape >= 0 ? (const void*)&smaps.a_lo[pt.shape] : (const void*)&map_a_lo;
        for (int tap = 0; tap < taps; ++tap) {
          const int dy = tap / p.kw, dx = tap - dy * p.kw;
          for (int cb = 0; cb < p.cblocks; ++cb) {
            for (int term = 0; term < p.terms; ++term) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              const uint32_t sa = smem_base + stage * p.stage_bytes;
              const uint32_t sb = sa + p.a_bytes;
              const uint32_t fb = bar_full + 8 * stage;
              mbar_expect_tx(fb, (JCM_DBG(p) & 4) ? a_tx : a_tx + b_tx);
              tma_load_4d(sa, term == 1 ? ma_plan_lo : ma_plan, fb, cb * p.kc, x0 + dx, y0 + dy, img);
              if (!(JCM_DBG(p) & 4)) tma_load_3d(sb, term == 2 ? (const void*)&map_b_lo : mb_hi, fb, cb * p.kc, nt * p.block_n + tid.n_off, tap);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N = block_n, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t row_bytes = p.kc * 2;                  // 128, 64 or 32 = the swizzle span
      const uint32_t layout = (p.kc == 64) ? 2u : (p.kc == 32 ? 4u : 6u);  // SWIZZLE_128B : SWIZZLE_64B : SWIZZLE_32B
      const uint32_t sbo = 8 * row_bytes;
      const int kk = p.kc / 16;
      const uint64_t desc_hi = make_smem_desc(0, sbo, layout);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int vt = blockIdx.x; vt < virt_tiles; vt += gridDim.x) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accumulate = 0, mcount = 0;     // bit j: accumulator j of this tile has been written; MMAs are dealt round-robin
        int tile_kb = num_kb;
        const TileId tid = decode_tile(p, vt, m_tiles, total_tiles);
        const uint32_t idesc_t = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(tid.n_cols >> 3) << 17);    // N of this tile (tail sub-tiles: block_n / tail_split)
        if (p.grp) {
          const Patch pt = decode_patch(p, tid.mt);
          const KRange kr = tile_krange(p, pt.img, pt.ty);
          tile_kb = (kr.tap_hi - kr.tap_lo) * (kr.cb_hi - kr.cb_lo);
        }
        for (int kb = 0; kb < tile_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint64_t adesc = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);                 // only the start address changes per stage
          const uint64_t bdesc = desc_hi | (uint64_t)(((sa + p.a_bytes) >> 4) & 0x3FFF);
          if (JCM_DBG(p) & 2) {
          } else if (JCM_SPLITN(p)) {
            // experiment: two independent half-N MMAs per K step (different TMEM columns, B rows n/2.. of the same stage)
            const uint32_t hn = p.block_n / 2;
            const uint32_t idesc_h = (idesc & ~(0x3Fu << 17)) | ((hn >> 3) << 17);
            const uint64_t bdesc2 = make_smem_desc(sa + p.a_bytes + hn * row_bytes, sbo, layout);
            for (int k = 0; k < kk; ++k) {
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_h, (kb | k) != 0);
              tc_mma_bf16(d_tmem + hn, adesc + (uint64_t)(2 * k), bdesc2 + (uint64_t)(2 * k), idesc_h, (kb | k) != 0);
            }
          } else
          for (int k = 0; k < kk; ++k) {
            // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in the 16-byte address field
            // (measured: a specialised nacc == 1 path without the round-robin bookkeeping compiles to a SLOWER issue loop - kept as is)
            const int j = (mcount++) & (p.nacc - 1);
            tc_mma_bf16(d_tmem + j * p.block_n, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_t, (accumulate >> j) & 1u);
            accumulate |= 1u << j;
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;          // accumulator row == pixel index inside the patch
    const int ly = row / p.TW, lx = row - ly * p.TW;
    int acc = 0;
    uint32_t acc_phase = 0;
    int epi_chunk = 0;
    // work items of this kernel's roles: tile groups in halo mode (mgroup tiles share the weight stages and one accumulator
    // stage), virtual tiles (N-split tail) otherwise
    const int m_groups = (m_tiles + p.mgroup - 1) / p.mgroup;
    const int epi_items = p.halo ? m_groups * p.n_tiles : virt_tiles;
    for (int wi = blockIdx.x; wi < epi_items; wi += gridDim.x) {
      TileId tid;
      int n_in = 1;
      if (p.halo) {
        tid.nt = wi / m_groups;
        tid.mt = (wi - tid.nt * m_groups) * p.mgroup;
        tid.n_off = 0;
        tid.n_cols = p.block_n;
        n_in = min(p.mgroup, m_tiles - tid.mt);
      } else {
        tid = decode_tile(p, wi, m_tiles, total_tiles);
      }
      const int nt = tid.nt;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      for (int gt = 0; gt < n_in; ++gt) {
      const Patch pt = decode_patch(p, tid.mt + gt);
      const int img = pt.img;
      const int oy = pt.y0 + ly, ox = pt.x0 + lx;
      const void* my = pt.shape >= 0 ? (const void*)&smaps.y[pt.shape] : (const void*)&map_y;
      const int ncols = tid.n_cols, ncol0 = nt * p.block_n + tid.n_off;
      const bool valid = (oy < p.H) && (ox < p.W);
      float* yrow = p.y + ((size_t)((size_t)img * p.H + oy) * p.W + ox) * p.Cout;
      const uint32_t taddr0 = tmem_base + acc * 256 + gt * p.block_n + ((uint32_t)(q * 32) << 16);
      if (JCM_DBG(p) & 1) {
        // timing experiment: no stores
      } else if (p.tma_store && p.y_bf16) {
        // bf16 output: 64-column chunks (128-byte rows of bf16) through the same swizzled staging buffers, box {64 ch, TW, TH, 1}
        const uint32_t stage0 = p.halo ? smem_b0 + p.b_stages * p.b_stride : smem_base + p.stages * p.stage_bytes;
        for (int c0 = 0; c0 < ncols; c0 += 64) {
          const uint32_t buf = stage0 + (uint32_t)(epi_chunk & 1) * (kTileM * 128);
          if (threadIdx.x == 128) bulk_wait_read<1>();
          epi_bar();
          uint32_t v[32], u[32];
          tc_ld32(taddr0 + c0, v);
          tc_ld32(taddr0 + c0 + 32, u);
          tc_ld_wait();
          const int co0 = ncol0 + c0;
          const uint32_t rowaddr = buf + (uint32_t)row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {       // 16-byte chunk j = channels co0 + 8j .. 8j+7
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(j < 4 ? v[8 * j + e] : u[8 * (j - 4) + e]);
            if (p.bias && co0 + 8 * j < p.Cout) {
              const float4 b0 = *reinterpret_cast<const float4*>(p.bias + co0 + 8 * j);
              const float4 b1 = *reinterpret_cast<const float4*>(p.bias + co0 + 8 * j + 4);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              w[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((j ^ (row & 7)) << 4)), "r"(w[0]), "r"(w[1]),
                         "r"(w[2]), "r"(w[3])
                         : "memory");
          }
          fence_proxy_async();
          epi_bar();
          if (threadIdx.x == 128) {
            if (co0 < p.Cout) tma_store_4d(my, buf, co0, pt.x0, pt.y0, img);
            bulk_commit();
          }
          ++epi_chunk;
        }
      } else if (p.tma_store) {
        // coalesced epilogue: 32-column chunks go through two 16 KB swizzled staging buffers and out with TMA stores (the box
        // {32 ch, TW, TH, 1} has the A operand's pixel order, so accumulator row == staging row; ragged edges are clipped by TMA)
        const uint32_t stage0 = p.halo ? smem_b0 + p.b_stages * p.b_stride : smem_base + p.stages * p.stage_bytes;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          const uint32_t buf = stage0 + (uint32_t)(epi_chunk & 1) * (kTileM * 128);
          if (threadIdx.x == 128) bulk_wait_read<1>();     // the store that last read this buffer is done with it
          epi_bar();
          uint32_t v[32];
          tc_ld32(taddr0 + c0, v);
          tc_ld_wait();
          for (int j = 1; j < p.nacc; ++j) {
            uint32_t u[32];
            tc_ld32(taddr0 + j * p.block_n + c0, u);
            tc_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
          }
          const int co0 = ncol0 + c0;
          const uint32_t rowaddr = buf + (uint32_t)row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            if (p.bias && co0 + 4 * j < p.Cout) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + co0 + 4 * j);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((j ^ (row & 7)) << 4)), "f"(o.x), "f"(o.y),
                         "f"(o.z), "f"(o.w)
                         : "memory");
          }
          fence_proxy_async();
          epi_bar();
          if (threadIdx.x == 128) {
            if (co0 < p.Cout) tma_store_4d(my, buf, co0, pt.x0, pt.y0, img);
            bulk_commit();   // one group per chunk even when nothing is stored (padded columns), so wait_group.read 1 == "chunk - 2 done"
          }
          ++epi_chunk;
        }
      } else {
      for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          tc_ld32(taddr0 + c0, v);
          tc_ld_wait();
          for (int j = 1; j < p.nacc; ++j) {
            uint32_t u[32];
            tc_ld32(taddr0 + j * p.block_n + c0, u);
            tc_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(u[e]));
          }
          const int co0 = ncol0 + c0;
          if (valid) {
            if (((p.Cout & 3) == 0) && (co0 + 32 <= p.Cout)) {
  #pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 o;
                float4 b = p.bias ? *reinterpret_cast<const float4*>(p.bias + co0 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                o.x = __uint_as_float(v[j + 0]) + b.x;
                o.y = __uint_as_float(v[j + 1]) + b.y;
                o.z = __uint_as_float(v[j + 2]) + b.z;
                o.w = __uint_as_float(v[j + 3]) + b.w;
                if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(yrow + co0 + j) = o;
              }
            } else {
  #pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int co = co0 + j;
                if (co < p.Cout && (c0 + j) < ncols) {
                  float o = __uint_as_float(v[j]) + (p.bias ? p.bias[co] : 0.f);
                  if (p.relu) o = fmaxf(o, 0.f);
                  yrow[co] = o;
                }
              }
            }
          }
        }
      }
      }   // tiles of the group
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.tma_store && threadIdx.x == 128) bulk_wait_all();   // shared memory must outlive the last TMA store
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// =====================================================================================================================
// CTA-pair form (`cta_group::2`) of the plain mode for N = 256 tiles: the layers that hold 84 % of the part detector's FLOPs
// (conv4_* / conv5 forward and data gradient) and the 1x1 tap GEMMs of conv6.  Two CTAs of a cluster own two consecutive
// 128-pixel M tiles of the same N tile; each loads its own A tile and HALF of the weight tile (128 of the 256 rows), the leader
// issues one M = 256 MMA per K step that reads both CTAs' shared memory and writes each CTA's own TMEM.  Per CTA and k-block the
// L2->SM traffic drops from 48 KB to 32 KB and the B operand's shared-memory reads halve.  Measured on B200 (tests/gpu_diag.py cta2,
// profiles/r02/diag_cta2_{on,off}.txt, isolated launches at batch 64): conv5 9.16 -> 8.77 ms best / 10.04 -> 8.77 ms median
// (1462 -> 1673 TFLOP/s algorithmic), conv4_fullres 5.84 -> 5.03 ms: the step is power-capped, and this form needs less energy
// per FLOP.
// Protocol (PTX forms as in CUTLASS's sm100 2-SM collective, cute/arch/copy_sm100_tma.hpp, cutlass/arch/barrier.h):
//   * full[s] lives in the LEADER (cluster rank 0): its producer arms it with the bytes of both CTAs; both producers' TMA
//     loads carry `.cta_group::2` and the leader's barrier address (own address with the peer bit 24 cleared).
//   * empty[s] / tfull[a] exist in both CTAs at the same offset and are signalled by `tcgen05.commit.cta_group::2 ...
//     .multicast::cluster` (mask 0b11); tempty[a] lives in the leader and counts the 8 epilogue warps of both CTAs
//     (`mbarrier.arrive.shared::cluster` on the leader's address).
//   * TMEM: one warp of EACH CTA issues `tcgen05.alloc.cta_group::2` (same warp index, same destination offset).
// Work item = (pair of M tiles, N tile); the mixed-shape tile plan and the N-split tail of ConvParams apply (decode_tile with the
// pair count in place of the M tile count; M tile = 2 * pair + cluster rank).
// =====================================================================================================================
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerBitMask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_igemm_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                       const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                       const __grid_constant__ CUtensorMap map_y, const __grid_constant__ ShapeMaps smaps, const __grid_constant__ ConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 8);       // 4 epilogue warps of each CTA of the pair (used in the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // the peer's barriers are initialised before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int m_tiles = p.B * (p.plan_n ? p.plan_n : p.tiles_y * p.tiles_x);
  const int m_pairs = (m_tiles + 1) / 2;
  const int total_pairs = m_pairs * p.n_tiles;
  const int virt_pairs = total_pairs + p.tail_tiles * (p.tail_split - 1);      // tail_tiles counts PAIRS here
  const int pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;
  const int taps = p.ksize * p.kw;
  const int num_kb = taps * p.cblocks * p.terms;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair0; t < virt_pairs; t += pair_step) {
        const TileId tid = decode_tile(p, t, m_pairs, total_pairs);
        const int nt = tid.nt, half_n = tid.n_cols >> 1;
        // a pair's second tile may lie past the last M tile: image index B is out of bounds, TMA fills zeros / clips the store
        const Patch pt = decode_patch(p, 2 * tid.mt + (int)rank);
        const Patch pt_peer = decode_patch(p, 2 * tid.mt + 1 - (int)rank);
        const int img = pt.img;
        const int x0 = pt.x0 - p.pad_x, y0 = pt.y0 - p.pad;
        // the leader's barrier counts the bytes of BOTH CTAs: two A boxes (each of its own patch shape) + the two weight halves
        const uint32_t a_own = pt.shape >= 0 ? (uint32_t)p.psw[pt.shape] * p.psh[pt.shape] * p.kc * 2 : (uint32_t)p.a_bytes;
        const uint32_t a_peer = pt_peer.shape >= 0 ? (uint32_t)p.psw[pt_peer.shape] * p.psh[pt_peer.shape] * p.kc * 2 : (uint32_t)p.a_bytes;
        const uint32_t tx_bytes = a_own + a_peer + (uint32_t)tid.n_cols * p.kc * 2;
        const void* ma_plan = pt.shape >= 0 ? (const void*)&smaps.a[pt.shape] : (const void*)&map_a_hi;
        const void* mb_hi = tid.n_cols == p.block_n ? (const void*)&map_b_hi : (const void*)&smaps.b_tail;
        const void* ma_plan_lo = pt.shape >= 0 ? (const void*)&smaps.a_lo[pt.shape] : (const void*)&map_a_lo;
        for (int tap = 0; tap < taps; ++tap) {
          const int dy = tap / p.kw, dx = tap - dy * p.kw;
          for (int cb = 0; cb < p.cblocks; ++cb) {
            for (int term = 0; term < p.terms; ++term) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1);
              const uint32_t sa = smem_base + stage * p.stage_bytes;
              const uint32_t fb = bar_full + 8 * stage;
              if (rank == 0) mbar_expect_tx(fb, tx_bytes);                             // both CTAs' tiles land on the leader's barrier
              tma_load_4d_2sm(sa, term == 1 ? ma_plan_lo : ma_plan, fb, cb * p.kc, x0 + dx, y0 + dy, img);
              tma_load_3d_2sm(sa + p.a_bytes, term == 2 ? (const void*)&map_b_lo : mb_hi, fb, cb * p.kc,
                              nt * p.block_n + tid.n_off + (int)rank * half_n, tap);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && elect_one()) {
      // D = f32, A = B = bf16, K-major, N = block_n, M = 256 (the pair's two 128-row halves)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t row_bytes = p.kc * 2;
      const uint32_t layout = (p.kc == 64) ? 2u : (p.kc == 32 ? 4u : 6u);
      const uint32_t sbo = 8 * row_bytes;
      const int kk = p.kc / 16;
      const uint64_t desc_hi = make_smem_desc(0, sbo, layout);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = pair0; t < virt_pairs; t += pair_step) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        const TileId tid = decode_tile(p, t, m_pairs, total_pairs);
        const uint32_t idesc_t = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(tid.n_cols >> 3) << 17);    // N of this tile (tail: block_n / tail_split)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint64_t adesc = desc_hi | (uint64_t)((sa >> 4) & 0x3FFF);
          const uint64_t bdesc = desc_hi | (uint64_t)(((sa + p.a_bytes) >> 4) & 0x3FFF);
          for (int k = 0; k < kk; ++k)
            tc_mma_bf16_2sm(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_t, (uint32_t)((kb | k) != 0));
          tc_commit_2sm(bar_empty + 8 * stage);      // frees the stage in both CTAs
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(bar_tfull + 8 * acc);          // both CTAs' epilogues may drain their half
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs; TMA-store forms only) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int epi_chunk = 0;
    const uint32_t stage0 = smem_base + p.stages * p.stage_bytes;
    for (int t = pair0; t < virt_pairs; t += pair_step) {
      const TileId tid = decode_tile(p, t, m_pairs, total_pairs);
      const Patch pt = decode_patch(p, 2 * tid.mt + (int)rank);
      const int img = pt.img;
      const void* my = pt.shape >= 0 ? (const void*)&smaps.y[pt.shape] : (const void*)&map_y;
      const int ncols = tid.n_cols, ncol0 = tid.nt * p.block_n + tid.n_off;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      if (p.y_bf16) {
        for (int c0 = 0; c0 < ncols; c0 += 64) {
          const uint32_t buf = stage0 + (uint32_t)(epi_chunk & 1) * (kTileM * 128);
          if (threadIdx.x == 128) bulk_wait_read<1>();
          epi_bar();
          uint32_t v[32], u[32];
          tc_ld32(taddr0 + c0, v);
          tc_ld32(taddr0 + c0 + 32, u);
          tc_ld_wait();
          const int co0 = ncol0 + c0;
          const uint32_t rowaddr = buf + (uint32_t)row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(j < 4 ? v[8 * j + e] : u[8 * (j - 4) + e]);
            if (p.bias && co0 + 8 * j < p.Cout) {
              const float4 b0 = *reinterpret_cast<const float4*>(p.bias + co0 + 8 * j);
              const float4 b1 = *reinterpret_cast<const float4*>(p.bias + co0 + 8 * j + 4);
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w; f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (p.relu) {
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
            }
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
              w[e] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((j ^ (row & 7)) << 4)), "r"(w[0]), "r"(w[1]),
                         "r"(w[2]), "r"(w[3])
                         : "memory");
          }
          fence_proxy_async();
          epi_bar();
          if (threadIdx.x == 128) {
            if (co0 < p.Cout) tma_store_4d(my, buf, co0, pt.x0, pt.y0, img);
            bulk_commit();
          }
          ++epi_chunk;
        }
      } else {
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          const uint32_t buf = stage0 + (uint32_t)(epi_chunk & 1) * (kTileM * 128);
          if (threadIdx.x == 128) bulk_wait_read<1>();
          epi_bar();
          uint32_t v[32];
          tc_ld32(taddr0 + c0, v);
          tc_ld_wait();
          const int co0 = ncol0 + c0;
          const uint32_t rowaddr = buf + (uint32_t)row * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                   __uint_as_float(v[4 * j + 3]));
            if (p.bias && co0 + 4 * j < p.Cout) {
              const float4 b = *reinterpret_cast<const float4*>(p.bias + co0 + 4 * j);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((j ^ (row & 7)) << 4)), "f"(o.x), "f"(o.y),
                         "f"(o.z), "f"(o.w)
                         : "memory");
          }
          fence_proxy_async();
          epi_bar();
          if (threadIdx.x == 128) {
            if (co0 < p.Cout) tma_store_4d(my, buf, co0, pt.x0, pt.y0, img);
            bulk_commit();
          }
          ++epi_chunk;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (threadIdx.x == 128) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                         // neither CTA's shared memory / TMEM goes away while the pair still uses it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

int make_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
             const uint32_t* box, CUtensorMapSwizzle swz, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) {
    jcm_set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return JCM_ENOTSUP;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, dtype, (cuuint32_t)rank, const_cast<void*>(base),
                   (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    jcm_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu, box %u %u %u)", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], box[0],
                  box[1], box[2]);
    return JCM_EINVAL;
  }
  return JCM_OK;
}

}  // namespace

// Chooses the TW x TH (= 128 pixels) patch shape that wastes the fewest accumulator rows on an H x W map.
static void pick_patch(int H, int W, int* TW, int* TH) {
  int best = 1 << 30, bw = 16, bh = 8;
  const int cand[5][2] = {{16, 8}, {8, 16}, {32, 4}, {64, 2}, {128, 1}};
  for (int i = 0; i < 5; ++i) {
    int t = jcm_cdiv(W, cand[i][0]) * jcm_cdiv(H, cand[i][1]);
    if (t < best) { best = t; bw = cand[i][0]; bh = cand[i][1]; }
  }
  *TW = bw;
  *TH = bh;
#ifdef JCM_EXPERIMENTS
  static const int force = JCM_ENV_INT("JCM_CONV_PATCH_TW", 0);   // measurement switch
  if (force > 0 && 128 % force == 0) { *TW = force; *TH = 128 / force; }
#endif
}

extern "C" int jcm_conv2d_fwd(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                              void* y, int y_bf16, int B, int H, int W, int Cin, int Cout, int Cout_pad, int ksize, int kw, int relu,
                              void* stream) {
  if (kw <= 0) kw = ksize;
  JCM_CHECK_ARG(ksize > 0 && (ksize & 1) && (kw & 1), "jcm_conv2d_fwd: kernel extents must be odd (SAME, stride 1), got %d x %d", ksize, kw);
  ConvExArgs a;
  memset(&a, 0, sizeof(a));
  a.x_hi = x_hi; a.x_lo = x_lo; a.w_hi = w_hi; a.w_lo = w_lo; a.bias = bias; a.y = y; a.y_bf16 = y_bf16;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.Cout_pad = Cout_pad; a.ksize = ksize; a.kw = kw; a.relu = relu;
  a.pad_y = -1;
  a.stream = stream;
  return jcm_conv_igemm_ex(a);
}

// jcm_conv2d_fwd with the kernel variant forced (tests and measurements; every variant computes the same values):
// bit 0 = single-CTA kernel instead of the CTA pair, bit 1 = uniform tile grid instead of the mixed-shape plan, bit 2 = no N-split tail,
// bit 3 = halo mode with one M tile per weight stage instead of two, bit 4 = one tile per stage (tap groups) for N <= 64 only.
extern "C" int jcm_conv2d_fwd_variant(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                                      void* y, int y_bf16, int B, int H, int W, int Cin, int Cout, int Cout_pad, int ksize, int kw,
                                      int relu, int variant, void* stream) {
  if (kw <= 0) kw = ksize;
  JCM_CHECK_ARG(ksize > 0 && (ksize & 1) && (kw & 1), "jcm_conv2d_fwd_variant: kernel extents must be odd, got %d x %d", ksize, kw);
  ConvExArgs a;
  memset(&a, 0, sizeof(a));
  a.x_hi = x_hi; a.x_lo = x_lo; a.w_hi = w_hi; a.w_lo = w_lo; a.bias = bias; a.y = y; a.y_bf16 = y_bf16;
  a.B = B; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout; a.Cout_pad = Cout_pad; a.ksize = ksize; a.kw = kw; a.relu = relu;
  a.pad_y = -1;
  a.variant = variant;
  a.stream = stream;
  return jcm_conv_igemm_ex(a);
}

// The launcher behind jcm_conv2d_fwd, also called by the tensor-core spatial model (spatial_model.cu) with its grouped forms.
int jcm_conv_igemm_ex(const ConvExArgs& a) {
  const void *x_hi = a.x_hi, *x_lo = a.x_lo, *w_hi = a.w_hi, *w_lo = a.w_lo;
  const float* bias = a.bias;
  void* y = a.y;
  void* stream = a.stream;
  const int y_bf16 = a.y_bf16, B = a.B, H = a.H, W = a.W, Cin = a.Cin, Cout = a.Cout, Cout_pad = a.Cout_pad, ksize = a.ksize, kw = a.kw,
            relu = a.relu;
  JCM_CHECK_ARG(x_hi && w_hi && y, "jcm_conv2d_fwd: null pointer");
  JCM_CHECK_ARG((x_lo == nullptr) == (w_lo == nullptr), "jcm_conv2d_fwd: x_lo and w_lo must both be given (bf16x3) or both NULL (bf16)");
  JCM_CHECK_ARG(B > 0 && H > 0 && W > 0 && ksize > 0 && kw > 0, "jcm_conv2d_fwd: bad shape B=%d H=%d W=%d k=%dx%d", B, H, W, ksize, kw);
  JCM_CHECK_ARG(Cin >= 16 && (Cin % 16) == 0, "jcm_conv2d_fwd: Cin must be a multiple of 16, got %d", Cin);
  JCM_CHECK_ARG(Cout > 0 && Cout_pad >= Cout && (Cout_pad % 16) == 0, "jcm_conv2d_fwd: Cout_pad=%d must be >= Cout=%d and a multiple of 16", Cout_pad, Cout);
  JCM_CHECK_ARG((((uintptr_t)x_hi | (uintptr_t)w_hi | (uintptr_t)x_lo | (uintptr_t)w_lo | (uintptr_t)y) & 15) == 0,
                "jcm_conv2d_fwd: pointers must be 16-byte aligned");
  JCM_CHECK_ARG(a.grp == 0 || (x_lo == nullptr && kw == 1), "jcm_conv_igemm_ex: grouped forms are single-term with kw == 1");

  ConvParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W; p.Cout = Cout;
  pick_patch(H, W, &p.TW, &p.TH);
  p.tiles_x = jcm_cdiv(W, p.TW);
  p.tiles_y = jcm_cdiv(H, p.TH);
  p.block_n = Cout_pad <= 256 ? Cout_pad : 256;
  JCM_CHECK_ARG(Cout_pad % p.block_n == 0, "jcm_conv2d_fwd: Cout_pad=%d must be <= 256 or a multiple of 256", Cout_pad);
  p.n_tiles = Cout_pad / p.block_n;
  p.kc = (Cin % 64) == 0 ? 64 : ((Cin % 32) == 0 ? 32 : 16);
  p.cblocks = Cin / p.kc;
  p.ksize = ksize;
  p.kw = kw;
  p.pad = a.pad_y >= 0 ? a.pad_y : (ksize - 1) / 2;
  p.pad_x = (kw - 1) / 2;
  p.grp = a.grp; p.grp_taps = ksize * kw; p.a_div = a.a_div > 0 ? a.a_div : 1; p.k_rows = a.k_rows;
  p.sm_pad = a.sm_pad; p.sm_rows = a.sm_rows; p.sm_rows_in = a.sm_rows_in;
  p.a_map = (a.grp == 1 || (a.grp == 2 && a.map_on_a)) ? a.img_map : nullptr;
  p.b_map = (a.grp == 2 && !a.map_on_a) ? a.img_map : nullptr;
  JCM_CHECK_ARG(!a.img_map || a.map_images > 0, "jcm_conv_igemm_ex: img_map needs map_images");
  p.terms = x_lo ? 3 : 1;
  p.a_bytes = kTileM * p.kc * 2;
  p.b_bytes = p.block_n * p.kc * 2;
  p.stage_bytes = ((p.a_bytes + p.b_bytes + 1023) / 1024) * 1024;
  // (grouped forms only: a block_n that is a multiple of 16 but not of 32 - the 144-column tiles of the tensor-core spatial model at W = 128 - ends in a
  // half chunk: the TMEM read runs into unused accumulator columns and the TMA store clips at Cout)
  p.tma_store = (Cout % 4) == 0 && Cout >= 32 && ((p.block_n % 32) == 0 || (a.grp != 0 && (p.block_n % 16) == 0));
  const int epi_bytes = p.tma_store ? 2 * kTileM * 128 : 0;     // two staging buffers of 128 rows x 32 fp32
  p.stages = (225 * 1024 - epi_bytes) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  // halo mode where the A operand's L2 traffic is the limiter (few output channels per A tile) and the taps have vertical extent
  p.halo = 0;
  p.mgroup = 1;
  static const int splitn_env = JCM_ENV_INT("JCM_MMA_SPLITN", 0);
  p.mma_split_n = splitn_env && (p.block_n % 32) == 0;
  {
    static const int nacc_env = JCM_ENV_INT("JCM_CONV_NACC", 0);   // 2 enables (measurement)
    const int mmas_per_tile = ksize * kw * p.cblocks * p.terms * (p.kc / 16);
    // measured (profiles/r01/conv_sweep_r01.txt): once the MMA warp issues through elect.sync the round-robin accumulators no longer
    // help (the serialisation seen before was the issuing thread, not the tensor pipe), so they stay off unless JCM_CONV_NACC=2
    p.nacc = 1;
    if (nacc_env >= 2) {
      if (p.block_n <= 128 && mmas_per_tile >= 2) p.nacc = 2;
      if (p.block_n <= 64 && mmas_per_tile >= 4) p.nacc = 4;
    }
    if (p.mma_split_n) p.nacc = 1;
  }
  static const int dbg_env = JCM_ENV_INT("JCM_CONV_DBG", 0);
  p.dbg = dbg_env;
  p.a_stages = 0; p.b_stages = 0; p.a_stride = 0; p.b_stride = 0; p.a_halo_bytes = 0;
  static const int halo_env = JCM_ENV_INT("JCM_CONV_HALO", 1);   // 0 disables (for A/B measurements)
  if (halo_env && p.grp == 0 && p.terms == 1 && ksize >= 3 && p.block_n <= (halo_env > 1 ? 256 : 128) && (p.TW % 8) == 0) {
    p.a_halo_bytes = (p.TH + ksize - 1) * p.TW * p.kc * 2;
    static const int group_env = JCM_ENV_INT("JCM_CONV_BGROUP", 0);   // measurement: force taps per stage
    p.a_stride = ((p.a_halo_bytes + 1023) / 1024) * 1024;
    p.b_slot = ((p.b_bytes + 1023) / 1024) * 1024;
    p.a_stages = p.a_stride > 40 * 1024 ? 2 : 3;
    const int budget = 225 * 1024 - epi_bytes - p.a_stages * p.a_stride;
    p.b_group = budget / (3 * p.b_slot);                 // as many vertical taps per stage as still leave 3 weight stages
    if (p.b_group > ksize) p.b_group = ksize;
    if (p.block_n > 64) p.b_group = 1;                   // measured: grouping pays for N <= 64 (+20 %), not for N = 128
    if (group_env > 0 && group_env < p.b_group) p.b_group = group_env;
    if (p.b_group < 1) p.b_group = 1;
    p.b_stride = p.b_group * p.b_slot;
    p.b_stages = budget / p.b_stride;
    if (p.b_stages > kMaxStages) p.b_stages = kMaxStages;
    p.halo = p.b_stages >= 3;
    // Two M tiles per weight stage: the weights then stream from L2 once per PAIR of tiles (for the 5x5 layers they are 3/4 of a
    // tile's L2->SM traffic).  Needs four halo patches (two tiles, double-buffered) next to three weight stages, and both tiles'
    // accumulators in one 256-column TMEM stage.  Measured (tests/gpu_diag.py halo2, profiles/r02/diag_halo2.txt, isolated, batch 64):
    // 0.576 -> 0.440 ms on the 64->128 5x5 layer at 120x180, -23 % on the other N = 128 layers; for N <= 64 (one tap per weight
    // stage instead of a tap group) 0.843 -> 0.717 ms on 128->64 at 120x180, 0.430 -> 0.382 ms on conv1_fullres.  Variant bit 3 turns
    // it off; bit 4 keeps the tap groups of the N <= 64 layers instead (one tile per stage there).
    if (p.halo && !(a.variant & 8) && p.block_n <= 128 && (p.b_group == 1 || !(a.variant & 16))) {
      const int b_stride1 = p.b_slot;
      const int budget2 = 225 * 1024 - epi_bytes - 4 * p.a_stride;
      if (budget2 >= 3 * b_stride1) {
        p.mgroup = 2;
        p.a_stages = 4;
        p.b_group = 1;
        p.b_stride = b_stride1;
        p.b_stages = budget2 / p.b_stride;
        if (p.b_stages > kMaxStages) p.b_stages = kMaxStages;
      }
    }
    if (p.halo) p.nacc = 1;
  }
  // CTA-pair form (cta_group::2) of the plain mode for N = 256 tiles: the default wherever it applies.  a.variant (tests /
  // measurements, jcm_conv2d_fwd_variant): bit 0 = single-CTA kernel, bit 1 = uniform tile grid, bit 2 = no N-split tail.
  const bool pair = !(a.variant & 1) && !p.halo && p.grp == 0 && p.block_n == 256 && p.tma_store && p.nacc == 1 && !p.mma_split_n &&
                    !p.dbg && jcm_num_sms() >= 2;
  if (pair) {
    p.b_bytes = (p.block_n / 2) * p.kc * 2;            // each CTA of the pair holds half of the weight tile
    p.stage_bytes = ((p.a_bytes + p.b_bytes + 1023) / 1024) * 1024;
    p.stages = (225 * 1024 - epi_bytes) / p.stage_bytes;
    if (p.stages > kMaxStages) p.stages = kMaxStages;
  }
  // mixed-shape tile plan (tiling.cuh): plain mode with the TMA-store epilogue, when it needs fewer tiles than the uniform grid
  TilePlan plan;
  plan.n_tiles = 0;
  p.plan_n = 0;
  if (!(a.variant & 2) && !p.halo && p.grp == 0 && p.tma_store && p.nacc == 1 && !p.mma_split_n &&
      plan_tiles(H, W, kTileM, true, kIgemmShapes, p.tiles_x * p.tiles_y, &plan)) {
    p.plan_n = plan.n_tiles;
    memcpy(p.px0, plan.x0, sizeof(p.px0));
    memcpy(p.py0, plan.y0, sizeof(p.py0));
    memcpy(p.pshape, plan.shape, sizeof(p.pshape));
    for (int i = 0; i < kIgemmShapes; ++i) { p.psw[i] = i < plan.n_shapes ? plan.sw[i] : 0; p.psh[i] = i < plan.n_shapes ? plan.sh[i] : 0; }
  }
  // N-split tail (see ConvParams): work units of the persistent loop = (M tiles or M-tile pairs) x N tiles
  p.tail_tiles = 0;
  p.tail_split = 1;
  const int m_tiles_h = B * (p.plan_n ? p.plan_n : p.tiles_x * p.tiles_y);
  const int units = (pair ? (m_tiles_h + 1) / 2 : m_tiles_h) * p.n_tiles;
  const int workers = pair ? (jcm_num_sms() & ~1) / 2 : jcm_num_sms();
  if (!(a.variant & 4) && p.terms == 1 && !p.halo && p.grp == 0 && p.tma_store && p.block_n == 256 && p.nacc == 1 && !p.mma_split_n && units > workers &&
      units % workers != 0) {
    const int rem = units % workers;
    // Measured (tests/gpu_diag.py cta2, profiles/r02/diag_variants2.txt): a quarter tile (N = 64) costs far more than a quarter of a
    // full tile - the per-k-block handshake and the A-operand reads do not shrink with N - so splitting pays only as ONE sub-round
    // of half tiles (N = 128): the partial wave then costs ~0.6 of a round instead of 1.  Anything else is left unsplit.
    if (2 * rem <= workers) p.tail_split = 2;
    if (p.tail_split > 1) p.tail_tiles = rem;
  }
  p.relu = relu;
  p.bias = bias;
  p.y = (float*)y;
  p.y_bf16 = y_bf16;
  JCM_CHECK_ARG(!y_bf16 || (p.tma_store && (Cout % 8) == 0 && (p.block_n % 64) == 0),
                "jcm_conv2d_fwd: bf16 output needs Cout a multiple of 8 and N tiles that are multiples of 64 (Cout=%d)", Cout);

  const CUtensorMapSwizzle swz = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  static ShapeMaps smaps_zero;            // zero-initialised template (unused slots must still be valid kernel-parameter bytes)
  ShapeMaps smaps = smaps_zero;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)(p.a_map ? a.map_images : (a.grp == 2 ? B / p.a_div : B))};
    uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)p.kc, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    int rc = make_map(&ma_hi, x_hi, 4, dims, str, box, swz);
    if (rc) return rc;
    for (int i = 0; i < (p.plan_n ? plan.n_shapes : 0); ++i) {
      uint32_t pbox[4] = {(uint32_t)p.kc, plan.sw[i], plan.sh[i], 1};
      rc = make_map(&smaps.a[i], x_hi, 4, dims, str, pbox, swz);
      if (rc) return rc;
      if (x_lo) {
        rc = make_map(&smaps.a_lo[i], x_lo, 4, dims, str, pbox, swz);
        if (rc) return rc;
      }
    }
    if (p.halo) box[2] = (uint32_t)(p.TH + ksize - 1);   // halo mode: the (unused) lo slot carries the halo-box map of x_hi
    rc = make_map(&ma_lo, x_lo ? x_lo : x_hi, 4, dims, str, box, swz);
    if (rc) return rc;
  }
  {
    // weight planes: one per tap; grouped forms: B * taps planes (grp 1), B / a_div planes with their own K extent w_cin (grp 2)
    const uint64_t wc = a.grp == 2 ? (uint64_t)a.w_cin : (uint64_t)Cin;
    const uint64_t planes = a.grp == 1 ? (uint64_t)B * ksize * kw
                                        : (a.grp == 2 ? (uint64_t)(p.b_map ? a.map_images : B / p.a_div) : (uint64_t)(ksize * kw));
    uint64_t dims[3] = {wc, (uint64_t)Cout_pad, planes};
    uint64_t str[2] = {wc * 2, (uint64_t)Cout_pad * wc * 2};
    uint32_t box[3] = {(uint32_t)p.kc, (uint32_t)(pair ? p.block_n / 2 : p.block_n), 1};
    int rc = make_map(&mb_hi, w_hi, 3, dims, str, box, swz);
    if (rc) return rc;
    rc = make_map(&mb_lo, w_lo ? w_lo : w_hi, 3, dims, str, box, swz);
    if (rc) return rc;
    if (p.tail_split > 1) {
      box[1] = (uint32_t)((p.block_n / p.tail_split) / (pair ? 2 : 1));
      rc = make_map(&smaps.b_tail, w_hi, 3, dims, str, box, swz);
      if (rc) return rc;
    }
  }

  CUtensorMap my;
  memset(&my, 0, sizeof(my));
  if (p.tma_store) {
    const uint64_t es = y_bf16 ? 2 : 4;
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)Cout * es, (uint64_t)W * Cout * es, (uint64_t)H * W * Cout * es};
    uint32_t box[4] = {y_bf16 ? 64u : 32u, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    int rc = make_map(&my, y, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B,
                      y_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
    if (rc) return rc;
    for (int i = 0; i < (p.plan_n ? plan.n_shapes : 0); ++i) {
      uint32_t pbox[4] = {y_bf16 ? 64u : 32u, plan.sw[i], plan.sh[i], 1};
      rc = make_map(&smaps.y[i], y, 4, dims, str, pbox, CU_TENSOR_MAP_SWIZZLE_128B,
                    y_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
      if (rc) return rc;
    }
  }

  static bool attr_set[2][64] = {{false}, {false}};     // function attributes are per device: one flag per kernel and device ordinal
  int dev = 0;
  JCM_CUDA(cudaGetDevice(&dev));
  if (pair) {
    int grid2 = 2 * workers;
    if (grid2 > 2 * units) grid2 = 2 * units;
    const size_t smem2 = (size_t)p.stages * p.stage_bytes + epi_bytes + 1024;
    if (dev < 0 || dev >= 64 || !attr_set[1][dev]) {
      JCM_CUDA(cudaFuncSetAttribute(conv_igemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
      if (dev >= 0 && dev < 64) attr_set[1][dev] = true;
    }
    conv_igemm_pair_kernel<<<grid2, kThreads, smem2, (cudaStream_t)stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, my, smaps, p);
    JCM_LAUNCH_CHECK();
    return JCM_OK;
  }
  const int total_tiles = p.halo ? jcm_cdiv(m_tiles_h, p.mgroup) * p.n_tiles : units;
  int grid = jcm_num_sms();
  if (grid > total_tiles) grid = total_tiles;
  const size_t smem = (p.halo ? (size_t)p.a_stages * p.a_stride + (size_t)p.b_stages * p.b_stride : (size_t)p.stages * p.stage_bytes) + epi_bytes + 1024;
  if (dev < 0 || dev >= 64 || !attr_set[0][dev]) {
    JCM_CUDA(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
    if (dev >= 0 && dev < 64) attr_set[0][dev] = true;
  }
  conv_igemm_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, my, smaps, p);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

// =====================================================================================================================
// Weight gradient:  dW[tap, ci, co] = sum_{n,y,x} X[n, y+dy-pad, x+dx-pad, ci] * G[n, y, x, co]      (G = d loss / d conv output)
//
// Per tap this is a GEMM whose contraction runs over PIXELS, so both operands are "MN-major" for the tensor core: a TMA box
// {64 channels, TW, TH, 1} lands in shared memory as [64 pixels][128 B of channels] (SWIZZLE_128B) and is consumed directly
// as a UMMA operand with a_major = b_major = MN (16 pixels = one K=16 MMA step, channel blocks of 64 are LBO apart).
// The tensor with more channels provides M (tile 128, zero-filled by TMA out-of-bounds when it has fewer), the other one N.
// Work item = (k-split, tap, M tile, N tile): loops over its share of the 64-pixel patches; partial sums go to
// partial[split][tap][M][N] and are reduced deterministically by wgrad_reduce_kernel (no atomics).
// =====================================================================================================================
namespace {

constexpr int kWgPix = 64;   // pixels per k-block

struct WgradParams {
  int B, H, W;
  int TW, TH, tiles_x, tiles_y;
  int ksize, kw, pad, pad_x, taps;
  int m_tiles, n_tiles, block_n, m_pad, n_pad;
  int shift_m, shift_n;            // 1 when that operand is the (shifted) layer input X, 0 when it is the gradient G
  int kc_n, kc_m;                  // channels per N-side / M-side box: 64, 32 or 16
  int splits, total_patches;
  int terms, stages, a_bytes, b_bytes, stage_bytes;
  int tgroup;          // conv_wgrad_taps_kernel: taps per task (1 elsewhere)
  // mixed-shape patch plan (tiling.cuh, exact 64-pixel boxes; one term): plan_n > 0 patches per image replace the uniform grid
  int plan_n;
  uint8_t px0[kPlanMaxTiles], py0[kPlanMaxTiles], pshape[kPlanMaxTiles];
  float* partial;
};
constexpr int kWgradShapes = 4;
struct WgradShapeMaps {
  CUtensorMap m[kWgradShapes];     // M-side operand boxes {kc_m, sw, sh, 1, channel blocks} per plan shape
  CUtensorMap n[kWgradShapes];     // N-side operand boxes
};

__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_m_hi, const __grid_constant__ CUtensorMap map_m_lo,
                  const __grid_constant__ CUtensorMap map_n_hi, const __grid_constant__ CUtensorMap map_n_lo,
                  const __grid_constant__ WgradShapeMaps smaps, const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int tasks_per_split = p.taps * p.m_tiles * p.n_tiles;
  const int total_tasks = p.splits * tasks_per_split;
  const int patches_per_img = p.plan_n ? p.plan_n : p.tiles_x * p.tiles_y;
  const int m_boxes = kTileM / p.kc_m;
  const int n_boxes = p.block_n / p.kc_n;
  const int m_box_bytes = kWgPix * p.kc_m * 2;
  const int n_box_bytes = kWgPix * p.kc_n * 2;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int task = blockIdx.x; task < total_tasks; task += gridDim.x) {
        const int split = task / tasks_per_split;
        int r = task - split * tasks_per_split;
        const int tap = r / (p.m_tiles * p.n_tiles);
        r -= tap * (p.m_tiles * p.n_tiles);
        const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
        const int dy = tap / p.kw - p.pad, dx = tap % p.kw - p.pad_x;
        const int p0 = (int)((long)split * p.total_patches / p.splits);
        const int p1 = (int)((long)(split + 1) * p.total_patches / p.splits);
        int img = p0 / patches_per_img;
        int q = p0 - img * patches_per_img;
        int ty = p.plan_n ? 0 : q / p.tiles_x, tx = p.plan_n ? 0 : q - ty * p.tiles_x;
        const int mblk0 = mt * m_boxes, nblk0 = nt * n_boxes;   // first channel block (of kc_m / kc_n channels) of this tile
        const int sdx_m = p.shift_m * dx, sdy_m = p.shift_m * dy, sdx_n = p.shift_n * dx, sdy_n = p.shift_n * dy;
        for (int patch = p0; patch < p1; ++patch) {
          int x0 = tx * p.TW, y0 = ty * p.TH;
          const void *pm = (const void*)&map_m_hi, *pn = (const void*)&map_n_hi;
          if (p.plan_n) {
            x0 = p.px0[q];
            y0 = p.py0[q];
            pm = (const void*)&smaps.m[p.pshape[q]];
            pn = (const void*)&smaps.n[p.pshape[q]];
          }
          for (int term = 0; term < p.terms; ++term) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t sa = smem_base + stage * p.stage_bytes;
            const uint32_t sb = sa + p.a_bytes;
            const uint32_t fb = bar_full + 8 * stage;
            mbar_expect_tx(fb, p.a_bytes + p.b_bytes);
            const void* mm = (term == 1) ? (const void*)&map_m_lo : pm;
            const void* mn = (term == 2) ? (const void*)&map_n_lo : pn;
            // one 5-D box per operand: {kc channels, TW, TH, 1 image, all channel blocks of the tile}
            tma_load_5d(sa, mm, fb, 0, x0 + sdx_m, y0 + sdy_m, img, mblk0);
            tma_load_5d(sb, mn, fb, 0, x0 + sdx_n, y0 + sdy_n, img, nblk0);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if (p.plan_n) {
            if (++q == p.plan_n) { q = 0; ++img; }
          } else if (++tx == p.tiles_x) {
            tx = 0;
            if (++ty == p.tiles_y) { ty = 0; ++img; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      // D=f32, A=B=bf16, both MN-major (bits 15, 16), N = block_n, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t n_row_bytes = p.kc_n * 2, m_row_bytes = p.kc_m * 2;
      const uint32_t n_layout = (p.kc_n == 64) ? 2u : (p.kc_n == 32 ? 4u : 6u);
      const uint32_t m_layout = (p.kc_m == 64) ? 2u : (p.kc_m == 32 ? 4u : 6u);
      const uint64_t adesc_hi = (make_smem_desc(0, 8 * m_row_bytes, m_layout) & ~(((uint64_t)0x3FFF << 16) | 0x3FFF)) |
                                ((uint64_t)((m_box_bytes >> 4) & 0x3FFF) << 16);
      const uint64_t bdesc_hi = (make_smem_desc(0, 8 * n_row_bytes, n_layout) & ~(((uint64_t)0x3FFF << 16) | 0x3FFF)) |
                                ((uint64_t)((n_box_bytes >> 4) & 0x3FFF) << 16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int task = blockIdx.x; task < total_tasks; task += gridDim.x) {
        const int split = task / tasks_per_split;
        const int p0 = (int)((long)split * p.total_patches / p.splits);
        const int p1 = (int)((long)(split + 1) * p.total_patches / p.splits);
        const int num_kb = (p1 - p0) * p.terms;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t sb = sa + p.a_bytes;
          // MN-major descriptors: leading byte offset = distance between channel blocks (one box), stride byte offset = 8 pixel rows;
          // only the start address changes between stages / K steps, so the upper words are built once (adesc_hi / bdesc_hi)
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k) {
            const uint64_t adesc = adesc_hi | (uint64_t)(((sa + k * 16 * m_row_bytes) >> 4) & 0x3FFF);
            const uint64_t bdesc = bdesc_hi | (uint64_t)(((sb + k * 16 * n_row_bytes) >> 4) & 0x3FFF);
            tc_mma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          tc_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: accumulator -> partial[split][tap][m][n] =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int task = blockIdx.x; task < total_tasks; task += gridDim.x) {
      const int split = task / tasks_per_split;
      int r = task - split * tasks_per_split;
      const int tap = r / (p.m_tiles * p.n_tiles);
      r -= tap * (p.m_tiles * p.n_tiles);
      const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
      float* orow = p.partial + (((long)split * p.taps + tap) * p.m_pad + mt * kTileM + row) * p.n_pad + nt * p.block_n;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t v[32];
        tc_ld32(taddr0 + c0, v);
        tc_ld_wait();
        if (c0 + 32 <= p.block_n) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(orow + c0 + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < p.block_n) orow[c0 + j] = __uint_as_float(v[j]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// Tap-group form of the weight-gradient kernel for the narrow layers (N tile <= 64 channels: conv1_*, conv2_*; one term).  Per
// 64-pixel k-block the plain kernel loads a gradient box and an input box (24 KB for conv2) for 128 clocks of MMA - far more than
// L2 delivers - and every tap task re-reads the SAME gradient box.  Here one task owns `tgroup` consecutive taps: the unshifted
// operand (the gradient G) is loaded once per k-block, the shifted operand (the layer input X) once per tap, and the taps accumulate
// side by side in TMEM (tgroup x block_n <= 256 columns per accumulator stage).
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_taps_kernel(const __grid_constant__ CUtensorMap map_m_hi, const __grid_constant__ CUtensorMap map_n_hi,
                       const __grid_constant__ WgradShapeMaps smaps, const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int tap_groups = (p.taps + p.tgroup - 1) / p.tgroup;
  const int tasks_per_split = tap_groups * p.m_tiles * p.n_tiles;
  const int total_tasks = p.splits * tasks_per_split;
  const int patches_per_img = p.plan_n ? p.plan_n : p.tiles_x * p.tiles_y;
  const int m_boxes = kTileM / p.kc_m;
  const int n_boxes = p.block_n / p.kc_n;
  const int m_box_bytes = kWgPix * p.kc_m * 2;
  const int n_box_bytes = kWgPix * p.kc_n * 2;
  // stage layout: [M-side boxes][N-side boxes]; the shifted side (the layer input) has one box per tap of the group
  const uint32_t n_off = (uint32_t)(p.shift_m ? p.tgroup : 1) * p.a_bytes;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int task = blockIdx.x; task < total_tasks; task += gridDim.x) {
        const int split = task / tasks_per_split;
        int r = task - split * tasks_per_split;
        const int tg = r / (p.m_tiles * p.n_tiles);
        r -= tg * (p.m_tiles * p.n_tiles);
        const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
        const int tap0 = tg * p.tgroup, ntap = min(p.tgroup, p.taps - tap0);
        const int p0 = (int)((long)split * p.total_patches / p.splits);
        const int p1 = (int)((long)(split + 1) * p.total_patches / p.splits);
        int img = p0 / patches_per_img;
        int q = p0 - img * patches_per_img;
        int ty = p.plan_n ? 0 : q / p.tiles_x, tx = p.plan_n ? 0 : q - ty * p.tiles_x;
        const int mblk0 = mt * m_boxes, nblk0 = nt * n_boxes;
        const uint32_t tx_bytes = (uint32_t)(p.shift_m ? ntap : 1) * p.a_bytes + (uint32_t)(p.shift_n ? ntap : 1) * p.b_bytes;
        for (int patch = p0; patch < p1; ++patch) {
          int x0 = tx * p.TW, y0 = ty * p.TH;
          const void *pm = (const void*)&map_m_hi, *pn = (const void*)&map_n_hi;
          if (p.plan_n) {
            x0 = p.px0[q];
            y0 = p.py0[q];
            pm = (const void*)&smaps.m[p.pshape[q]];
            pn = (const void*)&smaps.n[p.pshape[q]];
          }
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t fb = bar_full + 8 * stage;
          mbar_expect_tx(fb, tx_bytes);
          if (!p.shift_m) tma_load_5d(sa, pm, fb, 0, x0, y0, img, mblk0);
          if (!p.shift_n) tma_load_5d(sa + n_off, pn, fb, 0, x0, y0, img, nblk0);
          for (int j = 0; j < ntap; ++j) {
            const int tap = tap0 + j;
            const int dy = tap / p.kw - p.pad, dx = tap % p.kw - p.pad_x;
            if (p.shift_m) tma_load_5d(sa + j * p.a_bytes, pm, fb, 0, x0 + dx, y0 + dy, img, mblk0);
            else tma_load_5d(sa + n_off + j * p.b_bytes, pn, fb, 0, x0 + dx, y0 + dy, img, nblk0);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (p.plan_n) {
            if (++q == p.plan_n) { q = 0; ++img; }
          } else if (++tx == p.tiles_x) {
            tx = 0;
            if (++ty == p.tiles_y) { ty = 0; ++img; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)(kTileM >> 4) << 24);
      const uint32_t n_row_bytes = p.kc_n * 2, m_row_bytes = p.kc_m * 2;
      const uint32_t n_layout = (p.kc_n == 64) ? 2u : (p.kc_n == 32 ? 4u : 6u);
      const uint32_t m_layout = (p.kc_m == 64) ? 2u : (p.kc_m == 32 ? 4u : 6u);
      const uint64_t adesc_hi = (make_smem_desc(0, 8 * m_row_bytes, m_layout) & ~(((uint64_t)0x3FFF << 16) | 0x3FFF)) |
                                ((uint64_t)((m_box_bytes >> 4) & 0x3FFF) << 16);
      const uint64_t bdesc_hi = (make_smem_desc(0, 8 * n_row_bytes, n_layout) & ~(((uint64_t)0x3FFF << 16) | 0x3FFF)) |
                                ((uint64_t)((n_box_bytes >> 4) & 0x3FFF) << 16);
      const uint32_t a_step = p.shift_m ? (uint32_t)p.a_bytes : 0u, b_step = p.shift_n ? (uint32_t)p.b_bytes : 0u;
      // The MMAs of the narrow layers are short (N <= 64: 32-48 clocks each) and ONE thread issues 16 of them per k-block, so the
      // address arithmetic of that thread is on the critical path.  Everything that does not depend on the stage is tabulated here:
      // the descriptors' high words are constant, their low words are (stage base >> 4) + a per-(tap, K step) offset - one 32-bit add
      // per operand and MMA (the 14-bit address field cannot carry: shared-memory addresses stay below 256 KB).
      const uint32_t a_hi = (uint32_t)(adesc_hi >> 32), b_hi = (uint32_t)(bdesc_hi >> 32);
      const uint32_t a_lo0 = (uint32_t)adesc_hi, b_lo0 = (uint32_t)bdesc_hi;       // low words without the start address
      uint32_t aoff[4][kWgPix / 16], boff[4][kWgPix / 16];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < kWgPix / 16; ++k) {
          aoff[j][k] = (j * a_step + k * 16 * m_row_bytes) >> 4;
          boff[j][k] = (j * b_step + k * 16 * n_row_bytes) >> 4;
        }
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int task = blockIdx.x; task < total_tasks; task += gridDim.x) {
        const int split = task / tasks_per_split;
        const int tg = (task - split * tasks_per_split) / (p.m_tiles * p.n_tiles);
        const int ntap = min(p.tgroup, p.taps - tg * p.tgroup);
        const int p0 = (int)((long)split * p.total_patches / p.splits);
        const int p1 = (int)((long)(split + 1) * p.total_patches / p.splits);
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        uint32_t accf = 0;
        for (int kb = 0; kb < p1 - p0; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t a_lo = a_lo0 | (sa >> 4), b_lo = b_lo0 | ((sa + n_off) >> 4);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < ntap) {
#pragma unroll
              for (int k = 0; k < kWgPix / 16; ++k) {
                uint64_t adesc, bdesc;
                asm("mov.b64 %0, {%1, %2};" : "=l"(adesc) : "r"(a_lo + aoff[j][k]), "r"(a_hi));
                asm("mov.b64 %0, {%1, %2};" : "=l"(bdesc) : "r"(b_lo + boff[j][k]), "r"(b_hi));
                tc_mma_bf16(d_tmem + j * p.block_n, adesc, bdesc, idesc, k ? 1u : accf);
              }
            }
          }
          accf = 1u;
          tc_commit(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int task = blockIdx.x; task < total_tasks; task += gridDim.x) {
      const int split = task / tasks_per_split;
      int r = task - split * tasks_per_split;
      const int tg = r / (p.m_tiles * p.n_tiles);
      r -= tg * (p.m_tiles * p.n_tiles);
      const int mt = r / p.n_tiles, nt = r - mt * p.n_tiles;
      const int tap0 = tg * p.tgroup, ntap = min(p.tgroup, p.taps - tap0);
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      for (int j = 0; j < ntap; ++j) {
        float* orow = p.partial + (((long)split * p.taps + tap0 + j) * p.m_pad + mt * kTileM + row) * p.n_pad + nt * p.block_n;
        const uint32_t taddr0 = tmem_base + acc * 256 + j * p.block_n + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          uint32_t v[32];
          tc_ld32(taddr0 + c0, v);
          tc_ld_wait();
          if (c0 + 32 <= p.block_n) {
#pragma unroll
            for (int jj = 0; jj < 32; jj += 4)
              *reinterpret_cast<float4*>(orow + c0 + jj) =
                  make_float4(__uint_as_float(v[jj]), __uint_as_float(v[jj + 1]), __uint_as_float(v[jj + 2]), __uint_as_float(v[jj + 3]));
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj)
              if (c0 + jj < p.block_n) orow[c0 + jj] = __uint_as_float(v[jj]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// CTA-pair form (cta_group::2) of the weight-gradient kernel, for layers with an even number of 128-channel M tiles and N tiles of at
// least 128 channels (conv3 .. conv5: 98 % of the weight-gradient FLOPs).  The two CTAs of a cluster own two adjacent M tiles of the
// same (k-split, tap, N tile); each loads its own M-side box and HALF of the N-side box (block_n / 2 channels), the leader issues one
// M = 256 MMA per 16 pixels that reads both CTAs' shared memory and accumulates into each CTA's own TMEM.  Per CTA and 64-pixel
// k-block the operand traffic drops from 48 KB to 32 KB; barrier protocol as in conv_igemm_pair_kernel.
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_wgrad_pair_kernel(const __grid_constant__ CUtensorMap map_m_hi, const __grid_constant__ CUtensorMap map_m_lo,
                       const __grid_constant__ CUtensorMap map_n_hi, const __grid_constant__ CUtensorMap map_n_lo,
                       const __grid_constant__ WgradShapeMaps smaps, const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages + 4];
  __shared__ uint32_t tmem_base_slot;

  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar_full = smem_u32(&bars[0]);
  const uint32_t bar_empty = smem_u32(&bars[kMaxStages]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bar_tempty = smem_u32(&bars[2 * kMaxStages + 2]);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_tfull + 8 * s, 1);
      mbar_init(bar_tempty + 8 * s, 8);       // 4 epilogue warps of each CTA of the pair (used in the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  const int m_pairs = p.m_tiles >> 1;
  const int tasks_per_split = p.taps * m_pairs * p.n_tiles;
  const int total_tasks = p.splits * tasks_per_split;
  const int task0 = blockIdx.x >> 1, task_step = gridDim.x >> 1;
  const int patches_per_img = p.plan_n ? p.plan_n : p.tiles_x * p.tiles_y;
  const int m_boxes = kTileM / p.kc_m;
  const int n_boxes_half = (p.block_n / p.kc_n) >> 1;          // N-side channel blocks each CTA loads
  const int m_box_bytes = kWgPix * p.kc_m * 2;
  const int n_box_bytes = kWgPix * p.kc_n * 2;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int task = task0; task < total_tasks; task += task_step) {
        const int split = task / tasks_per_split;
        int r = task - split * tasks_per_split;
        const int tap = r / (m_pairs * p.n_tiles);
        r -= tap * (m_pairs * p.n_tiles);
        const int mp = r / p.n_tiles, nt = r - mp * p.n_tiles;
        const int dy = tap / p.kw - p.pad, dx = tap % p.kw - p.pad_x;
        const int p0 = (int)((long)split * p.total_patches / p.splits);
        const int p1 = (int)((long)(split + 1) * p.total_patches / p.splits);
        int img = p0 / patches_per_img;
        int q = p0 - img * patches_per_img;
        int ty = p.plan_n ? 0 : q / p.tiles_x, tx = p.plan_n ? 0 : q - ty * p.tiles_x;
        const int mblk0 = (2 * mp + (int)rank) * m_boxes;                                      // this CTA's M tile
        const int nblk0 = nt * (p.block_n / p.kc_n) + (int)rank * n_boxes_half;               // this CTA's half of the N tile
        const int sdx_m = p.shift_m * dx, sdy_m = p.shift_m * dy, sdx_n = p.shift_n * dx, sdy_n = p.shift_n * dy;
        for (int patch = p0; patch < p1; ++patch) {
          int x0 = tx * p.TW, y0 = ty * p.TH;
          const void *pm = (const void*)&map_m_hi, *pn = (const void*)&map_n_hi;
          if (p.plan_n) {
            x0 = p.px0[q];
            y0 = p.py0[q];
            pm = (const void*)&smaps.m[p.pshape[q]];
            pn = (const void*)&smaps.n[p.pshape[q]];
          }
          for (int term = 0; term < p.terms; ++term) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t sa = smem_base + stage * p.stage_bytes;
            const uint32_t sb = sa + p.a_bytes;
            const uint32_t fb = bar_full + 8 * stage;
            if (rank == 0) mbar_expect_tx(fb, 2 * (p.a_bytes + p.b_bytes));      // both CTAs' boxes land on the leader's barrier
            const void* mm = (term == 1) ? (const void*)&map_m_lo : pm;
            const void* mn = (term == 2) ? (const void*)&map_n_lo : pn;
            tma_load_5d_2sm(sa, mm, fb, 0, x0 + sdx_m, y0 + sdy_m, img, mblk0);
            tma_load_5d_2sm(sb, mn, fb, 0, x0 + sdx_n, y0 + sdy_n, img, nblk0);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if (p.plan_n) {
            if (++q == p.plan_n) { q = 0; ++img; }
          } else if (++tx == p.tiles_x) {
            tx = 0;
            if (++ty == p.tiles_y) { ty = 0; ++img; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && elect_one()) {
      // D=f32, A=B=bf16, both MN-major (bits 15, 16), N = block_n, M = 256 (the pair's two 128-channel halves)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)(256 >> 4) << 24);
      const uint32_t n_row_bytes = p.kc_n * 2, m_row_bytes = p.kc_m * 2;
      const uint32_t n_layout = (p.kc_n == 64) ? 2u : (p.kc_n == 32 ? 4u : 6u);
      const uint32_t m_layout = (p.kc_m == 64) ? 2u : (p.kc_m == 32 ? 4u : 6u);
      const uint64_t adesc_hi = (make_smem_desc(0, 8 * m_row_bytes, m_layout) & ~(((uint64_t)0x3FFF << 16) | 0x3FFF)) |
                                ((uint64_t)((m_box_bytes >> 4) & 0x3FFF) << 16);
      const uint64_t bdesc_hi = (make_smem_desc(0, 8 * n_row_bytes, n_layout) & ~(((uint64_t)0x3FFF << 16) | 0x3FFF)) |
                                ((uint64_t)((n_box_bytes >> 4) & 0x3FFF) << 16);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int task = task0; task < total_tasks; task += task_step) {
        const int split = task / tasks_per_split;
        const int p0 = (int)((long)split * p.total_patches / p.splits);
        const int p1 = (int)((long)(split + 1) * p.total_patches / p.splits);
        const int num_kb = (p1 - p0) * p.terms;
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t sb = sa + p.a_bytes;
#pragma unroll
          for (int k = 0; k < kWgPix / 16; ++k) {
            const uint64_t adesc = adesc_hi | (uint64_t)(((sa + k * 16 * m_row_bytes) >> 4) & 0x3FFF);
            const uint64_t bdesc = bdesc_hi | (uint64_t)(((sb + k * 16 * n_row_bytes) >> 4) & 0x3FFF);
            tc_mma_bf16_2sm(d_tmem, adesc, bdesc, idesc, (uint32_t)((kb | k) != 0));
          }
          tc_commit_2sm(bar_empty + 8 * stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(bar_tfull + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs): own 128 accumulator rows -> partial[split][tap][m][n] =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int task = task0; task < total_tasks; task += task_step) {
      const int split = task / tasks_per_split;
      int r = task - split * tasks_per_split;
      const int tap = r / (m_pairs * p.n_tiles);
      r -= tap * (m_pairs * p.n_tiles);
      const int mp = r / p.n_tiles, nt = r - mp * p.n_tiles;
      float* orow = p.partial + (((long)split * p.taps + tap) * p.m_pad + (2 * mp + (int)rank) * kTileM + row) * p.n_pad + nt * p.block_n;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * 256 + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.block_n; c0 += 32) {
        uint32_t v[32];
        tc_ld32(taddr0 + c0, v);
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(orow + c0 + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_tempty + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// dw[tap][ci][co] = sum_split partial[split][tap][m][n],  (m, n) = (ci, co) when the input provided M, else (co, ci).
// 32 x 32 tiles through shared memory so that both the partial reads (n fastest) and the dw writes (co fastest) are coalesced
// in either orientation.  grid = (ceil(Cout/32), ceil(Cin/32), taps), block = (32, 8).
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int taps, int m_pad, int n_pad, int Cin, int Cout,
                                    int dw_cout_stride, int x_is_m, float* __restrict__ dw) {
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * 32;
  const long split_stride = (long)taps * m_pad * n_pad;
  const float* base = partial + (long)tap * m_pad * n_pad;
  // four rows per thread, the split loop unrolled: 16 independent loads in flight (the rolled one-row-at-a-time version ran at a
  // quarter of HBM speed: 0.83 ms per step for 1.2 GB); the additions of one element keep their split order
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  long off[4];
  bool ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = threadIdx.y + 8 * j;
    // x_is_m: rows m = ci, columns n = co (already co-fastest).  else: rows m = co, columns n = ci (read ci-fastest, transpose below)
    const int m = (x_is_m ? ci0 : co0) + r, n = (x_is_m ? co0 : ci0) + threadIdx.x;
    ok[j] = x_is_m ? (m < Cin && n < Cout) : (m < Cout && n < Cin);
    off[j] = (long)m * n_pad + n;
  }
#pragma unroll 4
  for (int sp = 0; sp < splits; ++sp) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ok[j]) s[j] += base[sp * split_stride + off[j]];
  }
  if (x_is_m) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + threadIdx.y + 8 * j, co = co0 + threadIdx.x;
      if (ok[j]) dw[((long)tap * Cin + ci) * dw_cout_stride + co] = s[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) tile[threadIdx.y + 8 * j][threadIdx.x] = s[j];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {
      const int ci = ci0 + r, co = co0 + threadIdx.x;
      if (ci < Cin && co < Cout) dw[((long)tap * Cin + ci) * dw_cout_stride + co] = tile[threadIdx.x][r];
    }
  }
}

void pick_patch64(int H, int W, int* TW, int* TH) {
  int best = 1 << 30, bw = 16, bh = 4;
  const int cand[5][2] = {{16, 4}, {8, 8}, {32, 2}, {64, 1}, {4, 16}};
  for (int i = 0; i < 5; ++i) {
    int t = jcm_cdiv(W, cand[i][0]) * jcm_cdiv(H, cand[i][1]);
    if (t < best) { best = t; bw = cand[i][0]; bh = cand[i][1]; }
  }
  *TW = bw;
  *TH = bh;
}

struct WgradPlan {
  int x_is_m, m_ch, n_ch, m_tiles, n_tiles, block_n, kc_n, kc_m, m_pad, n_pad, splits, TW, TH, tiles_x, tiles_y, total_patches;
  bool mixed;          // patches follow `tiles` (mixed-shape plan) instead of the uniform TW x TH grid
  bool pair;           // CTA-pair kernel (cta_group::2): even number of M tiles, N tiles of 128 or 256 channels
  int tgroup;          // > 1: tap-group kernel (narrow layers, one term): taps per task
  TilePlan tiles;
};

// variant: bit 0 = single-CTA kernel instead of the CTA pair, bit 1 = uniform patch grid, bit 2 = no tap groups
void plan_wgrad(int B, int H, int W, int Cin, int Gc, int ksize, int kw, bool one_term, int variant, WgradPlan* pl) {
  const bool allow_mixed = one_term && !(variant & 2), allow_pair = !(variant & 1), allow_taps = one_term && !(variant & 4);
  pl->x_is_m = Cin > Gc;       // ties (conv1's 64 -> 64 layers): the shifted operand X on the narrow N side, so that tap groups apply
  pl->m_ch = pl->x_is_m ? Cin : Gc;
  pl->n_ch = pl->x_is_m ? Gc : Cin;
  pl->m_tiles = jcm_cdiv(pl->m_ch, kTileM);
  pl->block_n = pl->n_ch <= 256 ? pl->n_ch : 256;
  pl->n_tiles = jcm_cdiv(pl->n_ch, pl->block_n);
  pl->kc_n = (pl->block_n % 64) == 0 ? 64 : ((pl->block_n % 32) == 0 ? 32 : 16);
  pl->kc_m = (pl->m_ch % 64) == 0 ? 64 : ((pl->m_ch % 32) == 0 ? 32 : 16);
  pl->m_pad = pl->m_tiles * kTileM;
  pl->n_pad = pl->n_tiles * pl->block_n;
  pick_patch64(H, W, &pl->TW, &pl->TH);
  pl->tiles_x = jcm_cdiv(W, pl->TW);
  pl->tiles_y = jcm_cdiv(H, pl->TH);
  pl->total_patches = B * pl->tiles_x * pl->tiles_y;
  pl->mixed = allow_mixed && plan_tiles(H, W, kWgPix, true, kWgradShapes, pl->tiles_x * pl->tiles_y, &pl->tiles);
  if (pl->mixed) pl->total_patches = B * pl->tiles.n_tiles;
  // k-split: the persistent grid runs ceil(tasks / SMs) waves of tasks that each stream total_patches / splits k-blocks (+ an
  // epilogue worth ~6 k-blocks).  Pick the split count that minimises waves x task length: e.g. conv5 (648 tasks = 4.4 waves)
  // wastes 12 % of the last wave unsplit, 0.5 % with 5 splits; the partial sums cost one extra read in wgrad_reduce_kernel.
  pl->pair = allow_pair && (pl->m_tiles % 2) == 0 && (pl->block_n == 128 || pl->block_n == 256) && (pl->block_n / pl->kc_n) % 2 == 0 &&
             jcm_num_sms() >= 2;
  // narrow layers (N tile <= 64, one term, several taps) whose SHIFTED operand - the layer input - is the narrow N side: tap groups
  // that fill one 256-column TMEM stage.  Measured (tests/gpu_diag.py wgpair, profiles/r02/diag_wgtaps.txt): conv2 at 120x180
  // 1.135 -> 1.066 ms, at 60x90 0.364 -> 0.320 ms; with the shifted operand on the (zero-padded, 128-row) M side - conv1's 64 -> 64
  // layers - the group's extra boxes cost more than the shared gradient box saves (0.496 -> 0.966 ms), so those keep one tap per task.
  pl->tgroup = 1;
  if (!pl->pair && allow_taps && !pl->x_is_m && pl->block_n <= 64 && ksize * kw > 1) {
    pl->tgroup = 256 / pl->block_n;
    if (pl->tgroup > 4) pl->tgroup = 4;
    if (pl->tgroup > ksize * kw) pl->tgroup = ksize * kw;
  }
  // work units and workers of the persistent loop: (tap [group], M tile, N tile) on every SM, or (tap, M-tile pair, N tile) on SM pairs
  const int base = jcm_cdiv(ksize * kw, pl->tgroup) * (pl->pair ? pl->m_tiles / 2 : pl->m_tiles) * pl->n_tiles;
  const int sms = pl->pair ? (jcm_num_sms() & ~1) / 2 : jcm_num_sms();
  int best_s = 1;
  double best_cost = 1e300;
  // (up to 2 x SMs splits: a layer with few (tap group, tile) units - conv1 with its three taps in one group has ONE - must still
  // fill the machine)
  for (int s = 1; s <= 2 * sms; ++s) {
    if (s > 1 && s > pl->total_patches / 8) break;
    const double waves = (double)jcm_cdiv(base * s, sms);
    // + the reduction pass over s partial copies (HBM-bound), in units of one k-block (~0.5 us)
    const double reduce = (double)s * ksize * kw * pl->m_pad * pl->n_pad * 4.0 / 6.4e12 / 0.5e-6;
    const double cost = waves * (jcm_cdiv(pl->total_patches, s) + 6.0) + reduce;
    if (cost < best_cost * 0.999) { best_cost = cost; best_s = s; }
  }
  pl->splits = best_s;
}

}  // namespace

// Test / measurement switch of jcm_conv2d_wgrad (every variant computes the same values): bit 0 = single-CTA kernel instead of the
// CTA pair, bit 1 = uniform patch grid instead of the mixed-shape plan, bit 2 = one tap per task instead of tap groups on the narrow
// layers.  Process-wide; the product never sets it.
static int g_wgrad_variant = 0;
extern "C" int jcm_debug_set_wgrad_variant(int variant) {
  const int old = g_wgrad_variant;
  g_wgrad_variant = variant & 7;
  return old;
}

// bytes of fp32 partial sums jcm_conv2d_wgrad needs in `workspace`
extern "C" long jcm_conv2d_wgrad_workspace(int B, int H, int W, int Cin, int Gc, int ksize, int kw) {
  if (kw <= 0) kw = ksize;
  // the largest split count over all kernel variants (one- / three-term form, test variants): they pick different split counts
  WgradPlan pl;
  int splits = 1;
  for (int one_term = 0; one_term < 2; ++one_term)
    for (int variant = 0; variant < 8; ++variant) {
      plan_wgrad(B, H, W, Cin, Gc, ksize, kw, one_term != 0, variant, &pl);
      if (pl.splits > splits) splits = pl.splits;
    }
  return (long)splits * ksize * kw * pl.m_pad * pl.n_pad * (long)sizeof(float);
}

// x planes [B,H,W,Cin] (layer input, Cin multiple of 16), g planes [B,H,W,Gc] (gradient w.r.t. the conv output, Gc = Cout padded
// to a multiple of 16) -> dw fp32 [k*k][Cin][dw_cout_stride] (first Cout columns written).  Gradient of tf.nn.conv2d w.r.t. its
// filter (TF autodiff of main.py:135).
extern "C" int jcm_conv2d_wgrad(const void* x_hi, const void* x_lo, const void* g_hi, const void* g_lo, float* dw, void* workspace,
                                long workspace_bytes, int B, int H, int W, int Cin, int Gc, int Cout, int dw_cout_stride, int ksize,
                                int kw, void* stream) {
  if (kw <= 0) kw = ksize;
  JCM_CHECK_ARG(x_hi && g_hi && dw && workspace, "jcm_conv2d_wgrad: null pointer");
  JCM_CHECK_ARG((x_lo == nullptr) == (g_lo == nullptr), "jcm_conv2d_wgrad: x_lo and g_lo must both be given or both NULL");
  JCM_CHECK_ARG(B > 0 && H > 0 && W > 0 && (ksize & 1) && (kw & 1), "jcm_conv2d_wgrad: bad shape");
  JCM_CHECK_ARG((Cin % 16) == 0 && (Gc % 16) == 0 && Cout <= Gc && Cout <= dw_cout_stride, "jcm_conv2d_wgrad: channel counts must be multiples of 16 (Cin=%d Gc=%d)", Cin, Gc);
  WgradPlan pl;
  plan_wgrad(B, H, W, Cin, Gc, ksize, kw, x_lo == nullptr, g_wgrad_variant, &pl);
  JCM_CHECK_ARG(pl.n_ch % pl.block_n == 0, "jcm_conv2d_wgrad: N-side channel count %d must be <= 256 or a multiple of 256", pl.n_ch);
  if (workspace_bytes < jcm_conv2d_wgrad_workspace(B, H, W, Cin, Gc, ksize, kw)) {
    jcm_set_error("jcm_conv2d_wgrad: workspace too small");
    return JCM_EWORKSPACE;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.H = H; p.W = W;
  p.TW = pl.TW; p.TH = pl.TH; p.tiles_x = pl.tiles_x; p.tiles_y = pl.tiles_y;
  p.ksize = ksize; p.kw = kw; p.pad = (ksize - 1) / 2; p.pad_x = (kw - 1) / 2; p.taps = ksize * kw;
  p.m_tiles = pl.m_tiles; p.n_tiles = pl.n_tiles; p.block_n = pl.block_n; p.m_pad = pl.m_pad; p.n_pad = pl.n_pad;
  p.shift_m = pl.x_is_m ? 1 : 0; p.shift_n = pl.x_is_m ? 0 : 1;
  p.kc_n = pl.kc_n; p.kc_m = pl.kc_m;
  p.splits = pl.splits; p.total_patches = pl.total_patches;
  p.terms = x_lo ? 3 : 1;
  p.a_bytes = kTileM * kWgPix * 2;
  p.b_bytes = (pl.pair ? pl.block_n / 2 : pl.block_n) * kWgPix * 2;      // pair: each CTA holds half of the N-side box
  p.tgroup = pl.tgroup;
  // tap groups: the shifted side (the layer input) has one box per tap in every stage
  p.stage_bytes = (((p.shift_m ? p.tgroup : 1) * p.a_bytes + (p.shift_n ? p.tgroup : 1) * p.b_bytes + 1023) / 1024) * 1024;
  p.stages = (200 * 1024) / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  JCM_CHECK_ARG(p.stages >= 2, "jcm_conv2d_wgrad: tap group of %d does not fit shared memory", p.tgroup);
  p.partial = (float*)workspace;
  p.plan_n = pl.mixed ? pl.tiles.n_tiles : 0;
  if (pl.mixed) {
    memcpy(p.px0, pl.tiles.x0, sizeof(p.px0));
    memcpy(p.py0, pl.tiles.y0, sizeof(p.py0));
    memcpy(p.pshape, pl.tiles.shape, sizeof(p.pshape));
  }

  const void* m_hi = pl.x_is_m ? x_hi : g_hi;
  const void* m_lo = pl.x_is_m ? x_lo : g_lo;
  const void* n_hi = pl.x_is_m ? g_hi : x_hi;
  const void* n_lo = pl.x_is_m ? g_lo : x_lo;
  const int m_c = pl.x_is_m ? Cin : Gc, n_c = pl.x_is_m ? Gc : Cin;
  // 5-D views {kc channels, W, H, B, C / kc channel blocks}: one TMA box {kc, TW, TH, 1, blocks per tile} lands in shared memory as
  // [block][TH][TW][kc] = per channel block a [64 pixels][kc] MN-major UMMA operand.  Blocks past the tensor's channel extent
  // (M tile of 128 over a 64-channel tensor) are zero-filled by TMA, which pads M.
  CUtensorMap mm_hi, mm_lo, mn_hi, mn_lo;
  static WgradShapeMaps wsm_zero;
  WgradShapeMaps wsm = wsm_zero;
  {
    uint64_t dims[5] = {(uint64_t)pl.kc_m, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)(m_c / pl.kc_m)};
    uint64_t str[4] = {(uint64_t)m_c * 2, (uint64_t)W * m_c * 2, (uint64_t)H * W * m_c * 2, (uint64_t)pl.kc_m * 2};
    uint32_t box[5] = {(uint32_t)pl.kc_m, (uint32_t)pl.TW, (uint32_t)pl.TH, 1, (uint32_t)(kTileM / pl.kc_m)};
    const CUtensorMapSwizzle swz = pl.kc_m == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (pl.kc_m == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    int rc = make_map(&mm_hi, m_hi, 5, dims, str, box, swz);
    if (rc) return rc;
    rc = make_map(&mm_lo, m_lo ? m_lo : m_hi, 5, dims, str, box, swz);
    if (rc) return rc;
    for (int i = 0; i < (pl.mixed ? pl.tiles.n_shapes : 0); ++i) {
      uint32_t pbox[5] = {(uint32_t)pl.kc_m, pl.tiles.sw[i], pl.tiles.sh[i], 1, (uint32_t)(kTileM / pl.kc_m)};
      rc = make_map(&wsm.m[i], m_hi, 5, dims, str, pbox, swz);
      if (rc) return rc;
    }
  }
  {
    uint64_t dims[5] = {(uint64_t)pl.kc_n, (uint64_t)W, (uint64_t)H, (uint64_t)B, (uint64_t)(n_c / pl.kc_n)};
    uint64_t str[4] = {(uint64_t)n_c * 2, (uint64_t)W * n_c * 2, (uint64_t)H * W * n_c * 2, (uint64_t)pl.kc_n * 2};
    const uint32_t n_blocks = (uint32_t)(pl.block_n / pl.kc_n) / (pl.pair ? 2u : 1u);
    uint32_t box[5] = {(uint32_t)pl.kc_n, (uint32_t)pl.TW, (uint32_t)pl.TH, 1, n_blocks};
    const CUtensorMapSwizzle swz = pl.kc_n == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (pl.kc_n == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    int rc = make_map(&mn_hi, n_hi, 5, dims, str, box, swz);
    if (rc) return rc;
    rc = make_map(&mn_lo, n_lo ? n_lo : n_hi, 5, dims, str, box, swz);
    if (rc) return rc;
    for (int i = 0; i < (pl.mixed ? pl.tiles.n_shapes : 0); ++i) {
      uint32_t pbox[5] = {(uint32_t)pl.kc_n, pl.tiles.sw[i], pl.tiles.sh[i], 1, n_blocks};
      rc = make_map(&wsm.n[i], n_hi, 5, dims, str, pbox, swz);
      if (rc) return rc;
    }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  if (pl.pair) {
    const int pair_tasks = p.splits * p.taps * (p.m_tiles / 2) * p.n_tiles;
    int grid2 = jcm_num_sms() & ~1;
    if (grid2 > 2 * pair_tasks) grid2 = 2 * pair_tasks;
    JCM_CUDA(cudaFuncSetAttribute(conv_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
    conv_wgrad_pair_kernel<<<grid2, kThreads, smem, (cudaStream_t)stream>>>(mm_hi, mm_lo, mn_hi, mn_lo, wsm, p);
  } else if (p.tgroup > 1) {
    const int total_tasks = p.splits * jcm_cdiv(p.taps, p.tgroup) * p.m_tiles * p.n_tiles;
    int grid = jcm_num_sms();
    if (grid > total_tasks) grid = total_tasks;
    JCM_CUDA(cudaFuncSetAttribute(conv_wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
    conv_wgrad_taps_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(mm_hi, mn_hi, wsm, p);
  } else {
    const int total_tasks = p.splits * p.taps * p.m_tiles * p.n_tiles;
    int grid = jcm_num_sms();
    if (grid > total_tasks) grid = total_tasks;
    JCM_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 512));
    conv_wgrad_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(mm_hi, mm_lo, mn_hi, mn_lo, wsm, p);
  }
  JCM_LAUNCH_CHECK();
  wgrad_reduce_kernel<<<dim3(jcm_cdiv(Cout, 32), jcm_cdiv(Cin, 32), p.taps), dim3(32, 8), 0, (cudaStream_t)stream>>>(
      p.partial, p.splits, p.taps, p.m_pad, p.n_pad, Cin, Cout, dw_cout_stride, pl.x_is_m, dw);
  JCM_LAUNCH_CHECK();
  return JCM_OK;
}

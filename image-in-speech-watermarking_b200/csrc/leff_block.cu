// The whole LeFF block (uformerWM/model.py:683-714) as ONE persistent tcgen05 kernel:
//     x += Linear2( GELU( DepthwiseConv3x3( GELU( Linear1( LN2(x) ) ) ) ) )        [+ the next block's LayerNorm of x]
// The 4C-wide hidden tensor - written and read twice by the separate kernels (32C of the block's 44C bytes per
// token) - never exists in HBM: it lives in shared memory, 64 channels at a time (the depthwise conv does not mix
// channels, so the hidden dimension is processed in independent 64-channel chunks).
//
// One output tile = 16 x 8 pixels of one image (128 tokens = one UMMA M tile); its hidden patch = 18 x 10 pixels
// (one-pixel halo, 180 rows = two UMMA M tiles of linear1; the halo is recomputed).
//
// The kernel is bound by the CUDA cores (tools/ubench/pipe_rates.cu: MUFU.TANH 16 / clk / SM, FFMA2 37 instr-lanes /
// clk / SM, FHFMA 82 / clk / SM; per chunk 12 288 + 8 192 GELUs and 73 728 conv FMAs against 768 tensor cycles), so every
// warp that is not a single-thread issuer is an identical WORKER and the work of a chunk is cut into items:
//   G item  (12 per chunk)  32 patch rows x 32 hidden channels: tcgen05.ld -> + b1 -> GELU -> 0 outside the image ->
//                           fp16 hidden patch in shared memory (row pitch 144 B: conflict free without a swizzle, so
//                           the conv's loads take immediate offsets).  Bound to the TMEM lane quarter of the warp.
//   V item  (16 per chunk)  one tile row (8 pixels) x 64 channels: depthwise 3x3 by mixed-precision FMAs (FHFMA: fp16
//                           patch value x fp16 weight + fp32 accumulator, no unpacking) + GELU -> the A operand of
//                           linear2 in the 128B-swizzled K-major layout UMMA expects.  Claimed dynamically (one
//                           shared counter), which also evens out the unequal G load of the four lane quarters.
//   E item  (4 x C/32 per tile) 32 pixels x 32 output channels: tcgen05.ld -> + b2 + fp32 residual -> x
// A worker runs G(chunk g), then V items of chunk g-1, then (once per tile) its E items of the tile before: the
// stages are skewed by one chunk and meet only through mbarriers.
//   warp 0   TMA: the LayerNorm-2 output patch (4-D box, zero padding = out-of-bounds fill) once per tile, the W1
//            chunk [64 x C] and the W2 chunk [C x 64] per hidden chunk (mbarrier rings)
//   warp 1   linear1 issuer: patch (2 M tiles) x W1 chunk -> TMEM accumulator [2][128 x 64] (double buffered)
//   warp 2   linear2 issuer: A operand x W2 chunk accumulated over the chunks into TMEM [128 x C] (double buffered
//            across tiles)
//   warp 3   TMEM allocation
//   warps 4-19  workers
//
// PRECISE (the WMK_PREC_MIXED extractor): dense weights arrive as (hi + lo) fp16 pairs - two MMA groups per product -,
// the depthwise weights as (hi + lo) fp16 pairs - two FHFMA per tap -, and both GELUs are the erf form; otherwise one
// weight tile and the tanh form on pre-halved weights (uformer_plan.cu).
#include "tc_ptx.cuh"

namespace wmk {

namespace {

using namespace tc;

constexpr int LB_W0 = 4, LB_WORKERS = 16;
constexpr int LB_THREADS = 32 * (LB_W0 + LB_WORKERS);            // 640
constexpr int LB_PW = 10, LB_PH = 18, LB_PROWS = LB_PW * LB_PH;  // 180 patch pixels
constexpr uint32_t LB_PATCH_BYTES = LB_PROWS * 128;              // one 64-channel k-block of the patch as TMA writes it
constexpr uint32_t LB_PATCH_STRIDE = 23 * 1024;                  // ... padded to the 1024-byte swizzle atom
constexpr uint32_t LB_HP_PITCH = 144;                            // hidden patch: 64 fp16 + 16 bytes of padding per pixel
constexpr uint32_t LB_HP_BYTES = 26 * 1024;                      // 180 x 144 = 25 920
constexpr uint32_t LB_A2_BYTES = 128 * 128;
constexpr int LB_NG = 12, LB_NV = 16;

struct LbGeom {
  int H, tiles_w, tiles_per_img, m_tiles;
  int KC, ksteps;          // 64-channel k-blocks of the patch, k-steps per k-block (2 when C = 32)
  int lg_nj;               // hidden chunks of 64 channels = 4C / 64 = 1 << lg_nj
  int ap_st, w1_st, w2_st, hpb;   // patch buffers, W1 / W2 ring depths, hidden patch buffers
  const float *b1, *dw_b, *b2;    // b1 / dw_b pre-halved unless PRECISE
  const uint16_t* dw16;    // depthwise weights [sets][9][4C] fp16 (pre-halved unless PRECISE; PRECISE: hi set, lo set)
  float* x;                // fp32 residual stream, updated in place
};

#ifdef LB_TIMING
__device__ unsigned long long lb_timing[8];      // [0..4] cycles waited at acc2full, acc1full, hpempty, hpfull, a2empty; [5] worker total
#define LB_T0 const long long _t0 = clock64();
#define LB_T1(i) if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd(&lb_timing[i], (unsigned long long)(clock64() - _t0));
#else
#define LB_T0
#define LB_T1(i)
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 24)) __trap();       // a lost arrival becomes a CUDA error, not a hang
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
  } while (!done);
}

// two mixed-precision FMAs: a0 += x.lo * w.lo, a1 += x.hi * w.hi (fp16 products are exact in fp32)
__device__ __forceinline__ void fhfma2(float& a0, float& a1, uint32_t x, uint32_t w) {
  asm("{\n\t.reg .b16 xl, xh, wl, wh;\n\t"
      "mov.b32 {xl, xh}, %2;\n\t"
      "mov.b32 {wl, wh}, %3;\n\t"
      "fma.rn.f32.f16 %0, xl, wl, %0;\n\t"
      "fma.rn.f32.f16 %1, xh, wh, %1;\n\t}"
      : "+f"(a0), "+f"(a1)
      : "r"(x), "r"(w));
}

template <int C, bool PRECISE, int HPB>
__global__ void __launch_bounds__(LB_THREADS, 1)
leff_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, LbGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr int WT = PRECISE ? 2 : 1;                              // weight tiles per product (hi, lo)
  constexpr int NJ = 4 * C / 64;
  const uint32_t AP_BYTES = (uint32_t)g.KC * LB_PATCH_STRIDE;      // one patch buffer
  const uint32_t W1_BYTES = (uint32_t)g.KC * WT * 8192u;           // [kc][hi, lo][64 rows x 128 B]   (C = 32 PRECISE: one [hi|lo] tile)
  constexpr uint32_t W2_TILE = (uint32_t)C * 128u;
  constexpr uint32_t W2_BYTES = WT * W2_TILE;
  const uint32_t apatch = base;
  const uint32_t w1s = apatch + (uint32_t)g.ap_st * AP_BYTES;
  const uint32_t w2s = w1s + (uint32_t)g.w1_st * W1_BYTES;
  const uint32_t a2s = w2s + (uint32_t)g.w2_st * W2_BYTES;         // [2][16 KB]
  const uint32_t hps = a2s + 2u * LB_A2_BYTES;                     // [hpb][26 KB] hidden patch, 64 channels
  const uint32_t bars = hps + (uint32_t)HPB * LB_HP_BYTES;
  // mbarriers (8 bytes each)
  auto apfull = [&](int i) { return bars + 8u * i; };              // [2]
  auto apempty = [&](int i) { return bars + 8u * (2 + i); };
  auto w1full = [&](int i) { return bars + 8u * (4 + i); };
  auto w1empty = [&](int i) { return bars + 8u * (6 + i); };
  auto w2full = [&](int i) { return bars + 8u * (8 + i); };
  auto w2empty = [&](int i) { return bars + 8u * (10 + i); };
  auto acc1full = [&](int i) { return bars + 8u * (12 + i); };
  auto acc1empty = [&](int i) { return bars + 8u * (14 + i); };
  auto a2full = [&](int i) { return bars + 8u * (16 + i); };
  auto a2empty = [&](int i) { return bars + 8u * (18 + i); };
  auto acc2full = [&](int i) { return bars + 8u * (20 + i); };
  auto acc2empty = [&](int i) { return bars + 8u * (22 + i); };
  auto hpfull = [&](int i) { return bars + 8u * (24 + i); };       // [3]
  auto hpempty = [&](int i) { return bars + 8u * (27 + i); };      // [3]
  const uint32_t tmem_slot = bars + 8u * 30;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - raw));
  uint32_t* vctr = (uint32_t*)(smem_raw + (tmem_slot + 8u - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = (int)gridDim.x;
  constexpr int E_SLABS = C / 32;                                  // 32-column slabs of the output tile
  constexpr int E_WORKERS = 4 * (E_SLABS < 4 ? E_SLABS : 4);       // workers that own at least one slab

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(apfull(i), 1); mbar_init(apempty(i), 1);
      mbar_init(w1full(i), 1); mbar_init(w1empty(i), 1);
      mbar_init(w2full(i), 1); mbar_init(w2empty(i), 1);
      mbar_init(acc1full(i), 1); mbar_init(acc1empty(i), LB_NG);
      mbar_init(a2full(i), LB_NV); mbar_init(a2empty(i), 1);
      mbar_init(acc2full(i), 1); mbar_init(acc2empty(i), E_WORKERS);
    }
    for (int i = 0; i < 3; ++i) { mbar_init(hpfull(i), LB_NG); mbar_init(hpempty(i), LB_NV); }
    *vctr = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns: linear1 accumulators [buffer b][M tile mt] at b*128 + mt*64 (64 columns each); linear2 accumulators
  // [buffer] at 256 + buffer*C
  auto tile_at = [&](int tile, int& b, int& h0, int& w0) {
    b = tile / g.tiles_per_img;
    const int rem = tile - b * g.tiles_per_img;
    const int th = rem / g.tiles_w;
    h0 = th * 16;
    w0 = (rem - th * g.tiles_w) * 8;
  };
  const int my_tiles = (int)blockIdx.x < g.m_tiles ? (g.m_tiles - 1 - (int)blockIdx.x) / grid + 1 : 0;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA
    if (lane == 0) {               // patch of the tile, then the W1 chunks
      int gch = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        int img, h0, w0;
        tile_at((int)blockIdx.x + lt * grid, img, h0, w0);
        const int ab = lt % g.ap_st;
        mbar_wait(apempty(ab), (((uint32_t)(lt / g.ap_st)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(apfull(ab), (uint32_t)g.KC * LB_PATCH_BYTES);
        for (int kc = 0; kc < g.KC; ++kc)
          tma_load_4d(apatch + (uint32_t)ab * AP_BYTES + (uint32_t)kc * LB_PATCH_STRIDE, &tmA, kc * 64, w0 - 1, h0 - 1, img, apfull(ab));
        for (int j = 0; j < NJ; ++j, ++gch) {
          const int s = gch % g.w1_st;
          mbar_wait(w1empty(s), (((uint32_t)(gch / g.w1_st)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(w1full(s), W1_BYTES);
          const uint32_t wb = w1s + (uint32_t)s * W1_BYTES;
          if (PRECISE && C == 32) {                                  // rows [hi(32) | lo(32)] = ONE 64-wide tile... x2 slots kept equal
            tma_load_2d(wb, &tmW1, 0, j * 64, w1full(s));
            tma_load_2d(wb + 8192u, &tmW1, 0, j * 64, w1full(s));    // (second slot unused by the issuer; keeps W1_BYTES uniform)
          } else {
            for (int kc = 0; kc < g.KC; ++kc)
              for (int t = 0; t < WT; ++t)
                tma_load_2d(wb + (uint32_t)(kc * WT + t) * 8192u, &tmW1, t * C + kc * 64, j * 64, w1full(s));
          }
        }
      }
    } else if (lane == 16) {       // the W2 chunks
      const int total = my_tiles * NJ;
      for (int gch = 0; gch < total; ++gch) {
        const int j = gch & (NJ - 1);
        const int s = gch % g.w2_st;
        mbar_wait(w2empty(s), (((uint32_t)(gch / g.w2_st)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(w2full(s), W2_BYTES);
        for (int t = 0; t < WT; ++t)
          tma_load_2d(w2s + (uint32_t)s * W2_BYTES + (uint32_t)t * W2_TILE, &tmW2, t * 4 * C + j * 64, 0, w2full(s));
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ linear1 issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(64, true);
      int gch = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int ab = lt % g.ap_st;
        mbar_wait(apfull(ab), ((uint32_t)(lt / g.ap_st)) & 1u);
        for (int j = 0; j < NJ; ++j, ++gch) {
          const int s = gch % g.w1_st, b = gch & 1;
          mbar_wait(w1full(s), ((uint32_t)(gch / g.w1_st)) & 1u);
          mbar_wait(acc1empty(b), (((uint32_t)(gch >> 1)) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t wb = w1s + (uint32_t)s * W1_BYTES;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t tacc = tmem_base + (uint32_t)(b * 128 + mt * 64);
            uint32_t first = 1u;
            for (int kc = 0; kc < g.KC; ++kc) {
              const uint64_t adesc = umma_desc_sw128(apatch + (uint32_t)ab * AP_BYTES + (uint32_t)kc * LB_PATCH_STRIDE + (uint32_t)mt * 16384u);
              if (PRECISE && C == 32) {
                const uint64_t wd = umma_desc_sw128(wb);
                for (int k = 0; k < 2; ++k) { tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc, first ? 0u : 1u); first = 0u; }
                for (int k = 0; k < 2; ++k) tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), wd + (uint64_t)(2 * (k + 2)), idesc, 1u);
              } else {
                for (int t = 0; t < WT; ++t) {
                  const uint64_t wd = umma_desc_sw128(wb + (uint32_t)(kc * WT + t) * 8192u);
                  for (int k = 0; k < g.ksteps; ++k) {
                    tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc, first ? 0u : 1u);
                    first = 0u;
                  }
                }
              }
            }
          }
          tcgen05_commit(w1empty(s));
          tcgen05_commit(acc1full(b));
        }
        tcgen05_commit(apempty(ab));           // the patch buffer is free once the last chunk's MMAs retire
      }
    }
  } else if (warp == 2) {
    // ------------------------------------------------------------------------------------------ linear2 issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(C, true);
      int gch = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int ac = lt & 1;
        mbar_wait(acc2empty(ac), (((uint32_t)(lt >> 1)) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(256 + ac * C);
        for (int j = 0; j < NJ; ++j, ++gch) {
          const int s = gch % g.w2_st, b = gch & 1;
          mbar_wait(w2full(s), ((uint32_t)(gch / g.w2_st)) & 1u);
          mbar_wait(a2full(b), ((uint32_t)(gch >> 1)) & 1u);
          tcgen05_fence_after();
          const uint64_t adesc = umma_desc_sw128(a2s + (uint32_t)b * LB_A2_BYTES);
#pragma unroll
          for (int t = 0; t < WT; ++t) {
            const uint64_t wd = umma_desc_sw128(w2s + (uint32_t)s * W2_BYTES + (uint32_t)t * W2_TILE);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc, (uint32_t)((j | t | k) != 0));
          }
          tcgen05_commit(a2empty(b));
          tcgen05_commit(w2empty(s));
        }
        tcgen05_commit(acc2full(ac));
      }
    }
  } else if (warp >= LB_W0) {
    // ------------------------------------------------------------------------------------------ workers
    const int wi = warp - LB_W0;
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int k4 = wi >> 2;                      // 0..3: which of the quarter's four workers
    // G item of this worker: lane quarter q of M tile g_mt, hidden channels g_ch*32 .. +31 of the chunk
    const bool has_g = q < 2 || k4 < 2;          // patch rows 192..255 do not exist
    const int g_mt = q < 2 ? (k4 >> 1) : 0;
    const int g_ch = q < 2 ? (k4 & 1) : k4;
    const int prow = g_mt * 128 + q * 32 + lane; // patch pixel of this thread
    const int py = prow / LB_PW, px = prow - py * LB_PW;
    const uint32_t g_tmem = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g_mt * 64 + g_ch * 32);
    const uint32_t g_hp = hps + (uint32_t)prow * LB_HP_PITCH + (uint32_t)g_ch * 64u;
    // V item geometry: lane = 4 channels (cg) x one half (4 pixels) of the tile row
    const int cg = lane & 15, half = lane >> 4;
    uint32_t a2off[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t m7 = (uint32_t)(half * 4 + i);
      a2off[i] = m7 * 128u + ((((uint32_t)cg >> 1) ^ m7) << 4) + (uint32_t)(cg & 1) * 8u;
    }
    const uint32_t v_hp = (uint32_t)(half * 4) * LB_HP_PITCH + (uint32_t)cg * 8u;
    const int T = my_tiles * NJ, totalV = T * LB_NV;
    int pending = -1;
    uint32_t keep = 0u;                          // all-ones when this thread's patch pixel lies inside the image
    int hb_g = 0;                                // hidden patch buffer of chunk g (g % hpb) and its use count parity
    uint32_t hp_par_g = 0u;

    auto e_items = [&](int lt) {                 // output tile lt of this CTA: + b2 + residual -> x
      if (k4 >= E_SLABS) return;
      const int ac = lt & 1;
      int tb, h0, w0;
      tile_at((int)blockIdx.x + lt * grid, tb, h0, w0);
      const int m = q * 32 + lane;               // tile pixel = TMEM lane
      const size_t tok = ((size_t)tb * g.H + h0 + (m >> 3)) * g.H + w0 + (m & 7);
      float* xrow = g.x + tok * C;
      bool waited = false;
#pragma unroll 1
      for (int s = k4; s < E_SLABS; s += 4) {
        float4 r[8];
        const float4* r4 = reinterpret_cast<const float4*>(xrow + s * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = r4[i];
        if (!waited) {
          { LB_T0 mbar_wait_sleep(acc2full(ac), ((uint32_t)(lt >> 1)) & 1u); LB_T1(0) }
          tcgen05_fence_after();
          waited = true;
        }
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 + ac * C + s * 32), v);
        const float4* b4 = reinterpret_cast<const float4*>(g.b2 + s * 32);
        float4* o4 = reinterpret_cast<float4*>(xrow + s * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(b4 + i);
          float4 o;
          o.x = __uint_as_float(v[4 * i]) + b.x + r[i].x;
          o.y = __uint_as_float(v[4 * i + 1]) + b.y + r[i].y;
          o.z = __uint_as_float(v[4 * i + 2]) + b.z + r[i].z;
          o.w = __uint_as_float(v[4 * i + 3]) + b.w + r[i].w;
          o4[i] = o;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2empty(ac));
    };

#ifdef LB_TIMING
    const long long _tw = clock64();
#endif
    for (int gi = 0; gi <= T; ++gi) {
      // ---------------------------------------------------------------- G item of chunk gi
      if (gi < T) {
        const int j = gi & (NJ - 1);
        if (j == 0) {
          int tb, h0, w0;
          tile_at((int)blockIdx.x + (gi >> g.lg_nj) * grid, tb, h0, w0);
          const int hh = h0 - 1 + py, ww = w0 - 1 + px;
          keep = (prow < LB_PROWS && hh >= 0 && hh < g.H && ww >= 0 && ww < g.H) ? 0xffffffffu : 0u;
          if (k4 < E_SLABS) {       // the residual rows this worker's E item will read: pull them into L2 now
            const int m = q * 32 + lane;
            const float* xr = g.x + (((size_t)tb * g.H + h0 + (m >> 3)) * g.H + w0 + (m & 7)) * C + k4 * 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xr));
          }
        }
        if (has_g) {
          const int b = gi & 1;
          float4 bias[8];
          const float4* b4 = reinterpret_cast<const float4*>(g.b1 + j * 64 + g_ch * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) bias[i] = __ldg(b4 + i);
          { LB_T0 mbar_wait_sleep(acc1full(b), ((uint32_t)(gi >> 1)) & 1u); LB_T1(1) }
          tcgen05_fence_after();
          uint32_t v[32];
          tmem_ld32(g_tmem + (uint32_t)(b * 128), v);
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc1empty(b));            // the accumulator quarter has been read
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float2 a0 = __fadd2_rn(make_float2(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])), make_float2(bias[i].x, bias[i].y));
            float2 a1 = __fadd2_rn(make_float2(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])), make_float2(bias[i].z, bias[i].w));
            if (PRECISE) { a0 = gelu_erf2(a0); a1 = gelu_erf2(a1); }
            else { a0 = gelu_tanh2_half_arg(a0); a1 = gelu_tanh2_half_arg(a1); }
            w[2 * i] = pack2_f16(a0.x, a0.y) & keep;           // the conv's zero padding acts on the hidden tensor
            w[2 * i + 1] = pack2_f16(a1.x, a1.y) & keep;
          }
          { LB_T0 mbar_wait_sleep(hpempty(hb_g), hp_par_g ^ 1u); LB_T1(2) }      // the V items that read this buffer last are done
          if (prow < LB_PROWS) {
            const uint32_t dst = g_hp + (uint32_t)hb_g * LB_HP_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) st_shared_v4(dst + 16u * i, w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(hpfull(hb_g));
        }
        if (++hb_g == HPB) { hb_g = 0; hp_par_g ^= 1u; }
      }
      // ---------------------------------------------------------------- V items of chunk gi - 1 (claimed dynamically)
      if (gi >= 1) {
        for (;;) {
          if (pending < 0) {
            int id = 0;
            if (lane == 0) id = (int)atomicAdd(vctr, 1u);
            pending = __shfl_sync(0xffffffffu, id, 0);
          }
          if (pending >= totalV) break;
          const int c = pending >> 4;                          // chunk of the item
          if (c > gi - 1) break;                               // belongs to a later chunk: keep it for then
          const int r = pending & 15;                          // tile row
          pending = -1;
          const int j = c & (NJ - 1);
          const int hb = c % HPB, b = c & 1;
          // depthwise weights of this thread's 4 channels (fp16 pairs) and the bias
          uint2 wv[PRECISE ? 2 : 1][9];
          const uint16_t* wp = g.dw16 + j * 64 + cg * 4;
#pragma unroll
          for (int sset = 0; sset < (PRECISE ? 2 : 1); ++sset)
#pragma unroll
            for (int t = 0; t < 9; ++t) wv[sset][t] = __ldg(reinterpret_cast<const uint2*>(wp + (size_t)(sset * 9 + t) * (4 * C)));
          const float4 bz = __ldg(reinterpret_cast<const float4*>(g.dw_b + j * 64 + cg * 4));
          { LB_T0 mbar_wait_sleep(hpfull(hb), ((uint32_t)(c / HPB)) & 1u); LB_T1(3) }
          const uint8_t* hp = smem_raw + (hps + (uint32_t)hb * LB_HP_BYTES - raw) + v_hp + (uint32_t)(r * LB_PW) * LB_HP_PITCH;
          uint2 win[3][6];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int kx = 0; kx < 6; ++kx) win[dy][kx] = *reinterpret_cast<const uint2*>(hp + (dy * LB_PW + kx) * LB_HP_PITCH);
          { LB_T0 mbar_wait_sleep(a2empty(b), (((uint32_t)(c >> 1)) & 1u) ^ 1u); LB_T1(4) }
          uint8_t* ab = smem_raw + (a2s + (uint32_t)b * LB_A2_BYTES - raw) + (uint32_t)r * 1024u;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float a0 = bz.x, a1 = bz.y, a2 = bz.z, a3 = bz.w;
#pragma unroll
            for (int sset = 0; sset < (PRECISE ? 2 : 1); ++sset)
#pragma unroll
              for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                  fhfma2(a0, a1, win[dy][i + dx].x, wv[sset][dy * 3 + dx].x);
                  fhfma2(a2, a3, win[dy][i + dx].y, wv[sset][dy * 3 + dx].y);
                }
            float2 g0 = make_float2(a0, a1), g1 = make_float2(a2, a3);
            if (PRECISE) { g0 = gelu_erf2(g0); g1 = gelu_erf2(g1); }
            else { g0 = gelu_tanh2_half_arg(g0); g1 = gelu_tanh2_half_arg(g1); }
            uint2 o;
            o.x = pack2_f16(g0.x, g0.y);
            o.y = pack2_f16(g1.x, g1.y);
            *reinterpret_cast<uint2*>(ab + a2off[i]) = o;
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(a2full(b));
            mbar_arrive(hpempty(hb));
          }
        }
      }
      // ---------------------------------------------------------------- E items of the tile before the current one
      if (gi >= NJ + 1 && ((gi - 1) & (NJ - 1)) == 0) e_items(((gi - 1) >> g.lg_nj) - 1);
    }
    if (my_tiles > 0) e_items(my_tiles - 1);
#ifdef LB_TIMING
    if (blockIdx.x == 0 && lane == 0) atomicAdd(&lb_timing[5], (unsigned long long)(clock64() - _tw));
#endif
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 3) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int C, bool PRECISE>
int launch_leff_block(const void* A, const void* W1, const void* W2, const float* b1, const uint16_t* dw16, const float* dw_b,
                      const float* b2, float* x, int n, int H, cudaStream_t st) {
  LbGeom g;
  g.H = H; g.tiles_w = H / 8; g.tiles_per_img = (H / 8) * (H / 16); g.m_tiles = n * g.tiles_per_img;
  g.KC = C >= 64 ? C / 64 : 1; g.ksteps = C >= 64 ? 4 : 2;
  g.lg_nj = C == 32 ? 1 : C == 64 ? 2 : 3;
  g.b1 = b1; g.dw16 = dw16; g.dw_b = dw_b; g.b2 = b2; g.x = x;
  constexpr int WT = PRECISE ? 2 : 1;
  const int ap_bytes = g.KC * (int)LB_PATCH_STRIDE, w1_bytes = g.KC * WT * 8192, w2_bytes = WT * C * 128;
  const int fixed = 1024 + 2 * (int)LB_A2_BYTES + 512;
  g.ap_st = 2; g.w1_st = 2; g.w2_st = 2; g.hpb = 3;
  auto total = [&]() { return fixed + g.ap_st * ap_bytes + g.w1_st * w1_bytes + g.w2_st * w2_bytes + g.hpb * (int)LB_HP_BYTES; };
  if (total() > 227 * 1024) g.hpb = 2;
  if (total() > 227 * 1024) g.w2_st = 1;
  if (total() > 227 * 1024) g.ap_st = 1;
  if (total() > 227 * 1024) g.w1_st = 1;
  WMK_REQUIRE(total() <= 227 * 1024, "leff_block: %d bytes of shared memory needed (C = %d)", total(), C);
  CUtensorMap tmA, tmW1, tmW2;
  {   // LayerNorm-2 output [n][H][H][C], 16-bit: the 18 x 10 pixel patch of 64 channels, zero fill outside the image
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)H, (uint64_t)H, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)H * C * 2, (uint64_t)H * H * C * 2};
    const uint32_t box[4] = {64, LB_PW, LB_PH, 1};
    WMK_TRY(make_tensor_map(&tmA, A, 4, dims, strides, box, false, 128));
  }
  {   // W1 [4C][C] (PRECISE: [4C][hi(C) | lo(C)]): 64 x 64 tiles
    const uint64_t dims[2] = {(uint64_t)(WT * C), (uint64_t)4 * C};
    const uint64_t strides[1] = {(uint64_t)(WT * C) * 2};
    const uint32_t box[2] = {64, 64};
    WMK_TRY(make_tensor_map(&tmW1, W1, 2, dims, strides, box, false, 128));
  }
  {   // W2 [C][4C] (PRECISE: [C][hi(4C) | lo(4C)]): C x 64 tiles
    const uint64_t dims[2] = {(uint64_t)(WT * 4 * C), (uint64_t)C};
    const uint64_t strides[1] = {(uint64_t)(WT * 4 * C) * 2};
    const uint32_t box[2] = {64, (uint32_t)C};
    WMK_TRY(make_tensor_map(&tmW2, W2, 2, dims, strides, box, false, 128));
  }
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(leff_block_kernel<C, PRECISE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    WMK_CHECK_CUDA(cudaFuncSetAttribute(leff_block_kernel<C, PRECISE, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = g.m_tiles < num_sms() ? g.m_tiles : num_sms();
  if (g.hpb == 3) leff_block_kernel<C, PRECISE, 3><<<grid, LB_THREADS, (size_t)total(), st>>>(tmA, tmW1, tmW2, g);
  else leff_block_kernel<C, PRECISE, 2><<<grid, LB_THREADS, (size_t)total(), st>>>(tmA, tmW1, tmW2, g);
  WMK_CHECK_LAUNCH("leff_block_kernel");
  return 0;
}

}  // namespace

#ifdef LB_TIMING
int leff_block_timing(unsigned long long* out) {
  cudaMemcpyFromSymbol(out, lb_timing, sizeof(unsigned long long) * 8);
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(lb_timing, z, sizeof(z));
  return 0;
}
#endif

// x[M][C] += LeFF(A) with A = LayerNorm-2 output [n][H][H][C] fp16, token layout.
//   plain   (precise = 0): W1 [4C][C], W2 [C][4C] fp16; b1, dw16 [9][4C] (fp16), dw_b PRE-HALVED (tanh-form GELU)
//   precise (precise = 1): W1 [4C][hi(C) | lo(C)], W2 [C][hi(4C) | lo(4C)] fp16 pairs; dw16 [2][9][4C] = hi set, lo set;
//                          plain b1 / dw_b, erf-form GELU
// C in {32, 64, 128}, H a power of two >= 16.
int leff_block(const void* A, const void* W1, const void* W2, const float* b1, const uint16_t* dw16, const float* dw_b,
               const float* b2, float* x, int n, int H, int C, int precise, cudaStream_t st) {
  WMK_REQUIRE(H >= 16 && H <= 128 && (H & (H - 1)) == 0, "leff_block: H=%d must be a power of two in [16,128]", H);
  WMK_REQUIRE(C == 32 || C == 64 || C == 128, "leff_block: covers C in {32,64,128}, got %d", C);
  const double M = (double)n * H * H;
  // algorithmic traffic: A 2C + residual 4C + x 4C bytes per token (+ weights); FLOPs: the two dense layers
  ProfScope prof(FAM_GEMM_HBM, M * C * 10 + 16.0 * C * C, st, 2.0 * M * C * 4 * C * 2);
  if (precise) {
    switch (C) {
      case 32: return launch_leff_block<32, true>(A, W1, W2, b1, dw16, dw_b, b2, x, n, H, st);
      case 64: return launch_leff_block<64, true>(A, W1, W2, b1, dw16, dw_b, b2, x, n, H, st);
      default: return launch_leff_block<128, true>(A, W1, W2, b1, dw16, dw_b, b2, x, n, H, st);
    }
  }
  switch (C) {
    case 32: return launch_leff_block<32, false>(A, W1, W2, b1, dw16, dw_b, b2, x, n, H, st);
    case 64: return launch_leff_block<64, false>(A, W1, W2, b1, dw16, dw_b, b2, x, n, H, st);
    default: return launch_leff_block<128, false>(A, W1, W2, b1, dw16, dw_b, b2, x, n, H, st);
  }
}

}  // namespace wmk

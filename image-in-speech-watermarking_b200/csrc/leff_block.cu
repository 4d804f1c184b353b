// The whole LeFF block (uformerWM/model.py:683-714) as ONE persistent tcgen05 kernel:
//     x += Linear2( GELU( DepthwiseConv3x3( GELU( Linear1( LN2(x) ) ) ) ) )
// The 4C-wide hidden tensor - written and read twice by the separate kernels (32C of the block's 44C bytes per
// token) - never exists in HBM: it lives in shared memory, 64 channels at a time (the depthwise conv does not mix
// channels, so the hidden dimension is processed in independent 64-channel chunks).
//
// One output tile = 16 x 8 pixels of one image (128 tokens = one UMMA M tile); its hidden patch = 18 x 10 pixels
// (one-pixel halo, 180 rows = two UMMA M tiles of linear1, the halo is recomputed: 1.4x of linear1's FLOPs).
//
//   warp 0       TMA: the LayerNorm-2 output patch (4-D box, zero padding = out-of-bounds fill) once per tile, the
//                W1 chunk [64 x C] and the W2 chunk [C x 64] per hidden chunk (mbarrier rings)
//   warp 1       linear1 issuer: patch (2 M tiles) x W1 chunk -> TMEM accumulator [2][128 x 64] (double buffered)
//   warps 3-10   "G": tcgen05.ld -> + b1 -> GELU -> 0 outside the image -> 16-bit hidden patch in shared memory;
//                after the last chunk of a tile the same warps run the output epilogue of the PREVIOUS tile
//                (+ b2 + fp32 residual -> swizzled staging -> TMA store of x), so the two never wait on each other
//   warps 11-18  "V": depthwise 3x3 (fp32, FFMA2) + GELU on the hidden patch -> the 128 x 64 A operand of linear2 in
//                the 128B-swizzled K-major layout UMMA expects (fence.proxy.async, mbarrier arrive)
//   warp 2       linear2 issuer: A operand x W2 chunk accumulated over the chunks into TMEM [128 x C] (double
//                buffered across tiles)
//
// PRECISE (the WMK_PREC_MIXED extractor): weights arrive as (hi + lo) fp16 pairs - two MMA groups per product - and
// both GELUs are the erf form; otherwise one weight tile and the tanh form on pre-halved weights (uformer_plan.cu).
#include "tc_ptx.cuh"

namespace wmk {

namespace {

using namespace tc;

constexpr int LB_G0 = 3, LB_NG = 8, LB_V0 = LB_G0 + LB_NG, LB_NV = 8;
constexpr int LB_THREADS = 32 * (LB_V0 + LB_NV);                 // 608
constexpr int LB_PW = 10, LB_PH = 18, LB_PROWS = LB_PW * LB_PH;  // 180 patch pixels
constexpr uint32_t LB_PATCH_BYTES = LB_PROWS * 128;              // one 64-channel k-block of the patch as TMA writes it
constexpr uint32_t LB_PATCH_STRIDE = 23 * 1024;                  // ... padded to the 1024-byte swizzle atom
constexpr uint32_t LB_A2_BYTES = 128 * 128;
constexpr uint32_t LB_STG_BYTES = 4096;

struct LbGeom {
  int H, tiles_w, tiles_per_img, m_tiles;
  int C, KC, ksteps;       // width, 64-channel k-blocks of the patch, k-steps per k-block (2 when C = 32)
  int NJ;                  // hidden chunks of 64 channels = 4C / 64
  int ap_st, w_st;         // patch buffers (1 / 2), W ring depth (1 / 2)
  int f16;                 // 16-bit tensors are fp16 (else bf16)
  const float *b1, *dw_w, *dw_b, *b2;   // b1 / dw_* pre-halved unless PRECISE
  const float* resid;      // x (fp32), also the output
  int M;
};

template <int C, bool PRECISE>
__global__ void __launch_bounds__(LB_THREADS, 1)
leff_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                  const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmC, LbGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr int WT = PRECISE ? 2 : 1;                              // weight tiles per product (hi, lo)
  const uint32_t AP_BYTES = (uint32_t)g.KC * LB_PATCH_STRIDE;      // one patch buffer
  const uint32_t W1_BYTES = (uint32_t)g.KC * WT * 8192u;           // [kc][hi, lo][64 rows x 128 B]   (C = 32 PRECISE: one [hi|lo] tile)
  constexpr uint32_t W2_TILE = (uint32_t)C * 128u;
  constexpr uint32_t W2_BYTES = WT * W2_TILE;
  const uint32_t apatch = base;
  const uint32_t w1s = apatch + (uint32_t)g.ap_st * AP_BYTES;
  const uint32_t w2s = w1s + (uint32_t)g.w_st * W1_BYTES;
  const uint32_t hps = w2s + (uint32_t)g.w_st * W2_BYTES;          // [2][LB_PATCH_STRIDE] hidden patch, 64 channels
  const uint32_t a2s = hps + 2u * LB_PATCH_STRIDE;                 // [2][16 KB]
  const uint32_t staging = a2s + 2u * LB_A2_BYTES;                 // [8 warps][4 KB]
  const uint32_t bars = staging + LB_NG * LB_STG_BYTES;
  // mbarriers (8 bytes each)
  auto apfull = [&](int i) { return bars + 8u * i; };              // [2]
  auto apempty = [&](int i) { return bars + 8u * (2 + i); };       // [2]
  auto w1full = [&](int i) { return bars + 8u * (4 + i); };
  auto w1empty = [&](int i) { return bars + 8u * (6 + i); };
  auto w2full = [&](int i) { return bars + 8u * (8 + i); };
  auto w2empty = [&](int i) { return bars + 8u * (10 + i); };
  auto acc1full = [&](int i) { return bars + 8u * (12 + i); };
  auto acc1empty = [&](int i) { return bars + 8u * (14 + i); };
  auto hpfull = [&](int i) { return bars + 8u * (16 + i); };
  auto hpempty = [&](int i) { return bars + 8u * (18 + i); };
  auto a2full = [&](int i) { return bars + 8u * (20 + i); };
  auto a2empty = [&](int i) { return bars + 8u * (22 + i); };
  auto acc2full = [&](int i) { return bars + 8u * (24 + i); };
  auto acc2empty = [&](int i) { return bars + 8u * (26 + i); };
  const uint32_t tmem_slot = bars + 8u * 28;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = (int)gridDim.x;
  constexpr int EPI_WARPS = C >= 64 ? 8 : 4;                      // G warps that own output columns (32 per warp and piece)
  constexpr int GELU_WARPS = 6;                                    // (M tile 0: 4 quarters) + (M tile 1: rows 128..159, 160..179)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(apfull(i), 1); mbar_init(apempty(i), 1);
      mbar_init(w1full(i), 1); mbar_init(w1empty(i), 1);
      mbar_init(w2full(i), 1); mbar_init(w2empty(i), 1);
      mbar_init(acc1full(i), 1); mbar_init(acc1empty(i), GELU_WARPS);
      mbar_init(hpfull(i), GELU_WARPS); mbar_init(hpempty(i), LB_NV);
      mbar_init(a2full(i), LB_NV); mbar_init(a2empty(i), 1);
      mbar_init(acc2full(i), 1); mbar_init(acc2empty(i), EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
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
  const int NJ = g.NJ;

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA
    if (lane == 0) {               // patch of the tile, then the W1 chunks
      int lt = 0, gch = 0;
      for (int tile = blockIdx.x; tile < g.m_tiles; tile += grid, ++lt) {
        int img, h0, w0;
        tile_at(tile, img, h0, w0);
        const int ab = lt % g.ap_st;
        mbar_wait(apempty(ab), (((uint32_t)(lt / g.ap_st)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(apfull(ab), (uint32_t)g.KC * LB_PATCH_BYTES);
        for (int kc = 0; kc < g.KC; ++kc)
          tma_load_4d(apatch + (uint32_t)ab * AP_BYTES + (uint32_t)kc * LB_PATCH_STRIDE, &tmA, kc * 64, w0 - 1, h0 - 1, img, apfull(ab));
        for (int j = 0; j < NJ; ++j, ++gch) {
          const int s = gch % g.w_st;
          mbar_wait(w1empty(s), (((uint32_t)(gch / g.w_st)) & 1u) ^ 1u);
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
      int gch = 0;
      for (int tile = blockIdx.x; tile < g.m_tiles; tile += grid) {
        for (int j = 0; j < NJ; ++j, ++gch) {
          const int s = gch % g.w_st;
          mbar_wait(w2empty(s), (((uint32_t)(gch / g.w_st)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(w2full(s), W2_BYTES);
          for (int t = 0; t < WT; ++t)
            tma_load_2d(w2s + (uint32_t)s * W2_BYTES + (uint32_t)t * W2_TILE, &tmW2, t * 4 * C + j * 64, 0, w2full(s));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ linear1 issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(64, g.f16 != 0);
      int lt = 0, gch = 0;
      for (int tile = blockIdx.x; tile < g.m_tiles; tile += grid, ++lt) {
        const int ab = lt % g.ap_st;
        mbar_wait(apfull(ab), ((uint32_t)(lt / g.ap_st)) & 1u);
        for (int j = 0; j < NJ; ++j, ++gch) {
          const int s = gch % g.w_st, b = gch & 1;
          mbar_wait(w1full(s), ((uint32_t)(gch / g.w_st)) & 1u);
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
      const uint32_t idesc = umma_idesc(C, g.f16 != 0);
      int lt = 0, gch = 0;
      for (int tile = blockIdx.x; tile < g.m_tiles; tile += grid, ++lt) {
        const int ac = lt & 1;
        mbar_wait(acc2empty(ac), (((uint32_t)(lt >> 1)) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(256 + ac * C);
        for (int j = 0; j < NJ; ++j, ++gch) {
          const int s = gch % g.w_st, b = gch & 1;
          mbar_wait(w2full(s), ((uint32_t)(gch / g.w_st)) & 1u);
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
  } else if (warp < LB_V0) {
    // ------------------------------------------------------------------------------------------ G warps
    const int gw = warp - LB_G0;                 // 0..7
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int mt = gw >> 2;                      // linear1 M tile (patch rows mt*128 ..) / output column half
    const int prow = mt * 128 + q * 32 + lane;   // patch pixel of this thread
    const bool gelu_warp = mt == 0 || q < 2;     // rows 192..255 do not exist
    const int py = prow / LB_PW, px = prow - py * LB_PW;
    const uint32_t buf = staging + (uint32_t)gw * LB_STG_BYTES;
    constexpr int CPW = C >= 64 ? C / 2 : 32;    // output columns per epilogue warp
    const bool epi_warp = gw < EPI_WARPS;
    const int col0 = (C >= 64 ? mt : 0) * CPW;

    auto epilogue = [&](int tile, int lt) {      // output tile `tile` (the lt-th of this CTA): + b2 + residual -> x
      if (!epi_warp) return;
      const int ac = lt & 1;
      int tb, h0, w0;
      tile_at(tile, tb, h0, w0);
      const int grow = tb * g.H + h0 + q * 4;                         // first image row of this warp in the [clip*H + h] view
      const int rowi = (grow + (lane >> 3)) * g.H + w0 + (lane & 7);  // token index of this thread's row
      float4 rpre[8];
      {
        const float4* r4 = reinterpret_cast<const float4*>(g.resid + (size_t)rowi * C + col0);
#pragma unroll
        for (int j = 0; j < 8; ++j) rpre[j] = r4[j];
      }
      mbar_wait(acc2full(ac), ((uint32_t)(lt >> 1)) & 1u);
      tcgen05_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(256 + ac * C + col0);
#pragma unroll 1
      for (int cc = 0; cc < CPW; cc += 32) {
        uint32_t v[32];
        tmem_ld32(tacc + (uint32_t)cc, v);
        const int n = col0 + cc;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        const float4* b4 = reinterpret_cast<const float4*>(g.b2 + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
        }
        if (cc == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[4 * j] += rpre[j].x; f[4 * j + 1] += rpre[j].y; f[4 * j + 2] += rpre[j].z; f[4 * j + 3] += rpre[j].w;
          }
        } else {
          const float4* r4 = reinterpret_cast<const float4*>(g.resid + (size_t)rowi * C + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 r = r4[j];
            f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
          }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(buf + (uint32_t)lane * 128u + (((uint32_t)j ^ (uint32_t)(lane & 7)) << 4),
                       __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                       __float_as_uint(f[4 * j + 3]));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tma_store_3d(&tmC, buf, n, w0, grow);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2empty(ac));
    };

    int lt = 0, gch = 0, prev_tile = -1;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += grid, ++lt) {
      int tb, h0, w0;
      tile_at(tile, tb, h0, w0);
      const int hh = h0 - 1 + py, ww = w0 - 1 + px;
      const bool inimg = prow < LB_PROWS && hh >= 0 && hh < g.H && ww >= 0 && ww < g.H;
      for (int j = 0; j < NJ; ++j, ++gch) {
        if (gelu_warp) {
          const int b = gch & 1;
          mbar_wait(acc1full(b), ((uint32_t)(gch >> 1)) & 1u);
          mbar_wait(hpempty(b), (((uint32_t)(gch >> 1)) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 128 + mt * 64);
          const uint32_t hrow = hps + (uint32_t)b * LB_PATCH_STRIDE + (uint32_t)prow * 128u;
#pragma unroll 1
          for (int cc = 0; cc < 64; cc += 32) {
            uint32_t v[32];
            tmem_ld32(tacc + (uint32_t)cc, v);
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            const float4* b4 = reinterpret_cast<const float4*>(g.b1 + j * 64 + cc);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 bb = __ldg(b4 + i);
              f[4 * i] += bb.x; f[4 * i + 1] += bb.y; f[4 * i + 2] += bb.z; f[4 * i + 3] += bb.w;
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float2 gq = PRECISE ? gelu_erf2(make_float2(f[2 * i], f[2 * i + 1])) : gelu_tanh2_half_arg(make_float2(f[2 * i], f[2 * i + 1]));
              f[2 * i] = inimg ? gq.x : 0.f;                       // the conv's zero padding acts on the hidden tensor
              f[2 * i + 1] = inimg ? gq.y : 0.f;
            }
            if (prow < LB_PROWS) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                uint32_t w0_, w1_, w2_, w3_;
                pack8_16(f + 8 * i, g.f16 != 0, w0_, w1_, w2_, w3_);
                const uint32_t c16 = (uint32_t)(cc / 8 + i);
                st_shared_v4(hrow + ((c16 ^ (uint32_t)(prow & 7)) << 4), w0_, w1_, w2_, w3_);
              }
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(acc1empty(b));
            mbar_arrive(hpfull(b));
          }
        }
        // the previous tile's output epilogue runs once this tile's first chunks are in flight
        if (prev_tile >= 0 && j == (NJ > 1 ? 1 : 0)) { epilogue(prev_tile, lt - 1); prev_tile = -1; }
      }
      prev_tile = tile;
    }
    if (prev_tile >= 0) epilogue(prev_tile, lt - 1);
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    // ------------------------------------------------------------------------------------------ V warps: conv
    const int ct = (int)threadIdx.x - 32 * LB_V0;
    const int cg = ct & 15, strip = ct >> 4;                   // 4 channels x one 8-pixel tile row
    const int m_first = strip * 8;
    int gch = 0;
    for (int tile = blockIdx.x; tile < g.m_tiles; tile += grid) {
      for (int j = 0; j < NJ; ++j, ++gch) {
        const int b = gch & 1;
        const int c = j * 64 + cg * 4;
        float2 wreg[9][2], bz[2];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(g.dw_w + (size_t)t * (4 * C) + c));
          wreg[t][0] = make_float2(a.x, a.y);
          wreg[t][1] = make_float2(a.z, a.w);
        }
        {
          const float4 a = __ldg(reinterpret_cast<const float4*>(g.dw_b + c));
          bz[0] = make_float2(a.x, a.y);
          bz[1] = make_float2(a.z, a.w);
        }
        mbar_wait(hpfull(b), ((uint32_t)(gch >> 1)) & 1u);
        mbar_wait(a2empty(b), (((uint32_t)(gch >> 1)) & 1u) ^ 1u);
        const uint8_t* hp = smem_raw + (hps + (uint32_t)b * LB_PATCH_STRIDE - raw);
        uint8_t* ab = smem_raw + (a2s + (uint32_t)b * LB_A2_BYTES - raw);
        float2 win[3][3][2];                                   // [column mod 3][dy][channel pair]
        auto load_col = [&](int jx, float2 (&dst)[3][2]) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const int p = (strip + dy) * LB_PW + jx;           // patch pixel
            const uint2 u = *reinterpret_cast<const uint2*>(hp + p * 128 + ((((uint32_t)cg >> 1) ^ (uint32_t)(p & 7)) << 4) + (cg & 1) * 8);
            if (g.f16) { dst[dy][0] = unpack2_16<true>(u.x); dst[dy][1] = unpack2_16<true>(u.y); }
            else { dst[dy][0] = unpack2_16<false>(u.x); dst[dy][1] = unpack2_16<false>(u.y); }
          }
        };
        load_col(0, win[0]);
        load_col(1, win[1]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          load_col(i + 2, win[(i + 2) % 3]);
          float2 a0 = bz[0], a1 = bz[1];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              a0 = __ffma2_rn(win[(i + dx) % 3][dy][0], wreg[dy * 3 + dx][0], a0);
              a1 = __ffma2_rn(win[(i + dx) % 3][dy][1], wreg[dy * 3 + dx][1], a1);
            }
          if (PRECISE) { a0 = gelu_erf2(a0); a1 = gelu_erf2(a1); }
          else { a0 = gelu_tanh2_half_arg(a0); a1 = gelu_tanh2_half_arg(a1); }
          const int m = m_first + i;
          uint2 o;
          if (g.f16) { o.x = pack2_f16(a0.x, a0.y); o.y = pack2_f16(a1.x, a1.y); }
          else { o.x = pack2_bf16(a0.x, a0.y); o.y = pack2_bf16(a1.x, a1.y); }
          *reinterpret_cast<uint2*>(ab + m * 128 + ((((uint32_t)cg >> 1) ^ (uint32_t)(m & 7)) << 4) + (cg & 1) * 8) = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(a2full(b));
          mbar_arrive(hpempty(b));
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int C, bool PRECISE>
int launch_leff_block(const void* A, const void* W1, const void* W2, const float* b1, const float* dw_w, const float* dw_b,
                      const float* b2, float* x, int n, int H, int f16, cudaStream_t st) {
  LbGeom g;
  g.H = H; g.tiles_w = H / 8; g.tiles_per_img = (H / 8) * (H / 16); g.m_tiles = n * g.tiles_per_img;
  g.C = C; g.KC = C >= 64 ? C / 64 : 1; g.ksteps = C >= 64 ? 4 : 2; g.NJ = 4 * C / 64;
  g.f16 = f16; g.b1 = b1; g.dw_w = dw_w; g.dw_b = dw_b; g.b2 = b2; g.resid = x; g.M = n * H * H;
  constexpr int WT = PRECISE ? 2 : 1;
  const int ap_bytes = g.KC * (int)LB_PATCH_STRIDE, w1_bytes = g.KC * WT * 8192, w2_bytes = WT * C * 128;
  const int fixed = 1024 + 2 * (int)LB_PATCH_STRIDE + 2 * (int)LB_A2_BYTES + LB_NG * (int)LB_STG_BYTES + 512;
  g.ap_st = 2; g.w_st = 2;
  auto total = [&]() { return fixed + g.ap_st * ap_bytes + g.w_st * (w1_bytes + w2_bytes); };
  if (total() > 227 * 1024) g.ap_st = 1;
  if (total() > 227 * 1024) g.w_st = 1;
  WMK_REQUIRE(total() <= 227 * 1024, "leff_block: %d bytes of shared memory needed (C = %d)", total(), C);
  CUtensorMap tmA, tmW1, tmW2, tmC;
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
  {   // x as [clip*H + h][w][C] fp32: one epilogue warp stores 4 image rows x 8 pixels x 32 channels
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)H, (uint64_t)n * H};
    const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)H * C * 4};
    const uint32_t box[3] = {32, 8, 4};
    WMK_TRY(make_tensor_map(&tmC, x, 3, dims, strides, box, true, 128));
  }
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(leff_block_kernel<C, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = g.m_tiles < num_sms() ? g.m_tiles : num_sms();
  leff_block_kernel<C, PRECISE><<<grid, LB_THREADS, (size_t)total(), st>>>(tmA, tmW1, tmW2, tmC, g);
  WMK_CHECK_LAUNCH("leff_block_kernel");
  return 0;
}

}  // namespace

// x[M][C] += LeFF(A) with A = LayerNorm-2 output [n][H][H][C] (16-bit: bf16, or fp16 when f16), token layout.
//   plain   (precise = 0): W1 [4C][C], W2 [C][4C] 16-bit; b1, dw_w [9][4C], dw_b PRE-HALVED (tanh-form GELU)
//   precise (precise = 1): W1 [4C][hi(C) | lo(C)], W2 [C][hi(4C) | lo(4C)] fp16 pairs; plain b1 / dw_w / dw_b, erf-form GELU
// C in {32, 64, 128}, H a power of two >= 16.
int leff_block(const void* A, const void* W1, const void* W2, const float* b1, const float* dw_w, const float* dw_b,
               const float* b2, float* x, int n, int H, int C, int f16, int precise, cudaStream_t st) {
  WMK_REQUIRE(H >= 16 && H <= 128 && (H & (H - 1)) == 0, "leff_block: H=%d must be a power of two in [16,128]", H);
  WMK_REQUIRE(C == 32 || C == 64 || C == 128, "leff_block: covers C in {32,64,128}, got %d", C);
  WMK_REQUIRE(!precise || f16, "leff_block: the precise form works on fp16 tensors");
  const double M = (double)n * H * H;
  // algorithmic traffic: A 2C + residual 4C + x 4C bytes per token (+ weights); FLOPs: the two dense layers
  ProfScope prof(FAM_GEMM_HBM, M * C * 10 + 16.0 * C * C, st, 2.0 * M * C * 4 * C * 2);
  if (precise) {
    switch (C) {
      case 32: return launch_leff_block<32, true>(A, W1, W2, b1, dw_w, dw_b, b2, x, n, H, f16, st);
      case 64: return launch_leff_block<64, true>(A, W1, W2, b1, dw_w, dw_b, b2, x, n, H, f16, st);
      default: return launch_leff_block<128, true>(A, W1, W2, b1, dw_w, dw_b, b2, x, n, H, f16, st);
    }
  }
  switch (C) {
    case 32: return launch_leff_block<32, false>(A, W1, W2, b1, dw_w, dw_b, b2, x, n, H, f16, st);
    case 64: return launch_leff_block<64, false>(A, W1, W2, b1, dw_w, dw_b, b2, x, n, H, f16, st);
    default: return launch_leff_block<128, false>(A, W1, W2, b1, dw_w, dw_b, b2, x, n, H, f16, st);
  }
}

}  // namespace wmk

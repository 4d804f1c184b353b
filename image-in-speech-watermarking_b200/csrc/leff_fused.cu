// Fused tail of the LeFF block (uformerWM/model.py:688-690,706-713):
//     x += Linear2( GELU( DepthwiseConv3x3( H1 ) ) )
// as ONE persistent tcgen05 kernel: the depthwise convolution is the PRODUCER of the GEMM's A operand, so
// the convolved hidden tensor (8C bytes per token written + read again by the separate kernels) never
// exists in HBM.
//
//   warp 0      TMA producer: per k-block (64 hidden channels) one 4-D box of H1 - the 128 tokens of the
//               M tile plus a one-pixel halo, zero padding = out-of-bounds fill - into a 2-deep patch ring,
//               and the [C x 64] slice of W2 into the operand stage
//   warps 2-9   convolution: thread = 4 channels x 8 consecutive pixels, 3x3 register window sliding along
//               the row, FFMA2 + MUFU.TANH GELU, result written as bf16 straight into the operand stage in
//               the 128-byte-swizzled K-major layout UMMA expects (fence.proxy.async, mbarrier arrive)
//   warp 1      MMA issuer: tcgen05.mma M128 x N=C x K16 into one of two TMEM accumulators
//   warps 10-17 epilogue: tcgen05.ld -> + bias + fp32 residual -> swizzled staging -> TMA store of x
//
// An M tile is a 16 x 8 pixel block of one image (two whole images when H = 8): the patch is 18 x 10 pixels
// (1.4x halo, served by L2) whatever the resolution, GEMM row m = 8 r + c is pixel (h0 + r, w0 + c), and the
// epilogue stores each warp's 32 rows = 4 image rows x 8 pixels with one 3-D TMA box over x seen as
// [clip*H + h][w][C].
#include "tc_ptx.cuh"

namespace wmk {

namespace {

using namespace tc;

constexpr int kConvWarps = 8, kEpiW = 8;
constexpr int kFThreads = 64 + 32 * kConvWarps + 32 * kEpiW;      // 576
constexpr int P_MAX = 4;                                          // patch ring depth (runtime: 2..4, what fits)
constexpr uint32_t F_STG_BYTES = 4096;                            // fp32 staging box: 32 rows x 128 B

struct LeffGeom {
  int H, lgH;          // image side (power of two, 8..128)
  int tiles_w, tiles_per_img;   // 8-wide, 16-high tiles (H >= 16); H = 8: one tile = two images
  int PW, PH, imgs;    // patch: PW = 10, PH = 18 (10 when H = 8), images per tile
  uint32_t patch_bytes;
  const float* dw_w;   // [9][Ch] tap-major
  const float* dw_b;   // [Ch]
  int Ch;              // 4C
};

__device__ __forceinline__ float2 bf2f(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

template <int BN>
__global__ void __launch_bounds__(kFThreads, 1)
leff_dwconv_linear2_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmW,
                           const __grid_constant__ CUtensorMap tmC, EpiParams p, LeffGeom g, int K, int m_tiles,
                           int n_stages, int P_ST) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t W_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE = A_BYTES + W_BYTES;
  const int kblocks = K / BK;
  const uint32_t stages = base;
  const uint32_t patches = stages + (uint32_t)n_stages * STAGE;                   // [P_ST][patch_bytes (1024-multiple)]
  const uint32_t pstride = (g.patch_bytes + 1023u) & ~1023u;
  const uint32_t staging = patches + (uint32_t)P_ST * pstride;                     // [8 warps][4096]
  const uint32_t bars = staging + kEpiW * F_STG_BYTES;
  auto wfull_bar = [&](int s) { return bars + 8u * s; };
  auto afull_bar = [&](int s) { return bars + 8u * (8 + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (16 + s); };
  auto pfull_bar = [&](int q) { return bars + 8u * (24 + q); };
  auto pempty_bar = [&](int q) { return bars + 8u * (28 + q); };
  const uint32_t tfull_bar = bars + 8u * 32;      // [2]
  const uint32_t tempty_bar = bars + 8u * 34;     // [2]
  const uint32_t tmem_slot = bars + 8u * 36;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - raw));

  constexpr int CPW = BN >= 64 ? BN / 2 : 32;     // accumulator columns per epilogue warp
  constexpr int SLABS = BN / CPW;
  constexpr int ACTIVE_EPI = SLABS * 4;
  constexpr int TCOLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(afull_bar(s), kConvWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int q = 0; q < P_MAX; ++q) {
      mbar_init(pfull_bar(q), 1);
      mbar_init(pempty_bar(q), kConvWarps);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8u * a, 1);
      mbar_init(tempty_bar + 8u * a, ACTIVE_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // tile -> (first image, first row, first column)
  auto tile_at = [&](int tile, int& b, int& h0, int& w0) {
    if (g.imgs == 2) { b = tile * 2; h0 = 0; w0 = 0; return; }
    b = tile / g.tiles_per_img;
    const int rem = tile - b * g.tiles_per_img;
    const int th = rem / g.tiles_w;
    h0 = th * 16;
    w0 = (rem - th * g.tiles_w) * 8;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producers (two independent lanes:
    // the patch ring may run further ahead than the operand stages)
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid) {
        int img, h0, w0;
        tile_at(tile, img, h0, w0);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int q = it % P_ST;
          mbar_wait(pempty_bar(q), (((uint32_t)(it / P_ST)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(pfull_bar(q), g.patch_bytes);
          tma_load_4d(patches + (uint32_t)q * pstride, &tmP, kb * BK, w0 - 1, h0 - 1, img, pfull_bar(q));
        }
      }
    } else if (lane == 16) {
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid) {
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % n_stages;
          mbar_wait(empty_bar(s), (((uint32_t)(it / n_stages)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(wfull_bar(s), W_BYTES);
          tma_load_2d(stages + (uint32_t)s * STAGE + A_BYTES, &tmW, kb * BK, 0, wfull_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BN);
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid, ++lt) {
        const int acc = lt & 1;
        mbar_wait(tempty_bar + 8u * acc, ((uint32_t)(lt >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
          mbar_wait(wfull_bar(s), ph);
          mbar_wait(afull_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t sa = stages + (uint32_t)s * STAGE;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          tcgen05_commit(empty_bar(s));
        }
        tcgen05_commit(tfull_bar + 8u * acc);
      }
    }
  } else if (warp < 2 + kConvWarps) {
    // ------------------------------------------------------------------ depthwise conv + GELU -> A operand
    const int ct = (int)threadIdx.x - 64;
    const int cg = ct & 15, strip = ct >> 4;                   // 4 channels x one 8-pixel tile row
    const int m_first = strip * 8;
    const int pimg = g.imgs == 2 ? strip >> 3 : 0, prow = g.imgs == 2 ? strip & 7 : strip;
    const uint32_t poff = (uint32_t)(((pimg * g.PH + prow) * g.PW) * 128 + cg * 8);
    const uint32_t rowb = (uint32_t)g.PW * 128u;
    int it = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += grid) {
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % n_stages, q = it % P_ST;
        const int c = kb * BK + cg * 4;
        float2 wreg[9][2], bz[2];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(g.dw_w + (size_t)t * g.Ch + c));
          wreg[t][0] = make_float2(a.x, a.y);
          wreg[t][1] = make_float2(a.z, a.w);
        }
        {
          const float4 a = __ldg(reinterpret_cast<const float4*>(g.dw_b + c));
          bz[0] = make_float2(a.x, a.y);
          bz[1] = make_float2(a.z, a.w);
        }
        mbar_wait(pfull_bar(q), ((uint32_t)(it / P_ST)) & 1u);
        mbar_wait(empty_bar(s), (((uint32_t)(it / n_stages)) & 1u) ^ 1u);
        const uint8_t* pb = smem_raw + (patches + (uint32_t)q * pstride - raw) + poff;
        uint8_t* ab = smem_raw + (stages + (uint32_t)s * STAGE - raw);
        float2 win[3][3][2];                                   // [column mod 3][dy][channel pair]
        auto load_col = [&](int j, float2 (&dst)[3][2]) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const uint2 u = *reinterpret_cast<const uint2*>(pb + dy * rowb + j * 128);
            dst[dy][0] = bf2f(u.x);
            dst[dy][1] = bf2f(u.y);
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
          a0 = gelu_tanh2(a0);
          a1 = gelu_tanh2(a1);
          const int m = m_first + i;
          uint2 o;
          o.x = pack_bf16x2(a0.x, a0.y);
          o.y = pack_bf16x2(a1.x, a1.y);
          *reinterpret_cast<uint2*>(ab + m * 128 + ((((uint32_t)cg >> 1) ^ (uint32_t)(m & 7)) << 4) + (cg & 1) * 8) = o;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(afull_bar(s));
          mbar_arrive(pempty_bar(q));
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: + bias + residual -> x
    const int ew = warp - (2 + kConvWarps);
    const int q = warp & 3;
    const int slab = ew >> 2;
    if (slab < SLABS) {
      const uint32_t buf = staging + (uint32_t)ew * F_STG_BYTES;
      int lt = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid, ++lt) {
        const int acc = lt & 1;
        int tb, h0, w0;
        tile_at(tile, tb, h0, w0);
        // GEMM row m = q*32 + lane is pixel (h0 + m/8, w0 + m%8); with H = 8 the rows run on into image tb + 1
        const int grow = tb * g.H + h0 + q * 4;                       // first row of this warp in the [clip*H + h] view
        const int row = (grow + (lane >> 3)) * g.H + w0 + (lane & 7); // token index
        float4 rpre[8];
        if (row < p.M) {
          const float4* r4 = reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ldc + slab * CPW);
#pragma unroll
          for (int j = 0; j < 8; ++j) rpre[j] = r4[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) rpre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(tfull_bar + 8u * acc, (uint32_t)(lt >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + slab * CPW);
#pragma unroll 1
        for (int cc = 0; cc < CPW; cc += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)cc, v);
          const int n = slab * CPW + cc;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          if (cc == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[4 * j] += rpre[j].x; f[4 * j + 1] += rpre[j].y; f[4 * j + 2] += rpre[j].z; f[4 * j + 3] += rpre[j].w;
            }
          } else if (row < p.M) {
            const float4* r4 = reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ldc + n);
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
        if (lane == 0) mbar_arrive(tempty_bar + 8u * acc);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
  }
}

// ---------------------------------------------------------------------------------------------
// Split-bf16 variant (the WMK_PREC_MIXED extractor): H1 is fp32 (the linear1 GEMM of a split plan writes
// fp32), every product is hi*hi + lo*hi + hi*lo.  A k-block = 32 hidden channels:
//   patch : TMA 4-D box {32 ch, 10, 18, imgs} of fp32 - the same 128-byte pixel rows as the bf16 kernel;
//   conv  : 256 threads = 8 channel groups (4 ch) x 32 half rows (4 consecutive pixels): fp32 taps, erf-form GELU
//           (gelu_fast), the result split into hi / lo and written as ONE 64-wide swizzled A row [hi(32) | lo(32)];
//   MMA   : that A tile against W tile 0 = [W2_hi | W2_hi] (4 k-steps) and W tile 1 = [W2_lo | 0] (2 k-steps),
//           all into the same fp32 TMEM accumulator (the K = 32 form of gemm_tcgen05.cu, p.split = 2);
//   epilogue: as above (+ bias + fp32 residual, TMA store of x).
// W2t: [C][(4C/32) * 128] bf16, k-block kb at columns kb*128: [hi(32) | hi(32) | lo(32) | 0(32)] (uformer_plan.cu).
// ---------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(kFThreads, 1)
leff_tail_split_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmC, EpiParams p, LeffGeom g, int kblocks, int m_tiles,
                       int n_stages, int P_ST) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t W_TILE = BN * BK * 2;
  constexpr uint32_t W_BYTES = 2 * W_TILE;
  constexpr uint32_t STAGE = A_BYTES + W_BYTES;
  const uint32_t stages = base;
  const uint32_t patches = stages + (uint32_t)n_stages * STAGE;
  const uint32_t pstride = (g.patch_bytes + 1023u) & ~1023u;
  const uint32_t staging = patches + (uint32_t)P_ST * pstride;
  const uint32_t bars = staging + kEpiW * F_STG_BYTES;
  auto wfull_bar = [&](int s) { return bars + 8u * s; };
  auto afull_bar = [&](int s) { return bars + 8u * (8 + s); };
  auto empty_bar = [&](int s) { return bars + 8u * (16 + s); };
  auto pfull_bar = [&](int q) { return bars + 8u * (24 + q); };
  auto pempty_bar = [&](int q) { return bars + 8u * (28 + q); };
  const uint32_t tfull_bar = bars + 8u * 32;      // [2]
  const uint32_t tempty_bar = bars + 8u * 34;     // [2]
  const uint32_t tmem_slot = bars + 8u * 36;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - raw));

  constexpr int CPW = BN >= 64 ? BN / 2 : 32;
  constexpr int SLABS = BN / CPW;
  constexpr int ACTIVE_EPI = SLABS * 4;
  constexpr int TCOLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(wfull_bar(s), 1);
      mbar_init(afull_bar(s), kConvWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int q = 0; q < P_MAX; ++q) {
      mbar_init(pfull_bar(q), 1);
      mbar_init(pempty_bar(q), kConvWarps);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8u * a, 1);
      mbar_init(tempty_bar + 8u * a, ACTIVE_EPI);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  auto tile_at = [&](int tile, int& b, int& h0, int& w0) {
    if (g.imgs == 2) { b = tile * 2; h0 = 0; w0 = 0; return; }
    b = tile / g.tiles_per_img;
    const int rem = tile - b * g.tiles_per_img;
    const int th = rem / g.tiles_w;
    h0 = th * 16;
    w0 = (rem - th * g.tiles_w) * 8;
  };

  if (warp == 0) {
    if (lane == 0) {                      // patch ring
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid) {
        int img, h0, w0;
        tile_at(tile, img, h0, w0);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int q = it % P_ST;
          mbar_wait(pempty_bar(q), (((uint32_t)(it / P_ST)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(pfull_bar(q), g.patch_bytes);
          tma_load_4d(patches + (uint32_t)q * pstride, &tmP, kb * 32, w0 - 1, h0 - 1, img, pfull_bar(q));
        }
      }
    } else if (lane == 16) {              // the two W2 tiles of every k-block
      int it = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid) {
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % n_stages;
          mbar_wait(empty_bar(s), (((uint32_t)(it / n_stages)) & 1u) ^ 1u);
          mbar_arrive_expect_tx(wfull_bar(s), W_BYTES);
          const uint32_t wb = stages + (uint32_t)s * STAGE + A_BYTES;
          tma_load_2d(wb, &tmW, kb * 128, 0, wfull_bar(s));
          tma_load_2d(wb + W_TILE, &tmW, kb * 128 + 64, 0, wfull_bar(s));
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BN);           // bf16 operands
      int it = 0, lt = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid, ++lt) {
        const int acc = lt & 1;
        mbar_wait(tempty_bar + 8u * acc, ((uint32_t)(lt >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
          mbar_wait(wfull_bar(s), ph);
          mbar_wait(afull_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t sa = stages + (uint32_t)s * STAGE;
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t b0 = umma_desc_sw128(sa + A_BYTES);
          const uint64_t b1 = umma_desc_sw128(sa + A_BYTES + W_TILE);
#pragma unroll
          for (int k = 0; k < 4; ++k)       // [hi | lo] x [W_hi | W_hi]
            tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
#pragma unroll
          for (int k = 0; k < 2; ++k)       // hi x W_lo
            tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), b1 + (uint64_t)(2 * k), idesc, 1u);
          tcgen05_commit(empty_bar(s));
        }
        tcgen05_commit(tfull_bar + 8u * acc);
      }
    }
  } else if (warp < 2 + kConvWarps) {
    // ------------------------------------------------------------------ depthwise conv + GELU -> split A operand
    const int ct = (int)threadIdx.x - 64;
    const int cg = ct & 7, hs = ct >> 3;                        // 4 channels x 4 consecutive pixels of tile row hs >> 1
    const int row = hs >> 1, half = hs & 1;
    const int m_first = row * 8 + half * 4;
    const int pimg = g.imgs == 2 ? row >> 3 : 0, prow = g.imgs == 2 ? row & 7 : row;
    const uint32_t poff = (uint32_t)(((pimg * g.PH + prow) * g.PW + half * 4) * 128 + cg * 16);
    const uint32_t rowb = (uint32_t)g.PW * 128u;
    int it = 0;
    for (int tile = blockIdx.x; tile < m_tiles; tile += grid) {
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % n_stages, q = it % P_ST;
        const int c = kb * 32 + cg * 4;
        float2 wreg[9][2], bz[2];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(g.dw_w + (size_t)t * g.Ch + c));
          wreg[t][0] = make_float2(a.x, a.y);
          wreg[t][1] = make_float2(a.z, a.w);
        }
        {
          const float4 a = __ldg(reinterpret_cast<const float4*>(g.dw_b + c));
          bz[0] = make_float2(a.x, a.y);
          bz[1] = make_float2(a.z, a.w);
        }
        mbar_wait(pfull_bar(q), ((uint32_t)(it / P_ST)) & 1u);
        mbar_wait(empty_bar(s), (((uint32_t)(it / n_stages)) & 1u) ^ 1u);
        const uint8_t* pb = smem_raw + (patches + (uint32_t)q * pstride - raw) + poff;
        uint8_t* ab = smem_raw + (stages + (uint32_t)s * STAGE - raw);
        float2 win[3][3][2];                                   // [column mod 3][dy][channel pair]
        auto load_col = [&](int j, float2 (&dst)[3][2]) {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
            const float4 u = *reinterpret_cast<const float4*>(pb + dy * rowb + j * 128);
            dst[dy][0] = make_float2(u.x, u.y);
            dst[dy][1] = make_float2(u.z, u.w);
          }
        };
        load_col(0, win[0]);
        load_col(1, win[1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          load_col(i + 2, win[(i + 2) % 3]);
          float2 a0 = bz[0], a1 = bz[1];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              a0 = __ffma2_rn(win[(i + dx) % 3][dy][0], wreg[dy * 3 + dx][0], a0);
              a1 = __ffma2_rn(win[(i + dx) % 3][dy][1], wreg[dy * 3 + dx][1], a1);
            }
          const int m = m_first + i;
          uint2 hi, lo;
          split_pack2(gelu_fast(a0.x), gelu_fast(a0.y), hi.x, lo.x);
          split_pack2(gelu_fast(a1.x), gelu_fast(a1.y), hi.y, lo.y);
          // A row m = 64 bf16 = [hi(32) | lo(32)]: channel group cg -> 16-byte chunk cg >> 1 (hi) / 4 + (cg >> 1) (lo)
          uint8_t* rowp = ab + m * 128 + (cg & 1) * 8;
          *reinterpret_cast<uint2*>(rowp + ((((uint32_t)cg >> 1) ^ (uint32_t)(m & 7)) << 4)) = hi;
          *reinterpret_cast<uint2*>(rowp + (((4u + ((uint32_t)cg >> 1)) ^ (uint32_t)(m & 7)) << 4)) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(afull_bar(s));
          mbar_arrive(pempty_bar(q));
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: + bias + residual -> x
    const int ew = warp - (2 + kConvWarps);
    const int q = warp & 3;
    const int slab = ew >> 2;
    if (slab < SLABS) {
      const uint32_t buf = staging + (uint32_t)ew * F_STG_BYTES;
      int lt = 0;
      for (int tile = blockIdx.x; tile < m_tiles; tile += grid, ++lt) {
        const int acc = lt & 1;
        int tb, h0, w0;
        tile_at(tile, tb, h0, w0);
        const int grow = tb * g.H + h0 + q * 4;
        const int rowi = (grow + (lane >> 3)) * g.H + w0 + (lane & 7);
        float4 rpre[8];
        if (rowi < p.M) {
          const float4* r4 = reinterpret_cast<const float4*>(p.resid + (size_t)rowi * p.ldc + slab * CPW);
#pragma unroll
          for (int j = 0; j < 8; ++j) rpre[j] = r4[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) rpre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(tfull_bar + 8u * acc, (uint32_t)(lt >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + slab * CPW);
#pragma unroll 1
        for (int cc = 0; cc < CPW; cc += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)cc, v);
          const int n = slab * CPW + cc;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(b4 + j);
              f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
            }
          }
          if (cc == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              f[4 * j] += rpre[j].x; f[4 * j + 1] += rpre[j].y; f[4 * j + 2] += rpre[j].z; f[4 * j + 3] += rpre[j].w;
            }
          } else if (rowi < p.M) {
            const float4* r4 = reinterpret_cast<const float4*>(p.resid + (size_t)rowi * p.ldc + n);
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
        if (lane == 0) mbar_arrive(tempty_bar + 8u * acc);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TCOLS));
  }
}

template <int BN>
int launch_leff_split(const float* H1, const float* dw_w, const float* dw_b, const __nv_bfloat16* W2t, const float* b2,
                      float* x, int n, int H, int C, cudaStream_t st) {
  const int M = n * H * H, K = 4 * C;
  LeffGeom g;
  g.H = H;
  g.lgH = 0;
  while ((1 << g.lgH) < H) ++g.lgH;
  g.PW = 10;
  if (H == 8) { g.PH = 10; g.imgs = 2; g.tiles_w = 1; g.tiles_per_img = 1; }
  else { g.PH = 18; g.imgs = 1; g.tiles_w = H / 8; g.tiles_per_img = (H / 8) * (H / 16); }
  g.patch_bytes = (uint32_t)(128 * g.PW * g.PH * g.imgs);
  g.dw_w = dw_w; g.dw_b = dw_b; g.Ch = K;
  const int kblocks = K / 32;
  CUtensorMap tmP, tmW, tmC;
  {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)H, (uint64_t)H, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)K * 4, (uint64_t)H * K * 4, (uint64_t)H * H * K * 4};
    const uint32_t box[4] = {32, (uint32_t)g.PW, (uint32_t)g.PH, (uint32_t)g.imgs};
    WMK_TRY(make_tensor_map(&tmP, H1, 4, dims, strides, box, true, 0));
  }
  {
    const uint64_t dims[2] = {(uint64_t)kblocks * 128, (uint64_t)C};
    const uint64_t strides[1] = {(uint64_t)kblocks * 128 * 2};
    const uint32_t box[2] = {64, (uint32_t)BN};
    WMK_TRY(make_tensor_map(&tmW, W2t, 2, dims, strides, box, false, 128));
  }
  {
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)H, (uint64_t)n * H};
    const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)H * C * 4};
    const uint32_t box[3] = {32, 8, 4};
    WMK_TRY(make_tensor_map(&tmC, x, 3, dims, strides, box, true, 128));
  }
  const int m_tiles = H == 8 ? (n + 1) / 2 : n * g.tiles_per_img;
  const uint32_t pstride = (g.patch_bytes + 1023u) & ~1023u;
  const int stage = BM * BK * 2 + 2 * BN * BK * 2;
  int n_stages = 2;
  int n_patch = (226 * 1024 - 1024 - 512 - kEpiW * (int)F_STG_BYTES - n_stages * stage) / (int)pstride;
  if (n_patch > P_MAX) n_patch = P_MAX;
  WMK_REQUIRE(n_patch >= 2, "leff(split): not enough shared memory for C=%d H=%d", C, H);
  const int fixed = 1024 + n_patch * (int)pstride + kEpiW * (int)F_STG_BYTES + 512;
  while (n_stages < 4 && fixed + (n_stages + 1) * stage <= 226 * 1024) ++n_stages;
  const size_t smem = (size_t)fixed + (size_t)n_stages * stage;
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(leff_tail_split_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  EpiParams p{b2, x, x, M, C, C, EPI_BIAS_RESID, 0, 0, 0, 0};
  const int grid = m_tiles < num_sms() ? m_tiles : num_sms();
  leff_tail_split_kernel<BN><<<grid, kFThreads, smem, st>>>(tmP, tmW, tmC, p, g, kblocks, m_tiles, n_stages, n_patch);
  WMK_CHECK_LAUNCH("leff_tail_split_kernel");
  return 0;
}

template <int BN>
int launch_leff(const __nv_bfloat16* H1, const float* dw_w, const float* dw_b, const __nv_bfloat16* W2, const float* b2,
                float* x, int n, int H, int C, cudaStream_t st) {
  const int M = n * H * H, K = 4 * C;
  LeffGeom g;
  g.H = H;
  g.lgH = 0;
  while ((1 << g.lgH) < H) ++g.lgH;
  g.PW = 10;
  if (H == 8) { g.PH = 10; g.imgs = 2; g.tiles_w = 1; g.tiles_per_img = 1; }
  else { g.PH = 18; g.imgs = 1; g.tiles_w = H / 8; g.tiles_per_img = (H / 8) * (H / 16); }
  g.patch_bytes = (uint32_t)(128 * g.PW * g.PH * g.imgs);
  g.dw_w = dw_w; g.dw_b = dw_b; g.Ch = K;
  CUtensorMap tmP, tmW, tmC;
  {
    const uint64_t dims[4] = {(uint64_t)K, (uint64_t)H, (uint64_t)H, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)K * 2, (uint64_t)H * K * 2, (uint64_t)H * H * K * 2};
    const uint32_t box[4] = {64, (uint32_t)g.PW, (uint32_t)g.PH, (uint32_t)g.imgs};
    WMK_TRY(make_tensor_map(&tmP, H1, 4, dims, strides, box, false, 0));
  }
  {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)C};
    const uint64_t strides[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)BN};
    WMK_TRY(make_tensor_map(&tmW, W2, 2, dims, strides, box, false, 128));
  }
  {   // x as [clip*H + h][w][C]: one epilogue warp stores 4 image rows x 8 pixels x 32 channels
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)H, (uint64_t)n * H};
    const uint64_t strides[2] = {(uint64_t)C * 4, (uint64_t)H * C * 4};
    const uint32_t box[3] = {32, 8, 4};
    WMK_TRY(make_tensor_map(&tmC, x, 3, dims, strides, box, true, 128));
  }
  const int m_tiles = H == 8 ? (n + 1) / 2 : n * g.tiles_per_img;
  const uint32_t pstride = (g.patch_bytes + 1023u) & ~1023u;
  const int stage = (BM + BN) * BK * 2;
  // 2 operand stages are enough (the convolution, not the MMA, is the slow consumer); the rest of the shared
  // memory goes to the patch ring, whose depth hides the HBM latency of the patch loads
  int n_stages = 2;
  int n_patch = (226 * 1024 - 1024 - 512 - kEpiW * (int)F_STG_BYTES - n_stages * stage) / (int)pstride;
  if (n_patch > P_MAX) n_patch = P_MAX;
  WMK_REQUIRE(n_patch >= 2, "leff: not enough shared memory for C=%d H=%d", C, H);
  const int fixed = 1024 + n_patch * (int)pstride + kEpiW * (int)F_STG_BYTES + 512;
  while (n_stages < 4 && fixed + (n_stages + 1) * stage <= 226 * 1024) ++n_stages;
  const size_t smem = (size_t)fixed + (size_t)n_stages * stage;
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(leff_dwconv_linear2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  EpiParams p{b2, x, x, M, C, C, EPI_BIAS_RESID, 0, 0, 0, 0};
  const int grid = m_tiles < num_sms() ? m_tiles : num_sms();
  leff_dwconv_linear2_kernel<BN><<<grid, kFThreads, smem, st>>>(tmP, tmW, tmC, p, g, K, m_tiles, n_stages, n_patch);
  WMK_CHECK_LAUNCH("leff_dwconv_linear2_kernel");
  return 0;
}

}  // namespace

// x[M][C] += Linear2(GELU(dwconv3x3(H1) + dw_b)) + b2, H1 [n][H][H][4C] bf16 (token layout), W2 [C][4C] bf16.
int leff_dwconv_linear2_bf16(const __nv_bfloat16* H1, const float* dw_w, const float* dw_b, const __nv_bfloat16* W2,
                             const float* b2, float* x, int n, int H, int C, cudaStream_t st) {
  WMK_REQUIRE(H >= 8 && H <= 128 && (H & (H - 1)) == 0, "leff: H=%d must be a power of two in [8,128]", H);
  WMK_REQUIRE(C == 32 || C == 64 || C == 128 || C == 256, "leff: fused path covers C in {32,64,128,256}, got %d", C);
  const double M = (double)n * H * H;
  ProfScope prof(FAM_GEMM_HBM, M * 4 * C * 2 + 8.0 * C * C + M * C * 8, st, 2.0 * M * C * 4 * C);
  switch (C) {
    case 32: return launch_leff<32>(H1, dw_w, dw_b, W2, b2, x, n, H, C, st);
    case 64: return launch_leff<64>(H1, dw_w, dw_b, W2, b2, x, n, H, C, st);
    case 128: return launch_leff<128>(H1, dw_w, dw_b, W2, b2, x, n, H, C, st);
    default: return launch_leff<256>(H1, dw_w, dw_b, W2, b2, x, n, H, C, st);
  }
}

// Split-bf16 tail: x[M][C] += Linear2(GELU(dwconv3x3(H1) + dw_b)) + b2 with H1 [n][H][H][4C] fp32 (token layout) and
// W2t the packed split weight [C][(4C/32)*128] (see leff_tail_split_kernel).  C in {32, 64, 128}.
int leff_dwconv_linear2_split(const float* H1, const float* dw_w, const float* dw_b, const __nv_bfloat16* W2t, const float* b2,
                              float* x, int n, int H, int C, cudaStream_t st) {
  WMK_REQUIRE(H >= 8 && H <= 128 && (H & (H - 1)) == 0, "leff(split): H=%d must be a power of two in [8,128]", H);
  WMK_REQUIRE(C == 32 || C == 64 || C == 128, "leff(split): fused path covers C in {32,64,128}, got %d", C);
  const double M = (double)n * H * H;
  ProfScope prof(FAM_GEMM_HBM, M * 4 * C * 4 + 16.0 * C * C + M * C * 8, st, 2.0 * M * C * 4 * C);
  switch (C) {
    case 32: return launch_leff_split<32>(H1, dw_w, dw_b, W2t, b2, x, n, H, C, st);
    case 64: return launch_leff_split<64>(H1, dw_w, dw_b, W2t, b2, x, n, H, C, st);
    default: return launch_leff_split<128>(H1, dw_w, dw_b, W2t, b2, x, n, H, C, st);
  }
}

}  // namespace wmk

// LeFF depthwise 3x3 convolution (pad 1) + GELU on the bf16 hidden tensor [B][H][W][Ch]
// (uformerWM/model.py:688-689,706) for sm_100a:
//   * persistent CTAs (4 per SM) walk (channel slab of 64, spatial tile of 8 x 16 pixels) items;
//   * the 10 x 18 x 64-channel input patch of the NEXT item is fetched by one TMA
//     cp.async.bulk.tensor.4d while the current one is convolved (2-stage mbarrier pipeline); the
//     conv's zero padding is the tensor map's out-of-bounds fill, so there is no border code;
//   * a thread owns 4 channels of TWO adjacent tile columns and walks down the patch rows in scatter form:
//     each input row (4 columns, loaded and unpacked once) is accumulated into the three output rows it
//     touches, a finished row goes through the GELU and out.  Per 4 outputs: 2 conflict-free 8-byte shared
//     loads, 8 unpack ops, 18 packed FFMA2, 2 packed GELUs (the kernel is instruction-issue / FMA-pipe
//     bound, not HBM bound: measured 57 % issue-active at 64 % of the HBM peak with one column per thread);
//   * the 0.5 of the GELU is folded into the (pre-halved) weights and bias: gelu_tanh2_half_arg.
// Algorithmic traffic: 4 B per element (bf16 in + out); the 1.4x patch halo is served by L2.
// H = 8 (the bottleneck stage) keeps the one-column kernel with 8 x 8 tiles.
#include "tc_ptx.cuh"

namespace wmk {

namespace {

using namespace tc;

constexpr int DW_TW = 8, DW_PW = DW_TW + 2;
constexpr int kDwThreads = 128;

struct DwGeom {
  int H, Ch, ncs, tiles_w, tiles_h, n_items;
};

template <int TH, bool F16, bool EXACT = false>
__global__ void __launch_bounds__(kDwThreads, 4)
dwconv3x3_gelu_tma_kernel(const __grid_constant__ CUtensorMap tm, uint16_t* __restrict__ out,
                          const float* __restrict__ wt, const float* __restrict__ bias, DwGeom g) {
  constexpr int PH = TH + 2;
  constexpr uint32_t kPatchBytes = PH * DW_PW * 128;
  __shared__ __align__(128) uint8_t patch[2][kPatchBytes];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x;
  const int cg = tid & 15, col = tid >> 4;

  auto decode = [&](int item, int& slab, int& b, int& h0, int& w0) {
    slab = item % g.ncs;
    const int sp = item / g.ncs;
    const int per = g.tiles_w * g.tiles_h;
    b = sp / per;
    const int rem = sp - b * per;
    const int th = rem / g.tiles_w;
    h0 = th * TH;
    w0 = (rem - th * g.tiles_w) * DW_TW;
  };
  auto issue = [&](int item, int stage) {
    int slab, b, h0, w0;
    decode(item, slab, b, h0, w0);
    mbar_arrive_expect_tx(smem_u32(&bar[stage]), kPatchBytes);
    tma_load_4d(smem_u32(patch[stage]), &tm, slab * 64, w0 - 1, h0 - 1, b, smem_u32(&bar[stage]));
  };

  if (tid == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int item = blockIdx.x;
  if (item >= g.n_items) return;
  if (tid == 0) issue(item, 0);

  for (int it = 0; item < g.n_items; item += gridDim.x, ++it) {
    const int stage = it & 1;
    const int next = item + gridDim.x;
    if (tid == 0 && next < g.n_items) issue(next, stage ^ 1);   // that buffer was drained before the last barrier

    int slab, b, h0, w0;
    decode(item, slab, b, h0, w0);
    const int c = slab * 64 + cg * 4;
    float2 wreg[9][2], bz[2];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(wt + (size_t)t * g.Ch + c));
      wreg[t][0] = make_float2(a.x, a.y);
      wreg[t][1] = make_float2(a.z, a.w);
    }
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(bias + c));
      bz[0] = make_float2(a.x, a.y);
      bz[1] = make_float2(a.z, a.w);
    }
    mbar_wait(smem_u32(&bar[stage]), (it >> 1) & 1);

    // patch[r][x][64 ch] bf16; this thread reads columns col..col+2, channels 4 cg..4 cg+3
    const uint8_t* pb = patch[stage] + col * 128 + cg * 8;
    float2 win[3][3][2];
    auto load_row = [&](int r, float2 (&dst)[3][2]) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint2 u = *reinterpret_cast<const uint2*>(pb + (r * DW_PW + dx) * 128);
        dst[dx][0] = unpack2_16<F16>(u.x);
        dst[dx][1] = unpack2_16<F16>(u.y);
      }
    };
    load_row(0, win[0]);
    load_row(1, win[1]);
    uint16_t* op = out + (((size_t)b * g.H + h0) * g.H + w0 + col) * g.Ch + c;
#pragma unroll
    for (int r = 0; r < TH; ++r) {
      load_row(r + 2, win[(r + 2) % 3]);
      float2 a0 = bz[0], a1 = bz[1];
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          a0 = __ffma2_rn(win[(r + dy) % 3][dx][0], wreg[dy * 3 + dx][0], a0);
          a1 = __ffma2_rn(win[(r + dy) % 3][dx][1], wreg[dy * 3 + dx][1], a1);
        }
      if constexpr (EXACT) {
        a0 = gelu_erf2(a0);
        a1 = gelu_erf2(a1);
      } else {
        a0 = gelu_tanh2_half_arg(a0);
        a1 = gelu_tanh2_half_arg(a1);
      }
      uint2 o;
      o.x = pack2_16<F16>(a0.x, a0.y);
      o.y = pack2_16<F16>(a1.x, a1.y);
      *reinterpret_cast<uint2*>(op + (size_t)r * g.H * g.Ch) = o;
    }
    __syncthreads();                                // everyone has drained this stage's patch
  }
}

template <int TH, bool F16, bool EXACT = false>
int launch_dw(const void* in, void* out, const float* wt, const float* bias, int B, int H, int Ch, cudaStream_t st) {
  CUtensorMap tm;
  const uint64_t dims[4] = {(uint64_t)Ch, (uint64_t)H, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)Ch * 2, (uint64_t)H * Ch * 2, (uint64_t)H * H * Ch * 2};
  const uint32_t box[4] = {64, DW_PW, TH + 2, 1};
  WMK_TRY(make_tensor_map(&tm, in, 4, dims, strides, box, false, 0));
  DwGeom g;
  g.H = H; g.Ch = Ch; g.ncs = Ch / 64; g.tiles_w = H / DW_TW; g.tiles_h = H / TH;
  const long long items = (long long)B * g.tiles_w * g.tiles_h * g.ncs;
  WMK_REQUIRE(items < (1LL << 31), "dwconv: too many tiles (%lld)", items);
  g.n_items = (int)items;
  const int grid = (int)(items < 4LL * num_sms() ? items : 4LL * num_sms());
  dwconv3x3_gelu_tma_kernel<TH, F16, EXACT><<<grid, kDwThreads, 0, st>>>(tm, reinterpret_cast<uint16_t*>(out), wt, bias, g);
  WMK_CHECK_LAUNCH("dwconv3x3_gelu_tma_kernel");
  return 0;
}

// ---------------------------------------------------------------------------------- two columns per thread
#ifndef DW2_SCALAR_FMA
#define DW2_SCALAR_FMA 0
#endif
constexpr int DW2_TW = 16, DW2_PW = DW2_TW + 2, DW2_TH = 8, DW2_PH = DW2_TH + 2;

struct DwGeom2 {
  int H, Ch, lg_ncs, lg_tw, lg_th, n_items;      // channel slabs, tiles per row / column: powers of two
};

// EXACT: erf-form GELU (gelu_fast, 1.5e-7) on un-halved weights / bias instead of the tanh form (the precise extractor).
template <bool F16, bool EXACT = false>
__global__ void __launch_bounds__(kDwThreads, 4)
dwconv3x3_gelu_tma2_kernel(const __grid_constant__ CUtensorMap tm, uint16_t* __restrict__ out,
                           const float* __restrict__ wt, const float* __restrict__ bias, DwGeom2 g) {
  constexpr uint32_t kPatchBytes = DW2_PH * DW2_PW * 128;
  __shared__ __align__(128) uint8_t patch[2][kPatchBytes];
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x;
  const int cg = tid & 15, cp = tid >> 4;          // 4 channels x columns 2 cp, 2 cp + 1

  auto decode = [&](int item, int& slab, int& b, int& h0, int& w0) {
    slab = item & ((1 << g.lg_ncs) - 1);
    const int sp = item >> g.lg_ncs;
    w0 = (sp & ((1 << g.lg_tw) - 1)) * DW2_TW;
    h0 = ((sp >> g.lg_tw) & ((1 << g.lg_th) - 1)) * DW2_TH;
    b = sp >> (g.lg_tw + g.lg_th);
  };
  auto issue = [&](int item, int stage) {
    int slab, b, h0, w0;
    decode(item, slab, b, h0, w0);
    mbar_arrive_expect_tx(smem_u32(&bar[stage]), kPatchBytes);
    tma_load_4d(smem_u32(patch[stage]), &tm, slab * 64, w0 - 1, h0 - 1, b, smem_u32(&bar[stage]));
  };

  if (tid == 0) {
    mbar_init(smem_u32(&bar[0]), 1);
    mbar_init(smem_u32(&bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int item = blockIdx.x;
  if (item >= g.n_items) return;
  if (tid == 0) issue(item, 0);

  for (int it = 0; item < g.n_items; item += gridDim.x, ++it) {
    const int stage = it & 1;
    const int next = item + gridDim.x;
    if (tid == 0 && next < g.n_items) issue(next, stage ^ 1);   // that buffer was drained before the last barrier

    int slab, b, h0, w0;
    decode(item, slab, b, h0, w0);
    const int c = slab * 64 + cg * 4;
    float2 wreg[9][2], bz[2];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(wt + (size_t)t * g.Ch + c));
      wreg[t][0] = make_float2(a.x, a.y);
      wreg[t][1] = make_float2(a.z, a.w);
    }
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(bias + c));
      bz[0] = make_float2(a.x, a.y);
      bz[1] = make_float2(a.z, a.w);
    }
    float2 acc[3][2][2];                           // [output row mod 3][column][channel pair]
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int q = 0; q < 2; ++q) { acc[s][q][0] = bz[0]; acc[s][q][1] = bz[1]; }
    uint16_t* op = out + (((size_t)b * g.H + h0) * g.H + w0 + 2 * cp) * g.Ch + c;
    mbar_wait(smem_u32(&bar[stage]), (it >> 1) & 1);

    // patch[r][x][64 ch] 16-bit; this thread reads columns 2 cp .. 2 cp + 3, channels 4 cg .. 4 cg + 3
    const uint8_t* pb = patch[stage] + cp * 256 + cg * 8;
#pragma unroll
    for (int i = 0; i < DW2_PH; ++i) {
      float2 in[4][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint2 u = *reinterpret_cast<const uint2*>(pb + (i * DW2_PW + j) * 128);
        in[j][0] = unpack2_16<F16>(u.x);
        in[j][1] = unpack2_16<F16>(u.y);
      }
      // input row i is tap row dy of output row i - dy (same tap order per output as the gather form)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int o = i - dy;
        if (o < 0 || o >= DW2_TH) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
#if DW2_SCALAR_FMA
            // scalar FFMA: 87 FMA / clk / SM measured against 75 for FFMA2 (tools/ubench/pipe_rates.cu) at twice the issue slots
            acc[o % 3][q][0].x = fmaf(in[q + dx][0].x, wreg[dy * 3 + dx][0].x, acc[o % 3][q][0].x);
            acc[o % 3][q][0].y = fmaf(in[q + dx][0].y, wreg[dy * 3 + dx][0].y, acc[o % 3][q][0].y);
            acc[o % 3][q][1].x = fmaf(in[q + dx][1].x, wreg[dy * 3 + dx][1].x, acc[o % 3][q][1].x);
            acc[o % 3][q][1].y = fmaf(in[q + dx][1].y, wreg[dy * 3 + dx][1].y, acc[o % 3][q][1].y);
#else
            acc[o % 3][q][0] = __ffma2_rn(in[q + dx][0], wreg[dy * 3 + dx][0], acc[o % 3][q][0]);
            acc[o % 3][q][1] = __ffma2_rn(in[q + dx][1], wreg[dy * 3 + dx][1], acc[o % 3][q][1]);
#endif
          }
      }
      if (i >= 2) {                                 // output row i - 2 is complete
        const int o = i - 2;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float2 g0, g1;
          if constexpr (EXACT) {
            g0 = gelu_erf2(acc[o % 3][q][0]);
            g1 = gelu_erf2(acc[o % 3][q][1]);
          } else {
            g0 = gelu_tanh2_half_arg(acc[o % 3][q][0]);
            g1 = gelu_tanh2_half_arg(acc[o % 3][q][1]);
          }
          uint2 v;
          v.x = pack2_16<F16>(g0.x, g0.y);
          v.y = pack2_16<F16>(g1.x, g1.y);
          *reinterpret_cast<uint2*>(op + ((size_t)o * g.H + q) * g.Ch) = v;
          acc[o % 3][q][0] = bz[0];
          acc[o % 3][q][1] = bz[1];
        }
      }
    }
    __syncthreads();                                // everyone has drained this stage's patch
  }
}

template <bool F16, bool EXACT = false>
int launch_dw2(const void* in, void* out, const float* wt, const float* bias, int B, int H, int Ch, cudaStream_t st) {
  CUtensorMap tm;
  const uint64_t dims[4] = {(uint64_t)Ch, (uint64_t)H, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)Ch * 2, (uint64_t)H * Ch * 2, (uint64_t)H * H * Ch * 2};
  const uint32_t box[4] = {64, DW2_PW, DW2_PH, 1};
  WMK_TRY(make_tensor_map(&tm, in, 4, dims, strides, box, false, 0));
  auto lg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return l; };
  DwGeom2 g;
  g.H = H; g.Ch = Ch; g.lg_ncs = lg(Ch / 64); g.lg_tw = lg(H / DW2_TW); g.lg_th = lg(H / DW2_TH);
  const long long items = (long long)B * (H / DW2_TW) * (H / DW2_TH) * (Ch / 64);
  WMK_REQUIRE(items < (1LL << 31), "dwconv: too many tiles (%lld)", items);
  g.n_items = (int)items;
  const int grid = (int)(items < 4LL * num_sms() ? items : 4LL * num_sms());
  dwconv3x3_gelu_tma2_kernel<F16, EXACT><<<grid, kDwThreads, 0, st>>>(tm, reinterpret_cast<uint16_t*>(out), wt, bias, g);
  WMK_CHECK_LAUNCH("dwconv3x3_gelu_tma2_kernel");
  return 0;
}

}  // namespace

// in / out: [B][H][H][Ch] 16-bit (token layout): bf16 (f16 = 0) or fp16 (f16 = 1) with wt / bias [9][Ch] / [Ch] fp32 BOTH
// PRE-MULTIPLIED BY 0.5 (uformer_plan.cu pack_block, gelu_tanh2_half_arg); f16 = 2: fp16 tensors, plain wt / bias and the
// erf-form GELU (the precise extractor).
int dwconv3x3_gelu_op16(const void* in, void* out, const float* wt, const float* bias, int B, int H, int Ch, int f16,
                        cudaStream_t st) {
  WMK_REQUIRE(H >= 8 && (H & (H - 1)) == 0 && Ch >= 64 && (Ch & (Ch - 1)) == 0,
              "dwconv: H=%d must be a power of two >= 8 and Ch=%d a power of two >= 64", H, Ch);
  WMK_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "dwconv: buffers must be 16-byte aligned");
  static const int one_col = getenv("WMK_DW_ONECOL") ? atoi(getenv("WMK_DW_ONECOL")) : 0;
  if (f16 == 2) {        // fp16 tensors, erf-form GELU, wt / bias NOT pre-halved (the precise extractor)
    if (H % 16 == 0) return launch_dw2<true, true>(in, out, wt, bias, B, H, Ch, st);
    return launch_dw<8, true, true>(in, out, wt, bias, B, H, Ch, st);
  }
  if (f16) {
    if (H % 16 == 0 && !one_col) return launch_dw2<true>(in, out, wt, bias, B, H, Ch, st);
    if (H % 16 == 0) return launch_dw<16, true>(in, out, wt, bias, B, H, Ch, st);
    return launch_dw<8, true>(in, out, wt, bias, B, H, Ch, st);
  }
  if (H % 16 == 0 && !one_col) return launch_dw2<false>(in, out, wt, bias, B, H, Ch, st);
  if (H % 16 == 0) return launch_dw<16, false>(in, out, wt, bias, B, H, Ch, st);
  return launch_dw<8, false>(in, out, wt, bias, B, H, Ch, st);
}

}  // namespace wmk

// LeWin window attention (uformerWM/model.py:460-471 LinearProjection, :523-551 WindowAttention, :954-1012 roll /
// partition / shift mask / reverse) as ONE persistent tcgen05 kernel, fused with its q|k|v projection:
//     O = softmax( (LN1(x) Wq^T + bq) (LN1(x) Wk^T + bk)^T * scale + rel_pos_bias [+ shift mask] ) (LN1(x) Wv^T + bv)
// q, k, v (6C of the 16C bytes per token the separate projection + attention kernels move) never reach HBM.
//
// One tile = two horizontally adjacent 8x8 windows of one image = 128 tokens = one UMMA M tile.  The tokens of a
// window arrive as four 4x4 sub-blocks (TMA 4-D boxes of the [n][H][H][C] LayerNorm-1 output, 128B swizzle): with the
// cyclic shift of 4 a sub-block never wraps around the image edge, so roll + window partition are box coordinates.
// Row order inside a window is therefore "quad order" (sub-block, row in sub-block, column in sub-block); the
// relative-position bias table is packed in that order on the host and the shift-mask regions ARE the sub-blocks.
//
//   warp 0      TMA: the packed per-head weights [heads][q(32) | k(32) | v(32)][C] once (resident), the A tile per tile
//   warp 1      MMA issuer, per head h:   QKV_h [128 x 96]  = A [128 x C] . W_h^T                  (TMEM, double buffered)
//                                         S     [128 x 128] = Q_h . K_h^T  (both windows; only the two diagonal
//                                                             64 x 64 blocks are used)
//                                         O_h   [128 x 32]  = P . V_h      (P block diagonal, fp16; TMEM, double buffered)
//   warp 2      TMEM allocation
//   warps 4-19  workers, four per TMEM lane quarter (lane = token row, worker j4 = a quarter of the columns):
//               QKV_h -> + bias -> fp16 Q | K tile (K-major, swizzled) and V^T tile (transposed on the way) in shared memory;
//               S -> + bias (fp16 table in shared memory, added by mixed-precision FMAs) + mask -> row max (exchanged
//               between the four workers of a row through shared memory) -> exp2 -> P tile + partial row sums;
//               O_h -> / row sum -> fp16 -> global (token order: un-window + un-roll are address arithmetic)
// The scores are in the log2 domain: the packed q rows / bias table carry scale * log2(e) / log2(e) (pack_block).
#include "tc_ptx.cuh"

namespace wmk {

namespace {

using namespace tc;

constexpr int AB_W0 = 4, AB_WORKERS = 16;
constexpr int AB_THREADS = 32 * (AB_W0 + AB_WORKERS);     // 640
constexpr uint32_t AB_WTILE = 96 * 128;                   // one head's [96 x 64] weight k-block
constexpr uint32_t AB_ATILE = 128 * 128;                  // [128 x 64] 16-bit k-block
constexpr uint32_t AB_VT = 2 * 32 * 128;                  // V^T: two key k-blocks of [32 x 64]
constexpr float AB_MASK = -100.0f * 1.4426950408889634f;  // shift mask value (model.py:971) in the log2 domain

struct AbGeom {
  int H, lg_nw, shift, n_tiles;
  int KC, ksteps;
  int a_st;
  const float* bqkv;        // [heads][96]
  const uint16_t* bias;     // [heads][64][64] fp16, quad order, x log2(e)
  uint16_t* out;            // [tokens][C] fp16
};

__device__ __forceinline__ void ab_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    if (++spins > (1u << 24)) __trap();
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// s0 += lo(b) * 1, s1 += hi(b) * 1 : adds two fp16 values to fp32 accumulators without unpacking them
__device__ __forceinline__ void add_f16x2(float& s0, float& s1, uint32_t b) {
  asm("{\n\t.reg .b16 bl, bh, one;\n\t"
      "mov.b32 {bl, bh}, %2;\n\t"
      "mov.b16 one, 0x3c00;\n\t"
      "fma.rn.f32.f16 %0, bl, one, %0;\n\t"
      "fma.rn.f32.f16 %1, bh, one, %1;\n\t}"
      : "+f"(s0), "+f"(s1)
      : "r"(b));
}
__device__ __forceinline__ float ab_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int C>
__global__ void __launch_bounds__(AB_THREADS, 1)
attn_block_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, AbGeom g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr int NH = C / 32;
  const int KC = g.KC;
  const uint32_t wsm = base;                                               // [NH][KC][96 x 128 B]
  const uint32_t asm_ = wsm + (uint32_t)(NH * KC) * AB_WTILE;              // [a_st][KC][16 KB]
  constexpr bool BIAS_SMEM = C <= 64;                                      // C = 128: the table stays in global memory (L1 / L2)
  const uint32_t qksm = asm_ + (uint32_t)(g.a_st * KC) * AB_ATILE;         // [2][128 x 64]: q (cols 0-31) | k (cols 32-63) of a head
  const uint32_t vtsm = qksm + 2u * AB_ATILE;                              // [2] V^T
  const uint32_t psm = vtsm + 2u * AB_VT;                                  // P: two key k-blocks of [128 x 64]
  const uint32_t biassm = psm + 2u * AB_ATILE;                             // [NH][64 rows x 128 B], 16-byte chunks XOR (row & 7)
  const uint32_t smaxsm = biassm + (BIAS_SMEM ? (uint32_t)NH * 8192u : 0u);   // [2][4][128] float
  const uint32_t ssumsm = smaxsm + 4096u;                                  // [2][4][128] float
  const uint32_t bars = ssumsm + 4096u;
  const uint32_t wfull = bars;
  auto afull = [&](int i) { return bars + 8u * (1 + i); };
  auto aempty = [&](int i) { return bars + 8u * (3 + i); };
  auto qkvfull = [&](int i) { return bars + 8u * (5 + i); };
  auto qkvempty = [&](int i) { return bars + 8u * (7 + i); };
  auto ofull = [&](int i) { return bars + 8u * (9 + i); };
  auto oempty = [&](int i) { return bars + 8u * (11 + i); };
  const uint32_t qksmfull = bars + 8u * 13, sfull = bars + 8u * 14, sempty = bars + 8u * 15, pfull = bars + 8u * 16;
  const uint32_t tmem_slot = bars + 8u * 17;
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = (int)gridDim.x;
  const int my_tiles = (int)blockIdx.x < g.n_tiles ? (g.n_tiles - 1 - (int)blockIdx.x) / grid + 1 : 0;
  const int nW = 1 << g.lg_nw;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(afull(i), 1); mbar_init(aempty(i), 1);
      mbar_init(qkvfull(i), 1); mbar_init(qkvempty(i), AB_WORKERS);
      mbar_init(ofull(i), 1); mbar_init(oempty(i), AB_WORKERS);
    }
    mbar_init(qksmfull, AB_WORKERS); mbar_init(sfull, 1); mbar_init(sempty, AB_WORKERS); mbar_init(pfull, AB_WORKERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // the bias tables (fp16, 16-byte chunks swizzled by the row) and the never-written zero blocks of P
  {
    const uint4* src = reinterpret_cast<const uint4*>(g.bias);
    for (int i = threadIdx.x; BIAS_SMEM && i < NH * 64 * 8; i += AB_THREADS) {       // 16-byte chunks
      const int row = i >> 3, c = i & 7;
      *reinterpret_cast<uint4*>(smem_raw + (biassm - raw) + (uint32_t)row * 128u + (uint32_t)((c ^ (row & 7)) << 4)) = src[i];
    }
    for (int i = threadIdx.x; i < (int)(2 * AB_ATILE / 16); i += AB_THREADS)
      *reinterpret_cast<uint4*>(smem_raw + (psm - raw) + (uint32_t)i * 16u) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  // TMEM columns: q|k|v accumulators at 0 / 128 (96 columns each), S at 256 (128 columns), O_h at 384 / 416 (32 each)

  auto tile_at = [&](int lt, int& b, int& wy, int& wx0) {
    const int tile = (int)blockIdx.x + lt * grid;
    const int tpi = 1 << (2 * g.lg_nw - 1);           // tiles per image
    b = tile >> (2 * g.lg_nw - 1);
    const int rem = tile & (tpi - 1);
    wy = rem >> (g.lg_nw - 1);
    wx0 = (rem & ((nW >> 1) - 1)) << 1;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA
    if (lane == 0) {
      mbar_arrive_expect_tx(wfull, (uint32_t)(NH * KC) * AB_WTILE);
      for (int h = 0; h < NH; ++h)
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(wsm + (uint32_t)(h * KC + kc) * AB_WTILE, &tmW, kc * 64, h * 96, wfull);
      for (int lt = 0; lt < my_tiles; ++lt) {
        int b, wy, wx0;
        tile_at(lt, b, wy, wx0);
        const int ab = lt % g.a_st;
        mbar_wait(aempty(ab), (((uint32_t)(lt / g.a_st)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(afull(ab), (uint32_t)KC * AB_ATILE);
        for (int kc = 0; kc < KC; ++kc)
          for (int win = 0; win < 2; ++win)
            for (int sb = 0; sb < 4; ++sb) {
              const int hh = (wy * 8 + (sb >> 1) * 4 + g.shift) & (g.H - 1);
              const int ww = ((wx0 + win) * 8 + (sb & 1) * 4 + g.shift) & (g.H - 1);
              tma_load_4d(asm_ + (uint32_t)(ab * KC + kc) * AB_ATILE + (uint32_t)(win * 64 + sb * 16) * 128u, &tmA, kc * 64, ww, hh,
                          b, afull(ab));
            }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_qkv = umma_idesc(96, true), idesc_s = umma_idesc(128, true), idesc_pv = umma_idesc(32, true);
      mbar_wait(wfull, 0);
      const int T = my_tiles * NH;
      auto issue_qkv = [&](int gq) {
        const int lt = gq / NH, h = gq - lt * NH, ab = lt % g.a_st, b = gq & 1;
        if (h == 0) {
          mbar_wait(afull(ab), ((uint32_t)(lt / g.a_st)) & 1u);
          tcgen05_fence_after();
        }
        mbar_wait(qkvempty(b), (((uint32_t)(gq >> 1)) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(b * 128);
        uint32_t first = 1u;
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t adesc = umma_desc_sw128(asm_ + (uint32_t)(ab * KC + kc) * AB_ATILE);
          const uint64_t wdesc = umma_desc_sw128(wsm + (uint32_t)(h * KC + kc) * AB_WTILE);
          for (int k = 0; k < g.ksteps; ++k) {
            tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), wdesc + (uint64_t)(2 * k), idesc_qkv, first ? 0u : 1u);
            first = 0u;
          }
        }
        tcgen05_commit(qkvfull(b));
        if (h == NH - 1) tcgen05_commit(aempty(ab));          // the A tile is free once the last head's projection retires
      };
      int gq = 0;                                              // next projection to issue: two heads ahead of the attention
      for (; gq < T && gq < 2; ++gq) issue_qkv(gq);
      for (int gh = 0; gh < T; ++gh) {
        const int pb = gh & 1;
        // S = Q K^T
        mbar_wait(qksmfull, (uint32_t)gh & 1u);
        mbar_wait(sempty, ((uint32_t)gh & 1u) ^ 1u);
        tcgen05_fence_after();
        {
          const uint64_t qd = umma_desc_sw128(qksm + (uint32_t)pb * AB_ATILE);
#pragma unroll
          for (int k = 0; k < 2; ++k)
            tcgen05_mma_bf16(tmem_base + 256u, qd + (uint64_t)(2 * k), qd + (uint64_t)(4 + 2 * k), idesc_s, (uint32_t)(k != 0));
          tcgen05_commit(sfull);
        }
        if (gq < T) { issue_qkv(gq); ++gq; }                   // its accumulator (head gh's) has been drained: qksmfull
        // O_h = P V
        mbar_wait(pfull, (uint32_t)gh & 1u);
        mbar_wait(oempty(pb), (((uint32_t)(gh >> 1)) & 1u) ^ 1u);
        tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t pd = umma_desc_sw128(psm + (uint32_t)kb * AB_ATILE);
          const uint64_t vd = umma_desc_sw128(vtsm + (uint32_t)pb * AB_VT + (uint32_t)kb * 4096u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tcgen05_mma_bf16(tmem_base + (uint32_t)(384 + pb * 32), pd + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), idesc_pv,
                             (uint32_t)((kb | k) != 0));
        }
        tcgen05_commit(ofull(pb));
      }
    }
  } else if (warp >= AB_W0) {
    // ------------------------------------------------------------------------------------------ workers
    const int q = warp & 3;                       // TMEM lane quarter
    const int j4 = (warp - AB_W0) >> 2;           // column quarter
    const int r = q * 32 + lane;                  // tile row = token
    const int win = q >> 1, rw = r & 63, sb = rw >> 4;
    const int pi = ((sb >> 1) << 2) | ((rw >> 2) & 3), pj = ((sb & 1) << 2) | (rw & 3);   // pixel inside the window
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t sw = (uint32_t)(r & 7);
    auto gen = [&](uint32_t a) -> uint8_t* { return smem_raw + (a - raw); };     // shared-window address -> generic pointer
    float* smax = reinterpret_cast<float*>(smem_raw + (smaxsm - raw));
    float* ssum = reinterpret_cast<float*>(smem_raw + (ssumsm - raw));
    const int T = my_tiles * NH;
    size_t tok = 0;
    float maskadd = 0.f;

    auto o_evac = [&](int gq, size_t token) {    // O_h of head gq: / row sum -> fp16 -> global
      const int ob = gq & 1, h = gq & (NH - 1);
      ab_wait(ofull(ob), ((uint32_t)(gq >> 1)) & 1u);
      tcgen05_fence_after();
      uint32_t v[8];
      tmem_ld8(lane_addr + (uint32_t)(384 + ob * 32 + j4 * 8), v);
      tmem_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(oempty(ob));
      const float* ss = ssum + (gq & 1) * 512 + r;
      const float inv = 1.0f / (ss[0] + ss[128] + ss[256] + ss[384]);
      uint4 o;
      o.x = pack2_f16(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
      o.y = pack2_f16(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
      o.z = pack2_f16(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
      o.w = pack2_f16(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
      *reinterpret_cast<uint4*>(g.out + token * C + h * 32 + j4 * 8) = o;
    };

    auto qkv_evac = [&](int gq) {                 // q | k | v of head gq -> shared memory (buffer gq & 1)
      const int b = gq & 1, h = gq & (NH - 1);
      const float* bp = g.bqkv + h * 96 + j4 * 8;
      float4 bq0 = __ldg(reinterpret_cast<const float4*>(bp)), bq1 = __ldg(reinterpret_cast<const float4*>(bp + 4));
      float4 bk0 = __ldg(reinterpret_cast<const float4*>(bp + 32)), bk1 = __ldg(reinterpret_cast<const float4*>(bp + 36));
      float4 bv0 = __ldg(reinterpret_cast<const float4*>(bp + 64)), bv1 = __ldg(reinterpret_cast<const float4*>(bp + 68));
      ab_wait(qkvfull(b), ((uint32_t)(gq >> 1)) & 1u);
      tcgen05_fence_after();
      uint32_t vq[8], vk[8], vv[8];
      const uint32_t ta = lane_addr + (uint32_t)(b * 128 + j4 * 8);
      tmem_ld8(ta, vq);
      tmem_ld8(ta + 32u, vk);
      tmem_ld8(ta + 64u, vv);
      tmem_wait_ld();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(qkvempty(b));
      const uint32_t row = qksm + (uint32_t)b * AB_ATILE + (uint32_t)r * 128u;
      st_shared_v4(row + ((((uint32_t)j4) ^ sw) << 4),
                   pack2_f16(__uint_as_float(vq[0]) + bq0.x, __uint_as_float(vq[1]) + bq0.y),
                   pack2_f16(__uint_as_float(vq[2]) + bq0.z, __uint_as_float(vq[3]) + bq0.w),
                   pack2_f16(__uint_as_float(vq[4]) + bq1.x, __uint_as_float(vq[5]) + bq1.y),
                   pack2_f16(__uint_as_float(vq[6]) + bq1.z, __uint_as_float(vq[7]) + bq1.w));
      st_shared_v4(row + ((((uint32_t)(4 + j4)) ^ sw) << 4),
                   pack2_f16(__uint_as_float(vk[0]) + bk0.x, __uint_as_float(vk[1]) + bk0.y),
                   pack2_f16(__uint_as_float(vk[2]) + bk0.z, __uint_as_float(vk[3]) + bk0.w),
                   pack2_f16(__uint_as_float(vk[4]) + bk1.x, __uint_as_float(vk[5]) + bk1.y),
                   pack2_f16(__uint_as_float(vk[6]) + bk1.z, __uint_as_float(vk[7]) + bk1.w));
      // V^T[d][key r], d = j4*8 + e: key k-block r >> 6, 16-byte chunk (key & 63) >> 3 swizzled by d & 7
      const float vb[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
      const uint32_t kk = (uint32_t)(r & 63);
      uint8_t* vt = gen(vtsm + (uint32_t)b * AB_VT + (uint32_t)(r >> 6) * 4096u + (kk & 7u) * 2u);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t d = (uint32_t)(j4 * 8 + e);
        const __half hv = __float2half_rn(__uint_as_float(vv[e]) + vb[e]);
        *reinterpret_cast<__half*>(vt + d * 128u + (((kk >> 3) ^ (d & 7u)) << 4)) = hv;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(qksmfull);
    };

    size_t tok_prev = 0;
    uint4 nb0 = make_uint4(0u, 0u, 0u, 0u), nb1 = nb0;
    if constexpr (!BIAS_SMEM) {
      const uint4* bg = reinterpret_cast<const uint4*>(g.bias + (size_t)rw * 64 + j4 * 16);
      nb0 = __ldg(bg);
      nb1 = __ldg(bg + 1);
    }
    if (T > 0) qkv_evac(0);
    for (int gh = 0; gh < T; ++gh) {
      const int h = gh & (NH - 1);
      if (h == 0) {                               // geometry of the new tile
        int b, wy, wx0;
        tile_at(gh / NH, b, wy, wx0);
        const int wx = wx0 + win;
        tok_prev = tok;
        tok = ((size_t)b * g.H + ((wy * 8 + pi + g.shift) & (g.H - 1))) * g.H + ((wx * 8 + pj + g.shift) & (g.H - 1));
        const bool lastrow = g.shift > 0 && wy == nW - 1, lastcol = g.shift > 0 && wx == nW - 1;
        maskadd = ((lastrow && (sb >> 1) != (j4 >> 1)) || (lastcol && (sb & 1) != (j4 & 1))) ? AB_MASK : 0.f;
      }
      // ---------------------------------------------------------------- softmax of this row's 16 columns
      {
        uint4 b0, b1;                             // bias row: 16 fp16 values of this thread's columns
        if constexpr (BIAS_SMEM) {
          const uint8_t* brow = gen(biassm + (uint32_t)h * 8192u + (uint32_t)rw * 128u);
          b0 = *reinterpret_cast<const uint4*>(brow + ((((uint32_t)(j4 * 2)) ^ (uint32_t)(rw & 7)) << 4));
          b1 = *reinterpret_cast<const uint4*>(brow + ((((uint32_t)(j4 * 2 + 1)) ^ (uint32_t)(rw & 7)) << 4));
        } else {                                  // fetched from global memory one head ahead (L2 latency off the softmax path)
          b0 = nb0;
          b1 = nb1;
          const uint4* bg = reinterpret_cast<const uint4*>(g.bias + ((size_t)((h + 1) & (NH - 1)) * 64 + rw) * 64 + j4 * 16);
          nb0 = __ldg(bg);
          nb1 = __ldg(bg + 1);
        }
        ab_wait(sfull, (uint32_t)gh & 1u);
        tcgen05_fence_after();
        uint32_t v[16];
        tmem_ld16(lane_addr + (uint32_t)(256 + win * 64 + j4 * 16), v);
        tmem_wait_ld();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sempty);
        float s[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) s[i] = __uint_as_float(v[i]) + maskadd;      // maskadd = 0 except in border windows of shifted blocks
        add_f16x2(s[0], s[1], b0.x); add_f16x2(s[2], s[3], b0.y); add_f16x2(s[4], s[5], b0.z); add_f16x2(s[6], s[7], b0.w);
        add_f16x2(s[8], s[9], b1.x); add_f16x2(s[10], s[11], b1.y); add_f16x2(s[12], s[13], b1.z); add_f16x2(s[14], s[15], b1.w);
        float m = s[0];
#pragma unroll
        for (int i = 1; i < 16; ++i) m = fmaxf(m, s[i]);
        float* mx = smax + (gh & 1) * 512 + r;
        mx[j4 * 128] = m;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(128) : "memory");
        m = fmaxf(fmaxf(mx[0], mx[128]), fmaxf(mx[256], mx[384]));
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { s[i] = ab_ex2(s[i] - m); sum += s[i]; }
        ssum[(gh & 1) * 512 + j4 * 128 + r] = sum;
        if (gh > 0) ab_wait(ofull((gh - 1) & 1), ((uint32_t)((gh - 1) >> 1)) & 1u);      // P V of the previous head has read P (and its V^T)
        const uint32_t prow = psm + (uint32_t)win * AB_ATILE + (uint32_t)r * 128u;
        st_shared_v4(prow + ((((uint32_t)(j4 * 2)) ^ sw) << 4), pack2_f16(s[0], s[1]), pack2_f16(s[2], s[3]), pack2_f16(s[4], s[5]),
                     pack2_f16(s[6], s[7]));
        st_shared_v4(prow + ((((uint32_t)(j4 * 2 + 1)) ^ sw) << 4), pack2_f16(s[8], s[9]), pack2_f16(s[10], s[11]),
                     pack2_f16(s[12], s[13]), pack2_f16(s[14], s[15]));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(pfull);
      }
      if (gh + 1 < T) qkv_evac(gh + 1);           // the S MMA of the next head then runs under the output of the previous one
      if (gh > 0) o_evac(gh - 1, h == 0 ? tok_prev : tok);
    }
    if (T > 0) o_evac(T - 1, tok);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

template <int C>
int launch_attn_block(const void* A, const void* Wh, const float* bqkv, const uint16_t* bias, uint16_t* out, int n, int H,
                      int shift, cudaStream_t st) {
  constexpr int NH = C / 32;
  AbGeom g;
  g.H = H; g.lg_nw = 0;
  while ((8 << g.lg_nw) < H) ++g.lg_nw;
  g.shift = shift;
  g.n_tiles = n * (H / 8) * (H / 8) / 2;
  g.KC = C >= 64 ? C / 64 : 1; g.ksteps = C >= 64 ? 4 : 2;
  g.bqkv = bqkv; g.bias = bias; g.out = out;
  const int fixed = 1024 + 2 * (int)AB_ATILE + 2 * (int)AB_VT + 2 * (int)AB_ATILE + (C <= 64 ? NH * 8192 : 0) + 8192 + 256;
  g.a_st = 2;
  auto total = [&]() { return fixed + NH * g.KC * (int)AB_WTILE + g.a_st * g.KC * (int)AB_ATILE; };
  if (total() > 227 * 1024) g.a_st = 1;
  WMK_REQUIRE(total() <= 227 * 1024, "attn_block: %d bytes of shared memory needed (C = %d)", total(), C);
  CUtensorMap tmA, tmW;
  {   // LayerNorm-1 output [n][H][H][C] fp16: 4 x 4 pixel boxes of 64 channels
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)H, (uint64_t)H, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)C * 2, (uint64_t)H * C * 2, (uint64_t)H * H * C * 2};
    const uint32_t box[4] = {64, 4, 4, 1};
    WMK_TRY(make_tensor_map(&tmA, A, 4, dims, strides, box, false, 128));
  }
  {   // per-head weights [NH * 96][C]: 96-row tiles of 64 columns
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)NH * 96};
    const uint64_t strides[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {64, 96};
    WMK_TRY(make_tensor_map(&tmW, Wh, 2, dims, strides, box, false, 128));
  }
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(attn_block_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = g.n_tiles < num_sms() ? g.n_tiles : num_sms();
  attn_block_kernel<C><<<grid, AB_THREADS, (size_t)total(), st>>>(tmA, tmW, g);
  WMK_CHECK_LAUNCH("attn_block_kernel");
  return 0;
}

}  // namespace

// out[tokens][C] (fp16, token order) = window attention of A = LayerNorm-1 output [n][H][H][C] (fp16, incl. the modulator)
// with the q|k|v projection fused.  Wh: [C/32 heads][q(32) | k(32) | v(32)][C] fp16 (q rows x scale x log2 e);
// bqkv [heads][96] fp32 (q part scaled alike); bias [heads][64][64] fp16 in quad order x log2 e.
// C in {32, 64, 128}, H a power of two >= 16, shift 0 or 4.
int attn_block(const void* A, const void* Wh, const float* bqkv, const uint16_t* bias, uint16_t* out, int n, int H, int C,
               int shift, cudaStream_t st) {
  WMK_REQUIRE(H >= 16 && H <= 128 && (H & (H - 1)) == 0, "attn_block: H=%d must be a power of two in [16,128]", H);
  WMK_REQUIRE(C == 32 || C == 64 || C == 128, "attn_block: covers C in {32,64,128}, got %d", C);
  WMK_REQUIRE(shift == 0 || shift == 4, "attn_block: shift must be 0 or 4, got %d", shift);
  const double M = (double)n * H * H;
  // algorithmic traffic: A 2C + O 2C bytes per token; FLOPs: q|k|v projection + the two attention contractions
  ProfScope prof(FAM_ATTENTION, M * C * 4, st, 2.0 * M * C * 3 * C + 256.0 * C * M);
  switch (C) {
    case 32: return launch_attn_block<32>(A, Wh, bqkv, bias, out, n, H, shift, st);
    case 64: return launch_attn_block<64>(A, Wh, bqkv, bias, out, n, H, shift, st);
    default: return launch_attn_block<128>(A, Wh, bqkv, bias, out, n, H, shift, st);
  }
}

}  // namespace wmk

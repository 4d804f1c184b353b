// bf16 dense layer on the sm_100a 5th-generation tensor cores.
//
//   C[M][N] = epilogue( A[M][K] * W[N][K]^T + bias )       A, W bf16 K-major; fp32 accumulate
//
// One CTA computes one 128 x BN output tile:
//   warp 0     TMA producer : cp.async.bulk.tensor 2-D tiles (128B swizzle) of A and W into a
//                              ring of shared-memory stages, completion on mbarriers
//   warp 1     MMA issuer   : one thread issues tcgen05.mma (M=128, N=BN, K=16) into a TMEM
//                              accumulator; tcgen05.commit releases stages / signals the epilogue
//   warps 2-5  epilogue     : tcgen05.ld the accumulator (one TMEM lane = one output row),
//                              bias / GELU / residual / pixel-shuffle, 128-bit global stores
// Several CTAs are resident per SM (TMEM columns and shared memory permitting), so one CTA's
// epilogue overlaps another's main loop.  Replaces every nn.Linear of the LeWin blocks
// (uformerWM/model.py:455-456,518,686,690) and, through im2col / pixel-shuffle, the 4x4-s2 and
// transposed 2x2-s2 convolutions (model.py:763,789).
#include <cstdlib>
#include <mutex>

#include "tc_ptx.cuh"

namespace wmk {

namespace {

using namespace tc;
constexpr int kThreads = 192;


template <int BN>
__global__ void __launch_bounds__(kThreads)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    EpiParams p, int K, int n_tiles, int n_stages) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // 128B swizzle needs 1024B alignment
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t W_BYTES = BN * BK * 2;
  constexpr uint32_t STAGE = A_BYTES + W_BYTES;
  const uint32_t bars = base + (uint32_t)n_stages * STAGE;   // full[ns], empty[ns], tmem_full, tmem slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (n_stages + s); };
  const uint32_t tmem_full_bar = bars + 16u * n_stages;
  const uint32_t tmem_slot = tmem_full_bar + 8u;
  volatile uint32_t* tmem_slot_ptr =
      (volatile uint32_t*)(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int bn_idx = tile % n_tiles, bm_idx = tile / n_tiles;
  const int m0 = bm_idx * BM, n0 = bn_idx * BN;
  const int kblocks = (K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(tmem_cols(BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % n_stages;
        const uint32_t ph = (uint32_t)(kb / n_stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_arrive_expect_tx(full_bar(s), STAGE);
        const uint32_t sa = base + (uint32_t)s * STAGE;
        tma_load_2d(sa, &tmA, kb * BK, m0, full_bar(s));
        tma_load_2d(sa + A_BYTES, &tmW, kb * BK, n0, full_bar(s));
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BN, p.f16 != 0);
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % n_stages;
        const uint32_t ph = (uint32_t)(kb / n_stages) & 1u;
        mbar_wait(full_bar(s), ph);
        tcgen05_fence_after();
        const uint32_t sa = base + (uint32_t)s * STAGE;
        const uint64_t adesc = umma_desc_sw128(sa);
        const uint64_t bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
          tcgen05_mma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                           (uint32_t)((kb | k) != 0));
        }
        tcgen05_commit(empty_bar(s));          // frees the stage when these MMAs retire
      }
      tcgen05_commit(tmem_full_bar);           // accumulator complete
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const bool row_ok = row < p.M;
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      if (!row_ok) continue;
      const int n = n0 + c * 32;
      const size_t off = epi_row_offset(p, row, n);
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (p.bias) {
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = __ldg(b4 + j);
          f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w;
        }
      }
      if (p.epi == EPI_BIAS_GELU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
      } else if (p.epi == EPI_BIAS_RESID) {
        const float4* r4 = reinterpret_cast<const float4*>(p.resid + off);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 r = r4[j];
          f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
        }
      }
      if (p.out_bf16) {
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.C) + off);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 u;
          pack8_16(f + 8 * j, p.f16 != 0, u.x, u.y, u.z, u.w);
          o[j] = u;
        }
      } else {
        float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + off);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc),
                 "n"(tmem_cols(BN)));
  }
}


// ---------------------------------------------------------------------------------------------
// Persistent, warp-specialised version (all epilogues except the pixel-shuffle one):
//   one CTA per SM loops over output tiles (n fastest, so CTAs that run together share A in L2);
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-17 epilogue (four per TMEM lane quarter, each
//   owning a column slab - the GELU / residual epilogues are instruction-issue bound, so they get
//   16 warps).  Two TMEM accumulators: the epilogue of tile i overlaps the main loop of tile i+1.
//   Each epilogue warp stages its 32 rows x <=128 B in swizzled shared memory and writes them with a
//   TMA bulk-tensor store (coalesced, asynchronous, clips the M tail).
// ---------------------------------------------------------------------------------------------
constexpr int kEpiWarps = 16;
constexpr int kPThreads = 64 + 32 * kEpiWarps;

// columns of the accumulator each epilogue warp owns, and the TMA-store box width
// (128-wide tiles without the fused LayerNorm: 64 columns per warp, i.e. two teams of 8 warps)
#ifndef WMK_BN128_TEAMS
#define WMK_BN128_TEAMS 2
#endif
__host__ __device__ constexpr int epi_cpw(int bn, bool ln) {
  return bn >= 256 ? bn / 4 : (bn == 128 && !ln && WMK_BN128_TEAMS == 2) ? 64 : 32;
}
__host__ __device__ constexpr int epi_box_cols(int bn, bool out_bf16, bool ln) {
  return out_bf16 ? (epi_cpw(bn, ln) >= 64 ? 64 : 32) : 32;
}

// LN = true (EPI_BIAS_RESID, BN <= 128, n_tiles == 1): the epilogue also LayerNorms the finished row and writes
// it as bf16 through tmD - the operand of the next dense layer.  The SLABS warps that share a lane quarter
// exchange their partial (sum, sum of squares) through shared memory and a named barrier.
constexpr uint32_t STG2_BYTES = 2048;                       // 16-bit staging box of the fused LayerNorm: 32 rows x 64 B (two of them for split rows)
constexpr uint32_t LN_EXCH_BYTES = 2 * kEpiWarps * 32 * 8;  // [tile parity][slab][quarter][lane] float2

template <int BN, int EPI, bool OUT_BF16, bool LN = false>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_tcgen05_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD,
                               const __grid_constant__ CUtensorMap tmR,
                               EpiParams p, int K, int m_tiles, int n_tiles, int n_stages, int w_stationary) {
  // w_stationary: the CTA keeps its whole BN x K weight tile resident in shared memory and walks
  // down the M tiles of one N tile, so only A streams from L2 (for K <= 256 the weight tile would
  // otherwise be re-fetched for every output tile and the L2 -> SM path becomes the bound).
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t A_BYTES = BM * BK * 2;
  constexpr uint32_t W_BYTES = BN * BK * 2;
  // k-block schedule.  Plain: k-block kb = columns kb*64 of A and W, one A tile (+ one W tile) per stage.
  // Split-bf16 (p.split = 1, K % 64 == 0): A rows are [hi(K) | lo(K)], W rows [hi(K) | lo(K)]; a stage holds BOTH
  // A tiles (and both W tiles) of a logical k-block - each fetched once - and the issuer runs the three MMA groups
  // hi*hi, lo*hi, hi*lo on it.  p.split = 2 (K == 32): A rows [hi | lo] are ONE 64-wide tile, used against W tile
  // 0 = [hi | hi] (four k-steps) and W tile 1 = [lo | 0] (two k-steps).
  // p.split = 3 (W-only split, fp16): A is plain [M][K], W rows [hi(K) | lo(K)]: one A tile, two W tiles per stage, MMA
  // groups A*hi, A*lo.  p.split = 4 (the same, K == 32): W rows [hi(32) | lo(32)] are ONE tile; the zero-filled upper half
  // of the A tile is never read: k-steps 0-1 of A run against k-steps 0-1 (hi) and 2-3 (lo) of W.
  const int KB = (K + BK - 1) / BK;
  const int split = p.split;
  const int kblocks = split == 2 ? 1 : KB;                            // stages per output tile
  const int wblocks = (split == 1 || split == 3) ? 2 * KB : split == 2 ? 2 : KB;     // distinct W tiles (hi tiles first, then lo)
  const uint32_t A_STAGE = split == 1 ? 2 * A_BYTES : A_BYTES;
  const uint32_t W_STAGE = (split >= 1 && split <= 3) ? 2 * W_BYTES : W_BYTES;
  const uint32_t STAGE = w_stationary ? A_STAGE : A_STAGE + W_STAGE;
  const uint32_t wres = base;                                // resident weights: wblocks x W_BYTES
  const uint32_t stages = base + (w_stationary ? (uint32_t)wblocks * W_BYTES : 0u);
  // one staging box per epilogue warp: 32 rows x <= 128 B (bf16 output may use 64-byte rows, p.boxc = 32, when
  // a resident weight tile leaves too little shared memory for the A ring)
  const int BOXC = OUT_BF16 ? p.boxc : 32;
  const uint32_t STG_BYTES = OUT_BF16 ? 64u * (uint32_t)BOXC : 4096u;
  const uint32_t staging = stages + (uint32_t)n_stages * STAGE;          // [16 warps][STG_BYTES]
  // split-operand launches (the precise extractor) have no shared memory to spare for a second staging area: their fused
  // LayerNorm re-uses the warp's x staging buffer once its TMA store has been read (p.ln_reuse)
  const uint32_t stg2 = p.ln_reuse ? 0u : STG2_BYTES;
  const uint32_t staging2 = staging + kEpiWarps * STG_BYTES;             // LN: [16 warps][2048] + exchange
  const uint32_t ln_exch = staging2 + kEpiWarps * stg2;
  const uint32_t bars = LN ? ln_exch + LN_EXCH_BYTES : staging + kEpiWarps * STG_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (n_stages + s); };
  const uint32_t tfull_bar = bars + 16u * n_stages;         // [NACC <= 8]
  const uint32_t tempty_bar = tfull_bar + 64u;              // [NACC <= 8]
  const uint32_t wfull_bar = tempty_bar + 64u;
  const uint32_t tmem_slot = wfull_bar + 8u;
  const uint32_t rbar0 = tmem_slot + 8u;                    // [16 epilogue warps]: residual tile landed in the staging buffer
  volatile uint32_t* tmem_slot_ptr = (volatile uint32_t*)(smem_raw + (tmem_slot - raw));

  constexpr int CPW = epi_cpw(BN, LN);
  constexpr int SLABS = BN / CPW;                           // epilogue warps per lane quarter that share one tile
  constexpr int ACTIVE_EPI = SLABS * 4;                     // warps of one epilogue team
  // Narrow tiles (BN <= 64) finish their main loop faster than one epilogue pass (TMEM load -> residual ->
  // store [-> LayerNorm exchange -> store]) can run, so the 16 epilogue warps form 4 / 2 independent TEAMS
  // that take alternate tiles, with 2 TMEM accumulators per team: up to 8 tiles in flight.
  // 128-wide tiles whose epilogue is a latency chain (residual load -> TMEM load -> store) also run two teams
  // of 8 warps (two 32-column pieces per warp); the LayerNorm epilogue needs one piece per warp and keeps one team.
  constexpr int TEAMS = BN <= 32 ? 4 : BN <= 64 ? 2 : (BN == 128 && CPW == 64) ? 2 : 1;
  constexpr int NACC = 2 * TEAMS;
  constexpr int TCOLS = NACC * BN <= 32 ? 32 : NACC * BN <= 64 ? 64 : NACC * BN <= 128 ? 128 : NACC * BN <= 256 ? 256 : 512;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = m_tiles * n_tiles;
  // i-th tile of this CTA -> (m0, n0); false when the CTA has no i-th tile
  const int ws_groups = (int)gridDim.x / n_tiles;
  auto tile_at = [&](int i, int& m0, int& n0) -> bool {
    if (w_stationary) {
      const int mt = (int)blockIdx.x / n_tiles + i * ws_groups;
      m0 = mt * BM;
      n0 = ((int)blockIdx.x % n_tiles) * BN;
      return mt < m_tiles;
    }
    const int tile = (int)blockIdx.x + i * (int)gridDim.x;
    m0 = (tile / n_tiles) * BM;
    n0 = (tile % n_tiles) * BN;
    return tile < total;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    if (LN) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmD) : "memory");
    if (EPI == EPI_BIAS_RESID) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    mbar_init(wfull_bar, 1);
    for (int s = 0; s < n_stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(tfull_bar + 8u * a, 1);
      mbar_init(tempty_bar + 8u * a, ACTIVE_EPI);
    }
    for (int e = 0; e < kEpiWarps; ++e) mbar_init(rbar0 + 8u * e, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      int it = 0, m0, n0;
      if (w_stationary && tile_at(0, m0, n0)) {
        mbar_arrive_expect_tx(wfull_bar, (uint32_t)wblocks * W_BYTES);
        for (int wb = 0; wb < wblocks; ++wb) tma_load_2d(wres + (uint32_t)wb * W_BYTES, &tmW, wb * BK, n0, wfull_bar);
      }
      for (int i = 0; tile_at(i, m0, n0); ++i) {
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_arrive_expect_tx(full_bar(s), STAGE);
          const uint32_t sa = stages + (uint32_t)s * STAGE;
          if (p.conv_H > 0) {          // implicit GEMM: tap kb of the 3x3 window = the image row shifted by (dy-1, dx-1)
            const int rowi = m0 >> 7, img = rowi / p.conv_H, h = rowi - img * p.conv_H;
            tma_load_4d(sa, &tmA, 0, kb % 3 - 1, h + kb / 3 - 1, img, full_bar(s));
            if (!w_stationary) tma_load_2d(sa + A_STAGE, &tmW, kb * BK, n0, full_bar(s));
          } else if (p.dn_Ho > 0) {    // implicit GEMM over the space-to-depth tensor: k-block = (tap, 64-channel chunk)
            const int P = p.dn_Ho * p.dn_Ho;
            const int img = m0 / P, oh0 = (m0 - img * P) / p.dn_Ho;
            const int tap = kb / p.dn_chunks, c0 = (kb - tap * p.dn_chunks) * BK;
            tma_load_4d(sa, &tmA, c0, tap & 1, oh0 + (tap >> 1), img, full_bar(s));
            if (split == 1) tma_load_4d(sa + A_BYTES, &tmA, p.dn_chunks * BK + c0, tap & 1, oh0 + (tap >> 1), img, full_bar(s));   // lo
            tma_load_2d(sa + A_STAGE, &tmW, kb * BK, n0, full_bar(s));
            if (split == 1) tma_load_2d(sa + A_STAGE + W_BYTES, &tmW, K + kb * BK, n0, full_bar(s));
          } else if (split == 1) {
            tma_load_2d(sa, &tmA, kb * BK, m0, full_bar(s));                       // hi
            tma_load_2d(sa + A_BYTES, &tmA, K + kb * BK, m0, full_bar(s));         // lo
            if (!w_stationary) {
              tma_load_2d(sa + A_STAGE, &tmW, kb * BK, n0, full_bar(s));
              tma_load_2d(sa + A_STAGE + W_BYTES, &tmW, K + kb * BK, n0, full_bar(s));
            }
          } else {
            tma_load_2d(sa, &tmA, kb * BK, m0, full_bar(s));
            if (!w_stationary) {
              tma_load_2d(sa + A_STAGE, &tmW, kb * BK, n0, full_bar(s));
              if (split == 2) tma_load_2d(sa + A_STAGE + W_BYTES, &tmW, BK, n0, full_bar(s));
              if (split == 3) tma_load_2d(sa + A_STAGE + W_BYTES, &tmW, K + kb * BK, n0, full_bar(s));
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BN, p.f16 != 0 && (split == 0 || split >= 3));   // fully split operands are bf16 by construction
      int it = 0, m0, n0;
      if (w_stationary && tile_at(0, m0, n0)) mbar_wait(wfull_bar, 0);
      for (int lt = 0; tile_at(lt, m0, n0); ++lt) {
        const int acc = lt % NACC;
        mbar_wait(tempty_bar + 8u * acc, ((uint32_t)(lt / NACC) & 1u) ^ 1u);    // epilogue drained this accumulator
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % n_stages;
          const uint32_t ph = (uint32_t)(it / n_stages) & 1u;
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t sa = stages + (uint32_t)s * STAGE;
          const uint64_t adesc = umma_desc_sw128(sa);
          if (split == 0) {
            const uint64_t bdesc = umma_desc_sw128(w_stationary ? wres + (uint32_t)kb * W_BYTES : sa + A_STAGE);
            const int krem = K - kb * BK;
            const int ksteps = krem >= BK ? BK / UMMA_K : (krem + UMMA_K - 1) / UMMA_K;   // skip zero-filled K
            for (int k = 0; k < ksteps; ++k)
              tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
          } else {
            // W tile 0 / 1 of this k-block: hi / lo (split 1), [hi | hi] / [lo | 0] (split 2)
            const uint64_t w0 = umma_desc_sw128(w_stationary ? wres + (uint32_t)kb * W_BYTES : sa + A_STAGE);
            const uint64_t w1 = umma_desc_sw128(w_stationary ? wres + (uint32_t)((split == 2 ? 1 : KB) + kb) * W_BYTES
                                                             : sa + A_STAGE + W_BYTES);
            if (split == 3) {
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)     // A * hi
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w0 + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)     // A * lo
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w1 + (uint64_t)(2 * k), idesc, 1u);
            } else if (split == 4) {
#pragma unroll
              for (int k = 0; k < 2; ++k)               // A[0:32] * hi
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w0 + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
#pragma unroll
              for (int k = 0; k < 2; ++k)               // A[0:32] * lo (W columns 32..63)
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w0 + (uint64_t)(2 * (k + 2)), idesc, 1u);
            } else if (split == 1) {
              const uint64_t alo = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)     // hi * hi
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w0 + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)     // lo * hi
                tcgen05_mma_bf16(tacc, alo + (uint64_t)(2 * k), w0 + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)     // hi * lo
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w1 + (uint64_t)(2 * k), idesc, 1u);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)               // [hi | lo] * [hi | hi]
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w0 + (uint64_t)(2 * k), idesc, (uint32_t)(k != 0));
#pragma unroll
              for (int k = 0; k < 2; ++k)               // hi * lo
                tcgen05_mma_bf16(tacc, adesc + (uint64_t)(2 * k), w1 + (uint64_t)(2 * k), idesc, 1u);
            }
          }
          tcgen05_commit(empty_bar(s));
        }
        tcgen05_commit(tfull_bar + 8u * acc);
      }
    }
  } else {
    const int ew = warp - 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int team = (ew >> 2) / SLABS;           // which epilogue team
    const int slab = (ew >> 2) % SLABS;           // which CPW-wide column slab of the team's tile
    if (team < TEAMS) {
      const uint32_t buf = staging + (uint32_t)ew * STG_BYTES;
      // 16-bit staging box: this lane's row offset and swizzle key (hoisted out of the piece loop)
      const uint32_t sw_row = (uint32_t)lane * (BOXC == 64 ? 128u : 64u);
      const uint32_t sw_x = BOXC == 64 ? (uint32_t)(lane & 7) : ((uint32_t)(lane >> 1) & 3u);
      int m0, n0;
      // resid_tma: the residual rows of a tile are fetched by one TMA box into this warp's staging buffer (the buffer
      // alternates residual-in / result-out), issued as soon as the previous tile's store has been read - instead of
      // eight 16-byte global loads per lane, on which the epilogue warps were stalled (long_scoreboard)
      constexpr bool RT = EPI == EPI_BIAS_RESID && CPW == 32 && !OUT_BF16;      // one fp32 32-column piece per warp
      const bool rtma = RT && p.resid != nullptr;
      const uint32_t rbar = rbar0 + 8u * (uint32_t)ew;
      uint32_t rphase = 0;
      if (rtma && tile_at(team, m0, n0) && lane == 0) {
        mbar_arrive_expect_tx(rbar, 4096u);
        tma_load_2d(buf, &tmR, n0 + slab * CPW, m0 + q * 32, rbar);
      }
      for (int lt = team; tile_at(lt, m0, n0); lt += TEAMS) {
        const int acc = lt % NACC;
        const int row = m0 + q * 32 + lane;
        // residual rows do not depend on the accumulator: fetch the first 32-column piece while the MMAs run
        float4 rpre[RT ? 1 : 8];
        if constexpr (EPI == EPI_BIAS_RESID && !RT) {
          if (row < p.M && p.resid) {                     // resid == nullptr: plain bias epilogue
            const float4* r4 = reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ldc + n0 + slab * CPW);
#pragma unroll
            for (int j = 0; j < 8; ++j) rpre[j] = r4[j];
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) rpre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        if constexpr (EPI == EPI_BIAS_RESID && !RT) {
          // the residual rows of this team's NEXT tile: pull them from HBM into L2 now, so the loads at the top of
          // the next iteration see L2 latency (the epilogue chain of a tile is latency bound, not bandwidth bound)
          int m1, n1;
          if (p.resid_prefetch && p.resid && tile_at(lt + TEAMS, m1, n1)) {
            const int row1 = m1 + q * 32 + lane;
            if (row1 < p.M) {
              const float* r1 = p.resid + (size_t)row1 * p.ldc + n1 + slab * CPW;
#pragma unroll
              for (int cc = 0; cc < CPW; cc += 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(r1 + cc));
            }
          }
        }
        mbar_wait(tfull_bar + 8u * acc, (uint32_t)(lt / NACC) & 1u);
        tcgen05_fence_after();
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + slab * CPW);
#pragma unroll 1
        for (int cc = 0; cc < CPW; cc += 32) {
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)cc, v);
          const int n = n0 + slab * CPW + cc;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (p.bias) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
            for (int j = 0; j < 8; ++j) {                  // packed fp32x2 adds: 16 instead of 32 issue slots per piece
              const float4 b = __ldg(b4 + j);
              const float2 lo = __fadd2_rn(make_float2(f[4 * j], f[4 * j + 1]), make_float2(b.x, b.y));
              const float2 hi = __fadd2_rn(make_float2(f[4 * j + 2], f[4 * j + 3]), make_float2(b.z, b.w));
              f[4 * j] = lo.x; f[4 * j + 1] = lo.y; f[4 * j + 2] = hi.x; f[4 * j + 3] = hi.y;
            }
          }
          if constexpr (EPI == EPI_BIAS_GELU) {
            if (p.gelu_exact) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 gq = gelu_erf2(make_float2(f[2 * j], f[2 * j + 1]));
                f[2 * j] = gq.x; f[2 * j + 1] = gq.y;
              }
            } else if (p.gelu_half && p.gelu_mix) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 gq = (j & 1) ? gelu_poly2_half_arg(make_float2(f[2 * j], f[2 * j + 1]))
                                          : gelu_tanh2_half_arg(make_float2(f[2 * j], f[2 * j + 1]));
                f[2 * j] = gq.x; f[2 * j + 1] = gq.y;
              }
            } else if (p.gelu_half) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 gq = gelu_tanh2_half_arg(make_float2(f[2 * j], f[2 * j + 1]));
                f[2 * j] = gq.x; f[2 * j + 1] = gq.y;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float2 gq = gelu_tanh2(make_float2(f[2 * j], f[2 * j + 1]));
                f[2 * j] = gq.x; f[2 * j + 1] = gq.y;
              }
            }
          } else if constexpr (EPI == EPI_BIAS_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if constexpr (EPI == EPI_BIAS_RESID) {
            if constexpr (RT) {
              if (rtma) {                                  // resid == nullptr: plain bias epilogue (+ fused LayerNorm)
                mbar_wait(rbar, rphase);
                rphase ^= 1u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float4 r = ld_shared_f4(buf + (uint32_t)lane * 128u + (((uint32_t)j ^ (uint32_t)(lane & 7)) << 4));
                  f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
                }
              }
            } else if (cc == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[4 * j] += rpre[j].x; f[4 * j + 1] += rpre[j].y; f[4 * j + 2] += rpre[j].z; f[4 * j + 3] += rpre[j].w;
              }
            } else if (row < p.M && p.resid) {
              const float4* r4 = reinterpret_cast<const float4*>(p.resid + (size_t)row * p.ldc + n);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 r = r4[j];
                f[4 * j] += r.x; f[4 * j + 1] += r.y; f[4 * j + 2] += r.z; f[4 * j + 3] += r.w;
              }
            }
          }
          // stage this 32-column piece (one warp-private buffer; the previous store must have been read)
          const int cbox = cc & (BOXC - 1);               // BOXC is 32 or 64 (a runtime % cost an integer division per piece)
          const bool first_piece = OUT_BF16 ? cbox == 0 : true;
          if (first_piece && !rtma) {                   // rtma: the buffer was drained before the residual load was issued
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
          }
          if constexpr (OUT_BF16) {
            const uint32_t c16_0 = (uint32_t)(cbox >> 5) * 4u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              // SWIZZLE_128B (64-column box): row pitch 128 B, chunk ^ (row & 7); SWIZZLE_64B: pitch 64 B, chunk ^ ((row >> 1) & 3)
              const uint32_t off = sw_row + (((c16_0 + (uint32_t)j) ^ sw_x) << 4);
              uint32_t w0, w1, w2, w3;
              pack8_16(f + 8 * j, p.f16 != 0, w0, w1, w2, w3);
              st_shared_v4(buf + off, w0, w1, w2, w3);
            }
            if (cbox + 32 == BOXC) {
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0) tma_store_2d(&tmC, buf, n - cbox, m0 + q * 32);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              st_shared_v4(buf + (uint32_t)lane * 128u + (((uint32_t)j ^ (uint32_t)(lane & 7)) << 4),
                           __float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                           __float_as_uint(f[4 * j + 3]));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tma_store_2d(&tmC, buf, n, m0 + q * 32);
            if constexpr (LN) {
              // ---- fused LayerNorm of the finished rows (this warp holds 32 of the BN columns of row `row`)
              float s1 = 0.f, s2 = 0.f;
#pragma unroll
              for (int j = 0; j < 32; ++j) { s1 += f[j]; s2 = fmaf(f[j], f[j], s2); }
              if constexpr (SLABS > 1) {
                float2* ex = reinterpret_cast<float2*>(smem_raw + (ln_exch - raw));
                const int par = (lt / TEAMS) & 1;
                ex[((par * 4 + team * SLABS + slab) * 4 + q) * 32 + lane] = make_float2(s1, s2);
                asm volatile("bar.sync %0, %1;" ::"r"(1 + q + 4 * team), "r"(SLABS * 32) : "memory");
                s1 = 0.f; s2 = 0.f;
#pragma unroll
                for (int sb = 0; sb < SLABS; ++sb) {
                  const float2 e2 = ex[((par * 4 + team * SLABS + sb) * 4 + q) * 32 + lane];
                  s1 += e2.x; s2 += e2.y;
                }
              }
              const float mean = s1 * (1.0f / BN);
              const float rstd = rsqrtf(fmaxf(s2 * (1.0f / BN) - mean * mean, 0.f) + 1e-5f);
              const float* modrow = nullptr;
              if (p.ln_mod) {
                const int Hm = p.ln_H - 1, lgH = 31 - __clz(p.ln_H);
                const int hw = row & (p.ln_H * p.ln_H - 1);
                const int hs = ((hw >> lgH) - p.ln_shift) & Hm, ws = ((hw & Hm) - p.ln_shift) & Hm;
                modrow = p.ln_mod + (size_t)(((hs & 7) << 3) | (ws & 7)) * BN + n;
              }
              const float4* g4 = reinterpret_cast<const float4*>(p.ln_gamma + n);
              const float4* be4 = reinterpret_cast<const float4*>(p.ln_beta + n);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 gg = __ldg(g4 + j), bb = __ldg(be4 + j);
                f[4 * j] = fmaf((f[4 * j] - mean) * rstd, gg.x, bb.x);
                f[4 * j + 1] = fmaf((f[4 * j + 1] - mean) * rstd, gg.y, bb.y);
                f[4 * j + 2] = fmaf((f[4 * j + 2] - mean) * rstd, gg.z, bb.z);
                f[4 * j + 3] = fmaf((f[4 * j + 3] - mean) * rstd, gg.w, bb.w);
                if (modrow) {
                  const float4 mm = __ldg(reinterpret_cast<const float4*>(modrow) + j);
                  f[4 * j] += mm.x; f[4 * j + 1] += mm.y; f[4 * j + 2] += mm.z; f[4 * j + 3] += mm.w;
                }
              }
              uint32_t buf2 = staging2 + (uint32_t)ew * stg2;                 // free: every earlier store was waited for above
              if (p.ln_reuse) {                               // the x store issued above must have read `buf` first
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
                buf2 = buf;
              }
              if (p.ln_split) {                               // rows [hi(N) | lo(N)]: two boxes
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint32_t h0, h1, h2, h3, l0, l1, l2, l3;
                  split_pack2(f[8 * j], f[8 * j + 1], h0, l0);
                  split_pack2(f[8 * j + 2], f[8 * j + 3], h1, l1);
                  split_pack2(f[8 * j + 4], f[8 * j + 5], h2, l2);
                  split_pack2(f[8 * j + 6], f[8 * j + 7], h3, l3);
                  const uint32_t off = (uint32_t)lane * 64u + ((((uint32_t)j) ^ ((uint32_t)(lane >> 1) & 3u)) << 4);
                  st_shared_v4(buf2 + off, h0, h1, h2, h3);
                  st_shared_v4(buf2 + STG2_BYTES + off, l0, l1, l2, l3);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                  tma_store_2d(&tmD, buf2, n, m0 + q * 32);
                  tma_store_2d(&tmD, buf2 + STG2_BYTES, BN + n, m0 + q * 32);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint32_t w0, w1, w2, w3;
                  pack8_16(f + 8 * j, p.f16 != 0, w0, w1, w2, w3);
                  st_shared_v4(buf2 + (uint32_t)lane * 64u + ((((uint32_t)j) ^ ((uint32_t)(lane >> 1) & 3u)) << 4), w0, w1, w2, w3);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) tma_store_2d(&tmD, buf2, n, m0 + q * 32);
              }
            }
          }
        }
        // all tcgen05.ld of this accumulator have completed (tmem_ld32 waits): hand it back
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar + 8u * acc);
        if (rtma) {
          int m1, n1;
          if (tile_at(lt + TEAMS, m1, n1) && lane == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the stores have read buf (and buf2)
            mbar_arrive_expect_tx(rbar, 4096u);
            tma_load_2d(buf, &tmR, n1 + slab * CPW, m1 + q * 32, rbar);
          }
        }
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

// ------------------------------------------------------------------------------------- host
}  // namespace

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tensor_map(CUtensorMap* map, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, bool f32, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return WMK_ERR_CUDA;
  }
  cuuint64_t d[5], st[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  const CUtensorMapSwizzle swz = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                 : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                 : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank,
                  const_cast<void*>(ptr), d, st, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d), rank %d, dims %llu x %llu, box %u x %u", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return WMK_ERR_CUDA;
  }
  return 0;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace tc

namespace {

// 2-D row-major [rows][cols] tensor of bf16 (or fp32), box = box_cols x box_rows
int make_map_ex(CUtensorMap* map, const void* ptr, int rows, int cols, int box_rows, int box_cols, bool f32) {
  const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
  const uint64_t strides[1] = {(uint64_t)cols * (f32 ? 4 : 2)};
  const uint32_t box[2] = {(uint32_t)box_cols, (uint32_t)box_rows};
  return make_tensor_map(map, ptr, 2, dims, strides, box, f32, box_cols * (f32 ? 4 : 2) == 128 ? 128 : 64);
}
int make_map(CUtensorMap* map, const void* ptr, int rows, int K, int box_rows) {
  return make_map_ex(map, ptr, rows, K, box_rows, BK, false);
}

static int narrow_staging() {      // 0 off, 1 weight-stationary launches, 2 also streamed-weight launches with < 4 stages
  static const int v = getenv("WMK_GEMM_NARROW_STAGING") ? atoi(getenv("WMK_GEMM_NARROW_STAGING")) : 2;
  return v;
}

template <int BN, int EPI, bool OUT_BF16, bool LN = false>
int launch_persistent_t(const GemmArgs& g, cudaStream_t st) {
  CUtensorMap tmA, tmW, tmC, tmD;
  if (g.conv_H > 0) {
    const uint64_t dims[4] = {64, 128, (uint64_t)g.conv_H, (uint64_t)g.conv_B};
    const uint64_t strides[3] = {128, 128ull * 128, 128ull * 128 * (uint64_t)g.conv_H};
    const uint32_t box[4] = {64, 128, 1, 1};
    WMK_TRY(make_tensor_map(&tmA, g.A, 4, dims, strides, box, false, 128));
  } else if (g.dn_Ho > 0) {
    // S[dn_B][Ho + 1][Ho + 1][Ck], Ck = 4C (split: 8C = hi | lo); box = 64 channels x the 128 output pixels of a row tile
    const int Ho = g.dn_Ho, Ck = (g.K / 4) * (g.split ? 2 : 1);
    const int rows = Ho * Ho >= BM ? BM / Ho : Ho, imgs = BM / (Ho * rows);
    const uint64_t dims[4] = {(uint64_t)Ck, (uint64_t)(Ho + 1), (uint64_t)(Ho + 1), (uint64_t)g.dn_B};
    const uint64_t strides[3] = {(uint64_t)Ck * 2, (uint64_t)(Ho + 1) * Ck * 2, (uint64_t)(Ho + 1) * (Ho + 1) * Ck * 2};
    const uint32_t box[4] = {64, (uint32_t)Ho, (uint32_t)rows, (uint32_t)imgs};
    WMK_TRY(make_tensor_map(&tmA, g.A, 4, dims, strides, box, false, 128));
  } else {
    WMK_TRY(make_map(&tmA, g.A, g.M, g.split ? 2 * g.K : g.K, BM));
  }
  const int split_mode = g.split ? (g.K == 32 ? 2 : 1) : g.wsplit ? (g.K == 32 ? 4 : 3) : 0;
  WMK_TRY(make_map(&tmW, g.W, g.N, split_mode == 2 ? 128 : split_mode == 0 ? g.K : 2 * g.K, BN));
  const int KBl = cdiv(g.K, BK);
  const int kblocks = split_mode == 2 ? 1 : KBl;                                   // stages per output tile
  const int wblocks = (split_mode == 1 || split_mode == 3) ? 2 * KBl : split_mode == 2 ? 2 : KBl;       // distinct W tiles
  const int a_stage = (split_mode == 1 ? 2 : 1) * BM * BK * 2;
  const int w_stage = ((split_mode >= 1 && split_mode <= 3) ? 2 : 1) * BN * BK * 2;
  constexpr int budget = 226 * 1024;
  const int n_tiles = g.N / BN, m_tiles = cdiv(g.M, BM);
  const int w_bytes = wblocks * BN * BK * 2;
  int boxc = epi_box_cols(BN, OUT_BF16, LN), ws = 0, n_stages = 0;
  size_t smem = 0;
  for (;;) {
    const int stg = OUT_BF16 ? 64 * boxc : 4096;
    const bool ln_reuse = g.split || g.wsplit || g.ln_split;
    const int fixed = kEpiWarps * stg + 1024 + 512 + (LN ? (int)((ln_reuse ? 0 : kEpiWarps * STG2_BYTES) + LN_EXCH_BYTES) : 0);
    // weight-stationary when the whole BN x K tile + >= 2 A stages fit and every N tile gets >= 1 CTA
    ws = (g.conv_H == 0 && g.dn_Ho == 0 && wblocks <= (split_mode ? 8 : 4) && w_bytes + 2 * a_stage + fixed <= budget && n_tiles <= num_sms() &&
          m_tiles >= 2 * (num_sms() / n_tiles)) ? 1 : 0;
    const int stage = ws ? a_stage : a_stage + w_stage;
    const int avail = budget - fixed - (ws ? w_bytes : 0);
    n_stages = ws ? 6 : (kblocks < 6 ? (kblocks < 2 ? 2 : kblocks) : 6);
    while (n_stages > 2 && n_stages * stage > avail) --n_stages;
    smem = (size_t)n_stages * stage + (ws ? w_bytes : 0) + fixed;
    // a resident 128 KB weight tile leaves two 16 KB A stages = 32 KB in flight per SM, which bounds the kernel
    // by load latency (3.1 TB/s measured): halve the bf16 staging boxes to make room for a 4-deep ring
    if ((ws ? narrow_staging() >= 1 : narrow_staging() >= 2) && n_stages < 4 && OUT_BF16 && boxc == 64) { boxc = 32; continue; }
    break;
  }
  if (smem > 227 * 1024) {
    set_error("gemm_bf16: %zu bytes of shared memory needed (M %d N %d K %d, tile %d, split %d)", smem, g.M, g.N, g.K, BN, split_mode);
    return WMK_ERR_UNSUPPORTED;
  }
  WMK_TRY(make_map_ex(&tmC, g.C, g.M, g.ldc, 32, boxc, !OUT_BF16));
  if (LN) WMK_TRY(make_map_ex(&tmD, g.ln_out, g.M, g.ln_split ? 2 * g.N : g.N, 32, 32, false));
  else tmD = tmC;
  CUtensorMap tmR = tmC;                                   // residual tiles (fp32, same geometry as C; usually C itself)
  if (EPI == EPI_BIAS_RESID && g.resid && (const void*)g.resid != (const void*)g.C)
    WMK_TRY(make_map_ex(&tmR, g.resid, g.M, g.ldc, 32, 32, true));
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_persistent_kernel<BN, EPI, OUT_BF16, LN>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  EpiParams p{g.bias, g.resid, g.C, g.M, g.N, g.ldc, g.epi, g.out_bf16, g.up_h, g.up_w, g.up_cout};
  p.ln_gamma = g.ln_gamma; p.ln_beta = g.ln_beta; p.ln_mod = g.ln_mod; p.ln_H = g.ln_H; p.ln_shift = g.ln_shift;
  p.conv_H = g.conv_H;
  p.dn_Ho = g.dn_Ho;
  p.dn_chunks = g.dn_Ho > 0 ? g.K / 4 / BK : 0;
  p.boxc = boxc;
  static const int resid_pf = getenv("WMK_GEMM_RESID_PREFETCH") ? atoi(getenv("WMK_GEMM_RESID_PREFETCH")) : 1;
  p.resid_prefetch = resid_pf;
  p.gelu_half = g.gelu_half;
  p.gelu_exact = g.gelu_exact;
  static const int gelu_mix = getenv("WMK_GELU_MIX") ? atoi(getenv("WMK_GELU_MIX")) : 0;
  p.gelu_mix = gelu_mix;
  p.split = split_mode;
  p.f16 = g.f16 || g.wsplit;
  p.ln_split = g.ln_split;
  p.ln_reuse = (g.split || g.wsplit || g.ln_split) ? 1 : 0;

  const long long total = (long long)m_tiles * n_tiles;
  const int grid = ws ? (num_sms() / n_tiles) * n_tiles : (int)(total < num_sms() ? total : num_sms());
  gemm_tcgen05_persistent_kernel<BN, EPI, OUT_BF16, LN><<<grid, kPThreads, smem, st>>>(tmA, tmW, tmC, tmD, tmR, p, g.K, m_tiles,
                                                                                       n_tiles, n_stages, ws);
  WMK_CHECK_LAUNCH("gemm_tcgen05_persistent_kernel");
  return 0;
}

template <int BN>
int launch_persistent(const GemmArgs& g, cudaStream_t st) {
  if (g.epi == EPI_BIAS_GELU) {
    if (g.out_bf16) return launch_persistent_t<BN, EPI_BIAS_GELU, true>(g, st);
    return launch_persistent_t<BN, EPI_BIAS_GELU, false>(g, st);      // split-bf16 plans keep the hidden tensor in fp32
  }
  if (g.epi == EPI_BIAS_RELU) {
    if constexpr (BN <= 64) {
      if (g.out_bf16) return launch_persistent_t<BN, EPI_BIAS_RELU, true>(g, st);
    }
    set_error("gemm_bf16: the ReLU epilogue covers N = 32 / 64 with bf16 output");
    return WMK_ERR_UNSUPPORTED;
  }
  if (g.epi == EPI_BIAS_RESID) {
    if constexpr (BN <= 128) {
      if (g.ln_out) {
        if (g.N != BN || !g.ln_gamma || !g.ln_beta || (g.ln_mod && (g.ln_H & (g.ln_H - 1)))) {
          set_error("gemm_bf16: fused LayerNorm needs N == tile width (%d), gamma/beta and a power-of-two H", BN);
          return WMK_ERR_ARG;
        }
        return launch_persistent_t<BN, EPI_BIAS_RESID, false, true>(g, st);
      }
    }
    return launch_persistent_t<BN, EPI_BIAS_RESID, false>(g, st);
  }
  if (g.out_bf16) return launch_persistent_t<BN, EPI_BIAS, true>(g, st);
  return launch_persistent_t<BN, EPI_BIAS, false>(g, st);
}

template <int BN>
int launch(const GemmArgs& g, cudaStream_t st) {
  CUtensorMap tmA, tmW;
  WMK_TRY(make_map(&tmA, g.A, g.M, g.K, BM));
  WMK_TRY(make_map(&tmW, g.W, g.N, g.K, BN));
  const int kblocks = cdiv(g.K, BK);
  constexpr int stage = (BM + BN) * BK * 2;
  int n_stages = kblocks < 4 ? kblocks : 4;
  while (n_stages > 2 && n_stages * stage > 96 * 1024) --n_stages;   // keep >= 2 CTAs per SM
  const size_t smem = (size_t)n_stages * stage + 1024 /*align*/ + 16 * n_stages + 16;
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        200 * 1024));
    attr_set = true;
  }
  EpiParams p{g.bias, g.resid, g.C, g.M, g.N, g.ldc, g.epi, g.out_bf16, g.up_h, g.up_w, g.up_cout};
  p.f16 = g.f16;
  const int n_tiles = g.N / BN;
  const long long grid = (long long)cdiv(g.M, BM) * n_tiles;
  gemm_tcgen05_kernel<BN><<<(unsigned)grid, kThreads, smem, st>>>(tmA, tmW, p, g.K, n_tiles, n_stages);
  WMK_CHECK_LAUNCH("gemm_tcgen05_kernel");
  return 0;
}

}  // namespace

int gemm_bf16_tcgen05(const GemmArgs& g, cudaStream_t st) {
  WMK_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem %dx%dx%d", g.M, g.N, g.K);
  WMK_REQUIRE(g.K % 8 == 0, "gemm_bf16: K=%d must be a multiple of 8 (16-byte TMA rows)", g.K);
  WMK_REQUIRE(g.N % 32 == 0, "gemm_bf16: N=%d must be a multiple of 32", g.N);
  WMK_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0 && ((uintptr_t)g.C & 15) == 0,
              "gemm_bf16: operands must be 16-byte aligned");
  WMK_REQUIRE(g.ldc % 8 == 0, "gemm_bf16: ldc=%d must be a multiple of 8", g.ldc);
  if (g.conv_H > 0)
    WMK_REQUIRE(g.K == 576 && g.M == g.conv_B * g.conv_H * 128 && g.ldc == g.N && g.epi != EPI_UPSAMPLE,
                "gemm_bf16: implicit-GEMM conv needs K = 576 and M = B*H*128 (W = 128)");
  if (g.dn_Ho > 0)
    WMK_REQUIRE(g.conv_H == 0 && !g.wsplit && g.K % 256 == 0 && g.dn_Ho >= 8 && g.dn_Ho <= 128 && (g.dn_Ho & (g.dn_Ho - 1)) == 0 &&
                    g.M == g.dn_B * g.dn_Ho * g.dn_Ho && g.ldc == g.N && g.epi != EPI_UPSAMPLE,
                "gemm_bf16: implicit-GEMM downsample needs K = 16C (C %% 16 == 0), a power-of-two output side in [8, 128] and M = B*Ho*Ho");
  if (g.split || g.wsplit)
    WMK_REQUIRE((g.K == 32 || g.K % 64 == 0) && g.conv_H == 0 && g.epi != EPI_UPSAMPLE && g.ldc == g.N && !(g.split && g.wsplit),
                "gemm_bf16: split operands need K = 32 or K %% 64 == 0 (K = %d) and a plain row-major C", g.K);
  if (g.epi == EPI_UPSAMPLE)
    WMK_REQUIRE(g.up_cout % 32 == 0 && g.N == 4 * g.up_cout && g.M % (g.up_h * g.up_w) == 0,
                "gemm_bf16: bad upsample geometry");
  const GemmWork gw = gemm_work(g, 2);
  ProfScope prof(gw.family, gw.work, st, gw.work2);
  if (g.epi == EPI_BIAS && g.ln_out && !g.out_bf16) {
    // bias-only layer whose finished row is also LayerNormed (the downsample conv feeding the next stage's norm1):
    // the residual epilogue with no residual
    GemmArgs h = g;
    h.epi = EPI_BIAS_RESID;
    h.resid = nullptr;
    WMK_REQUIRE(h.ldc == h.N && (h.N == 32 || h.N == 64 || h.N == 128), "gemm_bf16: fused LayerNorm needs N in {32, 64, 128}");
    if (h.N == 128) return launch_persistent<128>(h, st);
    if (h.N == 64) return launch_persistent<64>(h, st);
    return launch_persistent<32>(h, st);
  }
  if (g.epi != EPI_UPSAMPLE && g.ldc == g.N) {
    const bool wide_ok = g.epi != EPI_BIAS_RESID || g.K >= 1024;   // fp32 residual tiles: one 32-column piece per warp
    // split operands double the W (and A) stage: 256-wide tiles would leave no room for a second stage
    if (g.N % 256 == 0 && g.N >= 256 && wide_ok && !g.split && !g.wsplit) return launch_persistent<256>(g, st);
    if (g.N % 128 == 0) return launch_persistent<128>(g, st);
    // N = 96 / 192 (the fused q|k|v projection of the C = 32 / 64 stages): one / two 96-wide tiles instead of three
    // 32- / 64-wide ones (A is fetched once, 12 of the 16 epilogue warps work instead of 4 / 8)
    if (g.N % 96 == 0 && g.epi == EPI_BIAS) return launch_persistent<96>(g, st);
    if (g.N % 64 == 0) return launch_persistent<64>(g, st);
    return launch_persistent<32>(g, st);
  }
  if (g.N % 128 == 0) return launch<128>(g, st);
  if (g.N % 96 == 0) return launch<96>(g, st);
  if (g.N % 64 == 0) return launch<64>(g, st);
  return launch<32>(g, st);
}

}  // namespace wmk

// fp32 dense layer on the CUDA cores: the reference-precision ("1e-3 parity") mode of every
// nn.Linear / im2col convolution of the Uformer blocks (uformerWM/model.py:455-456,518,686,690).
//   C[M][N] = epilogue( A[M][K] * W[N][K]^T + bias )      A, W fp32 K-major, fp32 accumulate
// 64x64 output tile per 256-thread CTA, 16-deep k-slices staged in shared memory (transposed so
// the inner product reads are conflict-free), 4x4 register micro-tile per thread.
#include "wmk_common.cuh"

namespace wmk {

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ void epi_store1(const EpiParams& p, int m, int n, float acc) {
  if (p.bias) acc += __ldg(p.bias + n);
  const size_t off = epi_row_offset(p, m, n);
  if (p.epi == EPI_BIAS_GELU) acc = gelu_erf(acc);
  else if (p.epi == EPI_BIAS_RESID) acc += p.resid[off];
  if (p.out_bf16) reinterpret_cast<__nv_bfloat16*>(p.C)[off] = __float2bfloat16(acc);
  else reinterpret_cast<float*>(p.C)[off] = acc;
}

__global__ void __launch_bounds__(256)
gemm_fp32_kernel(const float* __restrict__ A, const float* __restrict__ W, EpiParams p, int K, int n_tiles) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Ws[TK][TN + 4];
  const int tile = blockIdx.x;
  const int m0 = (tile / n_tiles) * TM, n0 = (tile % n_tiles) * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;          // 16 x 16 threads, 4x4 outputs each
  const int lr = tid >> 2, lc = (tid & 3) * 4;     // loader: 64 rows x 4 float4
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    {
      float4 a = make_float4(0, 0, 0, 0), w = make_float4(0, 0, 0, 0);
      const int m = m0 + lr, n = n0 + lr, k = k0 + lc;
      if (m < p.M && k < K) a = *reinterpret_cast<const float4*>(A + (size_t)m * K + k);
      if (n < p.N && k < K) w = __ldg(reinterpret_cast<const float4*>(W + (size_t)n * K + k));
      As[lc][lr] = a.x; As[lc + 1][lr] = a.y; As[lc + 2][lr] = a.z; As[lc + 3][lr] = a.w;
      Ws[lc][lr] = w.x; Ws[lc + 1][lr] = w.y; Ws[lc + 2][lr] = w.z; Ws[lc + 3][lr] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 w = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.N) epi_store1(p, m, n, acc[i][j]);
    }
  }
}

}  // namespace

int gemm_fp32_simt(const GemmArgs& g, cudaStream_t st) {
  WMK_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem %dx%dx%d", g.M, g.N, g.K);
  WMK_REQUIRE(g.K % 4 == 0, "gemm_fp32: K=%d must be a multiple of 4", g.K);
  WMK_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0, "gemm_fp32: operands must be 16-byte aligned");
  EpiParams p{g.bias, g.resid, g.C, g.M, g.N, g.ldc, g.epi, g.out_bf16, g.up_h, g.up_w, g.up_cout};
  const GemmWork gw = gemm_work(g, 4);
  ProfScope prof(gw.family, gw.work, st, gw.work2);
  const int n_tiles = cdiv(g.N, TN);
  const long long grid = (long long)cdiv(g.M, TM) * n_tiles;
  gemm_fp32_kernel<<<(unsigned)grid, 256, 0, st>>>(reinterpret_cast<const float*>(g.A),
                                                    reinterpret_cast<const float*>(g.W), p, g.K, n_tiles);
  WMK_CHECK_LAUNCH("gemm_fp32_kernel");
  return 0;
}

}  // namespace wmk

// Training-mode LeWin block: forward with the activations its gradient needs, and the full backward pass
// (uformerWM/model.py:937-1019 LeWinTransformerBlock, :460-471 / :523-551 projection + window attention with the
// relative-position table and the shift mask, :683-714 LeFF), fp32 on the CUDA cores - the building block of the
// UformerAudio training step (uformerWM/audio_uformer_stft.py:418-549).  Reference-precision kernels:
// they pin the gradient of every operator of the block against autograd of the oracle; the tensor-core forms of the
// two GEMM gradients (dX = dY W, dW = dY^T X) are the dense kernel run on transposed operands and are not built yet.
//
//   x -> LN1 (+ modulator by window position) -> q|k|v -> window attention -> proj -> + x = x1
//   x1 -> LN2 -> linear1 -> GELU -> depthwise 3x3 -> GELU -> linear2 -> + x1 = out
// Everything is token layout [n * H * H][C]; roll / window partition / reverse are index arithmetic (att_row).
#include <vector>

#include "uformer_kernels.cuh"

namespace wmk {
namespace {

constexpr float kInvSqrt2 = 0.70710678118654752440f, kInvSqrt2Pi = 0.39894228040143267794f;

__global__ void __launch_bounds__(256) gelu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = gelu_erf(x[i]);
}
// dx = dy * d/dx [x Phi(x)] = dy * (Phi(x) + x phi(x))
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const float* __restrict__ xpre, const float* __restrict__ dy, float* __restrict__ dx, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = xpre[i];
  const float cdf = 0.5f * (1.0f + erff(x * kInvSqrt2));
  dx[i] = dy[i] * (cdf + x * kInvSqrt2Pi * expf(-0.5f * x * x));
}

// depthwise 3x3, padding 1, token layout [B][H][H][Ch]; w [Ch][9]; flip: correlation with the reversed taps (data gradient)
__global__ void __launch_bounds__(256)
dwconv3x3_plain_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
                       const float* __restrict__ bias, int B, int H, int Ch, int flip) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * H * H * Ch) return;
  const int c = (int)(idx % Ch);
  const size_t pix = idx / Ch;
  const int wq = (int)(pix % H), h = (int)((pix / H) % H);
  const size_t b = pix / ((size_t)H * H);
  float a = bias ? bias[c] : 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int hh = h + t / 3 - 1, ww = wq + t % 3 - 1;
    if (hh < 0 || hh >= H || ww < 0 || ww >= H) continue;
    a = fmaf(in[((b * H + hh) * H + ww) * Ch + c], w[c * 9 + (flip ? 8 - t : t)], a);
  }
  out[idx] = a;
}
// dw[c][t] = sum_pixels in[pixel + tap t][c] dy[pixel][c];  db[c] = sum dy.  grid (chunks, Ch / 32), block (32 channels, 8 pixel lanes)
__global__ void __launch_bounds__(256)
dwconv3x3_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dy, float* __restrict__ dw,
                       float* __restrict__ db, int B, int H, int Ch) {
  const int c = blockIdx.y * 32 + (threadIdx.x & 31);
  const int sub = threadIdx.x >> 5;
  const size_t npix = (size_t)B * H * H;
  float acc[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t] = 0.f;
  for (size_t pix = (size_t)blockIdx.x * 8 + sub; pix < npix; pix += (size_t)gridDim.x * 8) {
    const int wq = (int)(pix % H), h = (int)((pix / H) % H);
    const size_t b = pix / ((size_t)H * H);
    const float d = dy[pix * Ch + c];
    acc[9] += d;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = wq + t % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= H) continue;
      acc[t] = fmaf(in[((b * H + hh) * H + ww) * Ch + c], d, acc[t]);
    }
  }
  __shared__ float red[8][32][10];
#pragma unroll
  for (int t = 0; t < 10; ++t) red[sub][threadIdx.x & 31][t] = acc[t];
  __syncthreads();
  if (sub == 0) {
#pragma unroll
    for (int t = 0; t < 10; ++t) {
      float s = 0.f;
      for (int q = 0; q < 8; ++q) s += red[q][threadIdx.x][t];
      if (t < 9) atomicAdd(dw + c * 9 + t, s);
      else atomicAdd(db + c, s);
    }
  }
}

// out[n] (+)= sum_m a[m][n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, float* __restrict__ out, int M, int N) {
  const int n = blockIdx.y * 32 + (threadIdx.x & 31), sub = threadIdx.x >> 5;
  float s = 0.f;
  if (n < N)
    for (int m = blockIdx.x * 8 + sub; m < M; m += gridDim.x * 8) s += a[(size_t)m * N + n];
  __shared__ float red[8][32];
  red[sub][threadIdx.x & 31] = s;
  __syncthreads();
  if (sub == 0 && n < N) {
    float t = 0.f;
    for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
    atomicAdd(out + n, t);
  }
}

__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ a, float* __restrict__ at, int R, int Cc) {
  __shared__ float t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < R && c0 + tx < Cc) t[i][tx] = a[(size_t)(r0 + i) * Cc + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < Cc && r0 + tx < R) at[(size_t)(c0 + i) * R + r0 + tx] = t[tx][i];
}

// C[N][K] += A^T B for A [M][N], B [M][K] (weight gradient dW = dY^T X): 64 x 64 output tile per CTA, 4 x 4 outputs per
// thread (two 16-byte shared-memory loads per 16 FMAs), 16 rows of M per step, M split over grid.z (atomics at the end)
__global__ void __launch_bounds__(256)
gemm_tn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K, int m_per) {
  __shared__ __align__(16) float As[16][68], Bs[16][68];
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int m_begin = blockIdx.z * m_per, m_end = min(M, m_begin + m_per);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;        // outputs (n0 + 4 ty + i, k0 + 4 tx + j)
  const int lr = threadIdx.x >> 4, lc = (threadIdx.x & 15) * 4;  // loader: row lr of the 16, columns lc .. lc + 3
  const bool vecA = (N & 3) == 0, vecB = (K & 3) == 0;
  float acc[4][4] = {};
  for (int m0 = m_begin; m0 < m_end; m0 += 16) {
    const int m = m0 + lr;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (m < m_end) {
      const float* ar = A + (size_t)m * N + n0 + lc;
      const float* br = B + (size_t)m * K + k0 + lc;
      if (vecA && n0 + lc + 3 < N) a = *reinterpret_cast<const float4*>(ar);
      else {
        if (n0 + lc < N) a.x = ar[0];
        if (n0 + lc + 1 < N) a.y = ar[1];
        if (n0 + lc + 2 < N) a.z = ar[2];
        if (n0 + lc + 3 < N) a.w = ar[3];
      }
      if (vecB && k0 + lc + 3 < K) b = *reinterpret_cast<const float4*>(br);
      else {
        if (k0 + lc < K) b.x = br[0];
        if (k0 + lc + 1 < K) b.y = br[1];
        if (k0 + lc + 2 < K) b.z = br[2];
        if (k0 + lc + 3 < K) b.w = br[3];
      }
    }
    *reinterpret_cast<float4*>(&As[lr][lc]) = a;
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = b;
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < 16; ++mm) {
      const float4 av = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[mm][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(C + (size_t)n * K + k, acc[i][j]);
    }
  }
}

// LayerNorm backward for y = xhat * gamma + beta (+ modulator[window position]): one warp per token.
//   dx_acc[m][:] += rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
//   dgamma += dy * xhat, dbeta += dy, dmod[pos][:] += dy   (CTA-level partial sums, then atomics)
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
              float* __restrict__ dx_acc, float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dmod,
              int M, int C, int H, int shift) {
  __shared__ float sg[512], sb[512];                 // the CTA's 8 tokens: partial dgamma / dbeta (C <= 512)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < C; c += 256) { sg[c] = 0.f; sb[c] = 0.f; }
  __syncthreads();
  const int token = blockIdx.x * 8 + warp;
  if (token < M) {
    const float* xr = x + (size_t)token * C;
    const float* dr = dy + (size_t)token * C;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += xr[c];
    s = warp_sum(s);
    const float mean = s / C;
    float q = 0.f;
    for (int c = lane; c < C; c += 32) { const float d = xr[c] - mean; q = fmaf(d, d, q); }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / C + 1e-5f);
    float g1 = 0.f, g2 = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = dr[c] * gamma[c], xh = (xr[c] - mean) * rstd;
      g1 += g;
      g2 = fmaf(g, xh, g2);
    }
    g1 = warp_sum(g1) / C;
    g2 = warp_sum(g2) / C;
    float* mrow = nullptr;
    if (dmod) {
      const int hw = token % (H * H);
      const int h = hw / H, w = hw - h * H;
      const int hs = (h - shift + H) % H, ws = (w - shift + H) % H;
      mrow = dmod + (size_t)(((hs & 7) << 3) | (ws & 7)) * C;
    }
    for (int c = lane; c < C; c += 32) {
      const float d = dr[c], xh = (xr[c] - mean) * rstd;
      dx_acc[(size_t)token * C + c] += rstd * (d * gamma[c] - g1 - xh * g2);
      atomicAdd(&sg[c], d * xh);
      atomicAdd(&sb[c], d);
      if (mrow) atomicAdd(mrow + c, d);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    atomicAdd(dgamma + c, sg[c]);
    atomicAdd(dbeta + c, sb[c]);
  }
}

// ---- window attention, one CTA (64 threads, thread = query row / key row) per (window, head); q|k|v UNSCALED,
// scores = scale * q k^T + table[rel_idx(i, j)][head] + mask
struct AttnTrainGeom { AttGeom g; int heads; float scale; };

__device__ __forceinline__ int rel_idx(int i, int j) { return ((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7); }

template <bool BWD>
__global__ void __launch_bounds__(64)
attn_train_kernel(const float* __restrict__ qkv, const float* __restrict__ table, float* __restrict__ out,
                  const float* __restrict__ dO, float* __restrict__ dqkv, float* __restrict__ dtable, AttnTrainGeom a) {
  extern __shared__ float sm[];
  float* Qs = sm;                  // [64][33]
  float* Ks = Qs + 64 * 33;
  float* Vs = Ks + 64 * 33;
  float* P = Vs + 64 * 33;         // [64][65] probabilities (BWD: then dS)
  float* Ds = P + 64 * 65;         // BWD: dO rows [64][33]
  __shared__ int tok[64], rid[64];
  const int win = blockIdx.x, head = blockIdx.y, i = threadIdx.x;
  const int C = a.g.C;
  {
    int t, r;
    att_row(a.g, win, i, t, r);
    tok[i] = t; rid[i] = r;
    const float* row = qkv + (size_t)t * 3 * C + head * 32;
    for (int d = 0; d < 32; ++d) { Qs[i * 33 + d] = row[d]; Ks[i * 33 + d] = row[C + d]; Vs[i * 33 + d] = row[2 * C + d]; }
    if (BWD) {
      const float* drow = dO + (size_t)t * C + head * 32;
      for (int d = 0; d < 32; ++d) Ds[i * 33 + d] = drow[d];
    }
  }
  __syncthreads();
  // row i of the scores
  float q[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) q[d] = Qs[i * 33 + d];
  float mx = -INFINITY;
  for (int j = 0; j < 64; ++j) {
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) s = fmaf(q[d], Ks[j * 33 + d], s);
    s = s * a.scale + table[rel_idx(i, j) * a.heads + head];
    if (a.g.shift > 0 && rid[i] != rid[j]) s -= 100.0f;
    P[i * 65 + j] = s;
    mx = fmaxf(mx, s);
  }
  float sum = 0.f;
  for (int j = 0; j < 64; ++j) { const float e = expf(P[i * 65 + j] - mx); P[i * 65 + j] = e; sum += e; }
  const float inv = 1.0f / sum;
  for (int j = 0; j < 64; ++j) P[i * 65 + j] *= inv;
  if (!BWD) {
    float o[32];
#pragma unroll
    for (int d = 0; d < 32; ++d) o[d] = 0.f;
    for (int j = 0; j < 64; ++j) {
      const float p = P[i * 65 + j];
#pragma unroll
      for (int d = 0; d < 32; ++d) o[d] = fmaf(p, Vs[j * 33 + d], o[d]);
    }
    float* orow = out + (size_t)tok[i] * C + head * 32;
#pragma unroll
    for (int d = 0; d < 32; ++d) orow[d] = o[d];
    return;
  }
  __syncthreads();                 // P complete (column reads below)
  // dV_i = sum_r P[r][i] dO_r   (thread i as key row)
  float acc[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
  for (int r = 0; r < 64; ++r) {
    const float p = P[r * 65 + i];
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = fmaf(p, Ds[r * 33 + d], acc[d]);
  }
  float* grow = dqkv + (size_t)tok[i] * 3 * C + head * 32;
#pragma unroll
  for (int d = 0; d < 32; ++d) grow[2 * C + d] = acc[d];
  __syncthreads();                 // everyone has read P's columns before rows are overwritten by dS
  // dS[i][j] = P[i][j] (dP[i][j] - sum_j dP[i][j] P[i][j]),  dP[i][j] = dO_i . V_j
  float dor[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) dor[d] = Ds[i * 33 + d];
  float dot = 0.f;
  float* Prow = P + i * 65;
  // first pass: dP into registers is too large; recompute in two passes over j
  for (int j = 0; j < 64; ++j) {
    float dp = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) dp = fmaf(dor[d], Vs[j * 33 + d], dp);
    dot = fmaf(dp, Prow[j], dot);
  }
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;            // dq_i
  for (int j = 0; j < 64; ++j) {
    float dp = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) dp = fmaf(dor[d], Vs[j * 33 + d], dp);
    const float ds = Prow[j] * (dp - dot);
    Prow[j] = ds;
    atomicAdd(dtable + rel_idx(i, j) * a.heads + head, ds);
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = fmaf(ds, Ks[j * 33 + d], acc[d]);
  }
#pragma unroll
  for (int d = 0; d < 32; ++d) grow[d] = acc[d] * a.scale;
  __syncthreads();                 // dS complete
  // dK_i = scale * sum_r dS[r][i] q_r
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
  for (int r = 0; r < 64; ++r) {
    const float ds = P[r * 65 + i];
#pragma unroll
    for (int d = 0; d < 32; ++d) acc[d] = fmaf(ds, Qs[r * 33 + d], acc[d]);
  }
#pragma unroll
  for (int d = 0; d < 32; ++d) grow[C + d] = acc[d] * a.scale;
}

__global__ void __launch_bounds__(256) add_kernel(float* __restrict__ a, const float* __restrict__ b, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += b[i];
}

// DropPath (stochastic depth, model.py:1016-1017) with a GIVEN per-sample factor s[b] in {0, 1 / keep}:
// out = base + s[b] * br (base NULL: out = s[b] * br); rows of sample b are m in [b * rows, (b + 1) * rows)
__global__ void __launch_bounds__(256)
row_scale_add_kernel(const float* __restrict__ base, const float* __restrict__ br, const float* __restrict__ s, float* __restrict__ out,
                     size_t n, size_t per_sample) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = s[i / per_sample] * br[i];
  out[i] = base ? base[i] + v : v;
}

int grid1(size_t n) { return (int)((n + 255) / 256); }

// stream-ordered temporaries; the device's default pool keeps what it has been given (release threshold = max): without it every
// synchronisation hands the pool back to the driver and the next step pays for ~3000 allocations again
void keep_pool_once() {
  static bool done = false;
  if (done) return;
  int dev = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done = true;
}
struct Scratch {
  cudaStream_t st;
  std::vector<void*> ptrs;
  float* get(size_t n) {
    keep_pool_once();
    void* p = nullptr;
    if (cudaMallocAsync(&p, n * sizeof(float), st) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return reinterpret_cast<float*>(p);
  }
  ~Scratch() { for (void* p : ptrs) cudaFreeAsync(p, st); }
};

int linear_fwd(const float* A, const float* W, const float* b, const float* resid, float* C, int M, int N, int K, cudaStream_t st) {
  GemmArgs g;
  g.A = A; g.W = W; g.bias = b; g.resid = resid; g.C = C; g.M = M; g.N = N; g.K = K; g.ldc = N;
  g.epi = resid ? EPI_BIAS_RESID : EPI_BIAS;
  return gemm_fp32_simt(g, st);
}

// dX (+)= dY W, dW = dY^T X, db = column sums of dY;  W [N][K], X [M][K], dY [M][N]
int linear_bwd(const float* X, const float* W, const float* dY, float* dX, bool dx_accumulate, float* dW, float* db, int M, int N, int K,
               Scratch& sc) {
  cudaStream_t st = sc.st;
  if (dX) {
    float* Wt = sc.get((size_t)N * K);               // [K][N]: the "weight" of the data-gradient GEMM
    if (!Wt) { set_error("lewin_train: scratch allocation failed"); return WMK_ERR_ALLOC; }
    transpose_kernel<<<dim3(cdiv(K, 32), cdiv(N, 32)), 256, 0, st>>>(W, Wt, N, K);
    WMK_CHECK_LAUNCH("transpose_kernel");
    WMK_TRY(linear_fwd(dY, Wt, nullptr, dx_accumulate ? dX : nullptr, dX, M, K, N, st));
  }
  WMK_CHECK_CUDA(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)N * K, st));
  WMK_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * N, st));
  const int tiles_nk = cdiv(N, 64) * cdiv(K, 64);
  int splits = (2 * 148 + tiles_nk - 1) / tiles_nk;                 // ~2 waves of CTAs
  if (splits > M / 64) splits = M / 64 > 0 ? M / 64 : 1;
  const int m_per = cdiv(cdiv(M, splits), 16) * 16;
  gemm_tn_kernel<<<dim3(cdiv(N, 64), cdiv(K, 64), cdiv(M, m_per)), 256, 0, st>>>(dY, X, dW, M, N, K, m_per);
  WMK_CHECK_LAUNCH("gemm_tn_kernel");
  colsum_kernel<<<dim3(M >= 2048 ? 64 : 8, cdiv(N, 32)), 256, 0, st>>>(dY, db, M, N);
  WMK_CHECK_LAUNCH("colsum_kernel");
  return 0;
}

}  // namespace
}  // namespace wmk

using namespace wmk;

// Parameter / gradient slots of one block, in this order (sizes for width C, heads = C / 32 ... any):
enum {
  LP_N1W = 0, LP_N1B, LP_MOD /* [64][C] or NULL */, LP_TABLE /* [225][heads] */, LP_QW /* [C][C] */, LP_QB, LP_KVW /* [2C][C] */, LP_KVB,
  LP_PW, LP_PB, LP_N2W, LP_N2B, LP_L1W /* [4C][C] */, LP_L1B, LP_DWW /* [4C][9] */, LP_DWB, LP_L2W /* [C][4C] */, LP_L2B, LP_COUNT
};

extern "C" int wmk_lewin_block_train_f32(const float* x, const float* dout, const float* const* params, float* const* grads,
                                         float* out, float* dx, int n, int H, int C, int heads, int shift, const float* drop_scales,
                                         void* stream) {
  WMK_REQUIRE(x && params && out && n > 0 && H >= 8 && (H & (H - 1)) == 0 && C >= 32 && C % 32 == 0 && heads * 32 == C &&
                  (C == 32 || C == 64 || C == 128 || C == 256 || C == 512) && (shift == 0 || shift == 4) && (!dout == !dx) && (!dout == !grads),
              "lewin_block_train: bad arguments (H a power of two >= 8, C in {32..512} = 32 heads, shift 0 or 4)");
  for (int i = 0; i < LP_COUNT; ++i)
    WMK_REQUIRE(i == LP_MOD || (params[i] && (!grads || grads[i])), "lewin_block_train: parameter / gradient slot %d is null", i);
  cudaStream_t st = (cudaStream_t)stream;
  if (H <= 8) shift = 0;                                    // model.py:892-894
  const int M = n * H * H, C3 = 3 * C, C4 = 4 * C;
  Scratch sc{st};
  float* a1 = sc.get((size_t)M * C);
  float* wqkv = sc.get((size_t)C3 * C);
  float* bqkv = sc.get(C3);
  float* qkv = sc.get((size_t)M * C3);
  float* O = sc.get((size_t)M * C);
  float* x1 = sc.get((size_t)M * C);
  float* a2 = sc.get((size_t)M * C);
  float* h1p = sc.get((size_t)M * C4);
  float* h1 = sc.get((size_t)M * C4);
  float* h2p = sc.get((size_t)M * C4);
  float* h2 = sc.get((size_t)M * C4);
  if (!a1 || !wqkv || !bqkv || !qkv || !O || !x1 || !a2 || !h1p || !h1 || !h2p || !h2) {
    set_error("lewin_block_train: scratch allocation failed");
    return WMK_ERR_ALLOC;
  }
  const size_t attn_smem = (size_t)(3 * 64 * 33 + 64 * 65 + 64 * 33) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(attn_train_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
    WMK_CHECK_CUDA(cudaFuncSetAttribute(attn_train_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)attn_smem));
    attr = true;
  }
  AttnTrainGeom ag{{C, H, shift, 31 - __builtin_clz((unsigned)(C >> 5)), 31 - __builtin_clz((unsigned)(H >> 3))}, heads,
                   1.0f / sqrtf((float)(C / heads))};
  const int n_windows = n * (H / 8) * (H / 8);
  // ---------------------------------------------------------------- forward
  launch_layernorm<float>(x, a1, params[LP_N1W], params[LP_N1B], params[LP_MOD], M, C, H, shift, st);
  WMK_CHECK_LAUNCH("layernorm_kernel");
  WMK_CHECK_CUDA(cudaMemcpyAsync(wqkv, params[LP_QW], sizeof(float) * (size_t)C * C, cudaMemcpyDeviceToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(wqkv + (size_t)C * C, params[LP_KVW], sizeof(float) * 2 * (size_t)C * C, cudaMemcpyDeviceToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(bqkv, params[LP_QB], sizeof(float) * C, cudaMemcpyDeviceToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(bqkv + C, params[LP_KVB], sizeof(float) * 2 * C, cudaMemcpyDeviceToDevice, st));
  WMK_TRY(linear_fwd(a1, wqkv, bqkv, nullptr, qkv, M, C3, C, st));
  attn_train_kernel<false><<<dim3(n_windows, heads), 64, attn_smem, st>>>(qkv, params[LP_TABLE], O, nullptr, nullptr, nullptr, ag);
  WMK_CHECK_LAUNCH("attn_train_kernel<fwd>");
  const size_t per_sample = (size_t)H * H * C;
  float* br = nullptr;                                      // DropPath: branch output before its per-sample factor
  if (drop_scales) {
    br = sc.get((size_t)M * C);
    if (!br) { set_error("lewin_block_train: scratch allocation failed"); return WMK_ERR_ALLOC; }
    WMK_TRY(linear_fwd(O, params[LP_PW], params[LP_PB], nullptr, br, M, C, C, st));
    row_scale_add_kernel<<<grid1((size_t)M * C), 256, 0, st>>>(x, br, drop_scales, x1, (size_t)M * C, per_sample);
    WMK_CHECK_LAUNCH("row_scale_add_kernel");
  } else {
    WMK_TRY(linear_fwd(O, params[LP_PW], params[LP_PB], x, x1, M, C, C, st));
  }
  launch_layernorm<float>(x1, a2, params[LP_N2W], params[LP_N2B], nullptr, M, C, H, 0, st);
  WMK_CHECK_LAUNCH("layernorm_kernel");
  WMK_TRY(linear_fwd(a2, params[LP_L1W], params[LP_L1B], nullptr, h1p, M, C4, C, st));
  gelu_fwd_kernel<<<grid1((size_t)M * C4), 256, 0, st>>>(h1p, h1, (size_t)M * C4);
  dwconv3x3_plain_kernel<<<grid1((size_t)M * C4), 256, 0, st>>>(h1, h2p, params[LP_DWW], params[LP_DWB], n, H, C4, 0);
  gelu_fwd_kernel<<<grid1((size_t)M * C4), 256, 0, st>>>(h2p, h2, (size_t)M * C4);
  WMK_CHECK_LAUNCH("leff forward kernels");
  if (drop_scales) {
    WMK_TRY(linear_fwd(h2, params[LP_L2W], params[LP_L2B], nullptr, br, M, C, C4, st));
    row_scale_add_kernel<<<grid1((size_t)M * C), 256, 0, st>>>(x1, br, drop_scales + n, out, (size_t)M * C, per_sample);
    WMK_CHECK_LAUNCH("row_scale_add_kernel");
  } else {
    WMK_TRY(linear_fwd(h2, params[LP_L2W], params[LP_L2B], x1, out, M, C, C4, st));
  }
  if (!dout) return 0;
  // ---------------------------------------------------------------- backward
  float* dh2 = sc.get((size_t)M * C4);
  float* dh = sc.get((size_t)M * C4);
  float* da = sc.get((size_t)M * C);
  float* dqkv = sc.get((size_t)M * C3);
  float* dwqkv = sc.get((size_t)C3 * C);
  float* dbqkv = sc.get(C3);
  if (!dh2 || !dh || !da || !dqkv || !dwqkv || !dbqkv) { set_error("lewin_block_train: scratch allocation failed"); return WMK_ERR_ALLOC; }
  // out = x1 + h2 W2^T + b2
  WMK_CHECK_CUDA(cudaMemcpyAsync(dx, dout, sizeof(float) * (size_t)M * C, cudaMemcpyDeviceToDevice, st));      // dx holds d(x1) for now
  const float* dbranch = dout;                              // gradient of the MLP branch output
  if (drop_scales) {
    row_scale_add_kernel<<<grid1((size_t)M * C), 256, 0, st>>>(nullptr, dout, drop_scales + n, br, (size_t)M * C, per_sample);
    WMK_CHECK_LAUNCH("row_scale_add_kernel");
    dbranch = br;
  }
  WMK_TRY(linear_bwd(h2, params[LP_L2W], dbranch, dh2, false, grads[LP_L2W], grads[LP_L2B], M, C, C4, sc));
  gelu_bwd_kernel<<<grid1((size_t)M * C4), 256, 0, st>>>(h2p, dh2, dh, (size_t)M * C4);                          // d(h2p)
  WMK_CHECK_CUDA(cudaMemsetAsync(grads[LP_DWW], 0, sizeof(float) * (size_t)C4 * 9, st));
  WMK_CHECK_CUDA(cudaMemsetAsync(grads[LP_DWB], 0, sizeof(float) * C4, st));
  dwconv3x3_wgrad_kernel<<<dim3(64, C4 / 32), 256, 0, st>>>(h1, dh, grads[LP_DWW], grads[LP_DWB], n, H, C4);
  dwconv3x3_plain_kernel<<<grid1((size_t)M * C4), 256, 0, st>>>(dh, dh2, params[LP_DWW], nullptr, n, H, C4, 1);  // d(h1)
  gelu_bwd_kernel<<<grid1((size_t)M * C4), 256, 0, st>>>(h1p, dh2, dh, (size_t)M * C4);                          // d(h1p)
  WMK_CHECK_LAUNCH("leff backward kernels");
  WMK_TRY(linear_bwd(a2, params[LP_L1W], dh, da, false, grads[LP_L1W], grads[LP_L1B], M, C4, C, sc));           // d(a2)
  for (int s : {LP_N2W, LP_N2B, LP_N1W, LP_N1B}) WMK_CHECK_CUDA(cudaMemsetAsync(grads[s], 0, sizeof(float) * C, st));
  ln_bwd_kernel<<<cdiv(M, 8), 256, 0, st>>>(x1, da, params[LP_N2W], dx, grads[LP_N2W], grads[LP_N2B], nullptr, M, C, H, 0);
  WMK_CHECK_LAUNCH("ln_bwd_kernel");
  // x1 = x + O Wp^T + bp
  dbranch = dx;                                             // gradient of the attention branch output
  if (drop_scales) {
    row_scale_add_kernel<<<grid1((size_t)M * C), 256, 0, st>>>(nullptr, dx, drop_scales, br, (size_t)M * C, per_sample);
    WMK_CHECK_LAUNCH("row_scale_add_kernel");
    dbranch = br;
  }
  WMK_TRY(linear_bwd(O, params[LP_PW], dbranch, da, false, grads[LP_PW], grads[LP_PB], M, C, C, sc));           // d(O)
  WMK_CHECK_CUDA(cudaMemsetAsync(grads[LP_TABLE], 0, sizeof(float) * 225 * heads, st));
  attn_train_kernel<true><<<dim3(n_windows, heads), 64, attn_smem, st>>>(qkv, params[LP_TABLE], nullptr, da, dqkv, grads[LP_TABLE], ag);
  WMK_CHECK_LAUNCH("attn_train_kernel<bwd>");
  WMK_TRY(linear_bwd(a1, wqkv, dqkv, da, false, dwqkv, dbqkv, M, C3, C, sc));                                    // d(a1)
  WMK_CHECK_CUDA(cudaMemcpyAsync(grads[LP_QW], dwqkv, sizeof(float) * (size_t)C * C, cudaMemcpyDeviceToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(grads[LP_KVW], dwqkv + (size_t)C * C, sizeof(float) * 2 * (size_t)C * C, cudaMemcpyDeviceToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(grads[LP_QB], dbqkv, sizeof(float) * C, cudaMemcpyDeviceToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(grads[LP_KVB], dbqkv + C, sizeof(float) * 2 * C, cudaMemcpyDeviceToDevice, st));
  if (params[LP_MOD]) WMK_CHECK_CUDA(cudaMemsetAsync(grads[LP_MOD], 0, sizeof(float) * 64 * C, st));
  ln_bwd_kernel<<<cdiv(M, 8), 256, 0, st>>>(x, da, params[LP_N1W], dx, grads[LP_N1W], grads[LP_N1B],
                                            params[LP_MOD] ? grads[LP_MOD] : nullptr, M, C, H, shift);
  WMK_CHECK_LAUNCH("ln_bwd_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// The other differentiable operators of EncoderTransformerWM (the extractor, uformerWM/model.py:1568-1583), fp32:
// layout changes, LeakyReLU, Downsample (Conv2d 4x4 stride 2), and the strided 8x8 head convolution.
// ------------------------------------------------------------------------------------------------------------------------
namespace wmk {
namespace {

// batched transpose: in [n][R][Cc] -> out [n][Cc][R]
__global__ void __launch_bounds__(256)
transpose_batched_kernel(const float* __restrict__ a, float* __restrict__ at, int R, int Cc) {
  __shared__ float t[32][33];
  const float* ab = a + (size_t)blockIdx.z * R * Cc;
  float* atb = at + (size_t)blockIdx.z * R * Cc;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < R && c0 + tx < Cc) t[i][tx] = ab[(size_t)(r0 + i) * Cc + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < Cc && r0 + tx < R) atb[(size_t)(c0 + i) * R + r0 + tx] = t[tx][i];
}

__global__ void __launch_bounds__(256)
leaky_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, size_t n, float slope) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  out[i] = dy ? (v > 0.f ? dy[i] : slope * dy[i]) : (v > 0.f ? v : slope * v);
}

// out = sigmoid(x) (dy == NULL), or out = dy * y (1 - y) with x = the forward OUTPUT y
__global__ void __launch_bounds__(256)
sigmoid_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ out, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  out[i] = dy ? dy[i] * v * (1.0f - v) : 1.0f / (1.0f + expf(-v));
}

// weight [2C][C][4][4] <-> GEMM order [2C][(kh, kw, ci)]
__global__ void __launch_bounds__(256)
down_w_reorder_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int to_gemm) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)2 * C * C * 16) return;
  const int tap = (int)(idx & 15), ci = (int)((idx >> 4) % C), co = (int)(idx / ((size_t)16 * C));
  const size_t g = ((size_t)co * 16 + tap) * C + ci;           // idx is the reference layout ((co * C + ci) * 16 + tap)
  if (to_gemm) dst[g] = src[idx];
  else dst[idx] = src[g];
}

// dx[b][ih][iw][ci] = sum over the (oh, kh), (ow, kw) pairs with 2 oh - 1 + kh = ih, 2 ow - 1 + kw = iw of dcol[(b, oh, ow)][(kh, kw, ci)]
__global__ void __launch_bounds__(256)
col2im_4x4s2_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int B, int H, int C) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * H * H * C) return;
  const int ci = (int)(idx % C);
  const size_t pix = idx / C;
  const int iw = (int)(pix % H), ih = (int)((pix / H) % H);
  const size_t b = pix / ((size_t)H * H);
  const int Ho = H >> 1;
  float a = 0.f;
  for (int kh = (ih + 1) & 1; kh < 4; kh += 2) {
    const int oh = (ih + 1 - kh) >> 1;
    if (oh < 0 || oh >= Ho) continue;
    for (int kw = (iw + 1) & 1; kw < 4; kw += 2) {
      const int ow = (iw + 1 - kw) >> 1;
      if (ow < 0 || ow >= Ho) continue;
      a += dcol[((b * Ho + oh) * Ho + ow) * (size_t)(16 * C) + (kh * 4 + kw) * C + ci];
    }
  }
  dx[idx] = a;
}

// head: Conv2d(1, 1, 8, stride = (16, 8)) over conv4 [B][64][512] -> feat [B][4][64]
__global__ void __launch_bounds__(256)
head_fwd_kernel(const float* __restrict__ conv4, float* __restrict__ out, const float* __restrict__ w,
                const float* __restrict__ bias, int B) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * 256) return;
  const int j = (int)(idx & 63), i = (int)((idx >> 6) & 3);
  const size_t b = idx >> 8;
  float a = bias[0];
  for (int u = 0; u < 8; ++u)
    for (int v = 0; v < 8; ++v) a = fmaf(conv4[(b * 64 + 16 * i + u) * 512 + 8 * j + v], w[u * 8 + v], a);
  out[idx] = a;
}
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ conv4, const float* __restrict__ dfeat, const float* __restrict__ w,
                float* __restrict__ dconv4, float* __restrict__ dw, float* __restrict__ db, int B) {
  // one thread per conv4 element: its gradient, and its contribution to dw (shared-memory partial sums per CTA)
  __shared__ float sw[65];
  if (threadIdx.x < 65) sw[threadIdx.x] = 0.f;
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < (size_t)B * 64 * 512) {
    const int col = (int)(idx & 511), row = (int)((idx >> 9) & 63);
    const size_t b = idx >> 15;
    const int i = row >> 4, u = row & 15, j = col >> 3, v = col & 7;
    float g = 0.f;
    if (u < 8) {
      const float d = dfeat[b * 256 + i * 64 + j];
      g = d * w[u * 8 + v];
      atomicAdd(&sw[u * 8 + v], d * conv4[idx]);
      if (u == 0 && v == 0) atomicAdd(&sw[64], d);
    }
    dconv4[idx] = g;
  }
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(dw + threadIdx.x, sw[threadIdx.x]);
  if (threadIdx.x == 64) atomicAdd(db, sw[64]);
}

}  // namespace
}  // namespace wmk

extern "C" int wmk_transpose_batched_f32(const float* in, float* out, int n, int R, int Cc, void* stream) {
  WMK_REQUIRE(in && out && n > 0 && n <= 65535 && R > 0 && Cc > 0, "transpose_batched: bad arguments");
  transpose_batched_kernel<<<dim3(cdiv(Cc, 32), cdiv(R, 32), n), 256, 0, (cudaStream_t)stream>>>(in, out, R, Cc);
  WMK_CHECK_LAUNCH("transpose_batched_kernel");
  return 0;
}

/* y = LeakyReLU(x) (dy == NULL), or dx = dy * LeakyReLU'(x) */
extern "C" int wmk_leaky_relu_f32(const float* x, const float* dy, float* out, size_t n, float slope, void* stream) {
  WMK_REQUIRE(x && out && n > 0, "leaky_relu: bad arguments");
  leaky_kernel<<<grid1(n), 256, 0, (cudaStream_t)stream>>>(x, dy, out, n, slope);
  WMK_CHECK_LAUNCH("leaky_kernel");
  return 0;
}

/* Downsample (uformerWM/model.py:763,768-775) on token layout, fp32: out [n * (H/2)^2][2C] = conv4x4s2(x [n * H * H][C]) with the
 * reference weight w [2C][C][4][4]; with dout also dx, dw (reference layout), db. */
extern "C" int wmk_downsample_train_f32(const float* x, const float* w, const float* b, float* out, const float* dout, float* dx,
                                        float* dw, float* db, int n, int H, int C, void* stream) {
  WMK_REQUIRE(x && w && b && out && n > 0 && H >= 2 && H % 2 == 0 && C % 8 == 0 && (!dout == !dx) && (!dout == !dw) && (!dout == !db),
              "downsample_train: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = H / 2, Mo = n * Ho * Ho, K = 16 * C, N = 2 * C;
  Scratch sc{st};
  float* col = sc.get((size_t)Mo * K);
  float* wg = sc.get((size_t)N * K);
  if (!col || !wg) { set_error("downsample_train: scratch allocation failed"); return WMK_ERR_ALLOC; }
  im2col_4x4s2_kernel<float><<<cdiv((size_t)Mo * 4 * (C / 8), 256), 256, 0, st>>>(x, col, n, H, C);
  down_w_reorder_kernel<<<grid1((size_t)N * K), 256, 0, st>>>(w, wg, C, 1);
  WMK_CHECK_LAUNCH("downsample forward kernels");
  WMK_TRY(linear_fwd(col, wg, b, nullptr, out, Mo, N, K, st));
  if (!dout) return 0;
  float* dcol = sc.get((size_t)Mo * K);
  float* dwg = sc.get((size_t)N * K);
  if (!dcol || !dwg) { set_error("downsample_train: scratch allocation failed"); return WMK_ERR_ALLOC; }
  WMK_TRY(linear_bwd(col, wg, dout, dcol, false, dwg, db, Mo, N, K, sc));
  down_w_reorder_kernel<<<grid1((size_t)N * K), 256, 0, st>>>(dwg, dw, C, 0);
  col2im_4x4s2_kernel<<<grid1((size_t)n * H * H * C), 256, 0, st>>>(dcol, dx, n, H, C);
  WMK_CHECK_LAUNCH("downsample backward kernels");
  return 0;
}

/* EncoderTransformerWM.conv2 (model.py:1566,1580-1582): feat [n][256] from conv4 [n][64][512]; with dfeat also dconv4, dw [64], db [1] */
extern "C" int wmk_extract_head_train_f32(const float* conv4, const float* w, const float* b, float* feat, const float* dfeat,
                                          float* dconv4, float* dw, float* db, int n, void* stream) {
  WMK_REQUIRE(conv4 && w && b && feat && n > 0 && (!dfeat == !dconv4) && (!dfeat == !dw) && (!dfeat == !db), "extract_head_train: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  head_fwd_kernel<<<cdiv((size_t)n * 256, 256), 256, 0, st>>>(conv4, feat, w, b, n);
  WMK_CHECK_LAUNCH("head_fwd_kernel");
  if (!dfeat) return 0;
  WMK_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 64, st));
  WMK_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float), st));
  head_bwd_kernel<<<cdiv((size_t)n * 64 * 512, 256), 256, 0, st>>>(conv4, dfeat, w, dconv4, dw, db, n);
  WMK_CHECK_LAUNCH("head_bwd_kernel");
  return 0;
}

/* out = sigmoid(x) (dy NULL), or out = dy * y (1 - y) where x holds the forward output y */
extern "C" int wmk_sigmoid_f32(const float* x, const float* dy, float* out, size_t n, void* stream) {
  WMK_REQUIRE(x && out && n > 0, "sigmoid: bad arguments");
  sigmoid_kernel<<<grid1(n), 256, 0, (cudaStream_t)stream>>>(x, dy, out, n);
  WMK_CHECK_LAUNCH("sigmoid_kernel");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// Embedder-side operators of the training step: the (16, 8) max-pool of the bottleneck (uformerWM/model.py:2398-2400) and the
// ADJOINTS of the in-model ISTFT -> STFT projection (model.py:2458-2463; n_fft 255, hop 63, rectangular window, centre
// reflect padding, one clip = 128 frames <-> 8002 samples).  Reference-precision direct-DFT kernels (255 x 128 table).
// ------------------------------------------------------------------------------------------------------------------------
namespace wmk {
namespace {

__global__ void __launch_bounds__(256)
maxpool16x8_kernel(const float* __restrict__ conv4, const float* __restrict__ dy, float* __restrict__ out, int B) {
  // forward: out [B][4][64] = max over 16 x 8 windows of conv4 [B][64][512]; backward (dy given): out = dconv4, gradient to the
  // first maximum in scan order
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * 256) return;
  const int j = (int)(idx & 63), i = (int)((idx >> 6) & 3);
  const size_t b = idx >> 8;
  float best = -INFINITY;
  int arg = 0;
  for (int r = 0; r < 16; ++r)
    for (int c = 0; c < 8; ++c) {
      const float v = conv4[(b * 64 + 16 * i + r) * 512 + 8 * j + c];
      if (v > best) { best = v; arg = r * 8 + c; }
    }
  if (!dy) { out[idx] = best; return; }
  for (int r = 0; r < 16; ++r)
    for (int c = 0; c < 8; ++c) out[(b * 64 + 16 * i + r) * 512 + 8 * j + c] = (r * 8 + c == arg) ? dy[idx] : 0.f;
}

__constant__ float c_cs255[2][255];      // cos / sin (2 pi j / 255)
int ensure_cs255() {
  static bool done = false;
  if (done) return 0;
  float h[2][255];
  for (int j = 0; j < 255; ++j) { h[0][j] = (float)cos(2.0 * M_PI * j / 255.0); h[1][j] = (float)sin(2.0 * M_PI * j / 255.0); }
  WMK_CHECK_CUDA(cudaMemcpyToSymbol(c_cs255, h, sizeof(h)));
  done = true;
  return 0;
}
constexpr int kT = 128, kL = 8002, kPadded = 255 + 63 * (kT - 1);      // 8256

// STFT adjoint: dS [B][2][128 bins][128 frames] -> gradient of the reflect-padded signal, folded back onto dwave [B][8002]
//   X[k][t] = sum_m xp[63 t + m] (cos - i sin)(2 pi k m / 255)  =>  dxp[q] = sum_{t, k} dXr cos - dXi sin
__global__ void __launch_bounds__(256)
stft_adjoint_kernel(const float* __restrict__ dS, float* __restrict__ dxp, int B) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (q >= kPadded) return;
  const float* gr = dS + (size_t)b * 2 * 128 * kT;
  const float* gi = gr + 128 * kT;
  float a = 0.f;
  int t_hi = q / 63;
  if (t_hi > kT - 1) t_hi = kT - 1;
  for (int t = t_hi; t >= 0 && q - 63 * t <= 254; --t) {
    const int m = q - 63 * t;
    int idx = 0;                                   // (k m) mod 255
    for (int k = 0; k < 128; ++k) {
      a += gr[k * kT + t] * c_cs255[0][idx] - gi[k * kT + t] * c_cs255[1][idx];
      idx += m;
      if (idx >= 255) idx -= 255;
    }
  }
  dxp[(size_t)b * kPadded + q] = a;
}
// xp[127 + j] = x[j]; xp[127 - d] = x[d], xp[127 + L - 1 + d] = x[L - 1 - d]  (d = 1 .. 127)
__global__ void __launch_bounds__(256)
reflect_fold_kernel(const float* __restrict__ dxp, float* __restrict__ dwave, int B) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (j >= kL) return;
  const float* p = dxp + (size_t)b * kPadded;
  float a = p[127 + j];
  if (j >= 1 && j <= 127) a += p[127 - j];
  const int d = kL - 1 - j;
  if (d >= 1 && d <= 127) a += p[127 + kL - 1 + d];
  dwave[(size_t)b * kL + j] = a;
}
// ISTFT adjoint: wave[j] = (1 / env) sum_t f_t[j + 127 - 63 t],  f_t[m] = (1/255) (Xr_0 + 2 sum_{k >= 1} Xr_k cos - Xi_k sin)
//   => dXr_k[t] = (c_k / 255) sum_m g[63 t + m] cos,  dXi_k[t] = -(c_k / 255) sum_m g[63 t + m] sin,  g = dwave / env (0 outside)
__global__ void __launch_bounds__(128)
istft_adjoint_kernel(const float* __restrict__ dwave, float* __restrict__ dSpec, int B) {
  __shared__ float g[255];
  const int t = blockIdx.x, b = blockIdx.y, k = threadIdx.x;
  for (int m = threadIdx.x; m < 255; m += 128) {
    const int q = 63 * t + m, j = q - 127;
    float v = 0.f;
    if (j >= 0 && j < kL) {
      int hi = q / 63;
      if (hi > kT - 1) hi = kT - 1;
      int lo = (q - 254 + 62) / 63;
      if (q - 254 < 0) lo = 0;
      v = dwave[(size_t)b * kL + j] / (float)(hi - lo + 1);
    }
    g[m] = v;
  }
  __syncthreads();
  float ar = 0.f, ai = 0.f;
  int idx = 0;
  for (int m = 0; m < 255; ++m) {
    ar = fmaf(g[m], c_cs255[0][idx], ar);
    ai = fmaf(g[m], c_cs255[1][idx], ai);
    idx += k;
    if (idx >= 255) idx -= 255;
  }
  const float c = (k == 0 ? 1.0f : 2.0f) / 255.0f;
  float* o = dSpec + (size_t)b * 2 * 128 * kT;
  o[k * kT + t] = c * ar;
  o[128 * kT + k * kT + t] = -c * ai;
}

}  // namespace
}  // namespace wmk

/* out [n][256] = MaxPool2d((16, 8)) of conv4 [n][64][512] (dy NULL), or out = dconv4 [n][64][512] from dy [n][256] */
extern "C" int wmk_maxpool16x8_f32(const float* conv4, const float* dy, float* out, int n, void* stream) {
  WMK_REQUIRE(conv4 && out && n > 0, "maxpool16x8: bad arguments");
  maxpool16x8_kernel<<<cdiv((size_t)n * 256, 256), 256, 0, (cudaStream_t)stream>>>(conv4, dy, out, n);
  WMK_CHECK_LAUNCH("maxpool16x8_kernel");
  return 0;
}

/* Adjoint of the in-model projection s = STFT(ISTFT(y)) for one-clip spectrograms [n][2][128][128] (model.py:2458-2463):
 * dy = ISTFT^T STFT^T ds.  scratch-free (stream-ordered temporaries). */
extern "C" int wmk_stft_projection_adjoint_f32(const float* ds, float* dy, int n, void* stream) {
  WMK_REQUIRE(ds && dy && n > 0 && n <= 65535, "stft_projection_adjoint: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  WMK_TRY(ensure_cs255());
  Scratch sc{st};
  float* dxp = sc.get((size_t)n * kPadded);
  float* dwave = sc.get((size_t)n * kL);
  if (!dxp || !dwave) { set_error("stft_projection_adjoint: scratch allocation failed"); return WMK_ERR_ALLOC; }
  stft_adjoint_kernel<<<dim3(cdiv(kPadded, 256), n), 256, 0, st>>>(ds, dxp, n);
  reflect_fold_kernel<<<dim3(cdiv(kL, 256), n), 256, 0, st>>>(dxp, dwave, n);
  istft_adjoint_kernel<<<dim3(kT, n), 128, 0, st>>>(dwave, dy, n);
  WMK_CHECK_LAUNCH("stft projection adjoint kernels");
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------------
// Upsample = ConvTranspose2d(Cin, Cout, 2, stride 2) on tokens (uformerWM/model.py:794-800) as a GEMM with a pixel-shuffle
// epilogue, and its gradients.  GEMM weight rows are (i, j, co), columns ci.
// ------------------------------------------------------------------------------------------------------------------------
namespace wmk {
namespace {

// reference w [Cin][Cout][2][2] <-> wg [(ij) * Cout + co][Cin]; to_gemm = 0 writes the reference layout from wg
__global__ void __launch_bounds__(256)
up_w_reorder_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cin, int Cout, int to_gemm) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Cin * Cout * 4) return;
  const int ij = (int)(idx & 3), co = (int)((idx >> 2) % Cout), ci = (int)(idx / ((size_t)4 * Cout));
  const size_t g = ((size_t)ij * Cout + co) * Cin + ci;          // idx = (ci * Cout + co) * 4 + ij is the reference layout
  if (to_gemm) dst[g] = src[idx];
  else dst[idx] = src[g];
}
// dyg [m = (b, h, w)][(ij) * Cout + co] = dout[token (b, 2h + i, 2w + j)][co]
__global__ void __launch_bounds__(256)
up_gather_kernel(const float* __restrict__ dout, float* __restrict__ dyg, int B, int h, int Cout) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * h * h * 4 * Cout) return;
  const int co = (int)(idx % Cout), ij = (int)((idx / Cout) & 3);
  const size_t m = idx / ((size_t)4 * Cout);
  const int wq = (int)(m % h), hq = (int)((m / h) % h);
  const size_t b = m / ((size_t)h * h);
  const size_t tok = (b * 2 * h + 2 * hq + (ij >> 1)) * (size_t)(2 * h) + 2 * wq + (ij & 1);
  dyg[idx] = dout[tok * Cout + co];
}
__global__ void __launch_bounds__(256) bias4_kernel(const float* __restrict__ b, float* __restrict__ b4, int Cout, int reduce) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (reduce) { if (i < Cout) b4[i] = (b[i] + b[Cout + i]) + (b[2 * Cout + i] + b[3 * Cout + i]); }      // db from the 4 Cout column sums
  else if (i < 4 * Cout) b4[i] = b[i % Cout];
}

}  // namespace
}  // namespace wmk

/* Upsample on tokens: out [n * (2h)^2][Cout] from x [n * h * h][Cin], reference weight w [Cin][Cout][2][2]; with dout also dx, dw, db */
extern "C" int wmk_upsample_train_f32(const float* x, const float* w, const float* b, float* out, const float* dout, float* dx,
                                      float* dw, float* db, int n, int h, int Cin, int Cout, void* stream) {
  WMK_REQUIRE(x && w && b && out && n > 0 && h > 0 && Cin % 4 == 0 && Cout > 0 && (!dout == !dx) && (!dout == !dw) && (!dout == !db),
              "upsample_train: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int M = n * h * h, N = 4 * Cout;
  Scratch sc{st};
  float* wg = sc.get((size_t)N * Cin);
  float* b4 = sc.get(N);
  if (!wg || !b4) { set_error("upsample_train: scratch allocation failed"); return WMK_ERR_ALLOC; }
  up_w_reorder_kernel<<<grid1((size_t)N * Cin), 256, 0, st>>>(w, wg, Cin, Cout, 1);
  bias4_kernel<<<cdiv(N, 256), 256, 0, st>>>(b, b4, Cout, 0);
  WMK_CHECK_LAUNCH("upsample forward kernels");
  GemmArgs g;
  g.A = x; g.W = wg; g.bias = b4; g.C = out; g.M = M; g.N = N; g.K = Cin; g.ldc = Cout;
  g.epi = EPI_UPSAMPLE; g.up_h = h; g.up_w = h; g.up_cout = Cout;
  WMK_TRY(gemm_fp32_simt(g, st));
  if (!dout) return 0;
  float* dyg = sc.get((size_t)M * N);
  float* dwg = sc.get((size_t)N * Cin);
  float* dbg = sc.get(N);
  if (!dyg || !dwg || !dbg) { set_error("upsample_train: scratch allocation failed"); return WMK_ERR_ALLOC; }
  up_gather_kernel<<<grid1((size_t)M * N), 256, 0, st>>>(dout, dyg, n, h, Cout);
  WMK_CHECK_LAUNCH("up_gather_kernel");
  WMK_TRY(linear_bwd(x, wg, dyg, dx, false, dwg, dbg, M, N, Cin, sc));
  up_w_reorder_kernel<<<grid1((size_t)N * Cin), 256, 0, st>>>(dwg, dw, Cin, Cout, 0);
  bias4_kernel<<<cdiv(Cout, 256), 256, 0, st>>>(dbg, db, Cout, 1);
  WMK_CHECK_LAUNCH("upsample backward kernels");
  return 0;
}

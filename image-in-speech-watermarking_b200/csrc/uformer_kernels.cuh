// Memory-bound kernels of the LeWin / Uformer blocks (everything that is not a dense layer).
// Token layout throughout: activations are [tokens][channels] with tokens = (clip, h, w)
// row-major, i.e. channels-last; the reference's NCHW <-> token transposes
// (uformerWM/model.py:703,709,773-774,798-799,826,861) disappear.
// OpT is the operand type of the dense layers: float (fp32 mode) or __nv_bfloat16 (bf16 mode).
#pragma once
#include "wmk_common.cuh"

namespace wmk {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-5) over channels, one warp per token, + optional modulator add
// (uformerWM/model.py:982, 996-999, 1017).  The modulator row is the token's index inside its
// (cyclically shifted) 8x8 window, so the add commutes with roll + window_partition.
// ------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, OpT* __restrict__ out, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ modulator, int M, int C, int H,
                 int shift) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* xr = x + (size_t)warp * C;
  const int per = C >> 5;                       // C in {32..512} -> 1..16 values per lane
  float v[16];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < per) { v[i] = xr[i * 32 + lane]; s += v[i]; }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < per) { const float d = v[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
  const float* mod = nullptr;
  if (modulator) {
    const int hw = warp % (H * H);
    const int h = hw / H, w = hw - h * H;
    const int hs = (h - shift + H) % H, ws = (w - shift + H) % H;
    mod = modulator + (size_t)(((hs & 7) << 3) | (ws & 7)) * C;
  }
  OpT* o = out + (size_t)warp * C;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    if (i < per) {
      const int c = i * 32 + lane;
      float y = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
      if (mod) y += __ldg(mod + c);
      o[c] = from_f<OpT>(y);
    }
}

// ------------------------------------------------------------------------------------------
// Window attention for one (8x8 window, head): softmax(q k^T + rel-pos bias + shift mask) v
// (uformerWM/model.py:523-551 with the roll / window_partition / window_reverse of :986-1012
// folded into the token addressing, and the 0/-100 shift mask of :954-972 computed from the
// token's region in the shifted image).  q is pre-scaled (scale folded into the packed weights).
// qkv: [tokens][3C] = [q | k | v], head h owns channels h*32..h*32+31 of each third.
// ------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(128)
window_attention_kernel(const OpT* __restrict__ qkv, OpT* __restrict__ out, const float* __restrict__ bias,
                        int C, int H, int shift) {
  __shared__ float Qs[64][33], Ks[64][33], Vs[64][33];
  __shared__ float S[64][65];
  __shared__ int tok[64], rid[64];
  const int head = blockIdx.y;
  const int nwin_side = H >> 3;
  const int win = blockIdx.x;
  const int b = win / (nwin_side * nwin_side);
  const int wrem = win - b * nwin_side * nwin_side;
  const int wh = wrem / nwin_side, ww = wrem - wh * nwin_side;
  const int tid = threadIdx.x;
  if (tid < 64) {
    const int hs = wh * 8 + (tid >> 3), ws = ww * 8 + (tid & 7);       // coordinates in the shifted image
    const int h = (hs + shift) % H, w = (ws + shift) % H;              // roll(-shift): shifted[i] = x[i+shift]
    tok[tid] = (b * H + h) * H + w;
    const int rh = hs < H - 8 ? 0 : (hs < H - shift ? 1 : 2);
    const int rw = ws < H - 8 ? 0 : (ws < H - shift ? 1 : 2);
    rid[tid] = rh * 3 + rw;
  }
  __syncthreads();
  for (int e = tid; e < 64 * 32; e += 128) {
    const int r = e >> 5, d = e & 31;
    const OpT* base = qkv + (size_t)tok[r] * (3 * C) + head * 32 + d;
    Qs[r][d] = to_f<OpT>(base[0]);
    Ks[r][d] = to_f<OpT>(base[C]);
    Vs[r][d] = to_f<OpT>(base[2 * C]);
  }
  const float* bh = bias + (size_t)head * 4096;
  for (int e = tid; e < 4096; e += 128) S[e >> 6][e & 63] = __ldg(bh + e);
  __syncthreads();

  const int i = tid >> 1, half = tid & 1;
  float qreg[32];
#pragma unroll
  for (int d = 0; d < 32; ++d) qreg[d] = Qs[i][d];
  const int my_rid = rid[i];
  float mx = -INFINITY;
  for (int jj = 0; jj < 32; ++jj) {
    const int j = half * 32 + jj;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < 32; ++d) a = fmaf(qreg[d], Ks[j][d], a);
    a += S[i][j];
    if (shift > 0 && rid[j] != my_rid) a += -100.0f;
    S[i][j] = a;
    mx = fmaxf(mx, a);
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  float sum = 0.f;
  for (int jj = 0; jj < 32; ++jj) {
    const int j = half * 32 + jj;
    const float e = expf(S[i][j] - mx);
    S[i][j] = e;
    sum += e;
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  const float inv = 1.0f / sum;
  __syncwarp();                                   // both halves of row i live in the same warp
  float o[16];
#pragma unroll
  for (int d = 0; d < 16; ++d) o[d] = 0.f;
  for (int j = 0; j < 64; ++j) {
    const float pj = S[i][j];
#pragma unroll
    for (int d = 0; d < 16; ++d) o[d] = fmaf(pj, Vs[j][half * 16 + d], o[d]);
  }
  OpT* orow = out + (size_t)tok[i] * C + head * 32 + half * 16;
#pragma unroll
  for (int d = 0; d < 16; ++d) orow[d] = from_f<OpT>(o[d] * inv);
}

// ------------------------------------------------------------------------------------------
// LeFF depthwise 3x3 conv (pad 1) + GELU on the [B][H][W][Ch] hidden tensor
// (uformerWM/model.py:688-689,706).  wt: [9][Ch] (tap-major), one thread = one pixel x 4 channels.
// ------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256)
dwconv3x3_gelu_kernel(const OpT* __restrict__ in, OpT* __restrict__ out, const float* __restrict__ wt,
                      const float* __restrict__ bias, int B, int H, int Ch) {
  const int cg = Ch >> 2;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * H * H * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 4;
  const size_t pix = idx / cg;
  const int w = (int)(pix % H);
  const int h = (int)((pix / H) % H);
  const size_t b = pix / ((size_t)H * H);
  float acc[4];
  {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
    acc[0] = bb.x; acc[1] = bb.y; acc[2] = bb.z; acc[3] = bb.w;
  }
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int hh = h + dy;
    if (hh < 0 || hh >= H) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int wwp = w + dx;
      if (wwp < 0 || wwp >= H) continue;
      const OpT* src = in + ((b * H + hh) * H + wwp) * Ch + c;
      const float4 k = __ldg(reinterpret_cast<const float4*>(wt + ((dy + 1) * 3 + (dx + 1)) * Ch + c));
      float v0, v1, v2, v3;
      if constexpr (sizeof(OpT) == 4) {
        const float4 v = *reinterpret_cast<const float4*>(src);
        v0 = v.x; v1 = v.y; v2 = v.z; v3 = v.w;
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(src);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
        const __nv_bfloat162 bq = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        v0 = __low2float(a); v1 = __high2float(a); v2 = __low2float(bq); v3 = __high2float(bq);
      }
      acc[0] = fmaf(v0, k.x, acc[0]); acc[1] = fmaf(v1, k.y, acc[1]);
      acc[2] = fmaf(v2, k.z, acc[2]); acc[3] = fmaf(v3, k.w, acc[3]);
    }
  }
  OpT* dst = out + pix * Ch + c;
#pragma unroll
  for (int j = 0; j < 4; ++j) dst[j] = from_f<OpT>(gelu_erf(acc[j]));
}

// ------------------------------------------------------------------------------------------
// im2col for Downsample = Conv2d(C, 2C, k=4, s=2, p=1) (uformerWM/model.py:763,768-775):
// A[(b,oh,ow)][(kh,kw,ci)] = x[b][2oh-1+kh][2ow-1+kw][ci] (zero outside).  One thread = 4 ci.
// ------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256)
im2col_4x4s2_kernel(const float* __restrict__ x, OpT* __restrict__ A, int B, int H, int C) {
  const int cg = C >> 2;
  const int Ho = H >> 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Ho * Ho * 16 * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 4;
  size_t r = idx / cg;
  const int tap = (int)(r % 16);
  r /= 16;
  const int ow = (int)(r % Ho);
  const int oh = (int)((r / Ho) % Ho);
  const size_t b = r / ((size_t)Ho * Ho);
  const int ih = 2 * oh - 1 + (tap >> 2), iw = 2 * ow - 1 + (tap & 3);
  float4 v = make_float4(0, 0, 0, 0);
  if (ih >= 0 && ih < H && iw >= 0 && iw < H)
    v = *reinterpret_cast<const float4*>(x + ((b * H + ih) * H + iw) * C + c);
  OpT* dst = A + ((b * Ho + oh) * Ho + ow) * (size_t)(16 * C) + tap * C + c;
  dst[0] = from_f<OpT>(v.x); dst[1] = from_f<OpT>(v.y); dst[2] = from_f<OpT>(v.z); dst[3] = from_f<OpT>(v.w);
}

// fp32 -> OpT copy of a [rows][cols] matrix into a [rows][ld_dst] buffer at column offset col0
// (operand casts; the decoder's torch.cat([up, skip], -1) of uformerWM/model.py:1225-1237).
template <typename DstT>
__global__ void __launch_bounds__(256)
copy_cols_kernel(const float* __restrict__ src, DstT* __restrict__ dst, size_t rows, int cols, int ld_dst, int col0) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = cols >> 2;
  if (idx >= rows * cg) return;
  const size_t r = idx / cg;
  const int c = (int)(idx % cg) * 4;
  const float4 v = *reinterpret_cast<const float4*>(src + r * cols + c);
  DstT* d = dst + r * ld_dst + col0 + c;
  d[0] = from_f<DstT>(v.x); d[1] = from_f<DstT>(v.y); d[2] = from_f<DstT>(v.z); d[3] = from_f<DstT>(v.w);
}

// Bottleneck concat (uformerWM/model.py:2388-2389,2411): A[(b,r)][0..511] = feat[b][r%4][c%64],
// A[(b,r)][512..1023] = conv4[b][r][c-512].   feat: [B][4][64] (image-codec code), conv4 fp32.
template <typename OpT>
__global__ void __launch_bounds__(256)
bottleneck_concat_kernel(const float* __restrict__ feat, const float* __restrict__ conv4, OpT* __restrict__ A, int B) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * 64 * 1024) return;
  const int c = (int)(idx & 1023);
  const int r = (int)((idx >> 10) & 63);
  const size_t b = idx >> 16;
  float v;
  if (c < 512) v = feat[(b * 4 + (r & 3)) * 64 + (c & 63)];
  else v = conv4[(b * 64 + r) * 512 + (c - 512)];
  A[idx] = from_f<OpT>(v);
}

}  // namespace wmk

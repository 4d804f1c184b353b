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
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) {
  return __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));      // saturate instead of overflowing to inf
}

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
__device__ __forceinline__ void store4(OpT* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = u;
}

// One token is handled by G = min(32, C/4) lanes, each owning float4 chunks (C/4/G of them); every
// lane group processes TPT tokens with all of their loads issued up front (memory-level
// parallelism), so a warp covers 32/G*TPT tokens and every global access is a 16-byte (fp32) /
// 8-byte (bf16) vector.
template <typename OpT, int C>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, OpT* __restrict__ out, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ modulator, int M, int H, int shift) {
  constexpr int G = (C / 4 < 32) ? C / 4 : 32;      // lanes per token
  constexpr int NV = C / 4 / G;                     // float4 chunks per lane per token
  constexpr int TPT = NV >= 4 ? 1 : (NV == 2 ? 2 : 4);   // tokens per lane group
  constexpr int GPW = 32 / G;                       // lane groups per warp
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int sub = lane / G, gl = lane % G;
  const int token0 = (warp * GPW + sub) * TPT;
  float4 v[TPT][NV];
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    const int token = token0 + t < M ? token0 + t : M - 1;
    const float* xr = x + (size_t)token * C;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[t][i] = *reinterpret_cast<const float4*>(xr + (i * G + gl) * 4);
  }
  float4 g4[NV], b4[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    g4[i] = __ldg(reinterpret_cast<const float4*>(gamma + (i * G + gl) * 4));
    b4[i] = __ldg(reinterpret_cast<const float4*>(beta + (i * G + gl) * 4));
  }
#pragma unroll
  for (int t = 0; t < TPT; ++t) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[t][i].x + v[t][i].y) + (v[t][i].z + v[t][i].w);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[t][i].x - mean, b = v[t][i].y - mean, c = v[t][i].z - mean, d = v[t][i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / C) + 1e-5f);
    const int token = token0 + t;
    if (token >= M) continue;
    const float* mod = nullptr;
    if (modulator) {
      const int hw = token % (H * H);
      const int h = hw / H, w = hw - h * H;
      const int hs = (h - shift + H) % H, ws = (w - shift + H) % H;
      mod = modulator + (size_t)(((hs & 7) << 3) | (ws & 7)) * C;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = (i * G + gl) * 4;
      float y0 = (v[t][i].x - mean) * rstd * g4[i].x + b4[i].x, y1 = (v[t][i].y - mean) * rstd * g4[i].y + b4[i].y;
      float y2 = (v[t][i].z - mean) * rstd * g4[i].z + b4[i].z, y3 = (v[t][i].w - mean) * rstd * g4[i].w + b4[i].w;
      if (mod) {
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mod + c));
        y0 += m4.x; y1 += m4.y; y2 += m4.z; y3 += m4.w;
      }
      if constexpr (OpMode<OpT>::v == 2)
        split_store4(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)token * 2 * C, c, C, y0, y1, y2, y3);
      else
        store4<OpT>(out + (size_t)token * C + c, y0, y1, y2, y3);
    }
  }
}

template <>
__device__ __forceinline__ void store4<__half>(__half* p, float a, float b, float c, float d) {
  uint2 u;
  u.x = pack2_f16(a, b);
  u.y = pack2_f16(c, d);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename OpT>
inline void launch_layernorm(const float* x, OpT* out, const float* gamma, const float* beta, const float* modulator,
                             int M, int C, int H, int shift, cudaStream_t st) {
  const int G = (C / 4 < 32) ? C / 4 : 32;
  const int NV = C / 4 / G;
  const int tpt = NV >= 4 ? 1 : (NV == 2 ? 2 : 4);
  const int tpw = (32 / G) * tpt;
  const int warps = (M + tpw - 1) / tpw;
  const int blocks = (warps + 7) / 8;
  switch (C) {
    case 32: layernorm_kernel<OpT, 32><<<blocks, 256, 0, st>>>(x, out, gamma, beta, modulator, M, H, shift); break;
    case 64: layernorm_kernel<OpT, 64><<<blocks, 256, 0, st>>>(x, out, gamma, beta, modulator, M, H, shift); break;
    case 128: layernorm_kernel<OpT, 128><<<blocks, 256, 0, st>>>(x, out, gamma, beta, modulator, M, H, shift); break;
    case 256: layernorm_kernel<OpT, 256><<<blocks, 256, 0, st>>>(x, out, gamma, beta, modulator, M, H, shift); break;
    default: layernorm_kernel<OpT, 512><<<blocks, 256, 0, st>>>(x, out, gamma, beta, modulator, M, H, shift); break;
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
// bf16 window attention on the tensor cores (mma.sync m16n8k16, fp32 accumulate): same math and
// addressing as window_attention_kernel.  One CTA = one (window, head); warp w owns query rows
// 16w..16w+15: S = Q K^T in registers (bias + shift mask + fp32 softmax), P re-used directly as the
// A fragments of P V (no shared-memory round trip), O staged through smem for 16-byte stores.
// The op is HBM-bound (8C bytes per token for ~256C FLOPs), so the legacy mma path is sufficient.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int ATT_LD = 40;     // bf16 elements per smem row (80 B): conflict-free ldmatrix
constexpr int ATT_TILE = 64 * ATT_LD;   // one 64 x 32 operand tile

struct AttGeom {
  int C, H, shift, lg_heads, lg_nws;      // heads = C/32 and windows per side = H/8 are powers of two
};
// token index and shift-mask region of row r (0..63) of window `win`
__device__ __forceinline__ void att_row(const AttGeom& g, int win, int r, int& token, int& region) {
  const int b = win >> (2 * g.lg_nws);
  const int wrem = win & ((1 << (2 * g.lg_nws)) - 1);
  const int wh = wrem >> g.lg_nws, ww = wrem & ((1 << g.lg_nws) - 1);
  const int hs = wh * 8 + (r >> 3), ws = ww * 8 + (r & 7);
  const int h = (hs + g.shift) & (g.H - 1), w = (ws + g.shift) & (g.H - 1);
  token = (b * g.H + h) * g.H + w;
  const int rh = hs < g.H - 8 ? 0 : (hs < g.H - g.shift ? 1 : 2);
  const int rw = ws < g.H - 8 ? 0 : (ws < g.H - g.shift ? 1 : 2);
  region = rh * 3 + rw;
}

// Persistent: a CTA owns ONE head (blockIdx.x % heads) and loops over windows, prefetching the next
// window's q/k/v tiles with cp.async while the tensor cores work on the current one.
//  * The head's 64x64 relative-position bias sits in shared memory as bf16 (loaded once per CTA) and is
//    added by the tensor cores: S = I * Bias + Q K^T (one extra k-step with an identity A fragment)
//    instead of 16 global loads + 32 FADDs per thread per window (the kernel was L1/TEX bound on them).
//  * Scores arrive in the log2 domain: the packed q rows and the bias table carry a factor log2(e)
//    (uformer_plan.cu pack_block), so softmax is exp2(s - max) with one FADD + one MUFU.EX2 per score.
//  * The shift mask only exists in windows of the last window row / column: all others skip it.
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr int ATT_BLD = 72;    // bf16 elements per bias row in smem (144 B: conflict-free ldmatrix)

// F16: q / k / v / out (and the in-kernel bias and probability fragments) are IEEE fp16 instead of bf16.
// SPLIT_OUT: the output rows are split-bf16 [hi(C) | lo(C)] (the A operand of a split projection; the precise extractor).
// NST: depth of the cp.async ring of q / k / v tiles (2: one window ahead, 5 CTAs per SM = 60 KB in flight per SM;
// 3: two windows ahead, dynamic shared memory, 4 CTAs per SM = 96 KB in flight)
template <bool F16, bool SPLIT_OUT = false, int NST = 2>
static __global__ void __launch_bounds__(128)
window_attention_mma_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out,
                            const float* __restrict__ bias, int C, int H, int shift, int n_windows) {
  extern __shared__ __align__(16) uint16_t att_dyn[];           // [NST][3 * ATT_TILE]
  uint16_t (*sbuf)[3 * ATT_TILE] = reinterpret_cast<uint16_t (*)[3 * ATT_TILE]>(att_dyn);
  __shared__ __align__(16) uint16_t sbias[64 * ATT_BLD];
  auto mma = [](float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (F16) mma_f16_16816(d, a, b0, b1);
    else mma_bf16_16816(d, a, b0, b1);
  };
  __shared__ int s_tok[NST][64], s_rid[NST][64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  AttGeom g{C, H, shift, 31 - __clz(C >> 5), 31 - __clz(H >> 3)};
  const int heads = C >> 5;
  const int head = blockIdx.x & (heads - 1);
  const int wstep = gridDim.x >> g.lg_heads;
  const int nws_mask = (1 << g.lg_nws) - 1;

  auto prefetch = [&](int win, int buf) {
    if (tid < 64) {
      int t, r;
      att_row(g, win, tid, t, r);
      s_tok[buf][tid] = t;
      s_rid[buf][tid] = r;
    }
    // thread -> rows (tid>>2) and 32+(tid>>2), 16-byte chunk tid&3, for each of q, k, v
    int ta, tb, rr;
    att_row(g, win, tid >> 2, ta, rr);
    att_row(g, win, 32 + (tid >> 2), tb, rr);
    const uint16_t* pa = qkv + (size_t)ta * (3 * C) + head * 32 + (tid & 3) * 8;
    const uint16_t* pb = qkv + (size_t)tb * (3 * C) + head * 32 + (tid & 3) * 8;
    const uint32_t da = (uint32_t)__cvta_generic_to_shared(&sbuf[buf][(tid >> 2) * ATT_LD + (tid & 3) * 8]);
    const uint32_t db = da + 2u * 32 * ATT_LD;
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da + 2u * which * ATT_TILE), "l"(pa + which * C) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(db + 2u * which * ATT_TILE), "l"(pb + which * C) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int win = blockIdx.x >> g.lg_heads;
  if (win >= n_windows) return;
  prefetch(win, 0);
  if constexpr (NST == 3) {                       // second window of the ring (an empty group keeps the group count uniform)
    if (win + wstep < n_windows) prefetch(win + wstep, 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }
  {
    const float* bh = bias + (size_t)head * 4096;
    for (int e = tid; e < 2048; e += 128) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(bh) + e);
      *reinterpret_cast<uint32_t*>(&sbias[(e >> 5) * ATT_BLD + (e & 31) * 2]) = pack2_16<F16>(v.x, v.y);
    }
  }
  const int gq = lane >> 2, t = lane & 3;
  // identity A fragment (16x16): thread (gq, t) holds A[gq][2t..2t+1] and A[gq+8][2t+8..2t+9]
  const uint32_t ident = pack2_16<F16>(gq == 2 * t ? 1.f : 0.f, gq == 2 * t + 1 ? 1.f : 0.f);
  const uint32_t aI[4] = {ident, 0u, 0u, ident};
  const uint32_t bs_addr = (uint32_t)__cvta_generic_to_shared(sbias);
  for (int it = 0; win < n_windows; win += wstep, ++it) {
    const int buf = NST == 3 ? it % 3 : (it & 1);
    if constexpr (NST == 3) {
      const int next2 = win + 2 * wstep;
      if (next2 < n_windows) prefetch(next2, (it + 2) % 3);          // the buffer of iteration it - 1 (freed by its closing barrier)
      else asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 2;" ::: "memory");           // groups complete in order: this window's tiles have landed
    } else {
      const int next = win + wstep;
      if (next < n_windows) {
        prefetch(next, buf ^ 1);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      }
    }
    __syncthreads();
    uint16_t* Qs = &sbuf[buf][0];
    const int* tok = s_tok[buf];
    const int* rid = s_rid[buf];
    const uint32_t qs = (uint32_t)__cvta_generic_to_shared(Qs), ks = qs + 2u * ATT_TILE, vs = qs + 4u * ATT_TILE;
    const int r0 = warp * 16;
    // ---- S = I * Bias  (rows r0..r0+15 of the bias as the B operand)
    float sacc[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) { sacc[n][0] = sacc[n][1] = sacc[n][2] = sacc[n][3] = 0.f; }
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, bs_addr + 2u * ((r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_BLD + np * 16 + (lane >> 4) * 8));
      mma(sacc[2 * np], aI, bb[0], bb[1]);
      mma(sacc[2 * np + 1], aI, bb[2], bb[3]);
    }
    // ---- S += Q K^T
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t a[4];
      ldmatrix_x4(a, qs + 2u * ((r0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LD + kk * 16 + (lane >> 4) * 8));
#pragma unroll
      for (int np = 0; np < 4; ++np) {           // two n-tiles (16 keys) per ldmatrix.x4
        uint32_t bq[4];
        ldmatrix_x4(bq, ks + 2u * ((np * 16 + (lane & 7) + (lane >> 4) * 8) * ATT_LD + kk * 16 + ((lane >> 3) & 1) * 8));
        mma(sacc[2 * np], a, bq[0], bq[1]);
        mma(sacc[2 * np + 1], a, bq[2], bq[3]);
      }
    }
    // ---- shift mask (only windows of the last window row / column hold more than one region), softmax
    const int row0 = r0 + gq, row1 = row0 + 8;
    const int wrem = win & ((1 << (2 * g.lg_nws)) - 1);
    const bool edge = shift > 0 && (((wrem >> g.lg_nws) == nws_mask) || ((wrem & nws_mask) == nws_mask));
    float m0 = -INFINITY, m1 = -INFINITY;
    if (edge) {
      const int rid0 = rid[row0], rid1 = rid[row1];
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int col = n * 8 + 2 * t;
        const int c0 = rid[col], c1 = rid[col + 1];
        if (c0 != rid0) sacc[n][0] -= 100.0f * kLog2e;
        if (c1 != rid0) sacc[n][1] -= 100.0f * kLog2e;
        if (c0 != rid1) sacc[n][2] -= 100.0f * kLog2e;
        if (c1 != rid1) sacc[n][3] -= 100.0f * kLog2e;
      }
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      m0 = fmaxf(m0, fmaxf(sacc[n][0], sacc[n][1]));
      m1 = fmaxf(m1, fmaxf(sacc[n][2], sacc[n][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      sacc[n][0] = ex2_approx(sacc[n][0] - m0); sacc[n][1] = ex2_approx(sacc[n][1] - m0);
      sacc[n][2] = ex2_approx(sacc[n][2] - m1); sacc[n][3] = ex2_approx(sacc[n][3] - m1);
      s0 += sacc[n][0] + sacc[n][1];
      s1 += sacc[n][2] + sacc[n][3];
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float inv0 = 1.0f / s0, inv1 = 1.0f / s1;
    // ---- O = P V   (P fragments come straight from the S accumulators)
    float oacc[4][4];
#pragma unroll
    for (int n = 0; n < 4; ++n) { oacc[n][0] = oacc[n][1] = oacc[n][2] = oacc[n][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {             // 16 keys per step
      uint32_t a[4];
      a[0] = pack2_16<F16>(sacc[2 * kk][0], sacc[2 * kk][1]);
      a[1] = pack2_16<F16>(sacc[2 * kk][2], sacc[2 * kk][3]);
      a[2] = pack2_16<F16>(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
      a[3] = pack2_16<F16>(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {           // two n-tiles (16 dims) per ldmatrix.x4.trans
        uint32_t bv[4];
        ldmatrix_x4_trans(bv, vs + 2u * ((kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ATT_LD + np * 16 + (lane >> 4) * 8));
        mma(oacc[2 * np], a, bv[0], bv[1]);
        mma(oacc[2 * np + 1], a, bv[2], bv[3]);
      }
    }
    // ---- stage O (this warp's 16 rows) in the Q tile, then 16-byte stores to the token rows
    __syncwarp();
    if constexpr (SPLIT_OUT) {
#pragma unroll
      for (int part = 0; part < 2; ++part) {       // hi rows, then lo rows, through the same staging rows
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          uint32_t h0, l0, h1, l1;
          split_pack2(oacc[n][0] * inv0, oacc[n][1] * inv0, h0, l0);
          split_pack2(oacc[n][2] * inv1, oacc[n][3] * inv1, h1, l1);
          *reinterpret_cast<uint32_t*>(Qs + row0 * ATT_LD + n * 8 + 2 * t) = part ? l0 : h0;
          *reinterpret_cast<uint32_t*>(Qs + row1 * ATT_LD + n * 8 + 2 * t) = part ? l1 : h1;
        }
        __syncwarp();
#pragma unroll
        for (int e = lane; e < 64; e += 32) {
          const int r = r0 + (e >> 2), ch = e & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(Qs + r * ATT_LD + ch * 8);
          *reinterpret_cast<uint4*>(out + (size_t)tok[r] * (2 * C) + part * C + head * 32 + ch * 8) = v;
        }
        __syncwarp();
      }
    } else {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        *reinterpret_cast<uint32_t*>(Qs + row0 * ATT_LD + n * 8 + 2 * t) = pack2_16<F16>(oacc[n][0] * inv0, oacc[n][1] * inv0);
        *reinterpret_cast<uint32_t*>(Qs + row1 * ATT_LD + n * 8 + 2 * t) = pack2_16<F16>(oacc[n][2] * inv1, oacc[n][3] * inv1);
      }
      __syncwarp();
#pragma unroll
      for (int e = lane; e < 64; e += 32) {
        const int r = r0 + (e >> 2), ch = e & 3;
        const uint4 v = *reinterpret_cast<const uint4*>(Qs + r * ATT_LD + ch * 8);
        *reinterpret_cast<uint4*>(out + (size_t)tok[r] * C + head * 32 + ch * 8) = v;
      }
    }
    __syncthreads();        // everyone is done with this buffer before it is refilled
  }
}

// ------------------------------------------------------------------------------------------
// LeFF depthwise 3x3 conv (pad 1) + GELU on the [B][H][W][Ch] hidden tensor
// (uformerWM/model.py:688-689,706).  wt: [9][Ch] (tap-major), one thread = one pixel x 4 channels.
// ------------------------------------------------------------------------------------------
// Register sliding window: one thread owns a column strip of 8 pixels x 4 channels of an 8x8
// tile, keeps its 36 weights and a 3x3x4 input window in registers and loads only the 3 new
// pixels of each row step (3 loads per output instead of 9; no shared memory, no barrier).
// A warp = 2 adjacent columns x 64 channels, so every load / store instruction covers full
// 128-byte lines; channel slabs are the fastest CTA index, then tiles in raster order, so halo
// pixels shared with neighbouring CTAs are L2 hits.
template <typename OpT>
__device__ __forceinline__ void dw_load4(const OpT* p, bool valid, float (&v)[4]) {
  if (!valid) { v[0] = v[1] = v[2] = v[3] = 0.f; return; }
  if constexpr (sizeof(OpT) == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
    const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
    v[0] = __low2float(lo); v[1] = __high2float(lo); v[2] = __low2float(hi); v[3] = __high2float(hi);
  }
}

template <typename OpT>
__global__ void __launch_bounds__(128)
dwconv3x3_gelu_kernel(const OpT* __restrict__ in, OpT* __restrict__ out, const float* __restrict__ wt,
                      const float* __restrict__ bias, int B, int H, int Ch) {
  const int ncs = Ch >> 6;
  const int slab = blockIdx.x % ncs;
  const int sp = blockIdx.x / ncs;
  const int tiles = H >> 3;
  const int b = sp / (tiles * tiles);
  const int trem = sp - b * tiles * tiles;
  const int h0 = (trem / tiles) * 8;
  const int w = (trem % tiles) * 8 + (threadIdx.x >> 4);
  const int c = slab * 64 + (threadIdx.x & 15) * 4;
  float wreg[9][4], bz[4];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(wt + (size_t)t * Ch + c));
    wreg[t][0] = a.x; wreg[t][1] = a.y; wreg[t][2] = a.z; wreg[t][3] = a.w;
  }
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias + c));
    bz[0] = a.x; bz[1] = a.y; bz[2] = a.z; bz[3] = a.w;
  }
  const OpT* img = in + (size_t)b * H * H * Ch + c;
  OpT* oimg = out + (size_t)b * H * H * Ch + c;
  const bool lv = w > 0, rv = w < H - 1;
  float win[3][3][4];
  auto load_row = [&](int hh, float (&dst)[3][4]) {
    const bool hv = hh >= 0 && hh < H;
    const OpT* rowp = img + ((size_t)hh * H + w) * Ch;
    dw_load4<OpT>(rowp - Ch, hv && lv, dst[0]);
    dw_load4<OpT>(rowp, hv, dst[1]);
    dw_load4<OpT>(rowp + Ch, hv && rv, dst[2]);
  };
  load_row(h0 - 1, win[0]);
  load_row(h0, win[1]);
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    load_row(h0 + r + 1, win[(r + 2) % 3]);
    float acc[4] = {bz[0], bz[1], bz[2], bz[3]};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(win[(r + dy) % 3][dx][j], wreg[dy * 3 + dx][j], acc[j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = sizeof(OpT) == 2 ? gelu_fast(acc[j]) : gelu_erf(acc[j]);
    store4<OpT>(oimg + ((size_t)(h0 + r) * H + w) * Ch, acc[0], acc[1], acc[2], acc[3]);
  }
}

// bf16 specialisation: all 30 input loads of the thread's strip are issued before any math (memory
// level parallelism: the sliding version above is latency bound at ~1 TB/s), channel pairs are
// processed with packed fp32x2 FMAs (sm_100 FFMA2).
__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

static __global__ void __launch_bounds__(128, 4)
dwconv3x3_gelu_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                           const float* __restrict__ wt, const float* __restrict__ bias, int B, int H, int Ch) {
  const int ncs = Ch >> 6;
  const int slab = blockIdx.x % ncs;
  const int sp = blockIdx.x / ncs;
  const int tiles = H >> 3;
  const int b = sp / (tiles * tiles);
  const int trem = sp - b * tiles * tiles;
  const int h0 = (trem / tiles) * 8;
  const int w = (trem % tiles) * 8 + (threadIdx.x >> 4);
  const int c = slab * 64 + (threadIdx.x & 15) * 4;
  const __nv_bfloat16* img = in + (size_t)b * H * H * Ch + c;
  __nv_bfloat16* oimg = out + (size_t)b * H * H * Ch + c;
  const bool lv = w > 0, rv = w < H - 1;
  uint2 raw[10][3];
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const int hh = h0 + r - 1;
    const bool hv = hh >= 0 && hh < H;
    const __nv_bfloat16* rowp = img + ((size_t)hh * H + w) * Ch;
    raw[r][0] = (hv && lv) ? *reinterpret_cast<const uint2*>(rowp - Ch) : make_uint2(0u, 0u);
    raw[r][1] = hv ? *reinterpret_cast<const uint2*>(rowp) : make_uint2(0u, 0u);
    raw[r][2] = (hv && rv) ? *reinterpret_cast<const uint2*>(rowp + Ch) : make_uint2(0u, 0u);
  }
  float2 wreg[9][2], bz[2];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(wt + (size_t)t * Ch + c));
    wreg[t][0] = make_float2(a.x, a.y);
    wreg[t][1] = make_float2(a.z, a.w);
  }
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(bias + c));
    bz[0] = make_float2(a.x, a.y);
    bz[1] = make_float2(a.z, a.w);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float2 a0 = bz[0], a1 = bz[1];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const uint2 u = raw[r + dy][dx];
        a0 = __ffma2_rn(bf16x2_to_float2(u.x), wreg[dy * 3 + dx][0], a0);
        a1 = __ffma2_rn(bf16x2_to_float2(u.y), wreg[dy * 3 + dx][1], a1);
      }
    a0 = gelu_tanh2(a0);
    a1 = gelu_tanh2(a1);
    uint2 o;
    o.x = pack_bf16(a0.x, a0.y);
    o.y = pack_bf16(a1.x, a1.y);
    *reinterpret_cast<uint2*>(oimg + ((size_t)(h0 + r) * H + w) * Ch) = o;
  }
}

// ------------------------------------------------------------------------------------------
// im2col for Downsample = Conv2d(C, 2C, k=4, s=2, p=1) (uformerWM/model.py:763,768-775):
// A[(b,oh,ow)][(kh,kw,ci)] = x[b][2oh-1+kh][2ow-1+kw][ci] (zero outside).  One thread = 4 ci.
// ------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256)
im2col_4x4s2_kernel(const float* __restrict__ x, OpT* __restrict__ A, int B, int H, int C) {
  // one thread = (output token, kh, 8 input channels): the 4 kw taps -> 8 independent 16-byte loads
  const int cg = C >> 3;
  const int Ho = H >> 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Ho * Ho * 4 * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 8;
  size_t r = idx / cg;
  const int kh = (int)(r & 3);
  r >>= 2;
  const int ow = (int)(r % Ho);
  const int oh = (int)((r / Ho) % Ho);
  const size_t b = r / ((size_t)Ho * Ho);
  const int ih = 2 * oh - 1 + kh;
  const bool hv = ih >= 0 && ih < H;
  float4 v[4][2];
#pragma unroll
  for (int kw = 0; kw < 4; ++kw) {
    const int iw = 2 * ow - 1 + kw;
    if (hv && iw >= 0 && iw < H) {
      const float* src = x + ((b * H + ih) * H + iw) * C + c;
      v[kw][0] = *reinterpret_cast<const float4*>(src);
      v[kw][1] = *reinterpret_cast<const float4*>(src + 4);
    } else {
      v[kw][0] = v[kw][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if constexpr (OpMode<OpT>::v == 2) {
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(A) + ((b * Ho + oh) * Ho + ow) * (size_t)(32 * C);
#pragma unroll
    for (int kw = 0; kw < 4; ++kw) {
      const int col = (kh * 4 + kw) * C + c;
      split_store4(row, col, 16 * C, v[kw][0].x, v[kw][0].y, v[kw][0].z, v[kw][0].w);
      split_store4(row, col + 4, 16 * C, v[kw][1].x, v[kw][1].y, v[kw][1].z, v[kw][1].w);
    }
    return;
  }
  OpT* dst = A + ((b * Ho + oh) * Ho + ow) * (size_t)(16 * C) + (kh * 4) * C + c;
#pragma unroll
  for (int kw = 0; kw < 4; ++kw) {
    if constexpr (OpPlain16<OpT>::v) {
      constexpr bool F16 = OpMode<OpT>::v == 3;
      uint4 u;
      u.x = pack2_16<F16>(v[kw][0].x, v[kw][0].y); u.y = pack2_16<F16>(v[kw][0].z, v[kw][0].w);
      u.z = pack2_16<F16>(v[kw][1].x, v[kw][1].y); u.w = pack2_16<F16>(v[kw][1].z, v[kw][1].w);
      *reinterpret_cast<uint4*>(dst + kw * C) = u;
    } else {
      *reinterpret_cast<float4*>(dst + kw * C) = v[kw][0];
      *reinterpret_cast<float4*>(dst + kw * C + 4) = v[kw][1];
    }
  }
}

// ------------------------------------------------------------------------------------------
// Space-to-depth copy for the IMPLICIT-GEMM form of Downsample = Conv2d(C, 2C, k=4, s=2, p=1): with the padded image rows
// r = h + 1 and columns q = w + 1 paired as (sh, ph) = (r >> 1, r & 1), (sw, pw) = (q >> 1, q & 1), the 4x4 stride-2
// convolution is a 2x2 stride-1 convolution over S[b][sh][sw][(ph, pw, ci)] (sh <= H/2, sw <= H/2, 4C channels): output
// (oh, ow) reads cells (oh + a, ow + b), a, b in {0, 1}, i.e. kernel position (kh, kw) = (2a + ph, 2b + pw).  The dense
// kernel fetches those four shifted views as TMA boxes (gemm_tcgen05.cu, GemmArgs::dn_*), so every input element is
// written ONCE in 16-bit form (the im2col matrix wrote it four times).  Border cells (r = 0, r = H + 1, ...) are written
// as zeros here.  Split-bf16 operands: cell rows are [hi(4C) | lo(4C)].  One thread = 8 channels of one (cell, phase).
// ------------------------------------------------------------------------------------------
template <typename OpT>
__global__ void __launch_bounds__(256)
s2d_pad_kernel(const float* __restrict__ x, OpT* __restrict__ S, int B, int H, int C) {
  const int cg = C >> 3;
  const int Hs = (H >> 1) + 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)B * Hs * Hs * 4 * cg;
  if (idx >= total) return;
  const int c = (int)(idx % cg) * 8;
  size_t r = idx / cg;
  const int phw = (int)(r & 3);
  r >>= 2;
  const int sw = (int)(r % Hs);
  const int sh = (int)((r / Hs) % Hs);
  const size_t b = r / ((size_t)Hs * Hs);
  const int ih = 2 * sh + (phw >> 1) - 1, iw = 2 * sw + (phw & 1) - 1;
  float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
  if (ih >= 0 && ih < H && iw >= 0 && iw < H) {
    const float* src = x + ((b * H + ih) * H + iw) * C + c;
    v0 = *reinterpret_cast<const float4*>(src);
    v1 = *reinterpret_cast<const float4*>(src + 4);
  }
  const size_t cell = (b * Hs + sh) * Hs + sw;
  const int col = phw * C + c;
  if constexpr (OpMode<OpT>::v == 2) {
    __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(S) + cell * (size_t)(8 * C);
    split_store4(row, col, 4 * C, v0.x, v0.y, v0.z, v0.w);
    split_store4(row, col + 4, 4 * C, v1.x, v1.y, v1.z, v1.w);
  } else {
    constexpr bool F16 = OpMode<OpT>::v == 3;
    uint4 u;
    u.x = pack2_16<F16>(v0.x, v0.y); u.y = pack2_16<F16>(v0.z, v0.w);
    u.z = pack2_16<F16>(v1.x, v1.y); u.w = pack2_16<F16>(v1.z, v1.w);
    *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(S) + cell * (size_t)(4 * C) + col) = u;
  }
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

// Small direct kernels around the LeWin stacks: 3x3 projections, the image codec, pooling heads.
// Each is a few MFLOP per clip; they exist so that no host round trip or PyTorch op sits inside
// the embed / extract pass.
#pragma once
#include "uformer_kernels.cuh"

namespace wmk {

// InputProj (uformerWM/model.py:813-816,824-826): Conv2d(2,32,3,p=1) + LeakyReLU(0.01),
// NCHW [B][2][128][128] in -> token layout [B*16384][32] out.  One thread per pixel; the 576 weights
// travel as a kernel parameter, i.e. in the constant bank, so every FMA takes its weight as an
// immediate-like c[][] operand (no shared-memory traffic).  The 128-byte token rows are transposed
// through swizzled shared memory so that every store instruction of a warp writes 512 contiguous bytes
// (a thread storing its own row touches 32 half-filled sectors per instruction).
// ln_out != nullptr: the kernel also LayerNorms each finished token (norm1 of the stage-0 block that follows,
// uformerWM/model.py:982; gamma / beta ride in the same parameter block) and stores it as the bf16 operand of the
// QKV projection, so the separate LayerNorm pass over the largest activation of the model disappears.
struct InProjW { float w[32 * 18]; float b[32]; float ln_g[32]; float ln_b[32]; };

__global__ void __launch_bounds__(128)
input_proj_kernel(const float* __restrict__ x, float* __restrict__ out, const __grid_constant__ InProjW W, int B,
                  uint16_t* __restrict__ ln_out, int ln_fmt) {       // ln_fmt: 0 bf16, 1 fp16, 2 split-bf16 rows [hi(32) | lo(32)]
  __shared__ __align__(16) float4 stage[128 * 8];            // [pixel][8 float4], chunk index XOR (pixel & 7)
  const size_t pix0 = (size_t)blockIdx.x * blockDim.x;
  const size_t pix = pix0 + threadIdx.x;
  const size_t npix = (size_t)B * 16384;
  if (pix < npix) {
    const int wq = (int)(pix & 127), h = (int)((pix >> 7) & 127);
    const size_t b = pix >> 14;
    float in[18];
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int hh = h + dy - 1, wwp = wq + dx - 1;
          in[c * 9 + dy * 3 + dx] =
              (hh >= 0 && hh < 128 && wwp >= 0 && wwp < 128) ? x[((b * 2 + c) * 128 + hh) * 128 + wwp] : 0.f;
        }
    float v[32];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = g * 4 + j;
        float a = W.b[co];
#pragma unroll
        for (int t = 0; t < 18; ++t) a = fmaf(in[t], W.w[co * 18 + t], a);
        v[co] = a > 0.f ? a : 0.01f * a;
      }
      stage[threadIdx.x * 8 + (g ^ (threadIdx.x & 7))] = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
    }
    if (ln_out) {
      float mean = 0.f, var = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) mean += v[c];
      mean *= (1.0f / 32);
#pragma unroll
      for (int c = 0; c < 32; ++c) { const float d = v[c] - mean; var = fmaf(d, d, var); }
      const float rstd = rsqrtf(var * (1.0f / 32) + 1e-5f);
      uint4* o4 = reinterpret_cast<uint4*>(ln_out + pix * (ln_fmt == 2 ? 64 : 32));
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t pk[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = q * 8 + 2 * e;
          const float a0 = fmaf((v[c] - mean) * rstd, W.ln_g[c], W.ln_b[c]);
          const float a1 = fmaf((v[c + 1] - mean) * rstd, W.ln_g[c + 1], W.ln_b[c + 1]);
          if (ln_fmt == 2) split_pack2(a0, a1, pk[e], lo[e]);
          else pk[e] = ln_fmt == 1 ? pack2_f16(a0, a1) : pack2_bf16(a0, a1);
        }
        o4[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        if (ln_fmt == 2) o4[4 + q] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  }
  __syncthreads();
  float4* o = reinterpret_cast<float4*>(out + pix0 * 32);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int e = k * 128 + threadIdx.x;                       // float4 index inside the CTA's 16 KB of output
    const int p = e >> 3, g = e & 7;
    if (pix0 + p < npix) o[e] = stage[p * 8 + (g ^ (p & 7))];
  }
}

// OutputProj (uformerWM/model.py:845-847,857-862) Conv2d(64,2,3,p=1) on the token-layout decoder
// output + the residual y = x + noise (model.py:2419-2421).
// One CTA = an 8 x 32 pixel tile.  Phase 1: one thread per pixel of the 10 x 34 halo tile turns its 64
// channels into the 18 per-tap partial sums P[pixel][(dy,dx),o] = sum_c W[o][c][dy][dx] in[pixel][c] (each
// token row is read once per tile, 1.33x halo served by L2).  Phase 2: every output pixel gathers
// its 9 taps from shared memory; NCHW stores are coalesced along w.
constexpr int OP_TH = 8, OP_TW = 32, OP_PH = OP_TH + 2, OP_PW = OP_TW + 2, OP_NPIX = OP_PH * OP_PW;
constexpr int OP_THREADS = 352;      // >= OP_NPIX (340)
struct OutProjW { float w[64 * 18]; float b[2]; };     // [c][tap*2 + o]: constant-bank FMA operands
__global__ void __launch_bounds__(OP_THREADS)
output_proj_kernel(const float* __restrict__ tokens, const float* __restrict__ x, float* __restrict__ noise,
                   float* __restrict__ y, const __grid_constant__ OutProjW W, int B) {
  __shared__ float Ps[OP_NPIX * 19];                // [pixel][18 padded to 19]
  const int tiles_w = 128 / OP_TW, tiles_h = 128 / OP_TH;
  const int tile = blockIdx.x;
  const int b = tile / (tiles_w * tiles_h);
  const int trem = tile - b * tiles_w * tiles_h;
  const int h0 = (trem / tiles_w) * OP_TH, w0 = (trem % tiles_w) * OP_TW;
  if (threadIdx.x < OP_NPIX) {
    const int pr = threadIdx.x / OP_PW, pc = threadIdx.x - pr * OP_PW;
    const int hh = h0 + pr - 1, wwp = w0 + pc - 1;
    float acc[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) acc[k] = 0.f;
    if (hh >= 0 && hh < 128 && wwp >= 0 && wwp < 128) {
      const float4* src = reinterpret_cast<const float4*>(tokens + (((size_t)b * 128 + hh) * 128 + wwp) * 64);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = src[half * 8 + j];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float vv[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = (half * 8 + j) * 4 + e;
#pragma unroll
            for (int k = 0; k < 18; ++k) acc[k] = fmaf(vv[e], W.w[c * 18 + k], acc[k]);
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 18; ++k) Ps[threadIdx.x * 19 + k] = acc[k];
  }
  __syncthreads();
  if (threadIdx.x < OP_TH * OP_TW) {
    const int r = threadIdx.x / OP_TW, c = threadIdx.x - r * OP_TW;
    float a0 = W.b[0], a1 = W.b[1];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const float* pp = Ps + ((r + dy) * OP_PW + c + dx) * 19 + (dy * 3 + dx) * 2;
        a0 += pp[0];
        a1 += pp[1];
      }
    const size_t o0 = (((size_t)b * 2 + 0) * 128 + h0 + r) * 128 + w0 + c;
    const size_t o1 = o0 + 16384;
    if (noise) { noise[o0] = a0; noise[o1] = a1; }
    y[o0] = x[o0] + a0;
    y[o1] = x[o1] + a1;
  }
}

// Generic tiny NCHW 3x3 conv (pad 1) on 128x128 maps: stft_layer (uformerWM/model.py:2305-2309).
template <int CIN, int COUT, bool RELU>
__global__ void __launch_bounds__(256)
conv3x3_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w,
                    const float* __restrict__ bias, int B) {
  __shared__ float ws[COUT * CIN * 9];
  __shared__ float bs[COUT];
  for (int i = threadIdx.x; i < COUT * CIN * 9; i += blockDim.x) ws[i] = w[i];
  if (threadIdx.x < COUT) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (size_t)B * 16384) return;
  const int wq = (int)(pix & 127), h = (int)((pix >> 7) & 127);
  const size_t b = pix >> 14;
  float v[CIN * 9];
#pragma unroll
  for (int c = 0; c < CIN; ++c)
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int hh = h + dy - 1, wwp = wq + dx - 1;
        v[c * 9 + dy * 3 + dx] =
            (hh >= 0 && hh < 128 && wwp >= 0 && wwp < 128) ? in[((b * CIN + c) * 128 + hh) * 128 + wwp] : 0.f;
      }
#pragma unroll
  for (int co = 0; co < COUT; ++co) {
    float a = bs[co];
#pragma unroll
    for (int t = 0; t < CIN * 9; ++t) a = fmaf(v[t], ws[co * CIN * 9 + t], a);
    if (RELU) a = fmaxf(a, 0.f);
    out[((b * COUT + co) * 128 + h) * 128 + wq] = a;
  }
}

// ConvAutoencoder.encode (uformerWM/model.py:1720-1726): [1][32][32] -> [4][8][8]; one CTA per image.
// The image of clip `clip0 + blockIdx.x` is msg[msg_index(mm, clip)] (wmk_common.cuh MsgMap).
__global__ void __launch_bounds__(256)
wm_encode_kernel(const float* __restrict__ msg, MsgMap mm, int clip0, float* __restrict__ feat,
                 const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                 const float* __restrict__ b2) {
  __shared__ float img[34][34];
  __shared__ float hid[16][18][18];        // pooled conv1 output with a zero border
  __shared__ float sw1[16 * 9], sb1[16], sw2[4 * 16 * 9], sb2[4];
  const float* m = msg + msg_index(mm, clip0 + (int)blockIdx.x) * 1024;
  const int tid = threadIdx.x;
  for (int i = tid; i < 34 * 34; i += 256) {
    const int r = i / 34 - 1, c = i % 34 - 1;
    img[i / 34][i % 34] = (r >= 0 && r < 32 && c >= 0 && c < 32) ? m[r * 32 + c] : 0.f;
  }
  for (int i = tid; i < 16 * 18 * 18; i += 256) (&hid[0][0][0])[i] = 0.f;
  for (int i = tid; i < 144; i += 256) sw1[i] = w1[i];
  for (int i = tid; i < 576; i += 256) sw2[i] = w2[i];
  if (tid < 16) sb1[tid] = b1[tid];
  if (tid < 4) sb2[tid] = b2[tid];
  __syncthreads();
  for (int i = tid; i < 16 * 256; i += 256) {
    const int co = i >> 8, ph = (i >> 4) & 15, pw = i & 15;
    float best = -INFINITY;
    for (int sy = 0; sy < 2; ++sy)
      for (int sx = 0; sx < 2; ++sx) {
        const int r = 2 * ph + sy, c = 2 * pw + sx;
        float a = sb1[co];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) a = fmaf(img[r + dy][c + dx], sw1[co * 9 + dy * 3 + dx], a);
        best = fmaxf(best, a);
      }
    hid[co][ph + 1][pw + 1] = fmaxf(best, 0.f);
  }
  __syncthreads();
  {
    const int co = tid >> 6, ph = (tid >> 3) & 7, pw = tid & 7;
    float best = -INFINITY;
    for (int sy = 0; sy < 2; ++sy)
      for (int sx = 0; sx < 2; ++sx) {
        const int r = 2 * ph + sy, c = 2 * pw + sx;
        float a = sb2[co];
        for (int ci = 0; ci < 16; ++ci)
#pragma unroll
          for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              a = fmaf(hid[ci][r + dy][c + dx], sw2[(co * 16 + ci) * 9 + dy * 3 + dx], a);
        best = fmaxf(best, a);
      }
    feat[(size_t)blockIdx.x * 256 + tid] = fmaxf(best, 0.f);
  }
}

// ConvAutoencoder.decode (uformerWM/model.py:1711-1718) on feat (+ optional addend, the
// `feature_wm_ori + conv4_downsample` of model.py:2403): [4][8][8] -> logits / sigmoid [32][32].
__global__ void __launch_bounds__(256)
wm_decode_kernel(const float* __restrict__ feat, const float* __restrict__ addend, float* __restrict__ wm,
                 float* __restrict__ logits, const float* __restrict__ w1, const float* __restrict__ b1,
                 const float* __restrict__ w2, const float* __restrict__ b2) {
  __shared__ float f[4][8][8];
  __shared__ float hid[16][16][16];
  __shared__ float sw1[4 * 16 * 4], sb1[16], sw2[16 * 4];
  const int tid = threadIdx.x;
  const size_t b = blockIdx.x;
  (&f[0][0][0])[tid] = feat[b * 256 + tid] + (addend ? addend[b * 256 + tid] : 0.f);
  sw1[tid] = w1[tid];                       // [ci][co][i][j]
  if (tid < 16) sb1[tid] = b1[tid];
  if (tid < 64) sw2[tid] = w2[tid];         // [ci][0][i][j]
  __syncthreads();
  for (int e = tid; e < 4096; e += 256) {
    const int co = e >> 8, r = (e >> 4) & 15, c = e & 15;
    float a = sb1[co];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) a = fmaf(f[ci][r >> 1][c >> 1], sw1[((ci * 16 + co) * 2 + (r & 1)) * 2 + (c & 1)], a);
    hid[co][r][c] = fmaxf(a, 0.f);
  }
  __syncthreads();
  const float bias2 = b2[0];
  for (int e = tid; e < 1024; e += 256) {
    const int r = e >> 5, c = e & 31;
    float a = bias2;
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) a = fmaf(hid[ci][r >> 1][c >> 1], sw2[(ci * 2 + (r & 1)) * 2 + (c & 1)], a);
    if (logits) logits[b * 1024 + e] = a;
    if (wm) wm[b * 1024 + e] = 1.0f / (1.0f + expf(-a));
  }
}

// MaxPool2d((16,8)) over the bottleneck token matrix [B][64][512] -> [B][4][64]
// (uformerWM/model.py:2250,2398-2400).
__global__ void __launch_bounds__(256)
bottleneck_maxpool_kernel(const float* __restrict__ conv4, float* __restrict__ out, int B) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * 256) return;
  const int j = (int)(idx & 63), i = (int)((idx >> 6) & 3);
  const size_t b = idx >> 8;
  float best = -INFINITY;
  for (int r = 0; r < 16; ++r)
    for (int c = 0; c < 8; ++c) best = fmaxf(best, conv4[(b * 64 + 16 * i + r) * 512 + 8 * j + c]);
  out[idx] = best;
}

// EncoderTransformerWM.conv2 = Conv2d(1,1,8,stride=(16,8)) over [B][1][64][512] -> [B][4][64]
// (uformerWM/model.py:1566,1580-1582).
__global__ void __launch_bounds__(256)
extract_head_kernel(const float* __restrict__ conv4, float* __restrict__ out, const float* __restrict__ w,
                    const float* __restrict__ bias, int B) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * 256) return;
  const int j = (int)(idx & 63), i = (int)((idx >> 6) & 3);
  const size_t b = idx >> 8;
  float a = bias[0];
  for (int u = 0; u < 8; ++u)
    for (int v = 0; v < 8; ++v) a = fmaf(conv4[(b * 64 + 16 * i + u) * 512 + 8 * j + v], __ldg(w + u * 8 + v), a);
  out[idx] = a;
}

}  // namespace wmk

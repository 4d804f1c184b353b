// CNN building blocks and spectrogram-domain noise layers:
//   * 3x3 convolution / 2x2 transposed convolution / 2x2 max-pool on NCHW fp32 tensors with the
//     BatchNorm (eval) affine and the activation fused - the layers of ModelA
//     (uformerWM/model.py:3000-3066) and of the HiDDeN Decoder / ConvBNRelu
//     (hidden/model/decoder.py:12-40, hidden/model/conv_bn_relu.py:7-18);
//   * the HiDDeN noise layers (hidden/noise_layers/{crop,cropout,dropout,resize,quantization}.py).
// Channel counts here are 1..64 and the tensors are small: these are memory-bound direct kernels
// (input tile + halo staged in shared memory, weights in shared memory, 16 output channels per
// thread in registers).
#include <stdlib.h>

#include "uformer_kernels.cuh"

namespace wmk {
namespace {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SIGMOID = 3 };

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : slope * v;
  if (act == ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}


// output channels per CTA (COB) and input channels per shared-memory pass (CIB) are template parameters: the layers with
// 1 / 2 input or output channels (ModelA's first / last layers and their data gradients, all at 128 x 128) would waste
// 4x / 8x of the FMAs and weight loads in the generic 8 x 16 blocking

// y[b][co_off+co][h][w] = act( scale[co] * (sum_{ci,dy,dx} x[b][ci][h+dy-1][w+dx-1] w[co][ci][dy][dx] + bias[co]) + shift[co] )
// Tile = 32 columns x 8 PX rows; a warp covers one 32-pixel row segment (conflict-free shared-memory reads, 128-byte
// stores), a thread PX pixels (PX consecutive rows of one column): a weight fetched from shared memory feeds PX FMAs, which moves the
// COB = 16 layers from the load / store pipe (5 loads per 16 FMAs) to the FMA pipe.  The input tile (+ halo) arrives as
// 16-byte loads: row = [3 pad][halo][32 pixels][halo][3 pad].
constexpr int CTW = 32, CXP = 40;
template <int CIB, int COB, int PX>
__global__ void __launch_bounds__(256)
conv3x3_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
               const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
               int Cin, int Cout, int H, int W, int co_off, int Ctot, int act, float slope, int wflip) {
  constexpr int TH = 8 * PX;
  __shared__ __align__(16) float tile[CIB][TH + 2][CXP];
  __shared__ __align__(16) float ws[CIB][9][COB];
  const int tiles_w = (W + CTW - 1) / CTW;
  const int th = blockIdx.x / tiles_w, tw = blockIdx.x % tiles_w;
  const int co0 = blockIdx.y * COB;
  const int b = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int wq = tw * CTW + tx;
  const bool vec_ok = (W & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  float acc[PX][COB];
#pragma unroll
  for (int p = 0; p < PX; ++p)
#pragma unroll
    for (int j = 0; j < COB; ++j) acc[p][j] = 0.f;
  for (int c0 = 0; c0 < Cin; c0 += CIB) {
    __syncthreads();
#pragma unroll 4
    for (int e = threadIdx.x; e < CIB * (TH + 2) * 10; e += 256) {      // per row: halo, 8 x 16 bytes, halo
      const int ci = e / ((TH + 2) * 10), rem = e - ci * ((TH + 2) * 10), r = rem / 10, k = rem - r * 10;
      const int hh = th * TH + r - 1;
      const bool ok = c0 + ci < Cin && hh >= 0 && hh < H;
      const float* src = x + (((size_t)b * Cin + c0 + ci) * H + (ok ? hh : 0)) * W + tw * CTW;
      float* dst = &tile[ci][r][0];
      if (k == 0) dst[3] = (ok && tw > 0) ? src[-1] : 0.f;
      else if (k == 9) dst[36] = (ok && tw * CTW + CTW < W) ? src[CTW] : 0.f;
      else {
        const int c = 4 * (k - 1), ww = tw * CTW + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) {
          if (vec_ok && ww + 3 < W) v = *reinterpret_cast<const float4*>(src + c);
          else {
            if (ww < W) v.x = src[c];
            if (ww + 1 < W) v.y = src[c + 1];
            if (ww + 2 < W) v.z = src[c + 2];
            if (ww + 3 < W) v.w = src[c + 3];
          }
        }
        *reinterpret_cast<float4*>(dst + 4 + c) = v;
      }
    }
    for (int e = threadIdx.x; e < CIB * 9 * COB; e += 256) {
      const int ci = e / (9 * COB), t = (e / COB) % 9, j = e % COB;
      float v = 0.f;
      // wflip (data gradient): w is the FORWARD layer's [Cin here = its Cout][Cout here = its Cin][3][3] tensor, read
      // transposed with the taps reversed - correlation with the flipped kernel - so no flipped copy is ever made
      if (c0 + ci < Cin && co0 + j < Cout)
        v = wflip ? w[((size_t)(c0 + ci) * Cout + co0 + j) * 9 + (8 - t)] : w[((size_t)(co0 + j) * Cin + c0 + ci) * 9 + t];
      ws[ci][t][j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CIB; ++ci) {
      // the thread's PX output rows are consecutive: its (PX + 2) x 3 input window is read once and shared by the taps
      float xr[PX + 2][3];
#pragma unroll
      for (int r = 0; r < PX + 2; ++r)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) xr[r][kx] = tile[ci][ty * PX + r][tx + kx + 3];
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        if constexpr (COB % 4 == 0) {
          const float4* w4 = reinterpret_cast<const float4*>(&ws[ci][t][0]);      // 4 weights per shared-memory load
#pragma unroll
          for (int j4 = 0; j4 < COB / 4; ++j4) {
            const float4 wv = w4[j4];
#pragma unroll
            for (int p = 0; p < PX; ++p) {
              const float v = xr[p + t / 3][t % 3];
              acc[p][4 * j4] = fmaf(v, wv.x, acc[p][4 * j4]);
              acc[p][4 * j4 + 1] = fmaf(v, wv.y, acc[p][4 * j4 + 1]);
              acc[p][4 * j4 + 2] = fmaf(v, wv.z, acc[p][4 * j4 + 2]);
              acc[p][4 * j4 + 3] = fmaf(v, wv.w, acc[p][4 * j4 + 3]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < COB; ++j) {
            const float wv = ws[ci][t][j];
#pragma unroll
            for (int p = 0; p < PX; ++p) acc[p][j] = fmaf(xr[p + t / 3][t % 3], wv, acc[p][j]);
          }
        }
      }
    }
  }
  if (wq >= W) return;
  float bj[COB], sc[COB], sh[COB];
#pragma unroll
  for (int j = 0; j < COB; ++j) {
    const int co = co0 + j < Cout ? co0 + j : Cout - 1;
    bj[j] = bias ? bias[co] : 0.f;
    sc[j] = scale ? scale[co] : 1.f;
    sh[j] = scale ? shift[co] : 0.f;
  }
#pragma unroll
  for (int p = 0; p < PX; ++p) {
    const int h = th * TH + ty * PX + p;
    if (h >= H) continue;
#pragma unroll
    for (int j = 0; j < COB; ++j) {
      const int co = co0 + j;
      if (co >= Cout) break;
      float v = acc[p][j] + bj[j];
      if (scale) v = v * sc[j] + sh[j];
      y[(((size_t)b * Ctot + co_off + co) * H + h) * W + wq] = apply_act(v, act, slope);
    }
  }
}

// ConvTranspose2d(k=2, s=2): y[b][co][2h+i][2w+j] = act(scale*(sum_ci x[b][ci][h][w] w[ci][co][i][j] + bias) + shift)
__global__ void __launch_bounds__(256)
convT2x2_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                int B, int Cin, int Cout, int H, int W, int act, float slope) {
  extern __shared__ float wsm[];             // [Cin][Cout][4]
  for (int e = threadIdx.x; e < Cin * Cout * 4; e += blockDim.x) wsm[e] = w[e];
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * H * W) return;
  const int wq = (int)(idx % W), h = (int)((idx / W) % H);
  const size_t b = idx / ((size_t)H * W);
  for (int co = 0; co < Cout; ++co) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ci = 0; ci < Cin; ++ci) {
      const float v = x[((b * Cin + ci) * H + h) * W + wq];
      const float* wp = wsm + ((size_t)ci * Cout + co) * 4;
      a[0] = fmaf(v, wp[0], a[0]); a[1] = fmaf(v, wp[1], a[1]); a[2] = fmaf(v, wp[2], a[2]); a[3] = fmaf(v, wp[3], a[3]);
    }
#pragma unroll
    for (int ij = 0; ij < 4; ++ij) {
      float v = a[ij] + (bias ? bias[co] : 0.f);
      if (scale) v = v * scale[co] + shift[co];
      y[((b * Cout + co) * (2 * H) + 2 * h + (ij >> 1)) * (size_t)(2 * W) + 2 * wq + (ij & 1)] = apply_act(v, act, slope);
    }
  }
}

// Register-blocked form for Cout <= COB (ModelA: 16 and 2): the input value of a pixel is read once per input channel and
// feeds 4 COB FMAs (the generic kernel above re-reads x for every output channel: 89 us for 76 MB at ConvTranspose2d(33, 16)).
template <int COB>
__global__ void __launch_bounds__(256)
convT2x2_rb_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ w,
                   const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                   int B, int Cin, int Cout, int H, int W, int act, float slope) {
  extern __shared__ __align__(16) float wsm[];             // [Cin][COB][4], output channels >= Cout zero
  for (int e = threadIdx.x; e < Cin * COB * 4; e += blockDim.x) {
    const int ci = e / (COB * 4), r = e - ci * COB * 4, co = r >> 2;
    wsm[e] = co < Cout ? w[((size_t)ci * Cout + co) * 4 + (r & 3)] : 0.f;
  }
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * H * W) return;
  const int wq = (int)(idx % W), h = (int)((idx / W) % H);
  const size_t b = idx / ((size_t)H * W);
  float a[COB][4];
#pragma unroll
  for (int co = 0; co < COB; ++co) a[co][0] = a[co][1] = a[co][2] = a[co][3] = 0.f;
  const float* xp = x + (b * Cin * H + h) * W + wq;
  for (int ci = 0; ci < Cin; ++ci) {
    const float v = xp[(size_t)ci * H * W];
    const float4* w4 = reinterpret_cast<const float4*>(wsm + ci * COB * 4);
#pragma unroll
    for (int co = 0; co < COB; ++co) {
      const float4 ww = w4[co];
      a[co][0] = fmaf(v, ww.x, a[co][0]); a[co][1] = fmaf(v, ww.y, a[co][1]);
      a[co][2] = fmaf(v, ww.z, a[co][2]); a[co][3] = fmaf(v, ww.w, a[co][3]);
    }
  }
#pragma unroll
  for (int co = 0; co < COB; ++co) {
    if (co >= Cout) break;
    const float bb = bias ? bias[co] : 0.f;
    float o[4];
#pragma unroll
    for (int ij = 0; ij < 4; ++ij) {
      float v = a[co][ij] + bb;
      if (scale) v = v * scale[co] + shift[co];
      o[ij] = apply_act(v, act, slope);
    }
    float* yp = y + ((b * Cout + co) * (2 * H) + 2 * h) * (size_t)(2 * W) + 2 * wq;
    *reinterpret_cast<float2*>(yp) = make_float2(o[0], o[1]);
    *reinterpret_cast<float2*>(yp + 2 * W) = make_float2(o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256)
maxpool2x2_kernel(const float* __restrict__ x, float* __restrict__ y, size_t planes, int H, int W) {
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= planes * Ho * Wo) return;
  const int wq = (int)(idx % Wo), h = (int)((idx / Wo) % Ho);
  const size_t p = idx / ((size_t)Ho * Wo);
  const float* s = x + (p * H + 2 * h) * W + 2 * wq;
  y[idx] = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[W], s[W + 1]));
}

// ---------------------------------------------------------------------------- noise layers
// Cropout (cropout.py:16-28) / Dropout (dropout.py:15-28): out = keep ? noised : cover.
// mask == nullptr: keep inside the rectangle [h0,h1) x [w0,w1); else keep = mask[h][w] != 0.
__global__ void __launch_bounds__(256)
mix_kernel(const float* __restrict__ noised, const float* __restrict__ cover, float* __restrict__ out, size_t planes,
           int H, int W, int h0, int h1, int w0, int w1, const float* __restrict__ mask) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= planes * H * W) return;
  const int wq = (int)(idx % W), h = (int)((idx / W) % H);
  const bool keep = mask ? mask[h * W + wq] != 0.f : (h >= h0 && h < h1 && wq >= w0 && wq < w1);
  out[idx] = keep ? noised[idx] : cover[idx];
}

// Crop (crop.py:63-75): out = in[:, :, h0:h1, w0:w1];  Resize nearest (resize.py:17-26,
// F.interpolate(mode='nearest', scale_factor=r)): src = floor(dst * (1/r)).
__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ in, float* __restrict__ out, size_t planes, int H, int W, int Ho, int Wo,
                int h0, int w0, float inv_scale, int nearest) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= planes * Ho * Wo) return;
  const int wq = (int)(idx % Wo), h = (int)((idx / Wo) % Ho);
  const size_t p = idx / ((size_t)Ho * Wo);
  int sh, sw;
  if (nearest) {
    sh = min((int)floorf(h * inv_scale), H - 1);
    sw = min((int)floorf(wq * inv_scale), W - 1);
  } else {
    sh = h0 + h;
    sw = w0 + wq;
  }
  out[idx] = in[(p * H + sh) * W + sw];
}

// Quantization (quantization.py:32-45): min-max to [0,255], x + sum_{n=1..10} (-1)^n/(pi n) sin(2 pi n x),
// min-max of the result back to the input's range.  Three passes: range, transform + range, rescale.
__device__ __forceinline__ float atomic_min_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) > v) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
  return __int_as_float(old);
}
__device__ __forceinline__ float atomic_max_f(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) < v) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
  return __int_as_float(old);
}
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, size_t n, float* __restrict__ mm) {
  float lo = INFINITY, hi = -INFINITY;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomic_min_f(mm, lo);
    atomic_max_f(mm + 1, hi);
  }
}
__global__ void __launch_bounds__(256)
quant_round_kernel(const float* __restrict__ x, float* __restrict__ t, size_t n, const float* __restrict__ mm) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lo = mm[0], hi = mm[1];
  float v = (x[i] - lo) / (hi - lo);
  v = v * 255.0f;                                   // transform(tensor, (0, 255))
  v = fminf(fmaxf(v, 0.f), 255.0f);
  // the reference's weights / scales are float64 tensors, so the rounding term is evaluated in fp64
  double z = 0.0;
#pragma unroll
  for (int k = 1; k <= 10; ++k) {
    const double wk = ((k & 1) ? -1.0 : 1.0) / (3.14159265358979323846 * k);       // (-1)^(n+1)/(pi (n+1)), n = k-1
    z += wk * sinpi(2.0 * k * (double)v);
  }
  t[i] = (float)((double)v + z);
}
__global__ void __launch_bounds__(256)
rescale_kernel(const float* __restrict__ t, float* __restrict__ out, size_t n, const float* __restrict__ mm_t,
               const float* __restrict__ mm_x) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = (t[i] - mm_t[0]) / (mm_t[1] - mm_t[0]);
  out[i] = v * (mm_x[1] - mm_x[0]) + mm_x[0];
}


// JpegCompression (hidden/noise_layers/jpeg_compression.py:128-160), 3-channel input as in the
// reference (:53-55): RGB->YUV, 8x8 DCT-II (the reference's 64-filter stride-8 conv, :40-41,93-101),
// keep the first 25 / 9 / 9 zig-zag coefficients of Y / U / V (:28-38), inverse (:44-46), YUV->RGB,
// zero padding of H, W to multiples of 8 and un-padding.  One CTA = one 8x8 block of one image:
// 192 threads = (channel, y, x); the three planes go through shared memory so the colour
// transforms are fused with the DCT and nothing but the output reaches HBM.
struct JpegKeep { unsigned long long m[3]; };             // bit (u*8+v) set = coefficient (u,v) kept; a kernel parameter

__global__ void __launch_bounds__(192)
jpeg_kernel(const float* __restrict__ in, float* __restrict__ out, int H, int W, const JpegKeep K) {
  __shared__ float pl[3][8][9], tmp[3][8][9];
  const int tid = threadIdx.x, c = tid >> 6, yy = (tid >> 3) & 7, xx = tid & 7;
  const int bx = blockIdx.x, by = blockIdx.y, b = blockIdx.z;
  const int h = by * 8 + yy, w = bx * 8 + xx;
  const bool inside = h < H && w < W;
  const size_t plane = (size_t)H * W;
  const float* src = in + (size_t)b * 3 * plane + (size_t)h * W + w;
  const float r = inside ? src[0] : 0.f, g = inside ? src[plane] : 0.f, bl = inside ? src[2 * plane] : 0.f;
  float v;
  if (c == 0) v = 0.299f * r + 0.587f * g + 0.114f * bl;
  else if (c == 1) v = -0.14713f * r + -0.28886f * g + 0.436f * bl;
  else v = 0.615f * r + -0.51499f * g + -0.10001f * bl;
  pl[c][yy][xx] = v;
  __syncthreads();
  // forward: D[u][v] = sum_{y,x} cos(pi/8 (y+1/2) u) cos(pi/8 (x+1/2) v) P[y][x]; rows first (u = yy role)
  {
    float a = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) a = fmaf(cospif((n + 0.5f) * xx * 0.125f), pl[c][yy][n], a);   // along x: v = xx
    tmp[c][yy][xx] = a;
  }
  __syncthreads();
  {
    float a = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) a = fmaf(cospif((n + 0.5f) * yy * 0.125f), tmp[c][n][xx], a);  // along y: u = yy
    const bool keep = (K.m[c] >> (yy * 8 + xx)) & 1ull;
    pl[c][yy][xx] = keep ? a : 0.f;
  }
  __syncthreads();
  // inverse: P[y][x] = sum_{u,v} i(u,y) i(v,x) D[u][v],  i(n,k) = ((n==0 ? -1/2 : 0) + cos(pi/8 (k+1/2) n)) / 4
  {
    float a = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) a = fmaf(((n == 0 ? -0.5f : 0.f) + cospif((xx + 0.5f) * n * 0.125f)) * 0.25f, pl[c][yy][n], a);
    tmp[c][yy][xx] = a;
  }
  __syncthreads();
  {
    float a = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) a = fmaf(((n == 0 ? -0.5f : 0.f) + cospif((yy + 0.5f) * n * 0.125f)) * 0.25f, tmp[c][n][xx], a);
    pl[c][yy][xx] = a;
  }
  __syncthreads();
  if (!inside) return;
  const float Y = pl[0][yy][xx], U = pl[1][yy][xx], V = pl[2][yy][xx];
  float o;
  if (c == 0) o = Y + 1.13983f * V;
  else if (c == 1) o = Y + -0.39465f * U + -0.58060f * V;
  else o = Y + 2.03211f * U;
  out[((size_t)b * 3 + c) * plane + (size_t)h * W + w] = o;
}

// Magnitude / phase view of a spectrogram clip (the 1-channel "STFT magnitudes" input of the HiDDeN
// flavour, BASELINE config 3; the reference keeps re/im and has no such op - SURVEY 8a note).
// spec [n][2][plane] (re, im) <-> mag [n][plane], phase [n][plane]
__global__ void __launch_bounds__(256)
magphase_split_kernel(const float* __restrict__ spec, float* __restrict__ mag, float* __restrict__ phase, size_t n, size_t plane) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * plane) return;
  const size_t b = i / plane, r = i - b * plane;
  const float re = spec[(2 * b) * plane + r], im = spec[(2 * b + 1) * plane + r];
  mag[i] = hypotf(re, im);
  if (phase) phase[i] = atan2f(im, re);
}
__global__ void __launch_bounds__(256)
magphase_merge_kernel(const float* __restrict__ mag, const float* __restrict__ phase, float* __restrict__ spec, size_t n, size_t plane) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * plane) return;
  const size_t b = i / plane, r = i - b * plane;
  float sn, cs;
  sincosf(phase[i], &sn, &cs);
  spec[(2 * b) * plane + r] = mag[i] * cs;
  spec[(2 * b + 1) * plane + r] = mag[i] * sn;
}

// ---------------------------------------------------------------------------- tensor-core path of the HiDDeN decoder
// (hidden/model/decoder.py:12-40): activations NHWC bf16, the 64 -> 64 / 64 -> message_length ConvBNRelu layers run as
// implicit GEMMs on tcgen05 (gemm_tcgen05.cu, conv_H mode); these are the layout-changing layers around them.
// first layer: Conv2d(1, 64, 3, p=1) + BN + ReLU, NCHW fp32 [B][1][H][W] -> NHWC bf16 [B][H][W][64]; thread = pixel x 8 channels
__global__ void __launch_bounds__(256)
conv3x3_c1_nhwc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ w,
                       const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                       size_t npix, int H, int W) {
  __shared__ float ws[64 * 9], sc[64], sh[64];
  for (int i = threadIdx.x; i < 576; i += 256) ws[i] = w[i];
  if (threadIdx.x < 64) {
    const float s = scale ? scale[threadIdx.x] : 1.f;
    sc[threadIdx.x] = s;
    sh[threadIdx.x] = (bias ? bias[threadIdx.x] : 0.f) * s + (shift ? shift[threadIdx.x] : 0.f);
  }
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npix * 8) return;
  const int g8 = (int)(idx & 7);
  const size_t pix = idx >> 3;
  const int wq = (int)(pix % W), h = (int)((pix / W) % H);
  const size_t b = pix / ((size_t)H * W);
  float v[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int hh = h + t / 3 - 1, ww = wq + t % 3 - 1;
    v[t] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(b * H + hh) * W + ww] : 0.f;
  }
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float r[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int co = g8 * 8 + 2 * j + e;
      float a = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) a = fmaf(v[t], ws[co * 9 + t], a);
      r[e] = fmaxf(fmaf(a, sc[co], sh[co]), 0.f);
    }
    o[j] = pack_bf16(r[0], r[1]);
  }
  *reinterpret_cast<uint4*>(y + pix * 64 + g8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
}

// MaxPool2d(2,2) on NHWC bf16 [B][H][W][C] (C multiple of 8); thread = output pixel x 8 channels
__global__ void __launch_bounds__(256)
maxpool2x2_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t total, int H, int W, int C) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cg = C >> 3, Ho = H >> 1, Wo = W >> 1;
  const int g8 = (int)(idx % cg);
  size_t r = idx / cg;
  const int wo = (int)(r % Wo), ho = (int)((r / Wo) % Ho);
  const size_t b = r / ((size_t)Wo * Ho);
  const __nv_bfloat16* s = x + (((b * H + 2 * ho) * W + 2 * wo) * C + g8 * 8);
  const uint4 a = *reinterpret_cast<const uint4*>(s), bq = *reinterpret_cast<const uint4*>(s + C);
  const uint4 c = *reinterpret_cast<const uint4*>(s + (size_t)W * C), d = *reinterpret_cast<const uint4*>(s + (size_t)W * C + C);
  const uint32_t av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {bq.x, bq.y, bq.z, bq.w}, cv[4] = {c.x, c.y, c.z, c.w}, dv[4] = {d.x, d.y, d.z, d.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 m = __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162*>(&av[j]), *reinterpret_cast<const __nv_bfloat162*>(&bv[j])),
                                     __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&cv[j]), *reinterpret_cast<const __nv_bfloat162*>(&dv[j])));
    o[j] = *reinterpret_cast<const uint32_t*>(&m);
  }
  *reinterpret_cast<uint4*>(y + (((b * Ho + ho) * Wo + wo) * C + g8 * 8)) = make_uint4(o[0], o[1], o[2], o[3]);
}

// last layer: Conv2d(Cin, 1, 3, p=1) + BN + ReLU from NHWC bf16 [B][H][W][Cp] (Cp = Cin padded to a multiple of 8, padding
// channels ignored) to NCHW fp32 [B][1][H][W]; wt: [9][Cp] fp32 (tap-major, zero in the padding channels); a warp-quarter
// (8 lanes) shares one pixel, each lane owning 8-channel groups
__global__ void __launch_bounds__(256)
conv3x3_nhwc_to1_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, const float* __restrict__ wt, float bias,
                        float scale, float shift, size_t npix, int H, int W, int Cp) {
  extern __shared__ float wsm[];                  // [9][Cp]
  for (int i = threadIdx.x; i < 9 * Cp; i += 256) wsm[i] = wt[i];
  __syncthreads();
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t pix = gid >> 2;
  const int part = (int)(gid & 3);                // 4 lanes per pixel
  float a = 0.f;
  if (pix < npix) {
    const int wq = (int)(pix % W), h = (int)((pix / W) % H);
    const size_t b = pix / ((size_t)H * W);
    for (int t = 0; t < 9; ++t) {
      const int hh = h + t / 3 - 1, ww = wq + t % 3 - 1;
      if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
      const __nv_bfloat16* s = x + ((b * H + hh) * W + ww) * Cp;
      for (int c8 = part * 8; c8 < Cp; c8 += 32) {
        const uint4 u = *reinterpret_cast<const uint4*>(s + c8);
        const uint32_t uv[4] = {u.x, u.y, u.z, u.w};
        const float* wp = wsm + t * Cp + c8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          a = fmaf(__uint_as_float(uv[j] << 16), wp[2 * j], a);
          a = fmaf(__uint_as_float(uv[j] & 0xffff0000u), wp[2 * j + 1], a);
        }
      }
    }
  }
  a += __shfl_xor_sync(0xffffffffu, a, 1);
  a += __shfl_xor_sync(0xffffffffu, a, 2);
  if (part == 0 && pix < npix) y[pix] = fmaxf(fmaf(a + bias, scale, shift), 0.f);
}

}  // namespace
}  // namespace wmk

using namespace wmk;

static int conv3x3_launch(const float* x, float* y, const float* w, const float* bias, const float* scale,
                          const float* shift, int B, int Cin, int Cout, int H, int W, int out_ch_offset,
                          int out_ch_total, int act, float slope, int wflip, void* stream) {
  WMK_REQUIRE(x && y && w && B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && out_ch_offset >= 0 &&
                  out_ch_offset + Cout <= out_ch_total && act >= 0 && act <= 3 && (!scale == !shift) && B <= 65535,
              "conv3x3: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 4.0 * B * H * W * (Cin + Cout), st);
  const int cob = Cout <= 2 ? 2 : 16;
  static const int px_max = getenv("WMK_CONV_PX") ? atoi(getenv("WMK_CONV_PX")) : 4;
  const int px = (H >= 32 && px_max >= 4) ? 4 : (H >= 16 && px_max >= 2) ? 2 : 1;
  dim3 grid(cdiv(H, 8 * px) * cdiv(W, CTW), cdiv(Cout, cob), B);
#define WMK_CONV_LAUNCH(CIB, COB)                                                                                                   \
  do {                                                                                                                              \
    if (px == 4) conv3x3_kernel<CIB, COB, 4><<<grid, 256, 0, st>>>(x, y, w, bias, scale, shift, Cin, Cout, H, W, out_ch_offset, out_ch_total, act, slope, wflip); \
    else if (px == 2) conv3x3_kernel<CIB, COB, 2><<<grid, 256, 0, st>>>(x, y, w, bias, scale, shift, Cin, Cout, H, W, out_ch_offset, out_ch_total, act, slope, wflip); \
    else conv3x3_kernel<CIB, COB, 1><<<grid, 256, 0, st>>>(x, y, w, bias, scale, shift, Cin, Cout, H, W, out_ch_offset, out_ch_total, act, slope, wflip); \
  } while (0)
  if (Cin <= 2 && Cout <= 2) WMK_CONV_LAUNCH(2, 2);
  else if (Cin <= 2) WMK_CONV_LAUNCH(2, 16);
  else if (Cout <= 2) WMK_CONV_LAUNCH(8, 2);
  else WMK_CONV_LAUNCH(8, 16);
#undef WMK_CONV_LAUNCH
  WMK_CHECK_LAUNCH("conv3x3_kernel");
  return 0;
}

extern "C" int wmk_conv3x3_f32(const float* x, float* y, const float* w, const float* bias, const float* scale,
                               const float* shift, int B, int Cin, int Cout, int H, int W, int out_ch_offset,
                               int out_ch_total, int act, float slope, void* stream) {
  return conv3x3_launch(x, y, w, bias, scale, shift, B, Cin, Cout, H, W, out_ch_offset, out_ch_total, act, slope, 0, stream);
}

// data gradient of Conv2d(Cin, Cout, 3, padding=1): dx [B][Cin][H][W] = correlation of dy [B][Cout][H][W] with the flipped,
// transposed forward weights w [Cout][Cin][3][3] (read in place)
extern "C" int wmk_conv3x3_dgrad_f32(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H, int W,
                                     void* stream) {
  return conv3x3_launch(dy, dx, w, nullptr, nullptr, nullptr, B, Cout, Cin, H, W, 0, Cin, 0, 0.f, 1, stream);
}

extern "C" int wmk_convT2x2_f32(const float* x, float* y, const float* w, const float* bias, const float* scale,
                                const float* shift, int B, int Cin, int Cout, int H, int W, int act, float slope,
                                void* stream) {
  WMK_REQUIRE(x && y && w && B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && act >= 0 && act <= 3 &&
                  (!scale == !shift) && (size_t)Cin * Cout * 16 <= 96 * 1024,
              "convT2x2: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 4.0 * B * H * W * (Cin + 4 * Cout), st);
  const size_t smem = (size_t)Cin * Cout * 16;
  if ((Cout <= 2 || (Cout > 4 && Cout <= 16)) && ((uintptr_t)y & 7) == 0 && (size_t)Cin * (Cout <= 2 ? 2 : 16) * 16 <= 48 * 1024) {      // register-blocked: x read once per input channel
    if (Cout <= 2) convT2x2_rb_kernel<2><<<cdiv((size_t)B * H * W, 256), 256, (size_t)Cin * 2 * 16, st>>>(x, y, w, bias, scale, shift, B, Cin, Cout, H, W, act, slope);
    else convT2x2_rb_kernel<16><<<cdiv((size_t)B * H * W, 256), 256, (size_t)Cin * 16 * 16, st>>>(x, y, w, bias, scale, shift, B, Cin, Cout, H, W, act, slope);
    WMK_CHECK_LAUNCH("convT2x2_rb_kernel");
    return 0;
  }
  if (smem > 48 * 1024) WMK_CHECK_CUDA(cudaFuncSetAttribute(convT2x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  convT2x2_kernel<<<cdiv((size_t)B * H * W, 256), 256, smem, st>>>(x, y, w, bias, scale, shift, B, Cin, Cout, H, W, act, slope);
  WMK_CHECK_LAUNCH("convT2x2_kernel");
  return 0;
}

extern "C" int wmk_maxpool2x2_f32(const float* x, float* y, int planes, int H, int W, void* stream) {
  WMK_REQUIRE(x && y && planes > 0 && H >= 2 && W >= 2, "maxpool2x2: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 5.0 * planes * H * W, st);
  maxpool2x2_kernel<<<cdiv((size_t)planes * (H / 2) * (W / 2), 256), 256, 0, st>>>(x, y, (size_t)planes, H, W);
  WMK_CHECK_LAUNCH("maxpool2x2_kernel");
  return 0;
}

extern "C" int wmk_noise_mix_f32(const float* noised, const float* cover, float* out, int planes, int H, int W, int h0,
                                 int h1, int w0, int w1, const float* mask_hw, void* stream) {
  WMK_REQUIRE(noised && cover && out && planes > 0 && H > 0 && W > 0, "noise_mix: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_ATTACK, 12.0 * planes * H * W, st);
  mix_kernel<<<cdiv((size_t)planes * H * W, 256), 256, 0, st>>>(noised, cover, out, (size_t)planes, H, W, h0, h1, w0, w1, mask_hw);
  WMK_CHECK_LAUNCH("mix_kernel");
  return 0;
}

extern "C" int wmk_noise_crop_f32(const float* in, float* out, int planes, int H, int W, int h0, int h1, int w0, int w1,
                                  void* stream) {
  WMK_REQUIRE(in && out && planes > 0 && 0 <= h0 && h0 < h1 && h1 <= H && 0 <= w0 && w0 < w1 && w1 <= W, "noise_crop: bad rectangle");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_ATTACK, 8.0 * planes * (h1 - h0) * (w1 - w0), st);
  resample_kernel<<<cdiv((size_t)planes * (h1 - h0) * (w1 - w0), 256), 256, 0, st>>>(in, out, (size_t)planes, H, W, h1 - h0,
                                                                                      w1 - w0, h0, w0, 1.f, 0);
  WMK_CHECK_LAUNCH("resample_kernel<crop>");
  return 0;
}

extern "C" int wmk_noise_resize_nearest_f32(const float* in, float* out, int planes, int H, int W, int Ho, int Wo,
                                            float scale, void* stream) {
  WMK_REQUIRE(in && out && planes > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && scale > 0.f, "noise_resize: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_ATTACK, 8.0 * planes * Ho * Wo, st);
  resample_kernel<<<cdiv((size_t)planes * Ho * Wo, 256), 256, 0, st>>>(in, out, (size_t)planes, H, W, Ho, Wo, 0, 0, 1.0f / scale, 1);
  WMK_CHECK_LAUNCH("resample_kernel<nearest>");
  return 0;
}

static __global__ void minmax_init_kernel(float* mm) {
  if (threadIdx.x < 4) mm[threadIdx.x] = (threadIdx.x & 1) ? -INFINITY : INFINITY;
}

extern "C" int wmk_noise_quantize_f32(const float* in, float* out, size_t n, void* stream) {
  WMK_REQUIRE(in && out && n > 0, "noise_quantize: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_ATTACK, 24.0 * n, st);
  float *mm = nullptr, *tmp = nullptr;
  WMK_CHECK_CUDA(cudaMallocAsync(&mm, 4 * sizeof(float), st));
  WMK_CHECK_CUDA(cudaMallocAsync(&tmp, n * sizeof(float), st));
  minmax_init_kernel<<<1, 32, 0, st>>>(mm);             // {+inf, -inf, +inf, -inf}: no host buffer, no synchronisation
  WMK_CHECK_LAUNCH("minmax_init_kernel");
  const int rb = (int)(cdiv(n, 256) < 1184 ? cdiv(n, 256) : 1184);
  minmax_kernel<<<rb, 256, 0, st>>>(in, n, mm);
  WMK_CHECK_LAUNCH("minmax_kernel");
  quant_round_kernel<<<cdiv(n, 256), 256, 0, st>>>(in, tmp, n, mm);
  WMK_CHECK_LAUNCH("quant_round_kernel");
  minmax_kernel<<<rb, 256, 0, st>>>(tmp, n, mm + 2);
  WMK_CHECK_LAUNCH("minmax_kernel");
  rescale_kernel<<<cdiv(n, 256), 256, 0, st>>>(tmp, out, n, mm + 2, mm);
  WMK_CHECK_LAUNCH("rescale_kernel");
  WMK_CHECK_CUDA(cudaFreeAsync(mm, st));
  WMK_CHECK_CUDA(cudaFreeAsync(tmp, st));
  return 0;
}

extern "C" int wmk_noise_jpeg_f32(const float* in, float* out, int B, int H, int W, int keep_y, int keep_u, int keep_v,
                                  void* stream) {
  WMK_REQUIRE(in && out && B > 0 && B <= 65535 && H > 0 && W > 0 && keep_y >= 0 && keep_y <= 64 && keep_u >= 0 &&
                  keep_u <= 64 && keep_v >= 0 && keep_v <= 64, "noise_jpeg: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  // zig-zag order of jpeg_compression.py:31-32: sort (x, y) by (x + y, -y if (x + y) odd else y)
  int order[64][2], n = 0;
  for (int sum = 0; sum < 15; ++sum)
    for (int k = 0; k < 8; ++k) {
      const int y = (sum & 1) ? 7 - k : k;          // odd sums: descending y; even sums: ascending y
      const int x = sum - y;
      if (x < 0 || x > 7) continue;
      order[n][0] = x; order[n][1] = y; ++n;
    }
  JpegKeep keep = {{0, 0, 0}};
  const int cnt[3] = {keep_y, keep_u, keep_v};
  for (int c = 0; c < 3; ++c)
    for (int i = 0; i < cnt[c]; ++i) keep.m[c] |= 1ull << (order[i][0] * 8 + order[i][1]);
  ProfScope prof(FAM_ATTACK, 24.0 * B * H * W, st);
  dim3 grid(cdiv(W, 8), cdiv(H, 8), B);
  jpeg_kernel<<<grid, 192, 0, st>>>(in, out, H, W, keep);
  WMK_CHECK_LAUNCH("jpeg_kernel");
  return 0;
}

extern "C" int wmk_magphase_split_f32(const float* spec, float* mag, float* phase, size_t n, size_t plane, void* stream) {
  WMK_REQUIRE(spec && mag && n > 0 && plane > 0, "magphase_split: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_ATTACK, (phase ? 16.0 : 12.0) * n * plane, st);
  magphase_split_kernel<<<cdiv(n * plane, 256), 256, 0, st>>>(spec, mag, phase, n, plane);
  WMK_CHECK_LAUNCH("magphase_split_kernel");
  return 0;
}

extern "C" int wmk_magphase_merge_f32(const float* mag, const float* phase, float* spec, size_t n, size_t plane, void* stream) {
  WMK_REQUIRE(spec && mag && phase && n > 0 && plane > 0, "magphase_merge: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_ATTACK, 16.0 * n * plane, st);
  magphase_merge_kernel<<<cdiv(n * plane, 256), 256, 0, st>>>(mag, phase, spec, n, plane);
  WMK_CHECK_LAUNCH("magphase_merge_kernel");
  return 0;
}

extern "C" int wmk_conv3x3_c1_nhwc_bf16(const float* x, void* y, const float* w, const float* bias, const float* scale,
                                        const float* shift, int B, int H, int W, void* stream) {
  WMK_REQUIRE(x && y && w && B > 0 && H > 0 && W > 0 && ((uintptr_t)y & 15) == 0, "conv3x3_c1_nhwc: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)B * H * W;
  ProfScope prof(FAM_SMALL, npix * (4.0 + 128.0), st);
  conv3x3_c1_nhwc_kernel<<<cdiv(npix * 8, 256), 256, 0, st>>>(x, reinterpret_cast<__nv_bfloat16*>(y), w, bias, scale, shift, npix, H, W);
  WMK_CHECK_LAUNCH("conv3x3_c1_nhwc_kernel");
  return 0;
}

extern "C" int wmk_conv3x3_nhwc_bf16_tc(const void* x, void* y, const void* w_packed, const float* bias, int B, int H, int W,
                                        int Cout, void* stream) {
  WMK_REQUIRE(x && y && w_packed && bias && B > 0 && H > 0 && W == 128 && (Cout == 32 || Cout == 64),
              "conv3x3_nhwc_bf16_tc: needs W = 128, 64 input channels and Cout in {32, 64}");
  GemmArgs g;
  g.A = x; g.W = w_packed; g.bias = bias; g.C = y; g.M = B * H * 128; g.N = Cout; g.K = 576; g.ldc = Cout;
  g.epi = EPI_BIAS_RELU; g.out_bf16 = 1; g.conv_H = H; g.conv_B = B;
  return gemm_bf16_tcgen05(g, (cudaStream_t)stream);
}

extern "C" int wmk_maxpool2x2_nhwc_bf16(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  WMK_REQUIRE(x && y && B > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "maxpool2x2_nhwc: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
  ProfScope prof(FAM_SMALL, 2.5 * B * H * W * C, st);
  maxpool2x2_nhwc_kernel<<<cdiv(total, 256), 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<__nv_bfloat16*>(y),
                                                           total, H, W, C);
  WMK_CHECK_LAUNCH("maxpool2x2_nhwc_kernel");
  return 0;
}

extern "C" int wmk_conv3x3_nhwc_to1_f32(const void* x, float* y, const float* wt, float bias, float scale, float shift, int B,
                                        int H, int W, int Cp, void* stream) {
  WMK_REQUIRE(x && y && wt && B > 0 && H > 0 && W > 0 && Cp % 8 == 0 && Cp <= 256, "conv3x3_nhwc_to1: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t npix = (size_t)B * H * W;
  ProfScope prof(FAM_SMALL, npix * (2.0 * Cp + 4.0), st);
  conv3x3_nhwc_to1_kernel<<<cdiv(npix * 4, 256), 256, 9 * Cp * sizeof(float), st>>>(reinterpret_cast<const __nv_bfloat16*>(x), y, wt, bias,
                                                                                   scale, shift, npix, H, W, Cp);
  WMK_CHECK_LAUNCH("conv3x3_nhwc_to1_kernel");
  return 0;
}

// Training-mode kernels of ModelA (uformerWM/model.py:3000-3066; step uformerWM/train_modelA.py:402-500):
// BatchNorm with batch statistics (+ fused activation) forward / backward, max-pool backward,
// weight / bias gradients of the 3x3 convolution and of the 2x2 stride-2 transposed convolution,
// its data gradient, the mean-squared-error loss with its gradient, and a fused Adam step.
// (The data gradient of the 3x3 convolution is the forward kernel of conv_noise.cu run with the
// flipped, transposed weights.)  NCHW float32; all of these are memory bound: every tensor is read
// once per kernel, per-channel reductions go through warp shuffles and fp64 atomics.
#include <stdlib.h>

#include "uformer_kernels.cuh"

namespace wmk {
namespace {

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SIGMOID = 3 };

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : slope * v;
  if (act == ACT_SIGMOID) return 1.0f / (1.0f + expf(-v));
  return v;
}
// derivative of the activation expressed through its OUTPUT y (sign(y) = sign(z) for (leaky) ReLU)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float slope) {
  if (act == ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (act == ACT_LEAKY) return y > 0.f ? 1.f : slope;
  if (act == ACT_SIGMOID) return y * (1.f - y);
  return 1.f;
}

// CTA-wide sum of two doubles, then ONE pair of atomics per CTA: the per-channel totals live at two addresses, and one
// atomic pair per warp (thousands per address) serialised in L2 - the statistics kernels ran at 0.8 / 1.6 TB/s
__device__ __forceinline__ void block_sum2_atomic(double d1, double d2, double* dst) {
  __shared__ double red[2][8];
  d1 = warp_sum(d1); d2 = warp_sum(d2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = d1; red[1][warp] = d2; }
  __syncthreads();
  if (warp == 0) {
    d1 = lane < 8 ? red[0][lane] : 0.0;
    d2 = lane < 8 ? red[1][lane] : 0.0;
    d1 = warp_sum(d1); d2 = warp_sum(d2);
    if (lane == 0) {
      atomicAdd(dst, d1);
      atomicAdd(dst + 1, d2);
    }
  }
}

// ---- BatchNorm2d (training): per-channel sums over (B, H, W)
// grid (chunks, C): stats[c] = {sum x, sum x^2}  (fp64 atomics); 16-byte loads (HW is a multiple of 4)
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ x, int B, int C, int HW, double* __restrict__ stats) {
  const int c = blockIdx.y;
  const int hw4 = HW >> 2;
  const size_t n4 = (size_t)B * hw4;
  float s1 = 0.f, s2 = 0.f;      // per-thread partials stay short; fp64 beyond
  double d1 = 0.0, d2 = 0.0;
  int cnt = 0;
  // (b, r) advance incrementally: a 64-bit division per 16-byte load was most of this kernel's instructions
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned step = gridDim.x * blockDim.x, step_b = step / (unsigned)hw4, step_r = step - step_b * (unsigned)hw4;
  size_t b = i0 / hw4;
  unsigned r = (unsigned)(i0 - b * hw4);
  for (size_t i = i0; i < n4; i += step, b += step_b, r += step_r) {
    if (r >= (unsigned)hw4) { r -= (unsigned)hw4; ++b; }
    const float4 v = reinterpret_cast<const float4*>(x + (b * C + c) * HW)[r];
    s1 += (v.x + v.y) + (v.z + v.w);
    s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
    if (++cnt == 256) { d1 += s1; d2 += s2; s1 = s2 = 0.f; cnt = 0; }
  }
  d1 += s1; d2 += s2;
  block_sum2_atomic(d1, d2, stats + 2 * c);
}

// mean / rstd from the sums; running statistics updated as nn.BatchNorm2d does (momentum, unbiased var)
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, double n, float eps, float momentum,
                                   float* __restrict__ mean_rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = stats[2 * c] / n;
  double var = stats[2 * c + 1] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[2 * c] = (float)mean;
  mean_rstd[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// y = act(gamma * (x - mean) * rstd + beta); one thread = 4 consecutive elements of one plane
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ mean_rstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, size_t total4, int C, int HW, int act,
                  float slope) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c = (int)((i / (HW >> 2)) % C);
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1], g = gamma[c], be = beta[c];
  float4 v = reinterpret_cast<const float4*>(x)[i];
  v.x = act_fwd(fmaf(g, (v.x - mean) * rstd, be), act, slope);
  v.y = act_fwd(fmaf(g, (v.y - mean) * rstd, be), act, slope);
  v.z = act_fwd(fmaf(g, (v.z - mean) * rstd, be), act, slope);
  v.w = act_fwd(fmaf(g, (v.w - mean) * rstd, be), act, slope);
  reinterpret_cast<float4*>(y)[i] = v;
}

// backward pass 1: dz = dy * act'(y); sums[c] = {sum dz, sum dz * xhat}
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                         const float* __restrict__ mean_rstd, int B, int C, int HW, int act, float slope,
                         double* __restrict__ sums) {
  const int c = blockIdx.y;
  const int hw4 = HW >> 2;
  const size_t n4 = (size_t)B * hw4;
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1];
  float s1 = 0.f, s2 = 0.f;
  double d1 = 0.0, d2 = 0.0;
  int cnt = 0;
  const size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned step = gridDim.x * blockDim.x, step_b = step / (unsigned)hw4, step_r = step - step_b * (unsigned)hw4;
  size_t b = i0 / hw4;
  unsigned r = (unsigned)(i0 - b * hw4);
  for (size_t i = i0; i < n4; i += step, b += step_b, r += step_r) {
    if (r >= (unsigned)hw4) { r -= (unsigned)hw4; ++b; }
    const size_t o = ((b * C + c) * HW >> 2) + r;
    const float4 xv = reinterpret_cast<const float4*>(x)[o], yv = reinterpret_cast<const float4*>(y)[o],
                 gv = reinterpret_cast<const float4*>(dy)[o];
    const float z0 = gv.x * act_grad_from_out(yv.x, act, slope), z1 = gv.y * act_grad_from_out(yv.y, act, slope);
    const float z2 = gv.z * act_grad_from_out(yv.z, act, slope), z3 = gv.w * act_grad_from_out(yv.w, act, slope);
    s1 += (z0 + z1) + (z2 + z3);
    s2 = fmaf(z0, (xv.x - mean) * rstd, fmaf(z1, (xv.y - mean) * rstd, fmaf(z2, (xv.z - mean) * rstd, fmaf(z3, (xv.w - mean) * rstd, s2))));
    if (++cnt == 256) { d1 += s1; d2 += s2; s1 = s2 = 0.f; cnt = 0; }
  }
  d1 += s1; d2 += s2;
  block_sum2_atomic(d1, d2, sums + 2 * c);
}

// backward pass 2: dx = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat)); the first C threads also write dgamma / dbeta
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                        float* __restrict__ dx, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                        const double* __restrict__ sums, size_t total4, int C, int HW, double n, int act, float slope,
                        float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)C) {
    dgamma[i] = (float)sums[2 * i + 1];
    dbeta[i] = (float)sums[2 * i];
  }
  if (i >= total4) return;
  const int c = (int)((i / (HW >> 2)) % C);
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1];
  const float m1 = (float)(sums[2 * c] / n), m2 = (float)(sums[2 * c + 1] / n), gr = gamma[c] * rstd;
  const float4 xv = reinterpret_cast<const float4*>(x)[i], yv = reinterpret_cast<const float4*>(y)[i],
               gv = reinterpret_cast<const float4*>(dy)[i];
  float4 o;
  o.x = gr * (gv.x * act_grad_from_out(yv.x, act, slope) - m1 - (xv.x - mean) * rstd * m2);
  o.y = gr * (gv.y * act_grad_from_out(yv.y, act, slope) - m1 - (xv.y - mean) * rstd * m2);
  o.z = gr * (gv.z * act_grad_from_out(yv.z, act, slope) - m1 - (xv.z - mean) * rstd * m2);
  o.w = gr * (gv.w * act_grad_from_out(yv.w, act, slope) - m1 - (xv.w - mean) * rstd * m2);
  reinterpret_cast<float4*>(dx)[i] = o;
}


// ---- BatchNorm2d (training) + activation + MaxPool2d(2,2) fused (ModelA's `Conv - BN - LeakyReLU - MaxPool` groups,
// uformerWM/model.py:3005-3013,3028-3037).  Forward: one pass writes the activation y (kept for the backward pass) AND
// its 2x2 maxima yp, so the pooling kernel's re-read of y disappears.  Backward: the gradient arrives at pooled
// resolution; the full-resolution gradient of the pooling layer (dyp routed to the first maximum of each window, in
// PyTorch's scan order) is never materialised - both passes recompute the argmax from y.
// A thread owns a 2-row x 4-column patch (two 16-byte loads per tensor) = two pooling windows.
struct Patch { float4 a, b; };       // rows 2r and 2r + 1
__device__ __forceinline__ int first_max4(float v0, float v1, float v2, float v3) {
  int k = 0;
  float m = v0;
  if (v1 > m) { m = v1; k = 1; }
  if (v2 > m) { m = v2; k = 2; }
  if (v3 > m) { k = 3; }
  return k;
}

__global__ void __launch_bounds__(256)
bn_act_pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ yp,
                       const float* __restrict__ mean_rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                       size_t units, int C, int H, int W, int act, float slope) {
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= units) return;
  const int w4 = W >> 2, h2 = H >> 1;
  const int c4 = (int)(u % w4), r2 = (int)((u / w4) % h2);
  const size_t plane = u / ((size_t)w4 * h2);
  const int c = (int)(plane % C);
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1], g = gamma[c], be = beta[c];
  const size_t o = (plane * H + 2 * r2) * W + 4 * c4;
  float4 a = *reinterpret_cast<const float4*>(x + o), b = *reinterpret_cast<const float4*>(x + o + W);
  a.x = act_fwd(fmaf(g, (a.x - mean) * rstd, be), act, slope); a.y = act_fwd(fmaf(g, (a.y - mean) * rstd, be), act, slope);
  a.z = act_fwd(fmaf(g, (a.z - mean) * rstd, be), act, slope); a.w = act_fwd(fmaf(g, (a.w - mean) * rstd, be), act, slope);
  b.x = act_fwd(fmaf(g, (b.x - mean) * rstd, be), act, slope); b.y = act_fwd(fmaf(g, (b.y - mean) * rstd, be), act, slope);
  b.z = act_fwd(fmaf(g, (b.z - mean) * rstd, be), act, slope); b.w = act_fwd(fmaf(g, (b.w - mean) * rstd, be), act, slope);
  *reinterpret_cast<float4*>(y + o) = a;
  *reinterpret_cast<float4*>(y + o + W) = b;
  float2 m;
  m.x = fmaxf(fmaxf(a.x, a.y), fmaxf(b.x, b.y));
  m.y = fmaxf(fmaxf(a.z, a.w), fmaxf(b.z, b.w));
  *reinterpret_cast<float2*>(yp + (plane * h2 + r2) * (size_t)(W >> 1) + 2 * c4) = m;
}

// gradient dz = dy * act'(y) of one patch from the pooled gradient (non-zero at the two window maxima only)
__device__ __forceinline__ void pooled_dz(const Patch& yv, float2 gp, int act, float slope, float dz[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) dz[i] = 0.f;
  const float y0[4] = {yv.a.x, yv.a.y, yv.b.x, yv.b.y}, y1[4] = {yv.a.z, yv.a.w, yv.b.z, yv.b.w};
  const int k0 = first_max4(y0[0], y0[1], y0[2], y0[3]), k1 = first_max4(y1[0], y1[1], y1[2], y1[3]);
  // patch order: 0..3 = row a (x, y, z, w), 4..7 = row b
  const int p0 = (k0 & 1) + ((k0 >> 1) << 2), p1 = 2 + (k1 & 1) + ((k1 >> 1) << 2);
  const float g0 = gp.x * act_grad_from_out(y0[k0], act, slope), g1 = gp.y * act_grad_from_out(y1[k1], act, slope);
#pragma unroll
  for (int i = 0; i < 8; ++i) dz[i] = i == p0 ? g0 : (i == p1 ? g1 : 0.f);
}

__global__ void __launch_bounds__(256)
bn_act_pool_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dyp,
                              const float* __restrict__ mean_rstd, int B, int C, int H, int W, int act, float slope,
                              double* __restrict__ sums) {
  const int c = blockIdx.y;
  const int w4 = W >> 2, h2 = H >> 1;
  const size_t per_img = (size_t)w4 * h2, n = (size_t)B * per_img;
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1];
  float s1 = 0.f, s2 = 0.f;
  double d1 = 0.0, d2 = 0.0;
  int cnt = 0;
  for (size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x; u < n; u += (size_t)gridDim.x * blockDim.x) {
    const size_t b = u / per_img;
    const int rem = (int)(u - b * per_img), r2 = rem / w4, c4 = rem - r2 * w4;
    const size_t plane = b * C + c;
    const size_t o = (plane * H + 2 * r2) * W + 4 * c4;
    Patch xv, yv;
    xv.a = *reinterpret_cast<const float4*>(x + o); xv.b = *reinterpret_cast<const float4*>(x + o + W);
    yv.a = *reinterpret_cast<const float4*>(y + o); yv.b = *reinterpret_cast<const float4*>(y + o + W);
    const float2 gp = *reinterpret_cast<const float2*>(dyp + (plane * h2 + r2) * (size_t)(W >> 1) + 2 * c4);
    float dz[8];
    pooled_dz(yv, gp, act, slope, dz);
    const float xs[8] = {xv.a.x, xv.a.y, xv.a.z, xv.a.w, xv.b.x, xv.b.y, xv.b.z, xv.b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s1 += dz[i];
      s2 = fmaf(dz[i], (xs[i] - mean) * rstd, s2);
    }
    if (++cnt == 128) { d1 += s1; d2 += s2; s1 = s2 = 0.f; cnt = 0; }
  }
  d1 += s1; d2 += s2;
  block_sum2_atomic(d1, d2, sums + 2 * c);
}

__global__ void __launch_bounds__(256)
bn_act_pool_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dyp,
                             float* __restrict__ dx, const float* __restrict__ mean_rstd, const float* __restrict__ gamma,
                             const double* __restrict__ sums, size_t units, int C, int H, int W, double n, int act, float slope,
                             float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < (size_t)C) {
    dgamma[u] = (float)sums[2 * u + 1];
    dbeta[u] = (float)sums[2 * u];
  }
  if (u >= units) return;
  const int w4 = W >> 2, h2 = H >> 1;
  const int c4 = (int)(u % w4), r2 = (int)((u / w4) % h2);
  const size_t plane = u / ((size_t)w4 * h2);
  const int c = (int)(plane % C);
  const float mean = mean_rstd[2 * c], rstd = mean_rstd[2 * c + 1];
  const float m1 = (float)(sums[2 * c] / n), m2 = (float)(sums[2 * c + 1] / n), gr = gamma[c] * rstd;
  const size_t o = (plane * H + 2 * r2) * W + 4 * c4;
  Patch xv, yv;
  xv.a = *reinterpret_cast<const float4*>(x + o); xv.b = *reinterpret_cast<const float4*>(x + o + W);
  yv.a = *reinterpret_cast<const float4*>(y + o); yv.b = *reinterpret_cast<const float4*>(y + o + W);
  const float2 gp = *reinterpret_cast<const float2*>(dyp + (plane * h2 + r2) * (size_t)(W >> 1) + 2 * c4);
  float dz[8];
  pooled_dz(yv, gp, act, slope, dz);
  const float xs[8] = {xv.a.x, xv.a.y, xv.a.z, xv.a.w, xv.b.x, xv.b.y, xv.b.z, xv.b.w};
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = gr * (dz[i] - m1 - (xs[i] - mean) * rstd * m2);
  *reinterpret_cast<float4*>(dx + o) = make_float4(r[0], r[1], r[2], r[3]);
  *reinterpret_cast<float4*>(dx + o + W) = make_float4(r[4], r[5], r[6], r[7]);
}

// MaxPool2d(2,2) backward: the gradient goes to the first maximum of each window (PyTorch's scan order)
__global__ void __launch_bounds__(256)
maxpool2x2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, size_t planes,
                      int H, int W) {
  const int Ho = H >> 1, Wo = W >> 1;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= planes * Ho * Wo) return;
  const int wq = (int)(idx % Wo), h = (int)((idx / Wo) % Ho);
  const size_t p = idx / ((size_t)Ho * Wo);
  const size_t o = (p * H + 2 * h) * W + 2 * wq;
  const float v[4] = {x[o], x[o + 1], x[o + W], x[o + W + 1]};
  int k = 0;
  if (v[1] > v[k]) k = 1;
  if (v[2] > v[k]) k = 2;
  if (v[3] > v[k]) k = 3;
  const float g = dy[idx];
  dx[o] = k == 0 ? g : 0.f;
  dx[o + 1] = k == 1 ? g : 0.f;
  dx[o + W] = k == 2 ? g : 0.f;
  dx[o + W + 1] = k == 3 ? g : 0.f;
}

// out = in * mask * scale   (Dropout forward and backward with the same mask)
__global__ void __launch_bounds__(256)
mask_scale_kernel(const float* __restrict__ in, const float* __restrict__ mask, float* __restrict__ out, size_t n, float scale) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * mask[i] * scale;
}

// ---- weight gradient of Conv2d(Cin, Cout, 3, padding=1):
//   dw[co][ci][ky][kx] = sum_{b,h,w} x[b][ci][h+ky-1][w+kx-1] dy[b][co][h][w],   db[co] = sum dy[b][co][h][w]
// Persistent CTAs walk 16x16 pixel tiles (all images): x tile (+halo) and dy tile in shared memory; a thread
// owns (pixel subset s, ci, block of CB output channels) = CB x 9 accumulators that stay in registers
// across ALL of the CTA's tiles, so the global gradient sees one atomic per accumulator per CTA
// (a few hundred CTAs) instead of one per tile (tens of thousands: measured 13.5 ms of atomic contention).
constexpr int WG_T = 16;
template <int CB>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                     float* __restrict__ db, int Cin, int Cout, int H, int W, int n_tiles) {
  extern __shared__ float sm[];
  float* xs = sm;                                        // [Cin][18][18]
  float* ds = sm + Cin * (WG_T + 2) * (WG_T + 2);        // [Cout][16][16]
  const int tiles_w = W / WG_T, tiles_img = tiles_w * (H / WG_T);
  const int n_cb = (Cout + CB - 1) / CB;
  const int pairs = n_cb * Cin;
  const int S = 256 / pairs > 0 ? 256 / pairs : 1;       // pixel subsets
  const int pair = threadIdx.x % pairs, s = threadIdx.x / pairs;
  const bool worker = s < S && threadIdx.x < pairs * S;
  const int ci = pair % Cin, cb = pair / Cin;
  float acc[CB][9], bsum[CB];
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    bsum[j] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_img, trem = tile - b * tiles_img;
    const int th = trem / tiles_w, tw = trem - th * tiles_w;
    __syncthreads();                                     // the previous tile has been consumed
    for (int e = threadIdx.x; e < Cin * (WG_T + 2) * (WG_T + 2); e += 256) {
      const int cc = e / ((WG_T + 2) * (WG_T + 2)), r = (e / (WG_T + 2)) % (WG_T + 2), c = e % (WG_T + 2);
      const int hh = th * WG_T + r - 1, ww = tw * WG_T + c - 1;
      xs[e] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(((size_t)b * Cin + cc) * H + hh) * W + ww] : 0.f;
    }
    for (int e = threadIdx.x; e < Cout * WG_T * WG_T; e += 256) {
      const int co = e / (WG_T * WG_T), r = (e / WG_T) % WG_T, c = e % WG_T;
      ds[e] = dy[(((size_t)b * Cout + co) * H + th * WG_T + r) * W + tw * WG_T + c];
    }
    __syncthreads();
    if (!worker) continue;
    for (int p = s; p < WG_T * WG_T; p += S) {
      const int r = p / WG_T, c = p % WG_T;
      float xv[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) xv[t] = xs[(ci * (WG_T + 2) + r + t / 3) * (WG_T + 2) + c + t % 3];
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        const int co = cb * CB + j;
        const float d = co < Cout ? ds[co * WG_T * WG_T + p] : 0.f;
        bsum[j] += d;
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[j][t] = fmaf(xv[t], d, acc[j][t]);
      }
    }
  }
  if (!worker) return;
#pragma unroll
  for (int j = 0; j < CB; ++j) {
    const int co = cb * CB + j;
    if (co >= Cout) break;
#pragma unroll
    for (int t = 0; t < 9; ++t) atomicAdd(dw + ((size_t)co * Cin + ci) * 9 + t, acc[j][t]);
    if (ci == 0 && db) atomicAdd(db + co, bsum[j]);
  }
}


// ---- register-blocked weight gradient (the product path for Cin*Cout <= 2048; the kernel above stays for wider layers)
// A thread owns CO_T x CI_T (output, input) channel pairs = CO_T * CI_T * 9 accumulators that live in registers across
// ALL tiles of its CTA, and walks 4-pixel row segments of the 16x16 tile: per segment and kernel row one x window of 6
// values per input channel (one scalar + one 16-byte + one scalar shared-memory load) feeds 12 * CO_T FMAs, so the
// kernel is bound by the FMA pipe and not by shared-memory bandwidth (the thread-per-pair kernel above: 13 loads per
// 36 FMAs).  Lanes of a warp hold DIFFERENT channel pairs of the SAME segment (shared-memory broadcasts; channels are
// dealt round-robin so that the distinct addresses of a 16-byte access fall into different banks: plane pitches 456 / 260
// words).  Per-CTA partial sums go to part[cta][...]: the fixed-order reduction below makes the result bit-reproducible.
constexpr int WG2_NT = 256;
constexpr int WG2_XP = 24;                     // x row: [3 pad][halo][16 pixels][halo][3 pad]
constexpr int WG2_XPL = 18 * WG2_XP + 24;      // x plane pitch (= 8 mod 32 words)
constexpr int WG2_DPL = 260;                   // dy plane pitch
template <int CO_T, int CI_T>
__global__ void __launch_bounds__(WG2_NT, 2)
conv3x3_wgrad2_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int Cin, int Cout,
                      int H, int W, int n_tiles, int n_cb, int n_cib, int want_db) {
  extern __shared__ __align__(16) float sm[];
  const int cin_pad = n_cib * CI_T, cout_pad = n_cb * CO_T;
  float* xs = sm;                                // [cin_pad][WG2_XPL]
  float* ds = sm + (size_t)cin_pad * WG2_XPL;    // [cout_pad][WG2_DPL]
  const int tiles_w = W / 16, tiles_img = tiles_w * (H / 16);
  const int pairs = n_cb * n_cib;
  const int S = WG2_NT / pairs;                  // pixel subsets (host guarantees pairs <= WG2_NT)
  const int pair = threadIdx.x % pairs, s = threadIdx.x / pairs;
  const bool worker = s < S;
  const int cib = pair % n_cib, cb = pair / n_cib;
  float acc[CO_T][CI_T][9], bsum[CO_T];
#pragma unroll
  for (int j = 0; j < CO_T; ++j) {
    bsum[j] = 0.f;
#pragma unroll
    for (int i = 0; i < CI_T; ++i)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[j][i][t] = 0.f;
  }
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_img, trem = tile - b * tiles_img;
    const int th = trem / tiles_w, tw = trem - th * tiles_w;
    __syncthreads();                             // the previous tile has been consumed
#pragma unroll 4
    for (int e = threadIdx.x; e < cin_pad * 108; e += WG2_NT) {          // 18 rows x (halo, 4 x 16 bytes, halo)
      const int ci = e / 108, rem = e - ci * 108, r = rem / 6, k = rem - r * 6;
      const int hh = th * 16 + r - 1;
      const bool ok = ci < Cin && hh >= 0 && hh < H;
      const float* src = x + (((size_t)b * Cin + (ok ? ci : 0)) * H + (ok ? hh : 0)) * W + tw * 16;
      float* dst = xs + ci * WG2_XPL + r * WG2_XP;
      if (k == 0) dst[3] = (ok && tw > 0) ? src[-1] : 0.f;
      else if (k == 5) dst[20] = (ok && tw + 1 < tiles_w) ? src[16] : 0.f;
      else *reinterpret_cast<float4*>(dst + 4 * k) = ok ? *reinterpret_cast<const float4*>(src + 4 * (k - 1)) : zero4;
    }
#pragma unroll 4
    for (int e = threadIdx.x; e < cout_pad * 64; e += WG2_NT) {
      const int co = e >> 6, r = (e >> 2) & 15, k = e & 3;
      *reinterpret_cast<float4*>(ds + co * WG2_DPL + r * 16 + 4 * k) =
          co < Cout ? *reinterpret_cast<const float4*>(dy + (((size_t)b * Cout + co) * H + th * 16 + r) * W + tw * 16 + 4 * k) : zero4;
    }
    __syncthreads();
    if (!worker) continue;
    for (int seg = s; seg < 64; seg += S) {
      const int r = seg >> 2, c4 = (seg & 3) * 4;
      float d[CO_T][4];
#pragma unroll
      for (int j = 0; j < CO_T; ++j) {
        const float4 v = *reinterpret_cast<const float4*>(ds + (cb + j * n_cb) * WG2_DPL + r * 16 + c4);
        d[j][0] = v.x; d[j][1] = v.y; d[j][2] = v.z; d[j][3] = v.w;
      }
      if (want_db && cib == 0) {
#pragma unroll
        for (int j = 0; j < CO_T; ++j) bsum[j] += (d[j][0] + d[j][1]) + (d[j][2] + d[j][3]);
      }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int i = 0; i < CI_T; ++i) {
          const float* xr = xs + (cib + i * n_cib) * WG2_XPL + (r + ky) * WG2_XP + c4 + 3;      // xr[0] = column c4 - 1
          const float4 m = *reinterpret_cast<const float4*>(xr + 1);
          const float xv[6] = {xr[0], m.x, m.y, m.z, m.w, xr[5]};
#pragma unroll
          for (int kx = 0; kx < 3; ++kx)
#pragma unroll
            for (int j = 0; j < CO_T; ++j)
#pragma unroll
              for (int px = 0; px < 4; ++px) acc[j][i][ky * 3 + kx] = fmaf(xv[px + kx], d[j][px], acc[j][i][ky * 3 + kx]);
        }
    }
  }
  // sum over the pixel subsets - lanes of a warp that hold the same pair by shuffles (pairs | 32), the rest through shared
  // memory (the tile buffers are free now) - then one partial row per CTA
  __syncthreads();
  constexpr int NACC = CO_T * CI_T * 9;
  const bool shfl = pairs < 32 && (32 % pairs) == 0;
  if (shfl) {
    for (int off = pairs; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < CO_T; ++j) {
        bsum[j] += __shfl_xor_sync(0xffffffffu, bsum[j], off);
#pragma unroll
        for (int i = 0; i < CI_T; ++i)
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[j][i][t] += __shfl_xor_sync(0xffffffffu, acc[j][i][t], off);
      }
    }
  }
  const int n_sub = shfl ? WG2_NT / 32 : S;       // partial sets left in shared memory
  const int my_sub = shfl ? (int)(threadIdx.x >> 5) : s;
  const bool writer = shfl ? (int)(threadIdx.x & 31) < pairs : worker;
  float* red = sm;                               // [n_sub][pairs][NACC + CO_T]
  if (writer) {
    float* mine = red + ((size_t)my_sub * pairs + pair) * (NACC + CO_T);
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
#pragma unroll
      for (int i = 0; i < CI_T; ++i)
#pragma unroll
        for (int t = 0; t < 9; ++t) mine[(j * CI_T + i) * 9 + t] = acc[j][i][t];
      mine[NACC + j] = bsum[j];
    }
  }
  __syncthreads();
  const int n_dw = Cout * Cin * 9;
  float* my_part = part + (size_t)blockIdx.x * (n_dw + Cout);
  for (int e = threadIdx.x; e < pairs * (NACC + CO_T); e += WG2_NT) {
    const int pr = e / (NACC + CO_T), k = e - pr * (NACC + CO_T);
    float v = 0.f;
    for (int q = 0; q < n_sub; ++q) v += red[((size_t)q * pairs + pr) * (NACC + CO_T) + k];
    const int pcib = pr % n_cib, pcb = pr / n_cib;
    if (k < NACC) {
      const int j = k / (CI_T * 9), i = (k / 9) % CI_T, t = k % 9;
      const int co = pcb + j * n_cb, ci = pcib + i * n_cib;
      if (co < Cout && ci < Cin) my_part[((size_t)co * Cin + ci) * 9 + t] = v;
    } else if (pcib == 0) {
      const int co = pcb + (k - NACC) * n_cb;
      if (co < Cout) my_part[n_dw + co] = v;
    }
  }
}

// out[o] = sum over the CTAs' partial rows in a fixed order: one warp per output, lane l adds rows l, l + 32, ..., then a
// butterfly (a: the first n_a outputs, b: the rest)
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ part, int n_parts, int n, float* __restrict__ a, int n_a, float* __restrict__ b) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= n) return;
  float v = 0.f;
  for (int p = lane; p < n_parts; p += 32) v += part[(size_t)p * n + o];
  v = warp_sum(v);
  if (lane == 0) {
    if (o < n_a) a[o] = v;
    else if (b) b[o - n_a] = v;
  }
}

// ---- ConvTranspose2d(Cin, Cout, 2, stride=2): data gradient
//   dx[b][ci][h][w] = sum_{co,i,j} w[ci][co][i][j] dy[b][co][2h+i][2w+j]
__global__ void __launch_bounds__(256)
convT2x2_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int Cin,
                      int Cout, int H, int W) {
  extern __shared__ float wsm[];             // [Cin][Cout][4]
  for (int e = threadIdx.x; e < Cin * Cout * 4; e += blockDim.x) wsm[e] = w[e];
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * H * W) return;
  const int wq = (int)(idx % W), h = (int)((idx / W) % H);
  const size_t b = idx / ((size_t)H * W);
  for (int ci = 0; ci < Cin; ++ci) {
    float a = 0.f;
    for (int co = 0; co < Cout; ++co) {
      const float* d = dy + ((b * Cout + co) * (size_t)(2 * H) + 2 * h) * (2 * W) + 2 * wq;
      const float2 d0 = *reinterpret_cast<const float2*>(d), d1 = *reinterpret_cast<const float2*>(d + 2 * W);
      const float* wp = wsm + ((size_t)ci * Cout + co) * 4;
      a = fmaf(wp[0], d0.x, fmaf(wp[1], d0.y, fmaf(wp[2], d1.x, fmaf(wp[3], d1.y, a))));
    }
    dx[((b * Cin + ci) * H + h) * W + wq] = a;
  }
}

// register-blocked form for Cout <= COB: the 4 COB gradient values of a pixel are read once (8-byte loads) and kept in
// registers for all input channels (the kernel above re-reads them for every ci)
template <int COB>
__global__ void __launch_bounds__(256)
convT2x2_dgrad_rb_kernel(const float* __restrict__ dy, const float* __restrict__ w, float* __restrict__ dx, int B, int Cin,
                         int Cout, int H, int W) {
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
  float d[COB][4];
#pragma unroll
  for (int co = 0; co < COB; ++co) {
    if (co < Cout) {
      const float* p = dy + ((b * Cout + co) * (size_t)(2 * H) + 2 * h) * (2 * W) + 2 * wq;
      const float2 d0 = *reinterpret_cast<const float2*>(p), d1 = *reinterpret_cast<const float2*>(p + 2 * W);
      d[co][0] = d0.x; d[co][1] = d0.y; d[co][2] = d1.x; d[co][3] = d1.y;
    } else {
      d[co][0] = d[co][1] = d[co][2] = d[co][3] = 0.f;
    }
  }
  float* xp = dx + (b * Cin * H + h) * W + wq;
  for (int ci = 0; ci < Cin; ++ci) {
    const float4* w4 = reinterpret_cast<const float4*>(wsm + ci * COB * 4);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int co = 0; co < COB; ++co) {
      const float4 ww = w4[co];
      a0 = fmaf(ww.x, d[co][0], fmaf(ww.y, d[co][1], a0));
      a1 = fmaf(ww.z, d[co][2], fmaf(ww.w, d[co][3], a1));
    }
    xp[(size_t)ci * H * W] = a0 + a1;
  }
}

// weight gradient: dw[ci][co][i][j] = sum_{b,h,w} x[b][ci][h][w] dy[b][co][2h+i][2w+j];  db[co] = sum dy
// Persistent CTAs walk 16x16 input tiles; a thread owns up to 3 (ci, co) pairs = 4 accumulators each, kept in
// registers across all of the CTA's tiles (one atomic per accumulator per CTA).
constexpr int TW_PAIRS = 3;        // pairs per thread: Cin * Cout <= 768
__global__ void __launch_bounds__(256)
convT2x2_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                      float* __restrict__ db, int Cin, int Cout, int H, int W, int n_tiles) {
  extern __shared__ float sm[];
  float* xs = sm;                              // [Cin][256]
  float* ds = sm + Cin * 256;                  // [Cout][32][32]
  const int tiles_w = W / 16, tiles_img = tiles_w * (H / 16);
  const int pairs = Cin * Cout;
  float a[TW_PAIRS][4], bs[TW_PAIRS];
#pragma unroll
  for (int q = 0; q < TW_PAIRS; ++q) { a[q][0] = a[q][1] = a[q][2] = a[q][3] = 0.f; bs[q] = 0.f; }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_img, trem = tile - b * tiles_img;
    const int th = trem / tiles_w, tw = trem - th * tiles_w;
    __syncthreads();
    for (int e = threadIdx.x; e < Cin * 256; e += 256) {
      const int ci = e >> 8, r = (e >> 4) & 15, c = e & 15;
      xs[e] = x[(((size_t)b * Cin + ci) * H + th * 16 + r) * W + tw * 16 + c];
    }
    for (int e = threadIdx.x; e < Cout * 1024; e += 256) {
      const int co = e >> 10, r = (e >> 5) & 31, c = e & 31;
      ds[e] = dy[(((size_t)b * Cout + co) * (2 * H) + th * 32 + r) * (size_t)(2 * W) + tw * 32 + c];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TW_PAIRS; ++q) {
      const int pair = threadIdx.x + q * 256;
      if (pair >= pairs) break;
      const int ci = pair / Cout, co = pair % Cout;
      for (int p = 0; p < 256; ++p) {
        const int r = p >> 4, c = p & 15;
        const float xv = xs[ci * 256 + p];
        const float* d = ds + co * 1024 + (2 * r) * 32 + 2 * c;
        a[q][0] = fmaf(xv, d[0], a[q][0]); a[q][1] = fmaf(xv, d[1], a[q][1]);
        a[q][2] = fmaf(xv, d[32], a[q][2]); a[q][3] = fmaf(xv, d[33], a[q][3]);
        bs[q] += (d[0] + d[1]) + (d[32] + d[33]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < TW_PAIRS; ++q) {
    const int pair = threadIdx.x + q * 256;
    if (pair >= pairs) break;
    const int ci = pair / Cout, co = pair % Cout;
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(dw + ((size_t)ci * Cout + co) * 4 + k, a[q][k]);
    if (ci == 0 && db) atomicAdd(db + co, bs[q]);
  }
}


// register-blocked form (product path): a thread owns CO_T x CI_T channel pairs (4 accumulators each) for a subset of the
// tile's pixels; lanes of a warp hold different pairs of the same pixel (broadcast loads, plane pitches 257 / 1026 words);
// per-CTA partial rows + the fixed-order reduction (reduce_partials_kernel).  The kernel above used one thread per pair for
// all 256 pixels of a tile: 32 of 256 threads busy for ConvTranspose2d(16, 2).
constexpr int TW2_XPL = 257, TW2_DPL = 1026;
template <int CO_T, int CI_T>
__global__ void __launch_bounds__(256, 2)
convT2x2_wgrad2_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int Cin, int Cout,
                       int H, int W, int n_tiles, int n_cb, int n_cib, int want_db) {
  extern __shared__ __align__(16) float sm[];
  const int cin_pad = n_cib * CI_T, cout_pad = n_cb * CO_T;
  float* xs = sm;                                     // [cin_pad][257]
  float* ds = sm + (((size_t)cin_pad * TW2_XPL + 1) & ~(size_t)1);   // [cout_pad][1026], 8-byte aligned
  const int tiles_w = W / 16, tiles_img = tiles_w * (H / 16);
  const int pairs = n_cb * n_cib;
  const int S = 256 / pairs;
  const int pair = threadIdx.x % pairs, s = threadIdx.x / pairs;
  const bool worker = s < S;
  const int cib = pair % n_cib, cb = pair / n_cib;
  float a[CO_T][CI_T][4], bs[CO_T];
#pragma unroll
  for (int j = 0; j < CO_T; ++j) {
    bs[j] = 0.f;
#pragma unroll
    for (int i = 0; i < CI_T; ++i) a[j][i][0] = a[j][i][1] = a[j][i][2] = a[j][i][3] = 0.f;
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / tiles_img, trem = tile - b * tiles_img;
    const int th = trem / tiles_w, tw = trem - th * tiles_w;
    __syncthreads();
    for (int e = threadIdx.x; e < cin_pad * 64; e += 256) {
      const int ci = e >> 6, r = (e >> 2) & 15, k = e & 3;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ci < Cin) v = *reinterpret_cast<const float4*>(x + (((size_t)b * Cin + ci) * H + th * 16 + r) * W + tw * 16 + 4 * k);
      float* dst = xs + ci * TW2_XPL + r * 16 + 4 * k;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    for (int e = threadIdx.x; e < cout_pad * 256; e += 256) {
      const int co = e >> 8, r = (e >> 3) & 31, k = e & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < Cout) v = *reinterpret_cast<const float4*>(dy + (((size_t)b * Cout + co) * (2 * H) + th * 32 + r) * (size_t)(2 * W) + tw * 32 + 4 * k);
      float2* dst = reinterpret_cast<float2*>(ds + co * TW2_DPL + r * 32 + 4 * k);
      dst[0] = make_float2(v.x, v.y); dst[1] = make_float2(v.z, v.w);
    }
    __syncthreads();
    if (!worker) continue;
    for (int p = s; p < 256; p += S) {
      const int r = p >> 4, c = p & 15;
      float xv[CI_T];
#pragma unroll
      for (int i = 0; i < CI_T; ++i) xv[i] = xs[(cib + i * n_cib) * TW2_XPL + p];
#pragma unroll
      for (int j = 0; j < CO_T; ++j) {
        const float* d = ds + (cb + j * n_cb) * TW2_DPL + (2 * r) * 32 + 2 * c;
        const float2 d0 = *reinterpret_cast<const float2*>(d), d1 = *reinterpret_cast<const float2*>(d + 32);
        if (want_db && cib == 0) bs[j] += (d0.x + d0.y) + (d1.x + d1.y);
#pragma unroll
        for (int i = 0; i < CI_T; ++i) {
          a[j][i][0] = fmaf(xv[i], d0.x, a[j][i][0]); a[j][i][1] = fmaf(xv[i], d0.y, a[j][i][1]);
          a[j][i][2] = fmaf(xv[i], d1.x, a[j][i][2]); a[j][i][3] = fmaf(xv[i], d1.y, a[j][i][3]);
        }
      }
    }
  }
  __syncthreads();
  constexpr int NACC = CO_T * CI_T * 4;
  float* red = sm;                               // [S][pairs][NACC + CO_T]
  if (worker) {
    float* mine = red + ((size_t)s * pairs + pair) * (NACC + CO_T);
#pragma unroll
    for (int j = 0; j < CO_T; ++j) {
#pragma unroll
      for (int i = 0; i < CI_T; ++i)
#pragma unroll
        for (int t = 0; t < 4; ++t) mine[(j * CI_T + i) * 4 + t] = a[j][i][t];
      mine[NACC + j] = bs[j];
    }
  }
  __syncthreads();
  const int n_dw = Cin * Cout * 4;
  float* my_part = part + (size_t)blockIdx.x * (n_dw + Cout);
  for (int e = threadIdx.x; e < pairs * (NACC + CO_T); e += 256) {
    const int pr = e / (NACC + CO_T), k = e - pr * (NACC + CO_T);
    float v = 0.f;
    for (int q = 0; q < S; ++q) v += red[((size_t)q * pairs + pr) * (NACC + CO_T) + k];
    const int pcib = pr % n_cib, pcb = pr / n_cib;
    if (k < NACC) {
      const int j = k / (CI_T * 4), i = (k / 4) % CI_T, t = k & 3;
      const int co = pcb + j * n_cb, ci = pcib + i * n_cib;
      if (co < Cout && ci < Cin) my_part[((size_t)ci * Cout + co) * 4 + t] = v;
    } else if (pcib == 0) {
      const int co = pcb + (k - NACC) * n_cb;
      if (co < Cout) my_part[n_dw + co] = v;
    }
  }
}

// any H, W (small tensors: the 8x8 / 16x16 planes of the image codec): one thread per input pixel and (ci, co) pair chunk, atomics
__global__ void __launch_bounds__(256)
convT2x2_wgrad_small_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, float* __restrict__ db,
                            int B, int Cin, int Cout, int H, int W) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)B * Cout * H * W) return;
  const int wq = (int)(idx % W), h = (int)((idx / W) % H), co = (int)((idx / ((size_t)W * H)) % Cout);
  const size_t b = idx / ((size_t)W * H * Cout);
  const float* d = dy + ((b * Cout + co) * (size_t)(2 * H) + 2 * h) * (2 * W) + 2 * wq;
  const float d00 = d[0], d01 = d[1], d10 = d[2 * W], d11 = d[2 * W + 1];
  if (db) atomicAdd(db + co, (d00 + d01) + (d10 + d11));
  for (int ci = 0; ci < Cin; ++ci) {
    const float xv = x[((b * Cin + ci) * H + h) * W + wq];
    float* o = dw + ((size_t)ci * Cout + co) * 4;
    atomicAdd(o, xv * d00); atomicAdd(o + 1, xv * d01); atomicAdd(o + 2, xv * d10); atomicAdd(o + 3, xv * d11);
  }
}

// ---- mean squared error (nn.MSELoss, train_modelA.py:435-445): loss += mean((a-b)^2) (fp64 accumulator),
// grad_a = grad_scale * 2 (a - b) / n
__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ grad_a, size_t n, float grad_scale,
           double* __restrict__ loss) {
  double acc = 0.0;
  float part = 0.f;
  int cnt = 0;
  const float gs = grad_scale * 2.0f / (float)n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    part = fmaf(d, d, part);
    if (grad_a) grad_a[i] = gs * d;
    if (++cnt == 1024) { acc += part; part = 0.f; cnt = 0; }
  }
  acc += part;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(loss, acc / (double)n);
}

// ---- Adam / AdamW over a flat parameter buffer (torch.optim.Adam semantics, train_modelA.py:234-236)
// step_dev (optional): device-resident step counter, so that the launch can be replayed from a CUDA graph: the
// kernel reads the count of COMPLETED steps, uses count + 1 for the bias corrections, and thread 0 of the LAST block
// to finish nothing - the increment is a separate 1-thread kernel launched after this one (adam_tick_kernel).
__global__ void adam_tick_kernel(int* step_dev) { *step_dev += 1; }

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
            float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2, float grad_scale,
            int decoupled, const int* __restrict__ step_dev) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (step_dev) {
    const float t = (float)(*step_dev + 1);
    bc1 = 1.f - powf(beta1, t);
    bc2 = 1.f - powf(beta2, t);
  }
  float grad = g[i] * grad_scale, w = p[i];
  if (decoupled) w *= 1.f - lr * weight_decay;        // AdamW
  else grad = fmaf(weight_decay, w, grad);            // Adam with L2 penalty
  const float mi = beta1 * m[i] + (1.f - beta1) * grad;
  const float vi = beta2 * v[i] + (1.f - beta2) * grad * grad;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
  p[i] = w - (lr / bc1) * (mi / denom);
}

// out = in * scale + shift (the audio_scale normalisation of the clips, uformerWM/audio_test.py:329-341,559-571,691-702)
__global__ void __launch_bounds__(256)
affine_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n4, size_t n, float scale, float shift) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 v = reinterpret_cast<const float4*>(in)[i];
    v.x = fmaf(v.x, scale, shift); v.y = fmaf(v.y, scale, shift); v.z = fmaf(v.z, scale, shift); v.w = fmaf(v.w, scale, shift);
    reinterpret_cast<float4*>(out)[i] = v;
  } else if (i < n4 + (n & 3)) {
    const size_t j = n4 * 4 + (i - n4);
    out[j] = fmaf(in[j], scale, shift);
  }
}

int grid_for(size_t n) { return (int)((n + 255) / 256); }

}  // namespace
}  // namespace wmk

using namespace wmk;

extern "C" int wmk_bn_train_fwd_f32(const float* x, float* y, const float* gamma, const float* beta, float* running_mean,
                                    float* running_var, float* mean_rstd, double* scratch, int B, int C, int HW, float eps,
                                    float momentum, int act, float slope, void* stream) {
  WMK_REQUIRE(x && y && gamma && beta && mean_rstd && scratch && B > 0 && C > 0 && HW > 0 && HW % 4 == 0 && act >= 0 &&
                  act <= 3 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0,
              "bn_train_fwd: bad arguments (H*W must be a multiple of 4, buffers 16-byte aligned)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * C * HW;
  ProfScope prof(FAM_SMALL, 12.0 * total, st);
  WMK_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
  const size_t per_c = (size_t)B * HW;
  int chunks = (int)((per_c / 4 + 256 * 8 - 1) / (256 * 8));
  const int max_chunks = (148 * 8 * 4 + C - 1) / C;          // ~4 waves of 8 resident CTAs per SM over all channels
  if (chunks > max_chunks) chunks = max_chunks;
  bn_stats_kernel<<<dim3(chunks, C), 256, 0, st>>>(x, B, C, HW, scratch);
  WMK_CHECK_LAUNCH("bn_stats_kernel");
  bn_finalize_kernel<<<cdiv(C, 64), 64, 0, st>>>(scratch, C, (double)per_c, eps, momentum, mean_rstd, running_mean, running_var);
  WMK_CHECK_LAUNCH("bn_finalize_kernel");
  bn_act_fwd_kernel<<<grid_for(total / 4), 256, 0, st>>>(x, y, mean_rstd, gamma, beta, total / 4, C, HW, act, slope);
  WMK_CHECK_LAUNCH("bn_act_fwd_kernel");
  return 0;
}

extern "C" int wmk_bn_train_bwd_f32(const float* x, const float* y, const float* dy, float* dx, const float* gamma,
                                    const float* mean_rstd, float* dgamma, float* dbeta, double* scratch, int B, int C,
                                    int HW, int act, float slope, void* stream) {
  WMK_REQUIRE(x && y && dy && dx && gamma && mean_rstd && dgamma && dbeta && scratch && B > 0 && C > 0 && HW > 0 && HW % 4 == 0 &&
                  (size_t)C <= (size_t)B * C * HW / 4,
              "bn_train_bwd: bad arguments (H*W must be a multiple of 4)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * C * HW;
  ProfScope prof(FAM_SMALL, 28.0 * total, st);
  WMK_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
  const size_t per_c = (size_t)B * HW;
  int chunks = (int)((per_c / 4 + 256 * 8 - 1) / (256 * 8));
  const int max_chunks = (148 * 8 * 4 + C - 1) / C;
  if (chunks > max_chunks) chunks = max_chunks;
  bn_act_bwd_reduce_kernel<<<dim3(chunks, C), 256, 0, st>>>(x, y, dy, mean_rstd, B, C, HW, act, slope, scratch);
  WMK_CHECK_LAUNCH("bn_act_bwd_reduce_kernel");
  bn_act_bwd_apply_kernel<<<grid_for(total / 4), 256, 0, st>>>(x, y, dy, dx, mean_rstd, gamma, scratch, total / 4, C, HW,
                                                           (double)per_c, act, slope, dgamma, dbeta);
  WMK_CHECK_LAUNCH("bn_act_bwd_apply_kernel");
  return 0;
}

extern "C" int wmk_bn_pool_train_fwd_f32(const float* x, float* y, float* y_pooled, const float* gamma, const float* beta,
                                         float* running_mean, float* running_var, float* mean_rstd, double* scratch, int B,
                                         int C, int H, int W, float eps, float momentum, int act, float slope, void* stream) {
  WMK_REQUIRE(x && y && y_pooled && gamma && beta && mean_rstd && scratch && B > 0 && C > 0 && H >= 2 && W >= 4 && H % 2 == 0 &&
                  W % 4 == 0 && act >= 0 && act <= 3 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                  ((uintptr_t)y_pooled & 7) == 0,
              "bn_pool_train_fwd: bad arguments (H even, W a multiple of 4, buffers 16-byte aligned)");
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  const size_t total = (size_t)B * C * HW;
  ProfScope prof(FAM_SMALL, 13.0 * total, st);
  WMK_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
  const size_t per_c = (size_t)B * HW;
  int chunks = (int)((per_c / 4 + 256 * 8 - 1) / (256 * 8));
  const int max_chunks = (148 * 8 * 4 + C - 1) / C;
  if (chunks > max_chunks) chunks = max_chunks;
  bn_stats_kernel<<<dim3(chunks, C), 256, 0, st>>>(x, B, C, HW, scratch);
  WMK_CHECK_LAUNCH("bn_stats_kernel");
  bn_finalize_kernel<<<cdiv(C, 64), 64, 0, st>>>(scratch, C, (double)per_c, eps, momentum, mean_rstd, running_mean, running_var);
  WMK_CHECK_LAUNCH("bn_finalize_kernel");
  bn_act_pool_fwd_kernel<<<grid_for(total / 8), 256, 0, st>>>(x, y, y_pooled, mean_rstd, gamma, beta, total / 8, C, H, W, act, slope);
  WMK_CHECK_LAUNCH("bn_act_pool_fwd_kernel");
  return 0;
}

extern "C" int wmk_bn_pool_train_bwd_f32(const float* x, const float* y, const float* dy_pooled, float* dx, const float* gamma,
                                         const float* mean_rstd, float* dgamma, float* dbeta, double* scratch, int B, int C,
                                         int H, int W, int act, float slope, void* stream) {
  WMK_REQUIRE(x && y && dy_pooled && dx && gamma && mean_rstd && dgamma && dbeta && scratch && B > 0 && C > 0 && H >= 2 && W >= 4 &&
                  H % 2 == 0 && W % 4 == 0 && (size_t)C <= (size_t)B * C * H * W / 8 && ((uintptr_t)x & 15) == 0 &&
                  ((uintptr_t)y & 15) == 0 && ((uintptr_t)dx & 15) == 0 && ((uintptr_t)dy_pooled & 7) == 0,
              "bn_pool_train_bwd: bad arguments (H even, W a multiple of 4, buffers 16-byte aligned)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)B * C * H * W;
  ProfScope prof(FAM_SMALL, 22.5 * total, st);
  WMK_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
  const size_t units_c = (size_t)B * H * W / 8;
  int chunks = (int)((units_c + 256 * 4 - 1) / (256 * 4));
  const int max_chunks = (148 * 8 * 4 + C - 1) / C;
  if (chunks > max_chunks) chunks = max_chunks;
  bn_act_pool_bwd_reduce_kernel<<<dim3(chunks, C), 256, 0, st>>>(x, y, dy_pooled, mean_rstd, B, C, H, W, act, slope, scratch);
  WMK_CHECK_LAUNCH("bn_act_pool_bwd_reduce_kernel");
  bn_act_pool_bwd_apply_kernel<<<grid_for(total / 8), 256, 0, st>>>(x, y, dy_pooled, dx, mean_rstd, gamma, scratch, total / 8, C, H, W,
                                                                    (double)B * H * W, act, slope, dgamma, dbeta);
  WMK_CHECK_LAUNCH("bn_act_pool_bwd_apply_kernel");
  return 0;
}

extern "C" int wmk_maxpool2x2_bwd_f32(const float* x, const float* dy, float* dx, int planes, int H, int W, void* stream) {
  WMK_REQUIRE(x && dy && dx && planes > 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "maxpool2x2_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 9.0 * planes * H * W, st);
  maxpool2x2_bwd_kernel<<<grid_for((size_t)planes * (H / 2) * (W / 2)), 256, 0, st>>>(x, dy, dx, (size_t)planes, H, W);
  WMK_CHECK_LAUNCH("maxpool2x2_bwd_kernel");
  return 0;
}

extern "C" int wmk_mask_scale_f32(const float* in, const float* mask, float* out, size_t n, float scale, void* stream) {
  WMK_REQUIRE(in && mask && out && n > 0, "mask_scale: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 12.0 * n, st);
  mask_scale_kernel<<<grid_for(n), 256, 0, st>>>(in, mask, out, n, scale);
  WMK_CHECK_LAUNCH("mask_scale_kernel");
  return 0;
}

namespace wmk {
namespace {
int num_sms_cached() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  return sms;
}

template <int CO_T, int CI_T>
int launch_wgrad2(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int H, int W, cudaStream_t st) {
  const int n_cb = (Cout + CO_T - 1) / CO_T, n_cib = (Cin + CI_T - 1) / CI_T, pairs = n_cb * n_cib;
  const bool shfl = pairs < 32 && (32 % pairs) == 0;
  const int n_sub = shfl ? WG2_NT / 32 : WG2_NT / pairs;
  const size_t tile_b = ((size_t)n_cib * CI_T * WG2_XPL + (size_t)n_cb * CO_T * WG2_DPL) * sizeof(float);
  const size_t red_b = (size_t)n_sub * pairs * (CO_T * CI_T * 9 + CO_T) * sizeof(float);
  const size_t smem = tile_b > red_b ? tile_b : red_b;
  WMK_REQUIRE(smem <= 220 * 1024, "conv3x3_wgrad: Cin=%d Cout=%d needs %zu bytes of shared memory", Cin, Cout, smem);
  const int n_tiles = (H / 16) * (W / 16) * B;
  auto kern = conv3x3_wgrad2_kernel<CO_T, CI_T>;
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  int per_sm = 1;
  WMK_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WG2_NT, smem));
  if (per_sm < 1) per_sm = 1;
  const int grid = n_tiles < per_sm * num_sms_cached() ? n_tiles : per_sm * num_sms_cached();
  const int n = Cout * Cin * 9 + Cout;
  float* part = nullptr;
  WMK_CHECK_CUDA(cudaMallocAsync(&part, (size_t)grid * n * sizeof(float), st));
  kern<<<grid, WG2_NT, smem, st>>>(x, dy, part, Cin, Cout, H, W, n_tiles, n_cb, n_cib, db != nullptr);
  WMK_CHECK_LAUNCH("conv3x3_wgrad2_kernel");
  reduce_partials_kernel<<<cdiv(n, 8), 256, 0, st>>>(part, grid, n, dw, Cout * Cin * 9, db);
  WMK_CHECK_LAUNCH("reduce_partials_kernel");
  WMK_CHECK_CUDA(cudaFreeAsync(part, st));
  return 0;
}
}  // namespace
}  // namespace wmk

extern "C" int wmk_conv3x3_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int H,
                                     int W, void* stream) {
  WMK_REQUIRE(x && dy && dw && B > 0 && Cin > 0 && Cout > 0 && H % 16 == 0 && W % 16 == 0 && ((uintptr_t)x & 15) == 0 &&
                  ((uintptr_t)dy & 15) == 0,
              "conv3x3_wgrad: bad arguments (H, W must be multiples of 16, buffers 16-byte aligned)");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 4.0 * B * H * W * (Cin + Cout), st);
  static const int legacy = getenv("WMK_WGRAD_LEGACY") ? atoi(getenv("WMK_WGRAD_LEGACY")) : 0;
  if (!legacy) {
    // register-blocked kernel: (output, input) channel blocking by layer shape
    if (Cout <= 2 && (Cin + 3) / 4 * Cout <= WG2_NT && Cin <= 96) return launch_wgrad2<1, 4>(x, dy, dw, db, B, Cin, Cout, H, W, st);
    if (((Cout + 3) / 4) * ((Cin + 1) / 2) <= WG2_NT && ((size_t)((Cin + 1) / 2 * 2) * WG2_XPL + (size_t)((Cout + 3) / 4 * 4) * WG2_DPL) * 4 <= 220 * 1024)
      return launch_wgrad2<4, 2>(x, dy, dw, db, B, Cin, Cout, H, W, st);
  }
  WMK_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * 9, st));
  if (db) WMK_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cout, st));
  const size_t smem = ((size_t)Cin * 18 * 18 + (size_t)Cout * 256) * sizeof(float);
  WMK_REQUIRE(smem <= 200 * 1024, "conv3x3_wgrad: Cin=%d Cout=%d needs %zu bytes of shared memory", Cin, Cout, smem);
  const int n_tiles = (H / 16) * (W / 16) * B;
  const int sms = num_sms_cached();
  const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;
  // CB output channels per thread: keep (Cout/CB)*Cin <= 256 thread slots
  if ((size_t)((Cout + 3) / 4) * Cin <= 256) {
    if (smem > 48 * 1024) WMK_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_wgrad_kernel<4><<<grid, 256, smem, st>>>(x, dy, dw, db, Cin, Cout, H, W, n_tiles);
  } else {
    WMK_REQUIRE((size_t)((Cout + 15) / 16) * Cin <= 256, "conv3x3_wgrad: Cin=%d x Cout=%d too large", Cin, Cout);
    if (smem > 48 * 1024) WMK_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv3x3_wgrad_kernel<16><<<grid, 256, smem, st>>>(x, dy, dw, db, Cin, Cout, H, W, n_tiles);
  }
  WMK_CHECK_LAUNCH("conv3x3_wgrad_kernel");
  return 0;
}

extern "C" int wmk_convT2x2_dgrad_f32(const float* dy, const float* w, float* dx, int B, int Cin, int Cout, int H, int W,
                                      void* stream) {
  WMK_REQUIRE(dy && w && dx && B > 0 && Cin > 0 && Cout > 0 && H > 0 && W > 0 && (size_t)Cin * Cout * 16 <= 96 * 1024,
              "convT2x2_dgrad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 4.0 * B * H * W * (Cin + 4 * Cout), st);
  const size_t smem = (size_t)Cin * Cout * 16;
  if ((Cout <= 2 || (Cout > 4 && Cout <= 16)) && ((uintptr_t)dy & 7) == 0 && (size_t)Cin * (Cout <= 2 ? 2 : 16) * 16 <= 48 * 1024) {
    if (Cout <= 2) convT2x2_dgrad_rb_kernel<2><<<grid_for((size_t)B * H * W), 256, (size_t)Cin * 2 * 16, st>>>(dy, w, dx, B, Cin, Cout, H, W);
    else convT2x2_dgrad_rb_kernel<16><<<grid_for((size_t)B * H * W), 256, (size_t)Cin * 16 * 16, st>>>(dy, w, dx, B, Cin, Cout, H, W);
    WMK_CHECK_LAUNCH("convT2x2_dgrad_rb_kernel");
    return 0;
  }
  if (smem > 48 * 1024) WMK_CHECK_CUDA(cudaFuncSetAttribute(convT2x2_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  convT2x2_dgrad_kernel<<<grid_for((size_t)B * H * W), 256, smem, st>>>(dy, w, dx, B, Cin, Cout, H, W);
  WMK_CHECK_LAUNCH("convT2x2_dgrad_kernel");
  return 0;
}

namespace wmk {
namespace {
template <int CO_T, int CI_T>
int launch_convT_wgrad2(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int H, int W, cudaStream_t st) {
  const int n_cb = (Cout + CO_T - 1) / CO_T, n_cib = (Cin + CI_T - 1) / CI_T, pairs = n_cb * n_cib;
  const int S = 256 / pairs;
  const size_t tile_b = ((((size_t)n_cib * CI_T * TW2_XPL + 1) & ~(size_t)1) + (size_t)n_cb * CO_T * TW2_DPL) * sizeof(float);
  const size_t red_b = (size_t)S * pairs * (CO_T * CI_T * 4 + CO_T) * sizeof(float);
  const size_t smem = tile_b > red_b ? tile_b : red_b;
  const int n_tiles = (H / 16) * (W / 16) * B;
  auto kern = convT2x2_wgrad2_kernel<CO_T, CI_T>;
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  int per_sm = 1;
  WMK_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
  if (per_sm < 1) per_sm = 1;
  const int grid = n_tiles < per_sm * num_sms_cached() ? n_tiles : per_sm * num_sms_cached();
  const int n = Cin * Cout * 4 + Cout;
  float* part = nullptr;
  WMK_CHECK_CUDA(cudaMallocAsync(&part, (size_t)grid * n * sizeof(float), st));
  kern<<<grid, 256, smem, st>>>(x, dy, part, Cin, Cout, H, W, n_tiles, n_cb, n_cib, db != nullptr);
  WMK_CHECK_LAUNCH("convT2x2_wgrad2_kernel");
  reduce_partials_kernel<<<cdiv(n, 8), 256, 0, st>>>(part, grid, n, dw, Cin * Cout * 4, db);
  WMK_CHECK_LAUNCH("reduce_partials_kernel");
  WMK_CHECK_CUDA(cudaFreeAsync(part, st));
  return 0;
}
}  // namespace
}  // namespace wmk

extern "C" int wmk_convT2x2_wgrad_f32(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int H,
                                      int W, void* stream) {
  WMK_REQUIRE(x && dy && dw && B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && H > 0 && W > 0, "convT2x2_wgrad: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 4.0 * B * H * W * (Cin + 4 * Cout), st);
  if (H % 16 != 0 || W % 16 != 0 || ((uintptr_t)x & 15) != 0 || ((uintptr_t)dy & 15) != 0) {      // small planes: plain atomic kernel
    WMK_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * 4, st));
    if (db) WMK_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cout, st));
    convT2x2_wgrad_small_kernel<<<grid_for((size_t)B * Cout * H * W), 256, 0, st>>>(x, dy, dw, db, B, Cin, Cout, H, W);
    WMK_CHECK_LAUNCH("convT2x2_wgrad_small_kernel");
    return 0;
  }
  const size_t smem = ((size_t)Cin * 256 + (size_t)Cout * 1024) * sizeof(float);
  WMK_REQUIRE(smem <= 200 * 1024, "convT2x2_wgrad: Cin=%d Cout=%d needs %zu bytes of shared memory", Cin, Cout, smem);
  static const int legacy = getenv("WMK_WGRAD_LEGACY") ? atoi(getenv("WMK_WGRAD_LEGACY")) : 0;
  if (!legacy) {
    if (Cout <= 2 && (Cin + 1) / 2 <= 256) return launch_convT_wgrad2<2, 2>(x, dy, dw, db, B, Cin, Cout, H, W, st);
    if (((Cout + 3) / 4) * ((Cin + 2) / 3) <= 256) return launch_convT_wgrad2<4, 3>(x, dy, dw, db, B, Cin, Cout, H, W, st);
  }
  WMK_CHECK_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * 4, st));
  if (db) WMK_CHECK_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * Cout, st));
  WMK_REQUIRE(Cin * Cout <= 256 * TW_PAIRS, "convT2x2_wgrad: Cin*Cout=%d exceeds %d", Cin * Cout, 256 * TW_PAIRS);
  const int n_tiles = (H / 16) * (W / 16) * B;
  const int sms = num_sms_cached();
  const int grid = n_tiles < 2 * sms ? n_tiles : 2 * sms;
  if (smem > 48 * 1024) WMK_CHECK_CUDA(cudaFuncSetAttribute(convT2x2_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  convT2x2_wgrad_kernel<<<grid, 256, smem, st>>>(x, dy, dw, db, Cin, Cout, H, W, n_tiles);
  WMK_CHECK_LAUNCH("convT2x2_wgrad_kernel");
  return 0;
}

extern "C" int wmk_mse_f32(const float* a, const float* b, float* grad_a, size_t n, float grad_scale, double* loss_accum,
                           void* stream) {
  WMK_REQUIRE(a && b && loss_accum && n > 0, "mse: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_STATS, (grad_a ? 12.0 : 8.0) * n, st);
  int blocks = grid_for(n);
  if (blocks > 148 * 8) blocks = 148 * 8;
  mse_kernel<<<blocks, 256, 0, st>>>(a, b, grad_a, n, grad_scale, loss_accum);
  WMK_CHECK_LAUNCH("mse_kernel");
  return 0;
}

extern "C" int wmk_adam_step_f32(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                 int decoupled, int* step_dev, void* stream) {
  WMK_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && (step >= 1 || step_dev), "adam_step: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_SMALL, 28.0 * n, st);
  const float bc1 = 1.f - powf(beta1, (float)(step >= 1 ? step : 1)), bc2 = 1.f - powf(beta2, (float)(step >= 1 ? step : 1));
  adam_kernel<<<grid_for(n), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                           grad_scale, decoupled, step_dev);
  WMK_CHECK_LAUNCH("adam_kernel");
  if (step_dev) {
    adam_tick_kernel<<<1, 1, 0, st>>>(step_dev);
    WMK_CHECK_LAUNCH("adam_tick_kernel");
  }
  return 0;
}

extern "C" int wmk_affine_f32(const float* in, float* out, size_t n, float scale, float shift, void* stream) {
  WMK_REQUIRE(in && out && n > 0 && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "affine: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ProfScope prof(FAM_ATTACK, 8.0 * n, st);
  const size_t n4 = n / 4;
  affine_kernel<<<grid_for(n4 + 3), 256, 0, st>>>(in, out, n4, n, scale, shift);
  WMK_CHECK_LAUNCH("affine_kernel");
  return 0;
}

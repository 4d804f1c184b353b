// Micro-benchmark: FP32 FMA issue rate on sm_100a by operand form (3 vector registers, packed FFMA2,
// a warp-uniform operand from the constant bank / a uniform register, a per-thread operand).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_forms fma_forms.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NACC = 8, NW = 36, ITERS = 512;

struct WParam { float w[1024]; };

// mode 0: weights in vector registers (3-register FFMA)
__global__ void __launch_bounds__(128, 4) k_vec(const float* __restrict__ wg, float* out, int n) {
  float w[NW], a[NACC], x[NACC];
  for (int i = 0; i < NW; ++i) w[i] = wg[(threadIdx.x & 15) * NW + i];
  for (int i = 0; i < NACC; ++i) { a[i] = 0.f; x[i] = threadIdx.x * 0.001f + i; }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < NW; ++j)
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = fmaf(x[(i + j) % NACC], w[j], a[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) x[i] = a[i] * 1e-9f + x[i];
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 1: packed FFMA2, weights in vector registers
__global__ void __launch_bounds__(128, 4) k_vec2(const float* __restrict__ wg, float* out, int n) {
  float2 w[NW / 2], a[NACC], x[NACC];
  for (int i = 0; i < NW / 2; ++i) w[i] = make_float2(wg[(threadIdx.x & 15) * NW + i], wg[(threadIdx.x & 15) * NW + i + 1]);
  for (int i = 0; i < NACC; ++i) { a[i] = make_float2(0.f, 0.f); x[i] = make_float2(threadIdx.x * 0.001f + i, i); }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < NW / 2; ++j)
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = __ffma2_rn(x[(i + j) % NACC], w[j], a[i]);
#pragma unroll
    for (int i = 0; i < NACC; ++i) { x[i].x = a[i].x * 1e-9f + x[i].x; x[i].y = a[i].y * 1e-9f + x[i].y; }
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 2: weights warp-uniform from a kernel parameter with a runtime (block-uniform) base index
__global__ void __launch_bounds__(128, 4) k_param(const __grid_constant__ WParam p, float* out, int n, int base) {
  float a[NACC], x[NACC];
  for (int i = 0; i < NACC; ++i) { a[i] = 0.f; x[i] = threadIdx.x * 0.001f + i; }
  const int b = (base + blockIdx.x * 40) & 511;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const float w = p.w[b + j];
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = fmaf(x[(i + j) % NACC], w, a[i]);
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) x[i] = a[i] * 1e-9f + x[i];
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 3: compile-time constant-bank offsets (weights at fixed param offsets)
__global__ void __launch_bounds__(128, 4) k_cbank(const __grid_constant__ WParam p, float* out, int n) {
  float a[NACC], x[NACC];
  for (int i = 0; i < NACC; ++i) { a[i] = 0.f; x[i] = threadIdx.x * 0.001f + i; }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < NW; ++j) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = fmaf(x[(i + j) % NACC], p.w[j], a[i]);
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) x[i] = a[i] * 1e-9f + x[i];
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 4: packed FFMA2 with a warp-uniform weight pair from the parameter bank
__global__ void __launch_bounds__(128, 4) k_param2(const __grid_constant__ WParam p, float* out, int n, int base) {
  float2 a[NACC], x[NACC];
  for (int i = 0; i < NACC; ++i) { a[i] = make_float2(0.f, 0.f); x[i] = make_float2(threadIdx.x * 0.001f + i, i); }
  const int b = (base + blockIdx.x * 40) & 510;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < NW / 2; ++j) {
      const float2 w = make_float2(p.w[b + 2 * j], p.w[b + 2 * j + 1]);
#pragma unroll
      for (int i = 0; i < NACC; ++i) a[i] = __ffma2_rn(x[(i + j) % NACC], w, a[i]);
    }
#pragma unroll
    for (int i = 0; i < NACC; ++i) { x[i].x = a[i].x * 1e-9f + x[i].x; x[i].y = a[i].y * 1e-9f + x[i].y; }
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 5: half the accumulators packed (FFMA2), half scalar (2 x FFMA), weights in vector registers
__global__ void __launch_bounds__(128, 4) k_mix(const float* __restrict__ wg, float* out, int n) {
  float2 w[NW / 2], a[NACC], x[NACC];
  for (int i = 0; i < NW / 2; ++i) w[i] = make_float2(wg[(threadIdx.x & 15) * NW + i], wg[(threadIdx.x & 15) * NW + i + 1]);
  for (int i = 0; i < NACC; ++i) { a[i] = make_float2(0.f, 0.f); x[i] = make_float2(threadIdx.x * 0.001f + i, i); }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < NW / 2; ++j)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (i & 1) {
          a[i].x = fmaf(x[(i + j) % NACC].x, w[j].x, a[i].x);
          a[i].y = fmaf(x[(i + j) % NACC].y, w[j].y, a[i].y);
        } else {
          a[i] = __ffma2_rn(x[(i + j) % NACC], w[j], a[i]);
        }
      }
#pragma unroll
    for (int i = 0; i < NACC; ++i) { x[i].x = a[i].x * 1e-9f + x[i].x; x[i].y = a[i].y * 1e-9f + x[i].y; }
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 6: one packed per three FMAs' worth (1 FFMA2 : 2 FFMA ... i % 4 == 0 packed)
__global__ void __launch_bounds__(128, 4) k_mix4(const float* __restrict__ wg, float* out, int n) {
  float2 w[NW / 2], a[NACC], x[NACC];
  for (int i = 0; i < NW / 2; ++i) w[i] = make_float2(wg[(threadIdx.x & 15) * NW + i], wg[(threadIdx.x & 15) * NW + i + 1]);
  for (int i = 0; i < NACC; ++i) { a[i] = make_float2(0.f, 0.f); x[i] = make_float2(threadIdx.x * 0.001f + i, i); }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int j = 0; j < NW / 2; ++j)
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (i & 3) {
          a[i].x = fmaf(x[(i + j) % NACC].x, w[j].x, a[i].x);
          a[i].y = fmaf(x[(i + j) % NACC].y, w[j].y, a[i].y);
        } else {
          a[i] = __ffma2_rn(x[(i + j) % NACC], w[j], a[i]);
        }
      }
#pragma unroll
    for (int i = 0; i < NACC; ++i) { x[i].x = a[i].x * 1e-9f + x[i].x; x[i].y = a[i].y * 1e-9f + x[i].y; }
  }
  float s = 0.f;
  for (int i = 0; i < NACC; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float *wg, *out;
  cudaMalloc(&wg, 16 * NW * 4);
  cudaMemset(wg, 0, 16 * NW * 4);
  const int grid = sms * 4, block = 128;
  cudaMalloc(&out, grid * block * 4);
  WParam hp;
  for (int i = 0; i < 1024; ++i) hp.w[i] = 1e-6f * i;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double fma_per_thread = (double)ITERS * NW * NACC;
  for (int mode = 0; mode < 7; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(e0);
      switch (mode) {
        case 0: k_vec<<<grid, block>>>(wg, out, ITERS); break;
        case 1: k_vec2<<<grid, block>>>(wg, out, ITERS); break;
        case 2: k_param<<<grid, block>>>(hp, out, ITERS, rep); break;
        case 3: k_cbank<<<grid, block>>>(hp, out, ITERS); break;
        case 4: k_param2<<<grid, block>>>(hp, out, ITERS, rep); break;
        case 5: k_mix<<<grid, block>>>(wg, out, ITERS); break;
        case 6: k_mix4<<<grid, block>>>(wg, out, ITERS); break;
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    const double fmas = fma_per_thread * grid * block * (mode == 1 || mode == 4 ? 1.0 : 1.0);
    const char* names[] = {"FFMA 3 vector regs", "FFMA2 3 vector regs", "FFMA uniform param (runtime index)",
                           "FFMA const bank (fixed offset)", "FFMA2 uniform param pair",
                           "1 FFMA2 : 2 FFMA, vector regs", "1 FFMA2 : 6 FFMA, vector regs"};
    printf("%-38s %8.3f ms  %7.1f FMA/clk/SM (at %d MHz nominal)  err=%s\n", names[mode], best,
           fmas / (best * 1e-3) / sms / (clk * 1e3), clk / 1000, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}

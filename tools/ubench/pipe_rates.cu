// Micro-benchmark: issue rates on sm_100a of the instructions the fused LeFF kernel's CUDA-core stages are made of
// (FHFMA = fma.rn.f32.f16 mixed-precision FMA, FFMA2, HADD2.F32 unpack, MUFU.TANH, cvt pack), per SM and clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;

__device__ __forceinline__ float fhfma(uint32_t a, uint32_t b, float c, int ha, int hb) {
  // a, b hold two fp16 each; ha / hb select the half
  float d;
  if (ha == 0 && hb == 0)
    asm volatile("{.reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n fma.rn.f32.f16 %0, al, bl, %3;}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  else
    asm volatile("{.reg .b16 al, ah, bl, bh;\n mov.b32 {al, ah}, %1;\n mov.b32 {bl, bh}, %2;\n fma.rn.f32.f16 %0, ah, bh, %3;}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}

// 0: FHFMA (8 independent accumulators, 4 x-registers, 4 w-registers)
__global__ void __launch_bounds__(256, 2) k_fhfma(const uint32_t* in, float* out, int n) {
  uint32_t x[4], w[4];
  float a[8];
  for (int i = 0; i < 4; ++i) { x[i] = in[threadIdx.x + 32 * i]; w[i] = in[threadIdx.x + 32 * i + 128]; }
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fhfma(x[(i + r) & 3], w[(i * 3 + r) & 3], a[i], i & 1, i & 1);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 1: FFMA2, three vector operands
__global__ void __launch_bounds__(256, 2) k_ffma2(const float* in, float* out, int n) {
  float2 x[4], w[4], a[8];
  for (int i = 0; i < 4; ++i) { x[i] = make_float2(in[threadIdx.x + 32 * i], in[threadIdx.x + 1]); w[i] = make_float2(in[threadIdx.x + 32 * i + 128], 0.5f); }
  for (int i = 0; i < 8; ++i) a[i] = make_float2(0.f, 0.f);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(x[(i + r) & 3], w[(i * 3 + r) & 3], a[i]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 2: HADD2.F32 unpack (fp16 -> fp32), consumed by an FADD to keep it alive
__global__ void __launch_bounds__(256, 2) k_unpack(const uint32_t* in, float* out, int n) {
  uint32_t x[8];
  float a[8];
  for (int i = 0; i < 8; ++i) { x[i] = in[threadIdx.x + 32 * i]; a[i] = 0.f; }
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&x[i]));
        x[i] = __float_as_uint(f.x + f.y) ^ (uint32_t)it;       // 2 unpacks + 1 FADD + 1 LOP per step
      }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += __uint_as_float(x[i]) + a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 3: MUFU.TANH
__global__ void __launch_bounds__(256, 2) k_tanh(const float* in, float* out, int n) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + 32 * i];
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4: the tanh-form GELU on pairs (5 packed FMA-pipe instructions + 2 MUFU per pair) + cvt pack
__device__ __forceinline__ float2 gelu_h(float2 h) {
  const float2 h2 = __fmul2_rn(h, h);
  float2 p = __ffma2_rn(h2, make_float2(32.f * -3.81889112e-04f, 32.f * -3.81889112e-04f), make_float2(8.f * 3.72153111e-02f, 8.f * 3.72153111e-02f));
  p = __ffma2_rn(p, h2, make_float2(2.f * 7.97237410e-01f, 2.f * 7.97237410e-01f));
  const float2 u = __fmul2_rn(p, h);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  return __ffma2_rn(h, t, h);
}
__global__ void __launch_bounds__(256, 2) k_gelu(const float* in, float* out, int n) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 1]);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = gelu_h(a[i]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 5: the same GELU in half2 arithmetic (HFMA2 + packed tanh.approx.f16x2)
__device__ __forceinline__ __half2 gelu_h2(__half2 h) {
  const __half2 h2 = __hmul2(h, h);
  __half2 p = __hfma2(h2, __float2half2_rn(32.f * -3.81889112e-04f), __float2half2_rn(8.f * 3.72153111e-02f));
  p = __hfma2(p, h2, __float2half2_rn(2.f * 7.97237410e-01f));
  const __half2 u = __hmul2(p, h);
  uint32_t t, uu = *reinterpret_cast<const uint32_t*>(&u);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(uu));
  return __hfma2(h, *reinterpret_cast<const __half2*>(&t), h);
}
__global__ void __launch_bounds__(256, 2) k_gelu_h2(const float* in, float* out, int n) {
  __half2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = __floats2half2_rn(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 1]);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = gelu_h2(a[i]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += __low2float(a[i]) + __high2float(a[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// 6: scalar FFMA, three register operands
__global__ void __launch_bounds__(256, 2) k_ffma(const float* in, float* out, int n) {
  float x[4], w[4], a[8];
  for (int i = 0; i < 4; ++i) { x[i] = in[threadIdx.x + 32 * i]; w[i] = in[threadIdx.x + 32 * i + 128]; }
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(x[(i + r) & 3], w[(i * 3 + r) & 3], a[i]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 7 / 8: MUFU.RCP, MUFU.EX2
__global__ void __launch_bounds__(256, 2) k_rcp(const float* in, float* out, int n) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + 32 * i];
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256, 2) k_ex2(const float* in, float* out, int n) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x + 32 * i];
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 9: the exact-class GELU of the precise extractor (log-odds polynomial: 10 packed + 2 EX2 + 2 RCP per pair)
__device__ __forceinline__ float2 gelu_x(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  float2 p = __ffma2_rn(x2, make_float2(-5.209925380000868e-09f, -5.209925380000868e-09f), make_float2(3.850457233056659e-07f, 3.850457233056659e-07f));
  p = __ffma2_rn(p, x2, make_float2(-1.1452440958237275e-05f, -1.1452440958237275e-05f));
  p = __ffma2_rn(p, x2, make_float2(1.5938949945848435e-04f, 1.5938949945848435e-04f));
  p = __ffma2_rn(p, x2, make_float2(9.559268073644489e-05f, 9.559268073644489e-05f));
  p = __ffma2_rn(p, x2, make_float2(-0.10483857989311218f, -0.10483857989311218f));
  p = __ffma2_rn(p, x2, make_float2(-2.3022072315216064f, -2.3022072315216064f));
  const float2 a = __fmul2_rn(p, x);
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
  const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return __fmul2_rn(x, r);
}
__global__ void __launch_bounds__(256, 2) k_gelu_x(const float* in, float* out, int n) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 1]);
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = gelu_x(a[i]);
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double run(F launch, int grid) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(grid, 16);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch(grid, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, grid = sms * 2;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  uint32_t* in; float* out;
  cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, (size_t)grid * 256 * 4);
  cudaMemset(in, 0x3c, 4096 * 4);
  const double threads = (double)grid * 256;
  auto report = [&](const char* name, double ms, double ops_per_thread_iter) {
    // rate per SM per clock at the NOMINAL max clock (the real clock under load is lower: compare the rows)
    const double ops = threads * ops_per_thread_iter * ITERS;
    printf("%-34s %8.3f ms   %7.1f ops/clk/SM @%d MHz nominal   %8.2f Tops/s\n", name, ms, ops / (ms * 1e-3) / sms / (khz * 1e3), khz / 1000, ops / ms * 1e-9);
  };
  report("FHFMA (fma.rn.f32.f16)", run([&](int g, int n) { k_fhfma<<<g, 256>>>(in, out, n); }, grid), 64);
  report("FFMA2 3-reg (FMAs = 2 x instr)", run([&](int g, int n) { k_ffma2<<<g, 256>>>((const float*)in, out, n); }, grid), 128);
  report("HADD2.F32 unpack (elements)", run([&](int g, int n) { k_unpack<<<g, 256>>>(in, out, n); }, grid), 128);
  report("MUFU.TANH", run([&](int g, int n) { k_tanh<<<g, 256>>>((const float*)in, out, n); }, grid), 64);
  report("GELU tanh-form fp32x2 (elements)", run([&](int g, int n) { k_gelu<<<g, 256>>>((const float*)in, out, n); }, grid), 64);
  report("GELU tanh-form half2 (elements)", run([&](int g, int n) { k_gelu_h2<<<g, 256>>>((const float*)in, out, n); }, grid), 64);
  report("FFMA scalar 3-reg", run([&](int g, int n) { k_ffma<<<g, 256>>>((const float*)in, out, n); }, grid), 64);
  report("MUFU.RCP", run([&](int g, int n) { k_rcp<<<g, 256>>>((const float*)in, out, n); }, grid), 64);
  report("MUFU.EX2", run([&](int g, int n) { k_ex2<<<g, 256>>>((const float*)in, out, n); }, grid), 64);
  report("GELU exact-class fp32x2 (elements)", run([&](int g, int n) { k_gelu_x<<<g, 256>>>((const float*)in, out, n); }, grid), 64);
  return 0;
}

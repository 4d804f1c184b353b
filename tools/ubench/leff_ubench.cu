// Stand-alone timing of the fused LeFF block kernel (csrc/leff_block.cu) on the shapes of the bench step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ubench/leff_ubench tools/ubench/leff_ubench.cu \
//        -L image-in-speech-watermarking_b200/csrc -lwmk -Xlinker -rpath -Xlinker '$ORIGIN/../../image-in-speech-watermarking_b200/csrc'
//   tools/ubench/leff_ubench [precise] [clips]
// Prints per shape: microseconds, cycles per 64-channel chunk per SM (at the nominal 1.7 GHz), effective TFLOP/s.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace wmk {
int leff_block(const void* A, const void* W1, const void* W2, const float* b1, const uint16_t* dw16, const float* dw_b,
               const float* b2, float* x, int n, int H, int C, int precise, cudaStream_t st);
}
extern "C" const char* wmk_last_error(void);

static __global__ void fill_half(__half* p, size_t n, float scale, unsigned seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)(i * 2654435761u) ^ seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  p[i] = __float2half(((h & 0xffff) / 32768.0f - 1.0f) * scale);
}
static __global__ void fill_float(float* p, size_t n, float scale, unsigned seed) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned h = (unsigned)(i * 2654435761u) ^ seed;
  h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
  p[i] = ((h & 0xffff) / 32768.0f - 1.0f) * scale;
}

int main(int argc, char** argv) {
  const int precise = argc > 1 ? atoi(argv[1]) : 0;
  const int clips = argc > 2 ? atoi(argv[2]) : 384;
  const int only = argc > 3 ? atoi(argv[3]) : -1;          // shape index, -1 = all
  struct Shape { int C, H; };
  const Shape shapes[] = {{32, 128}, {64, 64}, {128, 32}, {64, 128}, {128, 64}, {256, 16}, {256, 32}, {512, 16}, {512, 8}};
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  int idx = -1;
  for (const Shape& s : shapes) {
    if (++idx != only && only >= 0) continue;
    const int C = s.C, H = s.H, wt = precise ? 2 : 1;
    const size_t M = (size_t)clips * H * H, K4 = 4 * (size_t)C;
    __half *A, *W1, *W2;
    float *b1, *db, *b2, *x;
    __half* dw;
    cudaMalloc(&A, M * C * 2); cudaMalloc(&W1, K4 * C * 2 * wt); cudaMalloc(&W2, K4 * C * 2 * wt);
    cudaMalloc(&b1, K4 * 4); cudaMalloc(&dw, 9 * K4 * 2 * wt); cudaMalloc(&db, K4 * 4); cudaMalloc(&b2, C * 4);
    cudaMalloc(&x, M * C * 4);
    auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
    fill_half<<<blocks(M * C), 256>>>(A, M * C, 1.0f, 1);
    fill_half<<<blocks(K4 * C * wt), 256>>>(W1, K4 * C * wt, 0.1f, 2);
    fill_half<<<blocks(K4 * C * wt), 256>>>(W2, K4 * C * wt, 0.05f, 3);
    fill_float<<<blocks(K4), 256>>>(b1, K4, 0.1f, 4);
    fill_half<<<blocks(9 * K4 * wt), 256>>>(dw, 9 * K4 * wt, 0.2f, 5);
    fill_float<<<blocks(K4), 256>>>(db, K4, 0.1f, 6);
    fill_float<<<blocks(C), 256>>>(b2, C, 0.1f, 7);
    cudaMemset(x, 0, M * C * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int rc = 0;
    for (int i = 0; i < 2 && !rc; ++i) rc = wmk::leff_block(A, W1, W2, b1, (const uint16_t*)dw, db, b2, x, clips, H, C, precise, 0);
    if (rc || cudaDeviceSynchronize() != cudaSuccess) {
      printf("C=%d H=%d: not run (%s / %s)\n", C, H, wmk_last_error(), cudaGetErrorString(cudaGetLastError()));
    } else {
      const int iters = 5;
      cudaEventRecord(e0);
      for (int i = 0; i < iters; ++i) wmk::leff_block(A, W1, W2, b1, (const uint16_t*)dw, db, b2, x, clips, H, C, precise, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double us = ms * 1e3 / iters;
      const double chunks_per_sm = (double)M / 128 * (K4 / 64) / sms;
      const double flops = 2.0 * M * C * K4 * 2;
      printf("C=%3d H=%3d precise=%d: %9.1f us   %7.0f cycles/chunk/SM @1.7GHz   %6.1f TFLOP/s   %6.0f GB/s (10C B/token)\n", C, H,
             precise, us, us * 1e-6 * 1.7e9 / chunks_per_sm, flops / us * 1e-6, (double)M * C * 10 / us * 1e-3);
    }
    cudaFree(A); cudaFree(W1); cudaFree(W2); cudaFree(b1); cudaFree(dw); cudaFree(db); cudaFree(b2); cudaFree(x);
  }
  return 0;
}

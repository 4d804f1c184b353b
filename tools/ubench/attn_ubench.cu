// Stand-alone timing of the fused q|k|v projection + window attention kernel (csrc/attn_block.cu) on the shapes of the
// bench step.  Build like leff_ubench.cu (links libwmk.so).   usage: attn_ubench [clips] [shape index]
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace wmk {
int attn_block(const void* A, const void* Wh, const float* bqkv, const uint16_t* bias, uint16_t* out, int n, int H, int C,
               int shift, cudaStream_t st);
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
  const int clips = argc > 1 ? atoi(argv[1]) : 384;
  const int only = argc > 2 ? atoi(argv[2]) : -1;
  struct Shape { int C, H; };
  const Shape shapes[] = {{32, 128}, {64, 64}, {128, 32}, {64, 128}, {128, 64}};
  int idx = -1;
  for (const Shape& s : shapes) {
    if (++idx != only && only >= 0) continue;
    const int C = s.C, H = s.H, NH = C / 32;
    const size_t M = (size_t)clips * H * H;
    __half *A, *W, *bias, *O;
    float* b;
    cudaMalloc(&A, M * C * 2); cudaMalloc(&W, (size_t)NH * 96 * C * 2); cudaMalloc(&bias, (size_t)NH * 4096 * 2);
    cudaMalloc(&O, M * C * 2); cudaMalloc(&b, NH * 96 * 4);
    auto blocks = [](size_t n) { return (unsigned)((n + 255) / 256); };
    fill_half<<<blocks(M * C), 256>>>(A, M * C, 1.0f, 1);
    fill_half<<<blocks((size_t)NH * 96 * C), 256>>>(W, (size_t)NH * 96 * C, 0.1f, 2);
    fill_half<<<blocks((size_t)NH * 4096), 256>>>(bias, (size_t)NH * 4096, 0.05f, 3);
    fill_float<<<blocks(NH * 96), 256>>>(b, NH * 96, 0.1f, 4);
    for (int shift = 0; shift <= 4; shift += 4) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      int rc = 0;
      for (int i = 0; i < 2 && !rc; ++i) rc = wmk::attn_block(A, W, b, (const uint16_t*)bias, (uint16_t*)O, clips, H, C, shift, 0);
      const cudaError_t ce = cudaDeviceSynchronize();
      if (rc || ce != cudaSuccess) {
        printf("C=%d H=%d shift=%d: failed (%s / %s)\n", C, H, shift, wmk_last_error(), cudaGetErrorString(ce));
        return 1;
      }
      const int iters = 5;
      cudaEventRecord(e0);
      for (int i = 0; i < iters; ++i) wmk::attn_block(A, W, b, (const uint16_t*)bias, (uint16_t*)O, clips, H, C, shift, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double us = ms * 1e3 / iters;
      printf("C=%3d H=%3d shift=%d: %9.1f us   %6.0f GB/s (4C B/token)   %6.1f TFLOP/s\n", C, H, shift, us, (double)M * C * 4 / us * 1e-3,
             (2.0 * M * C * 3 * C + 256.0 * C * M) / us * 1e-6);
    }
    cudaFree(A); cudaFree(W); cudaFree(bias); cudaFree(O); cudaFree(b);
  }
  return 0;
}

// Where the fused LeFF kernel's worker warps spend their cycles (needs csrc/leff_block.cu compiled with -DLB_TIMING and linked into a
// private libwmk_timing.so: nvcc ... -DLB_TIMING -c leff_block.cu; nvcc -shared -o tools/ubench/libwmk_timing.so <that .o> <the other csrc/*.o> -lcudart;
// nvcc -o tools/ubench/leff_timing tools/ubench/leff_timing.cu -L tools/ubench -lwmk_timing -Xlinker -rpath -Xlinker '$ORIGIN').
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace wmk {
int leff_block(const void* A, const void* W1, const void* W2, const float* b1, const uint16_t* dw16, const float* dw_b,
               const float* b2, float* x, int n, int H, int C, int precise, cudaStream_t st);
int leff_block_timing(unsigned long long* out);
}
int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 128, H = argc > 2 ? atoi(argv[2]) : 32, clips = 384;
  const size_t M = (size_t)clips * H * H, K4 = 4 * (size_t)C;
  void *A, *W1, *W2, *dw; float *b1, *db, *b2, *x;
  cudaMalloc(&A, M * C * 2); cudaMalloc(&W1, K4 * C * 2); cudaMalloc(&W2, K4 * C * 2); cudaMalloc(&dw, 9 * K4 * 2);
  cudaMalloc(&b1, K4 * 4); cudaMalloc(&db, K4 * 4); cudaMalloc(&b2, C * 4); cudaMalloc(&x, M * C * 4);
  cudaMemset(A, 0, M * C * 2); cudaMemset(W1, 0, K4 * C * 2); cudaMemset(W2, 0, K4 * C * 2); cudaMemset(dw, 0, 9 * K4 * 2);
  cudaMemset(b1, 0, K4 * 4); cudaMemset(db, 0, K4 * 4); cudaMemset(b2, 0, C * 4); cudaMemset(x, 0, M * C * 4);
  unsigned long long t[8];
  wmk::leff_block(A, W1, W2, b1, (const uint16_t*)dw, db, b2, x, clips, H, C, 0, 0);
  cudaDeviceSynchronize();
  wmk::leff_block_timing(t);
  wmk::leff_block(A, W1, W2, b1, (const uint16_t*)dw, db, b2, x, clips, H, C, 0, 0);
  cudaDeviceSynchronize();
  wmk::leff_block_timing(t);
  const double tot = (double)t[5];
  printf("C=%d H=%d  worker-warp cycles (sum over 16 warps of CTA 0): total %.0f\n", C, H, tot);
  const char* names[5] = {"acc2full (E waits linear2)", "acc1full (G waits linear1)", "hpempty (G waits conv of chunk-2)", "hpfull (conv waits GELU of chunk)", "a2empty (conv waits linear2)"};
  double w = 0;
  for (int i = 0; i < 5; ++i) { printf("  wait %-36s %5.1f %%\n", names[i], 100.0 * t[i] / tot); w += t[i]; }
  printf("  waiting in total %.1f %%, busy %.1f %%\n", 100 * w / tot, 100 - 100 * w / tot);
  return 0;
}

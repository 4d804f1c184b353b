// Host emulation of the 16 x 16 Cooley-Tukey 256-point DFT used by the training-time STFT kernel
// (csrc/dft256.cuh compiles as plain C++): pass A for every n2, pass B for every k1, in kernel order.
// Built by tests/test_dft256_host.py with g++; no GPU involved.
#include <math.h>
struct tw2 { float x, y; };
#include "../image-in-speech-watermarking_b200/csrc/dft256.cuh"

using namespace wmk::dft256;

extern "C" void host_fft256_frame(const float* frame /* [256] */, float* out /* [2][128] */) {
  tw2 tw[256];
  for (int m = 0; m < 256; ++m) {
    const double a = -2.0 * 3.14159265358979323846 * m / 256.0;
    tw[m].x = (float)cos(a);
    tw[m].y = (float)sin(a);
  }
  static c32 work[16 * 9];                 // Hermitian half of pass A, as the kernel stores it
  for (int n2 = 0; n2 < 16; ++n2) {
    float xs[16];
    for (int n1 = 0; n1 < 16; ++n1) xs[n1] = frame[16 * n1 + n2];
    c32 v[16];
    pass_a(xs, v);
    for (int k1 = 0; k1 <= 8; ++k1) work[9 * n2 + k1] = v[k1];
  }
  for (int k1 = 0; k1 < 16; ++k1) {
    c32 v[16];
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = work[9 * n2 + pass_b_src(k1)];
    pass_b(v, k1, tw);
    for (int k2 = 0; k2 < 8; ++k2) {
      out[k1 + 16 * k2] = v[k2].re;
      out[128 + k1 + 16 * k2] = v[k2].im;
    }
  }
}

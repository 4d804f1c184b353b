// Host emulation of the prime-factor 255-point DFT phases used by the STFT / ISTFT kernels
// (csrc/dft255.cuh compiles as plain C++): every (role, lane) pair is run sequentially, phases in
// kernel order.  Built by tests/test_dft255_host.py with g++; no GPU involved.
#include <vector>
#include "../image-in-speech-watermarking_b200/csrc/dft255.cuh"

using namespace wmk::dft255;

extern "C" void host_stft_tile(const float* samp, float* out /* [256][32] */) {
  Tables tb;
  build_tables(&tb);
  std::vector<float2> SA(SA_FLOAT2), R(R_FLOAT2);
  for (int n2 = 0; n2 < 17; ++n2)
    for (int f = 0; f < FT; ++f) fwd_stage_a(samp, SA.data(), n2, f);
  for (int k1 = 0; k1 < 8; ++k1)
    for (int f = 0; f < FT; ++f) {
      fwd_stage_b<0>(SA.data(), R.data(), k1, f);
      fwd_stage_b<1>(SA.data(), R.data(), k1, f);
    }
  for (int bin = 0; bin < BINS; ++bin)
    for (int f = 0; f < FT; ++f) {
      const float2 X = fwd_stage_c(R.data(), tb.fwd[bin], f);
      out[bin * FT + f] = X.x;
      out[(BINS + bin) * FT + f] = X.y;
    }
}

extern "C" void host_istft_tile(const float* XS /* [256][32] */, float* FR /* [32][255] */) {
  Tables tb;
  build_tables(&tb);
  std::vector<float2> R(R_FLOAT2);
  for (int k1 = 0; k1 < 8; ++k1)
    for (int f = 0; f < FT; ++f) {
      inv_stage_b<0>(XS, tb.inv, R.data(), k1, f);
      inv_stage_b<1>(XS, tb.inv, R.data(), k1, f);
    }
  for (int n2 = 0; n2 < 17; ++n2)
    for (int f = 0; f < FT; ++f) inv_stage_a(R.data(), FR, n2, f);
}

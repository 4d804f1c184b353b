// Host emulation of the prime-factor 255-point DFT phases used by the STFT / ISTFT kernels
// (csrc/dft255.cuh compiles as plain C++): every (role, lane) pair is run sequentially, phases in
// kernel order.  Built by tests/test_dft255_host.py with g++; no GPU involved.
#include <vector>
#include "../image-in-speech-watermarking_b200/csrc/dft255.cuh"

using namespace wmk::dft255;

extern "C" void host_stft_tile(const float* samp, float* out /* [256][32] */) {
  Tables tb;
  build_tables(&tb);
  std::vector<float2> SA(SA_FLOAT2);
  for (int n2 = 0; n2 < 17; ++n2)
    for (int f = 0; f < FT; ++f) fwd_stage_a(samp, SA.data(), n2, f);
  for (int i = 0; i < 256 * FT; ++i) out[i] = -12345.f;        // every bin must be stored exactly once
  for (int k1 = 0; k1 < 8; ++k1)
    for (int f = 0; f < FT; ++f)
    {
      auto store = [&](int bin, float re, float im) {
        if (out[bin * FT + f] != -12345.f) out[bin * FT + f] = 1e30f;    // duplicate store -> test fails
        else out[bin * FT + f] = re;
        out[(BINS + bin) * FT + f] = im;
      };
      if (k1 & 1) {                  // exercise both the one-warp and the two-warp (split) form
        fwd_stage_b<-1>(SA.data(), tb.fwd, k1, f, store);
      } else {
        fwd_stage_b<0>(SA.data(), tb.fwd, k1, f, store);
        fwd_stage_b<1>(SA.data(), tb.fwd, k1, f, store);
      }
    }
}

extern "C" void host_istft_tile(const float* XS /* [256][32] */, float* FR /* [32][255] */) {
  Tables tb;
  build_tables(&tb);
  std::vector<float2> ZS(SA_FLOAT2);
  for (int k1 = 0; k1 < 8; ++k1)
    for (int f = 0; f < FT; ++f) {
      if (k1 & 1) {
        inv_stage_b<-1>(XS, tb.inv, ZS.data(), k1, f);
      } else {
        inv_stage_b<0>(XS, tb.inv, ZS.data(), k1, f);
        inv_stage_b<1>(XS, tb.inv, ZS.data(), k1, f);
      }
    }
  for (int n2 = 0; n2 < 17; ++n2)
    for (int f = 0; f < FT; ++f) inv_stage_a(ZS.data(), FR, n2, f);
}

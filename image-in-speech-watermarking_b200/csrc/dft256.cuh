// 256-point DFT of the training-time STFT (torch.stft(x, n_fft=256, hop_length=128, win_length=256),
// uformerWM/audio_test.py:465-469) as 16 x 16 Cooley-Tukey: with n = 16 n1 + n2, k = k1 + 16 k2
//     X[k1 + 16 k2] = sum_n2 W16^(n2 k2) * W256^(n2 k1) * ( sum_n1 x[16 n1 + n2] W16^(n1 k1) )
// pass A = the inner 16-point DFTs (one per n2; real input, so only k1 = 0..8 are kept), pass B = the W256
// twiddle and the outer 16-point DFTs (one per k1); every 16-point DFT is two radix-4 stages in registers with immediate coefficients.
// Compiles as plain C++ too, so tests/test_dft256_host.py checks the math against numpy without a GPU.
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WMK256_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define WMK256_HD static inline
#endif

namespace wmk {
namespace dft256 {

struct c32 { float re, im; };
WMK256_HD c32 cadd(c32 a, c32 b) { c32 r; r.re = a.re + b.re; r.im = a.im + b.im; return r; }
WMK256_HD c32 csub(c32 a, c32 b) { c32 r; r.re = a.re - b.re; r.im = a.im - b.im; return r; }
WMK256_HD c32 cmul(c32 a, float wr, float wi) { c32 r; r.re = a.re * wr - a.im * wi; r.im = a.re * wi + a.im * wr; return r; }
WMK256_HD c32 mul_mi(c32 a) { c32 r; r.re = a.im; r.im = -a.re; return r; }      // a * (-i)

// forward 4-point DFT, in place: y_c = sum_a v_a e^{-2 pi i a c / 4}
WMK256_HD void dft4(c32& v0, c32& v1, c32& v2, c32& v3) {
  const c32 s02 = cadd(v0, v2), d02 = csub(v0, v2), s13 = cadd(v1, v3), d13 = mul_mi(csub(v1, v3));
  v0 = cadd(s02, s13);
  v1 = cadd(d02, d13);
  v2 = csub(s02, s13);
  v3 = csub(d02, d13);
}

// forward 16-point DFT in registers: v[n] -> v[k], n = 4a + b, k = c + 4d
WMK256_HD void dft16(c32 (&v)[16]) {
  const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, R = 0.70710678118654752f;
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);      // over a: t[b][c] lives in v[4c + b]
  // t[b][c] *= W16^(b c)
  v[4 * 1 + 1] = cmul(v[4 * 1 + 1], C1, -S1);     // bc = 1
  v[4 * 1 + 2] = cmul(v[4 * 1 + 2], R, -R);       // 2
  v[4 * 1 + 3] = cmul(v[4 * 1 + 3], S1, -C1);     // 3
  v[4 * 2 + 1] = cmul(v[4 * 2 + 1], R, -R);       // 2
  v[4 * 2 + 2] = mul_mi(v[4 * 2 + 2]);            // 4
  v[4 * 2 + 3] = cmul(v[4 * 2 + 3], -R, -R);      // 6
  v[4 * 3 + 1] = cmul(v[4 * 3 + 1], S1, -C1);     // 3
  v[4 * 3 + 2] = cmul(v[4 * 3 + 2], -R, -R);      // 6
  v[4 * 3 + 3] = cmul(v[4 * 3 + 3], -C1, S1);     // 9
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // over b: X[c + 4d] in v[4c + d]
  // reorder v[4c + d] -> v[c + 4d]
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) { const c32 t = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = t; }
}

// pass A for one (n2, frame): x[n1] = sample 16 n1 + n2 of the frame; out[k1] = DFT16(x)[k1].  The input is
// real, so out[16 - k1] = conj(out[k1]): callers keep k1 = 0..8 only.
WMK256_HD void pass_a(const float (&x)[16], c32 (&out)[16]) {
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) { out[n1].re = x[n1]; out[n1].im = 0.f; }
  dft16(out);
}

// where pass B finds A[n2][k1] among the stored k1 = 0..8, and whether it is the conjugate
WMK256_HD int pass_b_src(int k1) { return k1 <= 8 ? k1 : 16 - k1; }

// pass B for one (k1, frame): v[n2] = A[n2][pass_b_src(k1)] as stored; afterwards v[k2] = X[k1 + 16 k2].
// tw = e^{-2 pi i m / 256} table (cos, sin pairs).
template <typename TW>
WMK256_HD void pass_b(c32 (&v)[16], int k1, const TW* tw) {
  const float sgn = k1 <= 8 ? 1.f : -1.f;
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) {
    v[n2].im *= sgn;
    v[n2] = cmul(v[n2], tw[(n2 * k1) & 255].x, tw[(n2 * k1) & 255].y);
  }
  dft16(v);
}

}  // namespace dft256
}  // namespace wmk

// 255-point real DFT / inverse by the Good-Thomas prime-factor algorithm, 255 = 15 x 17.
// The STFT front end of the reference is torch.stft(x, n_fft=255) / torch.istft(..., n_fft=255)
// (uformerWM/audio_test.py:315-316,598-600,677-678; uformerWM/model.py:2458,2463): rectangular
// window, hop 63, one-sided 128 bins.  15 and 17 are coprime, so with the index maps
//     n = (17 n1 + 15 n2) mod 255,      k = (136 k1 + 120 k2) mod 255
// the transform factors into 17 real 15-point DFTs followed by 8 complex 17-point DFTs with NO
// twiddle factors in between; Hermitian symmetry of the real input leaves k1 = 0..7 only
// (k1 = 8..14 are the mirrored bins 255-k).  Each short DFT uses the even/odd split
//     X[k], X[N-k] = x0 + sum_n (x[n]+x[N-n]) cos(2 pi k n / N)  -/+  i sum_n (x[n]-x[N-n]) sin(2 pi k n / N)
// so every product is real x (real|complex) and all coefficients are compile-time immediates.
// ~4.5 k FMA-class instructions per frame instead of the 65 k of the direct matrix product.
//
// The phase functions below are written per (role, lane): in the kernels `f` (the frame inside a
// 32-frame tile) is always the lane, so every shared-memory access is conflict free and every
// role index (n2, k1, bin) is warp uniform.  They also compile as plain C++ (g++) so that the
// index maps can be unit-tested on the host (tests/test_dft255_host.py).
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WMK_HD __host__ __device__ __forceinline__
#else
#include <math.h>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#define WMK_HD static inline
#endif

namespace wmk {
namespace dft255 {

constexpr int NFFT = 255, HOP = 63, PAD = 127, BINS = 128;
constexpr int FT = 32;                        // frames per tile = lanes of a warp
constexpr int SA_FLOAT2 = 8 * 17 * FT;        // stage-A output / inverse stage-B output  [k1][n2][f]

WMK_HD constexpr float cos15(int m) {
  constexpr float t[15] = {1.000000000e+00f, 9.135454297e-01f, 6.691306233e-01f, 3.090170026e-01f, -1.045284644e-01f,
                           -5.000000000e-01f, -8.090170026e-01f, -9.781476259e-01f, -9.781476259e-01f, -8.090170026e-01f,
                           -5.000000000e-01f, -1.045284644e-01f, 3.090170026e-01f, 6.691306233e-01f, 9.135454297e-01f};
  return t[m];
}
WMK_HD constexpr float sin15(int m) {
  constexpr float t[15] = {0.000000000e+00f, 4.067366421e-01f, 7.431448102e-01f, 9.510565400e-01f, 9.945219159e-01f,
                           8.660253882e-01f, 5.877852440e-01f, 2.079116851e-01f, -2.079116851e-01f, -5.877852440e-01f,
                           -8.660253882e-01f, -9.945219159e-01f, -9.510565400e-01f, -7.431448102e-01f, -4.067366421e-01f};
  return t[m];
}
WMK_HD constexpr float cos17(int m) {
  constexpr float t[17] = {1.000000000e+00f, 9.324722290e-01f, 7.390089035e-01f, 4.457383454e-01f, 9.226836264e-02f,
                           -2.736629844e-01f, -6.026346087e-01f, -8.502171636e-01f, -9.829730988e-01f, -9.829730988e-01f,
                           -8.502171636e-01f, -6.026346087e-01f, -2.736629844e-01f, 9.226836264e-02f, 4.457383454e-01f,
                           7.390089035e-01f, 9.324722290e-01f};
  return t[m];
}
WMK_HD constexpr float sin17(int m) {
  constexpr float t[17] = {0.000000000e+00f, 3.612416685e-01f, 6.736956239e-01f, 8.951632977e-01f, 9.957341552e-01f,
                           9.618256688e-01f, 7.980172038e-01f, 5.264321566e-01f, 1.837495118e-01f, -1.837495118e-01f,
                           -5.264321566e-01f, -7.980172038e-01f, -9.618256688e-01f, -9.957341552e-01f, -8.951632977e-01f,
                           -6.736956239e-01f, -3.612416685e-01f};
  return t[m];
}

// Bin tables (built once on the host, read warp-uniformly from __constant__ memory).
//   fwd[k1*17+k2] : the one-sided bin that output (k1,k2) of the forward transform is and the sign of
//                   its imaginary part (bin < 0 for the mirrored half of the real k1 = 0 transform,
//                   which is not stored)
//   inv[k1*17+k2] : the one-sided bin that feeds Y[k1][k2] of the inverse and the sign of its
//                   imaginary part (-1: conjugate, 0: the ignored imaginary part of DC)
struct InvEntry { int bin; float im_sign; };
struct Tables {
  InvEntry fwd[8 * 17];
  InvEntry inv[8 * 17];
};

static inline void build_tables(Tables* t) {
  for (int k1 = 0; k1 < 8; ++k1)
    for (int k2 = 0; k2 < 17; ++k2) {
      const int k = (136 * k1 + 120 * k2) % 255;
      const int bin = k <= 127 ? k : 255 - k;
      const int conj = k > 127;
      t->inv[k1 * 17 + k2].bin = bin;
      t->inv[k1 * 17 + k2].im_sign = bin == 0 ? 0.f : (conj ? -1.f : 1.f);
      t->fwd[k1 * 17 + k2].bin = (k1 == 0 && k2 > 8) ? -1 : bin;
      t->fwd[k1 * 17 + k2].im_sign = conj ? -1.f : 1.f;
    }
}

// 17-point complex DFT by the even/odd split; emit(k, Y[k]) is called once for every k = 0..16.
//   A_k = y0 + sum_{n=1..8} (y[n]+y[17-n]) cos(2 pi k n/17),   B_k = sum_{n=1..8} (y[n]-y[17-n]) sin(2 pi k n/17)
//   forward (e^-):  Y[k] = A - iB,  Y[17-k] = A + iB;     inverse (e^+): the two swap.
template <bool INVERSE, class Emit>
WMK_HD void dft17(const float2 (&y)[17], Emit&& emit) {
  float2 e[9], o[9];
#pragma unroll
  for (int n = 1; n <= 8; ++n) {
    e[n] = make_float2(y[n].x + y[17 - n].x, y[n].y + y[17 - n].y);
    o[n] = make_float2(y[n].x - y[17 - n].x, y[n].y - y[17 - n].y);
  }
  float2 s = y[0];
#pragma unroll
  for (int n = 1; n <= 8; ++n) { s.x += e[n].x; s.y += e[n].y; }
  emit(0, s);
#pragma unroll
  for (int k = 1; k <= 8; ++k) {
    float2 a = y[0], b = make_float2(0.f, 0.f);
#pragma unroll
    for (int n = 1; n <= 8; ++n) {
      const float c = cos17((k * n) % 17), sn = sin17((k * n) % 17);
      a.x = fmaf(e[n].x, c, a.x);
      a.y = fmaf(e[n].y, c, a.y);
      b.x = fmaf(o[n].x, sn, b.x);
      b.y = fmaf(o[n].y, sn, b.y);
    }
    const float2 m = make_float2(a.x + b.y, a.y - b.x);     // A - iB
    const float2 q = make_float2(a.x - b.y, a.y + b.x);     // A + iB
    emit(INVERSE ? 17 - k : k, m);
    emit(INVERSE ? k : 17 - k, q);
  }
}

// Good-Thomas input map n = (17 n1 + 15 n2) mod 255 for the 15 samples of residue n2.  n2 is warp
// uniform; the switch turns the 15 offsets into immediates (one uniform branch instead of a
// compare/select chain per sample).
template <int N2>
WMK_HD void gather15(const float* base, float (&v)[15]) {
#pragma unroll
  for (int n1 = 0; n1 < 15; ++n1) v[n1] = base[(17 * n1 + 15 * N2) % NFFT];
}
template <int N2>
WMK_HD void scatter15(float* base, const float (&v)[15]) {
#pragma unroll
  for (int n1 = 0; n1 < 15; ++n1) base[(17 * n1 + 15 * N2) % NFFT] = v[n1];
}
#define WMK_DFT255_SWITCH17(n2, CALL)                                                                     \
  switch (n2) {                                                                                           \
    case 0: CALL(0); break;   case 1: CALL(1); break;   case 2: CALL(2); break;   case 3: CALL(3); break;   \
    case 4: CALL(4); break;   case 5: CALL(5); break;   case 6: CALL(6); break;   case 7: CALL(7); break;   \
    case 8: CALL(8); break;   case 9: CALL(9); break;   case 10: CALL(10); break; case 11: CALL(11); break; \
    case 12: CALL(12); break; case 13: CALL(13); break; case 14: CALL(14); break; case 15: CALL(15); break; \
    default: CALL(16); break;                                                                             \
  }

// ---- forward stage A: 15-point real DFT over n1 for one (frame f, residue n2) -> k1 = 0..7
// samp: the tile's padded samples (frame f starts at samp[63 f]); SA[k1][n2][f].
WMK_HD void fwd_stage_a(const float* samp, float2* SA, int n2, int f) {
  float v[15];
#define WMK_CALL(N) gather15<N>(samp + HOP * f, v)
  WMK_DFT255_SWITCH17(n2, WMK_CALL)
#undef WMK_CALL
  float e[8], o[8];
#pragma unroll
  for (int n = 1; n <= 7; ++n) { e[n] = v[n] + v[15 - n]; o[n] = v[n] - v[15 - n]; }
  float s0 = v[0];
#pragma unroll
  for (int n = 1; n <= 7; ++n) s0 += e[n];
  SA[(0 * 17 + n2) * FT + f] = make_float2(s0, 0.f);
#pragma unroll
  for (int k1 = 1; k1 <= 7; ++k1) {
    float re = v[0], im = 0.f;
#pragma unroll
    for (int n = 1; n <= 7; ++n) {
      re = fmaf(e[n], cos15((k1 * n) % 15), re);
      im = fmaf(o[n], -sin15((k1 * n) % 15), im);
    }
    SA[(k1 * 17 + n2) * FT + f] = make_float2(re, im);
  }
}

// ---- forward stage B: the 17-point DFT over n2 for (k1, frame f); store(bin, re, im) receives each
// one-sided bin exactly once over k1 = 0..7.
template <class Store>
WMK_HD void fwd_stage_b(const float2* SA, const InvEntry* fwd_tab, int k1, int f, Store&& store) {
  float2 y[17];
#pragma unroll
  for (int n2 = 0; n2 < 17; ++n2) y[n2] = SA[(k1 * 17 + n2) * FT + f];
  dft17<false>(y, [&](int k2, float2 X) {
    const InvEntry e = fwd_tab[k1 * 17 + k2];
    if (e.bin >= 0) store(e.bin, X.x, X.y * e.im_sign);
  });
}

// ---- inverse stage B': the inverse 17-point DFT over k2 for (k1, frame f) -> ZS[k1][n2][f].
// XS[row][f], row = reim*128 + bin: the one-sided spectrum tile; the imaginary part of DC is ignored
// as in a C2R transform.
WMK_HD void inv_stage_b(const float* XS, const InvEntry* inv_tab, float2* ZS, int k1, int f) {
  float2 y[17];
#pragma unroll
  for (int k2 = 0; k2 < 17; ++k2) {
    const InvEntry e = inv_tab[k1 * 17 + k2];
    y[k2] = make_float2(XS[e.bin * FT + f], XS[(BINS + e.bin) * FT + f] * e.im_sign);
  }
  dft17<true>(y, [&](int n2, float2 Z) { ZS[(k1 * 17 + n2) * FT + f] = Z; });
}

// ---- inverse stage A': complex-to-real inverse 15-point DFT over k1 for (n2, frame f); writes the
// 15 time samples n = (17 n1 + 15 n2) mod 255 of frame f (already divided by 255) to FR[f][n].
WMK_HD void inv_stage_a(const float2* ZS, float* FR, int n2, int f) {
  float zr[8], zi[8];
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    const float2 z = ZS[(k1 * 17 + n2) * FT + f];
    zr[k1] = z.x;
    zi[k1] = z.y;
  }
  constexpr float s1 = 1.0f / 255.0f, s2 = 2.0f / 255.0f;
  float v[15];
  float p0 = zr[0] * s1;
#pragma unroll
  for (int k = 1; k <= 7; ++k) p0 = fmaf(zr[k], s2, p0);
  v[0] = p0;
#pragma unroll
  for (int n1 = 1; n1 <= 7; ++n1) {
    float P = zr[0] * s1, Q = 0.f;
#pragma unroll
    for (int k = 1; k <= 7; ++k) {
      P = fmaf(zr[k], s2 * cos15((k * n1) % 15), P);
      Q = fmaf(zi[k], s2 * sin15((k * n1) % 15), Q);
    }
    v[n1] = P - Q;
    v[15 - n1] = P + Q;
  }
#define WMK_CALL(N) scatter15<N>(FR + f * NFFT, v)
  WMK_DFT255_SWITCH17(n2, WMK_CALL)
#undef WMK_CALL
}

}  // namespace dft255
}  // namespace wmk

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
// role index (n2, k1, part, bin) is warp uniform.  They also compile as plain C++ (g++) so that the
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
constexpr int SA_FLOAT2 = 8 * 17 * FT;        // stage-A output  [k1][n2][f]
constexpr int R_FLOAT2 = 2 * 8 * 9 * FT;      // half-DFT-17 sums [part][k1][k][f]

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
//   fwd[bin]      : which half-sums make one-sided bin `bin`:  k1 | kk<<3 | minus<<7 | conj<<8
//   inv[k1*17+k2] : which one-sided bin feeds Y[k1][k2]:       bin | conj<<7
struct Tables {
  unsigned short fwd[128];
  unsigned char inv[8 * 17];
};

static inline void build_tables(Tables* t) {
  for (int k1 = 0; k1 < 8; ++k1)
    for (int k2 = 0; k2 < 17; ++k2) {
      const int k = (136 * k1 + 120 * k2) % 255;
      const int bin = k <= 127 ? k : 255 - k;
      const int conj = k > 127;
      t->inv[k1 * 17 + k2] = (unsigned char)(bin | (conj << 7));
      if (k1 == 0 && k2 > 8) continue;          // the mirrored half of the real k1 = 0 transform
      const int kk = k2 <= 8 ? k2 : 17 - k2;
      const int minus = k2 > 8;
      t->fwd[bin] = (unsigned short)(k1 | (kk << 3) | (minus << 7) | (conj << 8));
    }
}

// One half of a 17-point complex DFT by the even/odd split.
//   PART 0: acc[0] = y0 + sum_n y[n];  acc[k] = y0 + sum_{n=1..8} (y[n]+y[17-n]) cos(2 pi k n/17)
//   PART 1: acc[0] = 0;                acc[k] =      sum_{n=1..8} (y[n]-y[17-n]) sin(2 pi k n/17)
// forward  (e^-):  Y[k] = A - iB = (A.x + B.y, A.y - B.x),  Y[17-k] = A + iB = (A.x - B.y, A.y + B.x)
// inverse  (e^+):  roles of k and 17-k swap.
template <int PART>
WMK_HD void dft17_half(const float2 (&y)[17], float2 (&acc)[9]) {
  float2 d[9];
#pragma unroll
  for (int n = 1; n <= 8; ++n) {
    if (PART == 0) d[n] = make_float2(y[n].x + y[17 - n].x, y[n].y + y[17 - n].y);
    else d[n] = make_float2(y[n].x - y[17 - n].x, y[n].y - y[17 - n].y);
  }
  if (PART == 0) {
    float2 s = y[0];
#pragma unroll
    for (int n = 1; n <= 8; ++n) { s.x += d[n].x; s.y += d[n].y; }
    acc[0] = s;
  } else {
    acc[0] = make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int k = 1; k <= 8; ++k) {
    float2 a = PART == 0 ? y[0] : make_float2(0.f, 0.f);
#pragma unroll
    for (int n = 1; n <= 8; ++n) {
      const float c = PART == 0 ? cos17((k * n) % 17) : sin17((k * n) % 17);
      a.x = fmaf(d[n].x, c, a.x);
      a.y = fmaf(d[n].y, c, a.y);
    }
    acc[k] = a;
  }
}

// ---- forward stage A: 15-point real DFT over n1 for one (frame f, residue n2) -> k1 = 0..7
// samp: the tile's padded samples (frame f starts at samp[63 f]); SA[k1][n2][f].
WMK_HD void fwd_stage_a(const float* samp, float2* SA, int n2, int f) {
  float v[15];
  int off = 15 * n2;
#pragma unroll
  for (int n1 = 0; n1 < 15; ++n1) {
    v[n1] = samp[HOP * f + off];
    off += 17;
    if (off >= NFFT) off -= NFFT;
  }
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

// ---- forward stage B: one half of the 17-point DFT over n2 for (part, k1, frame f)
template <int PART>
WMK_HD void fwd_stage_b(const float2* SA, float2* R, int k1, int f) {
  float2 y[17], acc[9];
#pragma unroll
  for (int n2 = 0; n2 < 17; ++n2) y[n2] = SA[(k1 * 17 + n2) * FT + f];
  dft17_half<PART>(y, acc);
#pragma unroll
  for (int k = 0; k <= 8; ++k) R[((PART * 8 + k1) * 9 + k) * FT + f] = acc[k];
}

// ---- forward stage C: combine the two halves into one-sided bin `bin` of frame f
WMK_HD float2 fwd_stage_c(const float2* R, unsigned short entry, int f) {
  const int k1 = entry & 7, kk = (entry >> 3) & 15;
  const float2 a = R[((0 * 8 + k1) * 9 + kk) * FT + f];
  const float2 b = R[((1 * 8 + k1) * 9 + kk) * FT + f];
  const float sg = (entry & 0x80) ? -1.f : 1.f;
  float2 X = make_float2(fmaf(sg, b.y, a.x), fmaf(-sg, b.x, a.y));
  if (entry & 0x100) X.y = -X.y;
  return X;
}

// ---- inverse stage B': half of the inverse 17-point DFT over k2 for (part, k1, frame f).
// XS[row][f], row = reim*128 + bin: the one-sided spectrum tile; the imaginary part of DC is ignored
// as in a C2R transform.
template <int PART>
WMK_HD void inv_stage_b(const float* XS, const unsigned char* inv_tab, float2* R, int k1, int f) {
  float2 y[17], acc[9];
#pragma unroll
  for (int k2 = 0; k2 < 17; ++k2) {
    const int e = inv_tab[k1 * 17 + k2];
    const int bin = e & 127;
    const float re = XS[bin * FT + f];
    float im = XS[(BINS + bin) * FT + f];
    if (e & 128) im = -im;
    if (bin == 0) im = 0.f;
    y[k2] = make_float2(re, im);
  }
  dft17_half<PART>(y, acc);
#pragma unroll
  for (int k = 0; k <= 8; ++k) R[((PART * 8 + k1) * 9 + k) * FT + f] = acc[k];
}

// ---- inverse stage A': complex-to-real inverse 15-point DFT over k1 for (n2, frame f); writes the
// 15 time samples n = (17 n1 + 15 n2) mod 255 of frame f (already divided by 255) to FR[f][n].
WMK_HD void inv_stage_a(const float2* R, float* FR, int n2, int f) {
  const int n = n2 <= 8 ? n2 : 17 - n2;
  const float sg = n2 <= 8 ? 1.f : -1.f;
  float zr[8], zi[8];
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) {
    const float2 a = R[((0 * 8 + k1) * 9 + n) * FT + f];
    const float2 b = R[((1 * 8 + k1) * 9 + n) * FT + f];
    zr[k1] = fmaf(-sg, b.y, a.x);
    zi[k1] = fmaf(sg, b.x, a.y);
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
  int off = 15 * n2;
#pragma unroll
  for (int n1 = 0; n1 < 15; ++n1) {
    FR[f * NFFT + off] = v[n1];
    off += 17;
    if (off >= NFFT) off -= NFFT;
  }
}

}  // namespace dft255
}  // namespace wmk

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

// float2 arithmetic: on sm_100 these are single packed instructions (FADD2 / FFMA2, the constant pair
// lives in a uniform register), on the host plain scalar code.
WMK_HD float2 add2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
WMK_HD float2 sub2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// acc + x * c  (c real)
WMK_HD float2 fma2(float2 x, float c, float2 acc) {
#ifdef __CUDA_ARCH__
  return __ffma2_rn(x, make_float2(c, c), acc);
#else
  return make_float2(fmaf(x.x, c, acc.x), fmaf(x.y, c, acc.y));
#endif
}
WMK_HD float2 mul2(float2 x, float c) {
#ifdef __CUDA_ARCH__
  return __fmul2_rn(x, make_float2(c, c));
#else
  return make_float2(x.x * c, x.y * c);
#endif
}


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
// HALF: -1 = all outputs; 0 = k in {0, 1..4, 13..16}; 1 = k in {5..12}: lets two warps share one transform
// (each recomputes the 16 butterflies, the 256 multiply-adds are split).
template <bool INVERSE, int HALF, class Emit>
WMK_HD void dft17(const float2 (&y)[17], Emit&& emit) {
  constexpr int KLO = HALF == 1 ? 5 : 1, KHI = HALF == 0 ? 4 : 8;
  // The inverse kernel runs this with packed float2 instructions (FADD2 / FFMA2: measured +5 %), the forward
  // kernel with scalar ones (packed measured -7 % there: its stage B is FMA-pipe bound, not issue bound).
  float2 e[9], o[9];
#pragma unroll
  for (int n = 1; n <= 8; ++n) {
    if (INVERSE) {
      e[n] = add2(y[n], y[17 - n]);
      o[n] = sub2(y[n], y[17 - n]);
    } else {
      e[n] = make_float2(y[n].x + y[17 - n].x, y[n].y + y[17 - n].y);
      o[n] = make_float2(y[n].x - y[17 - n].x, y[n].y - y[17 - n].y);
    }
  }
  float2 s = y[0];
#pragma unroll
  for (int n = 1; n <= 8; ++n) {
    if (INVERSE) s = add2(s, e[n]);
    else { s.x += e[n].x; s.y += e[n].y; }
  }
  if (HALF != 1) emit(0, s);
#pragma unroll
  for (int k = KLO; k <= KHI; ++k) {
    float2 a = y[0], b = make_float2(0.f, 0.f);
#pragma unroll
    for (int n = 1; n <= 8; ++n) {
      const float c = cos17((k * n) % 17), sn = sin17((k * n) % 17);
      if (INVERSE) {
        a = fma2(e[n], c, a);
        b = fma2(o[n], sn, b);
      } else {
        a.x = fmaf(e[n].x, c, a.x);
        a.y = fmaf(e[n].y, c, a.y);
        b.x = fmaf(o[n].x, sn, b.x);
        b.y = fmaf(o[n].y, sn, b.y);
      }
    }
    if (INVERSE) {
      const float2 ib = make_float2(b.y, -b.x);            // -i B
      emit(17 - k, add2(a, ib));                           // A - iB
      emit(k, sub2(a, ib));                                // A + iB
    } else {
      emit(k, make_float2(a.x + b.y, a.y - b.x));          // A - iB
      emit(17 - k, make_float2(a.x - b.y, a.y + b.x));     // A + iB
    }
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

// 15-point transforms by a second Good-Thomas split, 15 = 3 x 5: n1 = (5a + 3b) mod 15,
// k1 = (10 ka + 6 kb) mod 15, again twiddle free: five 3-point DFTs, then one real and one complex
// 5-point DFT (the ka = 2 column is the conjugate mirror of ka = 1).  ~75 flops instead of 119.
constexpr float kC5_1 = 3.090169944e-01f, kC5_2 = -8.090169944e-01f;     // cos(2 pi/5), cos(4 pi/5)
constexpr float kS5_1 = 9.510565163e-01f, kS5_2 = 5.877852523e-01f;      // sin(2 pi/5), sin(4 pi/5)
constexpr float kS3 = 8.660254038e-01f;                                  // sin(2 pi/3)

// real v[n1] -> X[k1], k1 = 0..7 (forward, e^-)
WMK_HD void dft15_real(const float (&v)[15], float2 (&X)[8]) {
  float u0[5];
  float2 u1[5];
#pragma unroll
  for (int b = 0; b < 5; ++b) {
    const float x0 = v[(3 * b) % 15], x1 = v[(5 + 3 * b) % 15], x2 = v[(10 + 3 * b) % 15];
    const float t = x1 + x2;
    u0[b] = x0 + t;
    u1[b] = make_float2(fmaf(-0.5f, t, x0), kS3 * (x2 - x1));
  }
  {  // ka = 0: real 5-point DFT, kb = 0..2 -> k1 = 0, 6, 12 (= conj of 3)
    const float e1 = u0[1] + u0[4], e2 = u0[2] + u0[3], o1 = u0[1] - u0[4], o2 = u0[2] - u0[3];
    X[0] = make_float2(u0[0] + e1 + e2, 0.f);
    X[6] = make_float2(fmaf(kC5_2, e2, fmaf(kC5_1, e1, u0[0])), -fmaf(kS5_2, o2, kS5_1 * o1));
    X[3] = make_float2(fmaf(kC5_1, e2, fmaf(kC5_2, e1, u0[0])), fmaf(-kS5_1, o2, kS5_2 * o1));      // conj(U0[2])
  }
  {  // ka = 1: complex 5-point DFT, kb = 0..4 -> k1 = 10 (conj of 5), 1, 7, 13 (conj of 2), 4
    const float2 e1 = make_float2(u1[1].x + u1[4].x, u1[1].y + u1[4].y), e2 = make_float2(u1[2].x + u1[3].x, u1[2].y + u1[3].y);
    const float2 o1 = make_float2(u1[1].x - u1[4].x, u1[1].y - u1[4].y), o2 = make_float2(u1[2].x - u1[3].x, u1[2].y - u1[3].y);
    X[5] = make_float2(u1[0].x + e1.x + e2.x, -(u1[0].y + e1.y + e2.y));                            // conj(U1[0])
    const float2 a1 = make_float2(fmaf(kC5_2, e2.x, fmaf(kC5_1, e1.x, u1[0].x)), fmaf(kC5_2, e2.y, fmaf(kC5_1, e1.y, u1[0].y)));
    const float2 a2 = make_float2(fmaf(kC5_1, e2.x, fmaf(kC5_2, e1.x, u1[0].x)), fmaf(kC5_1, e2.y, fmaf(kC5_2, e1.y, u1[0].y)));
    const float2 b1 = make_float2(fmaf(kS5_2, o2.x, kS5_1 * o1.x), fmaf(kS5_2, o2.y, kS5_1 * o1.y));
    const float2 b2 = make_float2(fmaf(-kS5_1, o2.x, kS5_2 * o1.x), fmaf(-kS5_1, o2.y, kS5_2 * o1.y));
    X[1] = make_float2(a1.x + b1.y, a1.y - b1.x);                     // U1[1] = A1 - iB1
    X[4] = make_float2(a1.x - b1.y, a1.y + b1.x);                     // U1[4] = A1 + iB1
    X[7] = make_float2(a2.x + b2.y, a2.y - b2.x);                     // U1[2] = A2 - iB2
    X[2] = make_float2(a2.x - b2.y, -(a2.y + b2.x));                  // conj(U1[3]) = conj(A2 + iB2)
  }
}

// Hermitian z[k1] (k1 = 0..7, z[15-k] = conj z[k], Im z[0] ignored) -> real v[n1] = sum_k z[k] e^{+2 pi i n1 k/15} * scale
WMK_HD void dft15_c2r(const float2 (&z)[8], float scale, float (&v)[15]) {
  float u0[5];
  float2 u1[5];
  {  // ka = 0 column: U0[0] = z0, U0[1] = z6, U0[2] = conj z3  -> real 5-point inverse
    const float x1 = z[6].x, y1 = z[6].y, x2 = z[3].x, y2 = -z[3].y;
    u0[0] = fmaf(2.f, x1 + x2, z[0].x);
    const float p1 = fmaf(2.f * kC5_2, x2, fmaf(2.f * kC5_1, x1, z[0].x)), q1 = fmaf(2.f * kS5_2, y2, 2.f * kS5_1 * y1);
    const float p2 = fmaf(2.f * kC5_1, x2, fmaf(2.f * kC5_2, x1, z[0].x)), q2 = fmaf(-2.f * kS5_1, y2, 2.f * kS5_2 * y1);
    u0[1] = p1 - q1; u0[4] = p1 + q1;
    u0[2] = p2 - q2; u0[3] = p2 + q2;
  }
  {  // ka = 1 column: U1[0] = conj z5, U1[1] = z1, U1[2] = z7, U1[3] = conj z2, U1[4] = z4 -> complex 5-point inverse
    const float2 w0 = make_float2(z[5].x, -z[5].y), w1 = z[1], w2 = z[7], w3 = make_float2(z[2].x, -z[2].y), w4 = z[4];
    const float2 e1 = add2(w1, w4), e2 = add2(w2, w3);
    const float2 o1 = sub2(w1, w4), o2 = sub2(w2, w3);
    u1[0] = add2(add2(w0, e1), e2);
    const float2 a1 = fma2(e2, kC5_2, fma2(e1, kC5_1, w0));
    const float2 a2 = fma2(e2, kC5_1, fma2(e1, kC5_2, w0));
    const float2 b1 = fma2(o2, kS5_2, mul2(o1, kS5_1));
    const float2 b2 = fma2(o2, -kS5_1, mul2(o1, kS5_2));
    u1[1] = make_float2(a1.x - b1.y, a1.y + b1.x);                    // A1 + iB1
    u1[4] = make_float2(a1.x + b1.y, a1.y - b1.x);                    // A1 - iB1
    u1[2] = make_float2(a2.x - b2.y, a2.y + b2.x);
    u1[3] = make_float2(a2.x + b2.y, a2.y - b2.x);
  }
  const float r3 = 2.f * kS3 * scale;
#pragma unroll
  for (int b = 0; b < 5; ++b) {       // 3-point inverse over ka: u0 + 2 Re(u1 e^{+2 pi i a/3})
    const float base = u0[b] * scale;
    const float t = fmaf(-scale, u1[b].x, base);
    v[(3 * b) % 15] = fmaf(2.f * scale, u1[b].x, base);
    v[(5 + 3 * b) % 15] = fmaf(-r3, u1[b].y, t);
    v[(10 + 3 * b) % 15] = fmaf(r3, u1[b].y, t);
  }
}

// ---- forward stage A: 15-point real DFT over n1 for one (frame f, residue n2) -> k1 = 0..7
// samp: the tile's padded samples (frame f starts at samp[63 f]); SA[k1][n2][f].
WMK_HD void fwd_stage_a(const float* samp, float2* SA, int n2, int f) {
  float v[15];
#define WMK_CALL(N) gather15<N>(samp + HOP * f, v)
  WMK_DFT255_SWITCH17(n2, WMK_CALL)
#undef WMK_CALL
  float2 X[8];
  dft15_real(v, X);
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) SA[(k1 * 17 + n2) * FT + f] = X[k1];
}

// ---- forward stage B: the 17-point DFT over n2 for (k1, frame f); store(bin, re, im) receives each
// one-sided bin exactly once over k1 = 0..7.
template <int HALF, class Store>
WMK_HD void fwd_stage_b(const float2* SA, const InvEntry* fwd_tab, int k1, int f, Store&& store) {
  float2 y[17];
#pragma unroll
  for (int n2 = 0; n2 < 17; ++n2) y[n2] = SA[(k1 * 17 + n2) * FT + f];
  dft17<false, HALF>(y, [&](int k2, float2 X) {
    const InvEntry e = fwd_tab[k1 * 17 + k2];
    if (e.bin >= 0) store(e.bin, X.x, X.y * e.im_sign);
  });
}

// ---- inverse stage B': the inverse 17-point DFT over k2 for (k1, frame f) -> ZS[k1][n2][f].
// XS[row][f], row = reim*128 + bin: the one-sided spectrum tile; the imaginary part of DC is ignored
// as in a C2R transform.
template <int HALF>
WMK_HD void inv_stage_b(const float* XS, const InvEntry* inv_tab, float2* ZS, int k1, int f) {
  float2 y[17];
#pragma unroll
  for (int k2 = 0; k2 < 17; ++k2) {
    const InvEntry e = inv_tab[k1 * 17 + k2];
    y[k2] = make_float2(XS[e.bin * FT + f], XS[(BINS + e.bin) * FT + f] * e.im_sign);
  }
  dft17<true, HALF>(y, [&](int n2, float2 Z) { ZS[(k1 * 17 + n2) * FT + f] = Z; });
}

// ---- inverse stage A': complex-to-real inverse 15-point DFT over k1 for (n2, frame f); writes the
// 15 time samples n = (17 n1 + 15 n2) mod 255 of frame f (already divided by 255) to FR[f][n].
WMK_HD void inv_stage_a(const float2* ZS, float* FR, int n2, int f) {
  float2 z[8];
#pragma unroll
  for (int k1 = 0; k1 < 8; ++k1) z[k1] = ZS[(k1 * 17 + n2) * FT + f];
  float v[15];
  dft15_c2r(z, 1.0f / 255.0f, v);
#define WMK_CALL(N) scatter15<N>(FR + f * NFFT, v)
  WMK_DFT255_SWITCH17(n2, WMK_CALL)
#undef WMK_CALL
}

}  // namespace dft255
}  // namespace wmk

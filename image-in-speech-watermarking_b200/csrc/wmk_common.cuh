// Shared declarations of libwmk.so (internal; the public ABI is include/wmk.h).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/wmk.h"

namespace wmk {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define WMK_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::wmk::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                 \
                       cudaGetErrorString(_e));                                           \
      return WMK_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define WMK_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    ::wmk::count_launch();                                                                \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      ::wmk::set_error("%s:%d: launch of %s failed: %s", __FILE__, __LINE__, name,        \
                       cudaGetErrorString(_e));                                           \
      return WMK_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define WMK_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::wmk::set_error(__VA_ARGS__);                                                      \
      return WMK_ERR_ARG;                                                                 \
    }                                                                                     \
  } while (0)

#define WMK_TRY(expr)                                                                     \
  do {                                                                                    \
    int _s = (expr);                                                                      \
    if (_s != 0) return _s;                                                               \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Optional per-kernel-family timing with CUDA events on the launching stream (bench.py's live
// roofline numbers; wmk_profile_* in wmk.h).  When disabled a scope costs one branch.
// ---------------------------------------------------------------------------------------------
enum Family {
  FAM_GEMM = 0, FAM_ATTENTION, FAM_LAYERNORM, FAM_DWCONV, FAM_LAYOUT, FAM_SMALL, FAM_STFT, FAM_ISTFT,
  FAM_ATTACK, FAM_STATS, FAM_GEMM_HBM, FAM_COUNT
};
struct ProfScope {
  ProfScope(int family, double work, cudaStream_t st, double work2 = 0.0);
  ~ProfScope();
  int slot;
  cudaStream_t st;
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Dense layer (GEMM) interface shared by the fp32 SIMT kernel and the bf16 tcgen05 kernel.
//   C[m][n] = epilogue( sum_k A[m][k] * W[n][k] + bias[n] )
// A is [M][K] and W is [N][K], both K-contiguous ("K-major"): PyTorch's nn.Linear weight layout.
// ---------------------------------------------------------------------------------------------
enum Epilogue {
  EPI_BIAS = 0,       // C = acc + bias
  EPI_BIAS_GELU = 1,  // C = gelu(acc + bias)                      (LeFF linear1, model.py:686-687)
  EPI_BIAS_RESID = 2, // C = resid + acc + bias  (fp32 out; resid may alias C)  (model.py:1016-1017)
  EPI_UPSAMPLE = 3,   // ConvTranspose2d(k=2,s=2) pixel shuffle into the concat buffer (model.py:794-800,1225)
  EPI_BIAS_RELU = 4   // C = max(acc + bias, 0), bf16 out  (ConvBNRelu with the BatchNorm affine folded, hidden/model/conv_bn_relu.py:7-18)
};

struct GemmArgs {
  const void* A = nullptr;      // [M][K]   float (fp32 mode) or __nv_bfloat16 (bf16 mode)
  const void* W = nullptr;      // [N][K]   same type as A
  const float* bias = nullptr;  // [N] or nullptr
  const float* resid = nullptr; // [M][ldc] fp32 (EPI_BIAS_RESID)
  void* C = nullptr;            // [M][ldc] float or bf16 (out_bf16)
  int M = 0, N = 0, K = 0;
  int ldc = 0;                  // row stride of C / resid in elements
  int epi = EPI_BIAS;
  int out_bf16 = 0;
  int gelu_half = 0;            // EPI_BIAS_GELU: W and bias were pre-multiplied by 0.5 (gelu_tanh2_half_arg)
  // EPI_UPSAMPLE: A rows are (b, h, w) over an up_h x up_w grid; columns are (i, j, co) with
  // co < up_cout; element goes to token (b, 2h+i, 2w+j), channel co of a [.., ldc] buffer.
  int up_h = 0, up_w = 0, up_cout = 0;
  // Fused LayerNorm of the output row (EPI_BIAS_RESID, bf16 kernel, N in {32, 64, 128} = one tile per row):
  // ln_out[m][:] = bf16( LN(C[m][:]) * gamma + beta (+ modulator[window position of token m]) ), i.e. the
  // operand of the NEXT dense layer (uformerWM/model.py:982,996-999,1017) without re-reading C.
  void* ln_out = nullptr;          // [M][N] bf16
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  const float* ln_mod = nullptr;   // [64][N] or nullptr
  int ln_H = 0, ln_shift = 0;      // image side (power of two) and cyclic shift of the block that consumes ln_out
  int ln_split = 0;                // ln_out rows are split-bf16 [hi(N) | lo(N)] (the A operand of a split dense layer)
  // Implicit-GEMM 3x3 convolution (pad 1) over an NHWC bf16 image batch [conv_B][conv_H][128][64]: A is that tensor, a row
  // tile = the 128 pixels of one image row, k-block kb = tap (dy, dx) fetched as the TMA box shifted by (dy-1, dx-1)
  // with out-of-bounds zero fill as the padding; W is [N][9*64] with k = tap*64 + ci.  M = conv_B*conv_H*128, K = 576.
  int conv_H = 0, conv_B = 0;
  // Implicit-GEMM Downsample = Conv2d(C, 2C, k=4, s=2, p=1) (uformerWM/model.py:763,768-775): A is the space-to-depth
  // tensor S[dn_B][dn_Ho + 1][dn_Ho + 1][4C] of s2d_pad_kernel (split operands: cells [hi(4C) | lo(4C)]); a row tile =
  // 128 output pixels (whole rows of one image, or whole images when an image has 64 pixels), k-block kb = tap
  // (a, b) = kb / (4C/64) x 64-channel chunk, fetched as the TMA box shifted by (a, b) cells.  W is [N][16C] (split:
  // [hi | lo]) with k = ((a*2 + b)*4 + ph*2 + pw)*C + ci for kernel position (2a + ph, 2b + pw).  M = dn_B*dn_Ho^2, K = 16C.
  int dn_Ho = 0, dn_B = 0;
  // Split-bf16 ("bf16x3") operands: A is [M][2K] bf16, row = [hi(K) | lo(K)] with hi = bf16(v), lo = bf16(v - hi);
  // W is the packed split weight of pack_split_weight() ([N][2K] = [hi | lo]; K == 32: [N][128] = [hi | hi | lo | 0]).
  // The product is hi*hi + lo*hi + hi*lo, accumulated by three tcgen05 MMAs per k-step into the same fp32 TMEM
  // accumulator.  Outputs are fp32 (out_bf16 = 0).  gelu_exact: GELU by the 1.5e-7 erf form (gelu_fast).
  int split = 0;
  // W-only split ("fp16x2"): A is a plain fp16 [M][K] operand, W is [N][2K] fp16 = [hi(K) | lo(K)] with hi = fp16(w),
  // lo = fp16(w - hi); the product is A*hi + A*lo (two MMAs per k-step).  Weight rounding is what moves the extractor's
  // logits (2.5e-4 for fp16 weights vs 2e-5 for fp16 LeFF activations, tools/precision_study.py), so the LeFF layers
  // of the precise extractor keep 22-bit weights and 16-bit activations.  Implies f16 operands.
  int wsplit = 0;
  int gelu_exact = 0;
  // Plain (non-split) 16-bit operands - A, W, a 16-bit C (out_bf16) and ln_out - are IEEE fp16 instead of bf16:
  // same tcgen05 rate, 11 instead of 8 mantissa bits (the WMK_PREC_MIXED / WMK_PREC_F16 embedder).
  int f16 = 0;
};

// Roofline class of a dense-layer launch: algorithmic bytes (A + W + C, + the fp32 residual) against
// 2 M N K FLOPs; below the ridge point of the B200 (1401.6 TFLOP/s / 6548.8 GB/s = 214 FLOP/B) the
// launch is HBM bound.  `es` = operand element size (2 = bf16, 4 = fp32).
struct GemmWork { int family; double work, work2; };
static inline GemmWork gemm_work(const GemmArgs& g, int es) {
  if (g.split) es = 4;                   // hi + lo
  const double flops = 2.0 * g.M * g.N * g.K;
  // implicit-GEMM downsample: the space-to-depth tensor holds every input element once (K / 4 values per output pixel)
  const double a_bytes = (double)g.M * (g.dn_Ho > 0 ? g.K / 4 : g.K) * es;
  const double bytes = a_bytes + (double)g.N * g.K * es + (double)g.M * g.N * (g.out_bf16 ? 2 : 4) +
                       (g.epi == EPI_BIAS_RESID ? (double)g.M * g.N * 4 : 0.0);
  if (flops / bytes >= 214.0) return {FAM_GEMM, flops, bytes};
  return {FAM_GEMM_HBM, bytes, flops};
}

int gemm_fp32_simt(const GemmArgs& g, cudaStream_t st);
int gemm_bf16_tcgen05(const GemmArgs& g, cudaStream_t st);

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// Same function with erf from Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7, far below the bf16
// rounding of the value it feeds): one MUFU.RCP + one MUFU.EX2 + 8 FMA instead of erff's ~40
// instructions.  Used by every bf16-mode kernel; the fp32 parity mode keeps erff.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, __expf(-z * z), 1.0f);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}

// Exact-class GELU for PAIRS on the packed fp32x2 pipe: x * Phi(x) with Phi(x) = 1 / (1 + 2^(-L(x))), L = log2 of the
// odds Phi(x) / Phi(-x) - an odd function, fitted here by a degree-13 odd polynomial (weighted minimax on [0, 6] against
// scipy's log_ndtr, weight = the GELU's sensitivity x Phi (1 - Phi) ln 2: formula error 6.7e-8; monotone beyond the fit
// range, so large |x| saturate to x and -0 without a clamp).  10 packed FMA-pipe instructions + 2 MUFU.EX2 + 2 MUFU.RCP
// per pair (ex2.approx 2 ulp, rcp.approx 1 ulp) against 16 + 4 for the Abramowitz-Stegun 7.1.26 erf form this replaces
// (gelu_fast above keeps that form for single values); measured max |error| in fp32 8e-7 at |x| ~ 5 (the rounding of the
// result itself), 1.7e-7 for |x| < 1 - the same as the erf form.  Used by the precise extractor (GELU error must stay
// well under the 2^-11 of the fp16 tensors it feeds, and must not be systematic like the tanh form's 5e-5).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  // coefficients of -L(x) / x in x^2 (negated: the exponent below is -L)
  float2 p = __ffma2_rn(x2, make_float2(-5.209925380000868e-09f, -5.209925380000868e-09f), make_float2(3.850457233056659e-07f, 3.850457233056659e-07f));
  p = __ffma2_rn(p, x2, make_float2(-1.1452440958237275e-05f, -1.1452440958237275e-05f));
  p = __ffma2_rn(p, x2, make_float2(1.5938949945848435e-04f, 1.5938949945848435e-04f));
  p = __ffma2_rn(p, x2, make_float2(9.559268073644489e-05f, 9.559268073644489e-05f));
  p = __ffma2_rn(p, x2, make_float2(-0.10483857989311218f, -0.10483857989311218f));
  p = __ffma2_rn(p, x2, make_float2(-2.3022072315216064f, -2.3022072315216064f));
  const float2 a = __fmul2_rn(p, x);                                       // -L(x)
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
  const float2 d = __fadd2_rn(e, make_float2(1.0f, 1.0f));
  float2 r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
  return __fmul2_rn(x, r);
}

// GELU for pairs on the packed fp32x2 pipe (sm_100 FFMA2), MUFU-free: erf(z) on [-3,3] by an odd
// degree-15 minimax polynomial (|error| < 8.1e-5, fitted against scipy.special.erf), argument
// clamped to +-3 beyond.  |gelu error| < 2e-4 for |x| <= 4 (4e-5 * |x| beyond), i.e. below the
// bf16 rounding of the value it produces; used only where the result is stored as bf16.
__device__ __forceinline__ float2 gelu_poly2(float2 x) {
  float2 z = __fmul2_rn(x, make_float2(0.70710678118654752440f, 0.70710678118654752440f));
  z.x = fminf(fmaxf(z.x, -3.0f), 3.0f);
  z.y = fminf(fmaxf(z.y, -3.0f), 3.0f);
  const float2 z2 = __fmul2_rn(z, z);
  float2 p = make_float2(-4.055369516e-07f, -4.055369516e-07f);
  p = __ffma2_rn(p, z2, make_float2(1.715986036e-05f, 1.715986036e-05f));
  p = __ffma2_rn(p, z2, make_float2(-3.145957307e-04f, -3.145957307e-04f));
  p = __ffma2_rn(p, z2, make_float2(3.318712581e-03f, 3.318712581e-03f));
  p = __ffma2_rn(p, z2, make_float2(-2.268580347e-02f, -2.268580347e-02f));
  p = __ffma2_rn(p, z2, make_float2(1.077178344e-01f, 1.077178344e-01f));
  p = __ffma2_rn(p, z2, make_float2(-3.732314110e-01f, -3.732314110e-01f));
  p = __ffma2_rn(p, z2, make_float2(1.127895713e+00f, 1.127895713e+00f));
  const float2 e = __fmul2_rn(p, z);                 // ~erf(x / sqrt 2), odd in x
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(h, e, h);
}

// GELU for pairs: x * Phi(x) with Phi(x) = 0.5 (1 + tanh(a x + b x^3 + c x^5)); a, b, c fitted to
// the exact erf form (max |error| of the formula 5.4e-5 on [-8, 8], vs 4.7e-4 for the textbook
// two-term constants) and tanh from MUFU.TANH (relative error 2^-11).  Three packed issue slots +
// one MUFU per element; used where the result is stored as bf16 (whose own rounding is 4e-3).
__device__ __forceinline__ float2 gelu_tanh2(float2 x) {
  const float2 x2 = __fmul2_rn(x, x);
  float2 p = __ffma2_rn(x2, make_float2(-3.81889112e-04f, -3.81889112e-04f), make_float2(3.72153111e-02f, 3.72153111e-02f));
  p = __ffma2_rn(p, x2, make_float2(7.97237410e-01f, 7.97237410e-01f));
  const float2 u = __fmul2_rn(p, x);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 h = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  return __ffma2_rn(h, t, h);
}

// The same GELU for an argument that arrives HALVED (h = x / 2): the producer folds the 0.5 of
// x * 0.5 * (1 + tanh(u)) into its weights and bias (an exact power-of-two scaling), the polynomial takes
// the matching power-of-two multiples of its coefficients, and one packed multiply per pair disappears.
// Bit-identical to gelu_tanh2(2 h) outside the denormal range.
__device__ __forceinline__ float2 gelu_tanh2_half_arg(float2 h) {
  const float2 h2 = __fmul2_rn(h, h);
  float2 p = __ffma2_rn(h2, make_float2(32.f * -3.81889112e-04f, 32.f * -3.81889112e-04f),
                        make_float2(8.f * 3.72153111e-02f, 8.f * 3.72153111e-02f));
  p = __ffma2_rn(p, h2, make_float2(2.f * 7.97237410e-01f, 2.f * 7.97237410e-01f));
  const float2 u = __fmul2_rn(p, h);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  return __ffma2_rn(h, t, h);
}

// MUFU-free GELU for a HALVED argument (h = x / 2): gelu_poly2's degree-15 erf polynomial on z = x / sqrt(2) = h sqrt(2).
// Used for every other pair of the fp16 plan's linear1 epilogue (WMK_GELU_MIX=1) to move work from the XU pipe (MUFU.TANH +
// the fp32 -> fp16 packs) to the FMA pipe.
__device__ __forceinline__ float2 gelu_poly2_half_arg(float2 h) {
  float2 z = __fmul2_rn(h, make_float2(1.41421356237309505f, 1.41421356237309505f));
  z.x = fminf(fmaxf(z.x, -3.0f), 3.0f);
  z.y = fminf(fmaxf(z.y, -3.0f), 3.0f);
  const float2 z2 = __fmul2_rn(z, z);
  float2 p = make_float2(-4.055369516e-07f, -4.055369516e-07f);
  p = __ffma2_rn(p, z2, make_float2(1.715986036e-05f, 1.715986036e-05f));
  p = __ffma2_rn(p, z2, make_float2(-3.145957307e-04f, -3.145957307e-04f));
  p = __ffma2_rn(p, z2, make_float2(3.318712581e-03f, 3.318712581e-03f));
  p = __ffma2_rn(p, z2, make_float2(-2.268580347e-02f, -2.268580347e-02f));
  p = __ffma2_rn(p, z2, make_float2(1.077178344e-01f, 1.077178344e-01f));
  p = __ffma2_rn(p, z2, make_float2(-3.732314110e-01f, -3.732314110e-01f));
  p = __ffma2_rn(p, z2, make_float2(1.127895713e+00f, 1.127895713e+00f));
  const float2 e = __fmul2_rn(p, z);
  return __ffma2_rn(h, e, h);
}

// Split-bf16 ("bf16x3") operand format of the WMK_PREC_MIXED extractor: a value v travels as hi = bf16(v),
// lo = bf16(v - hi) (16 mantissa bits); a row of K values is stored as [hi(K) | lo(K)], i.e. 2K bf16 = 4K bytes.
// Tag type: sizeof 4 like the storage per element.
struct SplitBf16 { __nv_bfloat16 h, l; };
template <typename T> struct OpMode { static constexpr int v = 0; };                // 0 fp32, 1 bf16, 2 split-bf16, 3 fp16
template <> struct OpMode<__nv_bfloat16> { static constexpr int v = 1; };
template <> struct OpMode<SplitBf16> { static constexpr int v = 2; };
template <> struct OpMode<__half> { static constexpr int v = 3; };
template <typename T> struct OpPlain16 { static constexpr bool v = OpMode<T>::v == 1 || OpMode<T>::v == 3; };   // bf16 or fp16

// two fp32 -> one packed 16-bit pair.  fp16 saturates to +-65504 instead of overflowing to inf.
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack2_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <bool F16> __device__ __forceinline__ uint32_t pack2_16(float lo, float hi) {
  if constexpr (F16) return pack2_f16(lo, hi);
  else return pack2_bf16(lo, hi);
}
template <bool F16> __device__ __forceinline__ float2 unpack2_16(uint32_t u) {
  if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
  else return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
// eight consecutive values -> four packed words, format chosen at run time (uniform branch, hoisted out of the callers' loops)
__device__ __forceinline__ void pack8_16(const float* f, bool f16, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  if (f16) { a = pack2_f16(f[0], f[1]); b = pack2_f16(f[2], f[3]); c = pack2_f16(f[4], f[5]); d = pack2_f16(f[6], f[7]); }
  else { a = pack2_bf16(f[0], f[1]); b = pack2_bf16(f[2], f[3]); c = pack2_bf16(f[4], f[5]); d = pack2_bf16(f[6], f[7]); }
}

// (hi, lo) words of two consecutive values
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi << 16), rb = b - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// store 4 consecutive values of a split row: `row` points at the row's hi part, the lo part starts K elements later
__device__ __forceinline__ void split_store4(__nv_bfloat16* row, int c, int K, float a, float b, float c2, float d) {
  uint2 h, l;
  split_pack2(a, b, h.x, l.x);
  split_pack2(c2, d, h.y, l.y);
  *reinterpret_cast<uint2*>(row + c) = h;
  *reinterpret_cast<uint2*>(row + K + c) = l;
}

// Which message image a clip carries: clip c of the batch belongs to utterance c / cpu and carries that utterance's
// image (c % cpu) % mpu.  mpu = 1: one 32x32 image per utterance; mpu = 4: a 64x64 image as four tiles, tile j mod 4
// in clip j; cpu = 1: one image per clip; cpu >= the batch: one image for every clip.
struct MsgMap { int cpu, mpu; };
__host__ __device__ __forceinline__ size_t msg_index(MsgMap m, int clip) {
  const int u = clip / m.cpu;
  return (size_t)u * m.mpu + (clip - u * m.cpu) % m.mpu;
}

// Where element (m, n) of a GEMM result is stored.
struct EpiParams {
  const float* bias;
  const float* resid;
  void* C;
  int M, N, ldc, epi, out_bf16;
  int up_h, up_w, up_cout;
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  const float* ln_mod = nullptr;
  int ln_H = 0, ln_shift = 0;
  int conv_H = 0;      // > 0: implicit-GEMM 3x3 convolution (see GemmArgs)
  int dn_Ho = 0;       // > 0: implicit-GEMM 4x4 stride-2 downsample over the space-to-depth tensor (see GemmArgs)
  int dn_chunks = 0;   //      64-channel chunks per tap (4C / 64)
  int boxc = 64;       // persistent kernel, bf16 output: columns per TMA store box (64 or 32)
  int resid_prefetch = 0;   // persistent kernel: L2-prefetch the next tile's residual rows
  int gelu_half = 0;        // GELU epilogue: the accumulator holds x / 2 (weights and bias pre-halved)
  int gelu_exact = 0;       // GELU epilogue: erf form (gelu_fast, |error| 1.5e-7) instead of the tanh form
  int gelu_mix = 0;         // GELU epilogue (halved argument): every other pair by the MUFU-free polynomial
  int split = 0;            // 0: plain; 1 / 2: split-bf16 A and W (K % 64 == 0 / K == 32); 3 / 4: fp16 A, W = hi + lo fp16 (K % 64 == 0 / K == 32)
  int f16 = 0;              // plain 16-bit operands / outputs are fp16 (else bf16)
  int ln_split = 0;         // fused LayerNorm output as split-bf16 rows [hi(N) | lo(N)]
  int ln_reuse = 0;         // fused LayerNorm stages its output in the warp's x staging buffer (no second staging area)
};

__device__ __forceinline__ size_t epi_row_offset(const EpiParams& p, int m, int n_first) {
  if (p.epi == EPI_UPSAMPLE) {
    int hw = p.up_h * p.up_w;
    int b = m / hw;
    int r = m - b * hw;
    int h = r / p.up_w;
    int w = r - h * p.up_w;
    int ij = n_first / p.up_cout;
    int co = n_first - ij * p.up_cout;
    int i = ij >> 1, j = ij & 1;
    size_t tok = (size_t)b * 4 * hw + (size_t)(2 * h + i) * (2 * p.up_w) + (2 * w + j);
    return tok * p.ldc + co;
  }
  return (size_t)m * p.ldc + n_first;
}

}  // namespace wmk

// UformerAudio embedder / extractor executor (uformerWM/model.py:2225-2511, configuration
// uformerWM/utils/model_utils.py:83-85): weight packing + the launch sequence of one pass.
//
// Data layout in HBM (per pass of `chunk` clips):
//   residual streams E0..E4 (encoder stages / skips) and D0..D3 (decoder stages): fp32
//   [clips*tokens][C]; operand buffers bufA (LayerNorm out), bufQKV, bufO (attention out),
//   bufH1 / bufH2 (LeFF hidden, 4C wide): OpT = fp32 or bf16, reused by every block.
// Weights are packed once: every nn.Linear / conv-as-GEMM weight as a K-major [N][K] matrix in
// OpT, q/k/v fused into one [3C][C] matrix with the attention scale folded into the q rows, the
// relative-position bias gathered to [heads][64][64], depthwise weights tap-major.
#include <stdlib.h>
#include <map>
#include <string>
#include <vector>

#include "small_kernels.cuh"

namespace wmk {

namespace tc { int num_sms(); }
int stft_clips(const float* wave, int B, int L, float* clips, int n_clips, cudaStream_t st);
int istft_clips(const float* clips, int B, int n_clips, int T, float* wave, int length, cudaStream_t st);
int leff_block(const void* A, const void* W1, const void* W2, const float* b1, const uint16_t* dw16, const float* dw_b,
               const float* b2, float* x, int n, int H, int C, int precise, cudaStream_t st);
int attn_block(const void* A, const void* Wh, const float* bqkv, const uint16_t* bias, uint16_t* out, int n, int H, int C,
               int shift, cudaStream_t st);
int dwconv3x3_gelu_op16(const void* in, void* out, const float* wt, const float* bias, int B, int H, int Ch, int f16,
                        cudaStream_t st);

namespace {

const int kDepths[9] = {1, 2, 8, 8, 2, 8, 8, 2, 1};
const int kHeads[9] = {1, 2, 4, 8, 16, 16, 8, 4, 2};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

struct BlockW {
  int C = 0, heads = 0, H = 0, shift = 0;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr, *mod = nullptr;
  float* attn_bias = nullptr;
  void *w_qkv = nullptr, *w_proj = nullptr, *w_l1 = nullptr, *w_l2 = nullptr;
  float *b_qkv = nullptr, *b_proj = nullptr, *b_l1 = nullptr, *b_l2 = nullptr, *dw_w = nullptr, *dw_b = nullptr;
  float *dw_wh = nullptr, *dw_bh = nullptr;      // 0.5 x the depthwise weights / bias (dwconv_tma.cu folds the GELU's 0.5)
  // fused q|k|v projection + attention (attn_block.cu): per-head weights [heads][q|k|v (96)][C] fp16, biases [heads][96],
  // relative-position bias [heads][64][64] fp16 in quad order (x log2 e)
  void* w_qkv_heads = nullptr;
  float* b_qkv_heads = nullptr;
  uint16_t* bias_quad = nullptr;
  uint16_t* dw16 = nullptr;                      // fused LeFF (leff_block.cu): fp16 [9][4C] of dw_wh, or [2][9][4C] = hi, lo of dw_w (precise)
};

struct EncW {
  InProjW in_proj;                // travels as a kernel parameter (constant bank)
  bool in_proj_has_ln = false;    // in_proj.ln_g / ln_b hold norm1 of the stage-0 block
  std::vector<BlockW> stage[5];
  void* down_w[4] = {};
  float* down_b[4] = {};
};

}  // namespace
}  // namespace wmk

using namespace wmk;

struct wmk_plan {
  int precision = WMK_PREC_FP32;
  bool finalized = false;
  int chunk = 32;
  int device = 0;
  std::map<std::string, HostTensor> host;
  std::vector<void*> allocs;      // weights
  std::vector<void*> ws_allocs;   // workspace
  size_t ws_bytes = 0;

  EncW enc, ext;
  std::vector<BlockW> dec[4];
  void* up_w[4] = {};
  float* up_b[4] = {};
  OutProjW out_proj;              // kernel parameter (constant bank)
  float *codec_c1w = nullptr, *codec_c1b = nullptr, *codec_c2w = nullptr, *codec_c2b = nullptr;
  float *codec_t1w = nullptr, *codec_t1b = nullptr, *codec_t2w = nullptr, *codec_t2b = nullptr;
  float *head_w = nullptr, *head_b = nullptr;
  float *sl0_w = nullptr, *sl0_b = nullptr, *sl2_w = nullptr, *sl2_b = nullptr;

  // workspace
  float *E[5] = {}, *D[4] = {};
  void *bufA = nullptr, *bufQKV = nullptr, *bufO = nullptr, *bufH1 = nullptr, *bufH2 = nullptr;
  float *feat = nullptr, *pool = nullptr, *headout = nullptr, *ybuf = nullptr, *wavebuf = nullptr, *rt = nullptr,
        *rt2 = nullptr;
  bool ws_ready = false;

  bool taps_on = false;
  std::map<std::string, std::pair<float*, size_t>> taps;

  // operand mode of the two networks: 0 fp32 (SIMT), 1 bf16 (tcgen05), 2 split-bf16 (tcgen05, three MMAs per product),
  // 3 fp16 (tcgen05)
  int embed_mode() const { return precision == WMK_PREC_FP32 ? 0 : precision == WMK_PREC_BF16 ? 1 : 3; }
  int extract_mode() const {
    return precision == WMK_PREC_FP32 ? 0 : precision == WMK_PREC_BF16 ? 1 : precision == WMK_PREC_F16 ? 3 : 2;
  }
  size_t op_size() const { return (precision == WMK_PREC_BF16 || precision == WMK_PREC_F16) ? 2 : 4; }   // bytes per operand element (split: hi + lo)
};

namespace wmk {
namespace {

// ------------------------------------------------------------------------------------ packing
int upload_f32(wmk_plan* P, const std::vector<float>& v, float** out) {
  float* d = nullptr;
  if (cudaMalloc(&d, v.size() * 4) != cudaSuccess) { set_error("cudaMalloc of %zu bytes failed", v.size() * 4); return WMK_ERR_ALLOC; }
  P->allocs.push_back(d);
  WMK_CHECK_CUDA(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

// fp16 copy of a float vector (sets = 1), or its (hi, lo) fp16 pair as two consecutive sets (sets = 2)
int upload_f16_sets(wmk_plan* P, const std::vector<float>& v, int sets, uint16_t** out) {
  std::vector<__half> h(v.size() * (size_t)sets);
  for (size_t i = 0; i < v.size(); ++i) {
    const float c = v[i] > 65504.f ? 65504.f : (v[i] < -65504.f ? -65504.f : v[i]);
    const __half hi = __float2half_rn(c);
    h[i] = hi;
    if (sets == 2) h[v.size() + i] = __float2half_rn(c - __half2float(hi));
  }
  void* d = nullptr;
  if (cudaMalloc(&d, h.size() * 2) != cudaSuccess) { set_error("cudaMalloc of %zu bytes failed", h.size() * 2); return WMK_ERR_ALLOC; }
  P->allocs.push_back(d);
  WMK_CHECK_CUDA(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  *out = reinterpret_cast<uint16_t*>(d);
  return 0;
}

// Dense-layer weight [N][K] (K-major) in the operand format of `mode`.  Split-bf16 (mode 2): rows [hi(K) | lo(K)]
// with hi = bf16(w), lo = bf16(w - hi); K == 32: rows of 128 = [hi | hi | lo | 0] (gemm_tcgen05.cu, p.split = 2).
int upload_op(wmk_plan* P, const std::vector<float>& v, void** out, int mode, int K) {
  if (mode == 0) return upload_f32(P, v, reinterpret_cast<float**>(out));
  std::vector<__nv_bfloat16> h;
  if (mode == 2) {
    const size_t N = v.size() / (size_t)K;
    const size_t ld = K == 32 ? 128 : 2 * (size_t)K;
    h.assign(N * ld, __float2bfloat16(0.f));
    for (size_t n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const float w = v[n * K + k];
        const __nv_bfloat16 hi = __float2bfloat16(w);
        const __nv_bfloat16 lo = __float2bfloat16(w - __bfloat162float(hi));
        if (K == 32) { h[n * ld + k] = hi; h[n * ld + 32 + k] = hi; h[n * ld + 64 + k] = lo; }
        else { h[n * ld + k] = hi; h[n * ld + K + k] = lo; }
      }
  } else if (mode == 4) {          // W-only split: rows [hi(K) | lo(K)] in fp16 (gemm_tcgen05.cu, p.split = 3 / 4)
    const size_t N = v.size() / (size_t)K;
    h.resize(N * 2 * (size_t)K);
    for (size_t n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const float w = v[n * K + k];
        const float c = w > 65504.f ? 65504.f : (w < -65504.f ? -65504.f : w);
        const __half hi = __float2half_rn(c);
        const __half lo = __float2half_rn(c - __half2float(hi));
        h[n * 2 * K + k] = *reinterpret_cast<const __nv_bfloat16*>(&hi);
        h[n * 2 * K + K + k] = *reinterpret_cast<const __nv_bfloat16*>(&lo);
      }
  } else if (mode == 3) {          // IEEE fp16, saturating
    h.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) {
      const float c = v[i] > 65504.f ? 65504.f : (v[i] < -65504.f ? -65504.f : v[i]);
      const __half hv = __float2half_rn(c);
      h[i] = *reinterpret_cast<const __nv_bfloat16*>(&hv);      // 16-bit container
    }
  } else {
    h.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16(v[i]);
  }
  void* d = nullptr;
  if (cudaMalloc(&d, h.size() * 2) != cudaSuccess) { set_error("cudaMalloc of %zu bytes failed", h.size() * 2); return WMK_ERR_ALLOC; }
  P->allocs.push_back(d);
  WMK_CHECK_CUDA(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  *out = d;
  return 0;
}

// q|k|v operands of one LeWin block from the reference's tensors (uformerWM/model.py:455-471,489-507,526):
//   wqkv [3C][C] = [Wq * scale | Wk | Wv], bqkv [3C] alike, bias [heads][64][64] = table gathered by the relative-position
//   index (model.py:496-505).  log2_domain (the tensor-core kernels, softmax by exp2): scale and bias carry a factor log2(e).
void build_qkv_operands(const float* wq, const float* bq, const float* wkv, const float* bkv, const float* table, int C, int heads,
                        bool log2_domain, std::vector<float>& wqkv, std::vector<float>& bqkv, std::vector<float>& bias) {
  const float lg = log2_domain ? 1.4426950408889634f : 1.0f;
  bias.assign((size_t)heads * 4096, 0.f);
  for (int h = 0; h < heads; ++h)
    for (int i = 0; i < 64; ++i)
      for (int j = 0; j < 64; ++j) {
        const int idx = ((i >> 3) - (j >> 3) + 7) * 15 + ((i & 7) - (j & 7) + 7);      // model.py:496-505
        bias[((size_t)h * 64 + i) * 64 + j] = table[(size_t)idx * heads + h] * lg;
      }
  const float scale = lg / sqrtf((float)(C / heads));                // model.py:489,526
  wqkv.resize(3 * (size_t)C * C);
  bqkv.resize(3 * (size_t)C);
  for (size_t i = 0; i < (size_t)C * C; ++i) wqkv[i] = wq[i] * scale;
  for (size_t i = 0; i < 2 * (size_t)C * C; ++i) wqkv[(size_t)C * C + i] = wkv[i];
  for (int i = 0; i < C; ++i) bqkv[i] = bq[i] * scale;
  for (int i = 0; i < 2 * C; ++i) bqkv[C + i] = bkv[i];
}

// attn_block.cu operands: head h owns rows [q_h | k_h | v_h] of the projection; the relative-position bias in "quad order"
// (window rows as four 4x4 sub-blocks: row r -> sub-block r >> 4, pixel ((r >> 2) & 3, r & 3) inside it)
void pack_attn_heads(const std::vector<float>& wqkv, const std::vector<float>& bqkv, const std::vector<float>& bias, int C, int heads,
                     std::vector<float>& wh, std::vector<float>& bh, std::vector<float>& bquad) {
  wh.resize((size_t)heads * 96 * C);
  bh.resize((size_t)heads * 96);
  bquad.resize((size_t)heads * 4096);
  for (int h = 0; h < heads; ++h)
    for (int part = 0; part < 3; ++part)
      for (int d = 0; d < 32; ++d) {
        const size_t src = (size_t)part * C + h * 32 + d, dst = (size_t)h * 96 + part * 32 + d;
        for (int c = 0; c < C; ++c) wh[dst * C + c] = wqkv[src * C + c];
        bh[dst] = bqkv[src];
      }
  auto pos = [](int r, int& i, int& j) { const int sb = r >> 4; i = ((sb >> 1) << 2) | ((r >> 2) & 3); j = ((sb & 1) << 2) | (r & 3); };
  for (int h = 0; h < heads; ++h)
    for (int r = 0; r < 64; ++r)
      for (int c = 0; c < 64; ++c) {
        int ri, rj, ci, cj;
        pos(r, ri, rj);
        pos(c, ci, cj);
        bquad[((size_t)h * 64 + r) * 64 + c] = bias[((size_t)h * 64 + (ri * 8 + rj)) * 64 + (ci * 8 + cj)];
      }
}

int get(wmk_plan* P, const std::string& name, size_t numel, const HostTensor** out) {
  auto it = P->host.find(name);
  if (it == P->host.end()) { set_error("plan: missing tensor '%s'", name.c_str()); return WMK_ERR_STATE; }
  if (it->second.data.size() != numel) {
    set_error("plan: tensor '%s' has %zu elements, expected %zu", name.c_str(), it->second.data.size(), numel);
    return WMK_ERR_STATE;
  }
  *out = &it->second;
  return 0;
}

int get_f32(wmk_plan* P, const std::string& name, size_t numel, float** dev) {
  const HostTensor* t;
  WMK_TRY(get(P, name, numel, &t));
  return upload_f32(P, t->data, dev);
}

int pack_block(wmk_plan* P, const std::string& p, int C, int heads, int H, int shift, bool mod, BlockW* w, int mode) {
  w->C = C; w->heads = heads; w->H = H;
  w->shift = (H <= 8) ? 0 : shift;                                    // model.py:892-894
  WMK_TRY(get_f32(P, p + "norm1.weight", C, &w->ln1_w));
  WMK_TRY(get_f32(P, p + "norm1.bias", C, &w->ln1_b));
  WMK_TRY(get_f32(P, p + "norm2.weight", C, &w->ln2_w));
  WMK_TRY(get_f32(P, p + "norm2.bias", C, &w->ln2_b));
  if (mod) WMK_TRY(get_f32(P, p + "modulator.weight", 64 * (size_t)C, &w->mod));
  const HostTensor *tab, *wq, *bq, *wkv, *bkv, *dw;
  WMK_TRY(get(P, p + "attn.relative_position_bias_table", 225 * (size_t)heads, &tab));
  WMK_TRY(get(P, p + "attn.qkv.to_q.weight", (size_t)C * C, &wq));
  WMK_TRY(get(P, p + "attn.qkv.to_q.bias", C, &bq));
  WMK_TRY(get(P, p + "attn.qkv.to_kv.weight", 2 * (size_t)C * C, &wkv));
  WMK_TRY(get(P, p + "attn.qkv.to_kv.bias", 2 * (size_t)C, &bkv));
  std::vector<float> bias, wqkv, bqkv;
  build_qkv_operands(wq->data.data(), bq->data.data(), wkv->data.data(), bkv->data.data(), tab->data.data(), C, heads, mode != 0,
                     wqkv, bqkv, bias);
  WMK_TRY(upload_f32(P, bias, &w->attn_bias));
  WMK_TRY(upload_op(P, wqkv, &w->w_qkv, mode, C));
  WMK_TRY(upload_f32(P, bqkv, &w->b_qkv));
  if (mode == 3 && C <= 128) {
    std::vector<float> wh, bh, bquad;
    pack_attn_heads(wqkv, bqkv, bias, C, heads, wh, bh, bquad);
    WMK_TRY(upload_op(P, wh, &w->w_qkv_heads, 3, C));
    WMK_TRY(upload_f32(P, bh, &w->b_qkv_heads));
    WMK_TRY(upload_f16_sets(P, bquad, 1, &w->bias_quad));
  }
  const HostTensor* t;
  WMK_TRY(get(P, p + "attn.proj.weight", (size_t)C * C, &t));
  WMK_TRY(upload_op(P, t->data, &w->w_proj, mode, C));
  WMK_TRY(get_f32(P, p + "attn.proj.bias", C, &w->b_proj));
  WMK_TRY(get(P, p + "mlp.linear1.0.weight", 4 * (size_t)C * C, &t));
  if (mode == 1 || mode == 3) {
    // the GELU epilogue of linear1 takes x / 2 (gelu_tanh2_half_arg): halve W1 and b1, exact in bf16 / fp32
    const HostTensor* tb;
    WMK_TRY(get(P, p + "mlp.linear1.0.bias", 4 * (size_t)C, &tb));
    std::vector<float> wh(t->data), bh(tb->data);
    for (auto& v : wh) v *= 0.5f;
    for (auto& v : bh) v *= 0.5f;
    WMK_TRY(upload_op(P, wh, &w->w_l1, mode, C));
    WMK_TRY(upload_f32(P, bh, &w->b_l1));
  } else {
    WMK_TRY(upload_op(P, t->data, &w->w_l1, mode == 2 ? 4 : mode, C));      // precise extractor: fp16 activations x (hi + lo) fp16 weights
    WMK_TRY(get_f32(P, p + "mlp.linear1.0.bias", 4 * (size_t)C, &w->b_l1));
  }
  WMK_TRY(get(P, p + "mlp.linear2.0.weight", 4 * (size_t)C * C, &t));
  WMK_TRY(upload_op(P, t->data, &w->w_l2, mode == 2 ? 4 : mode, 4 * C));
  WMK_TRY(get_f32(P, p + "mlp.linear2.0.bias", C, &w->b_l2));
  WMK_TRY(get(P, p + "mlp.dwconv.0.weight", 36 * (size_t)C, &dw));
  std::vector<float> dwt(36 * (size_t)C);
  for (int c = 0; c < 4 * C; ++c)
    for (int tap = 0; tap < 9; ++tap) dwt[(size_t)tap * 4 * C + c] = dw->data[(size_t)c * 9 + tap];
  WMK_TRY(upload_f32(P, dwt, &w->dw_w));
  WMK_TRY(get_f32(P, p + "mlp.dwconv.0.bias", 4 * (size_t)C, &w->dw_b));
  if (mode == 2) WMK_TRY(upload_f16_sets(P, dwt, 2, &w->dw16));
  if (mode == 1 || mode == 3) {
    const HostTensor* db;
    WMK_TRY(get(P, p + "mlp.dwconv.0.bias", 4 * (size_t)C, &db));
    std::vector<float> bh(4 * (size_t)C);
    for (size_t i = 0; i < dwt.size(); ++i) dwt[i] *= 0.5f;
    for (size_t i = 0; i < bh.size(); ++i) bh[i] = db->data[i] * 0.5f;
    WMK_TRY(upload_f32(P, dwt, &w->dw_wh));
    WMK_TRY(upload_f32(P, bh, &w->dw_bh));
    if (mode == 3) WMK_TRY(upload_f16_sets(P, dwt, 1, &w->dw16));
  }
  return 0;
}

// Downsample as an implicit GEMM over the space-to-depth tensor (16-bit plans; WMK_IMPLICIT_DOWN=0 restores im2col)
static int implicit_down() {
  static const int v = getenv("WMK_IMPLICIT_DOWN") ? atoi(getenv("WMK_IMPLICIT_DOWN")) : 1;
  return v;
}

static int ext_down_wsplit() {
  static const int v = getenv("WMK_EXT_DOWN_WSPLIT") ? atoi(getenv("WMK_EXT_DOWN_WSPLIT")) : 0;
  return v;
}

int pack_encoder(wmk_plan* P, const std::string& p, const std::string& inproj, EncW* e, int mode) {
  {
    const HostTensor *tw, *tb;
    WMK_TRY(get(P, inproj + "proj.0.weight", 32 * 2 * 9, &tw));
    WMK_TRY(get(P, inproj + "proj.0.bias", 32, &tb));
    for (int i = 0; i < 576; ++i) e->in_proj.w[i] = tw->data[i];
    for (int i = 0; i < 32; ++i) e->in_proj.b[i] = tb->data[i];
    const HostTensor *lg, *lb;                       // norm1 of the stage-0 block (no modulator, no shift in encoders)
    WMK_TRY(get(P, p + "encoderlayer_0.blocks.0.norm1.weight", 32, &lg));
    WMK_TRY(get(P, p + "encoderlayer_0.blocks.0.norm1.bias", 32, &lb));
    for (int i = 0; i < 32; ++i) { e->in_proj.ln_g[i] = lg->data[i]; e->in_proj.ln_b[i] = lb->data[i]; }
    e->in_proj_has_ln = true;
  }
  for (int s = 0; s < 5; ++s) {
    const int C = 32 << s, H = 128 >> s;
    e->stage[s].resize(kDepths[s]);
    for (int i = 0; i < kDepths[s]; ++i) {
      const std::string bp = p + (s < 4 ? "encoderlayer_" + std::to_string(s) : std::string("conv")) + ".blocks." +
                             std::to_string(i) + ".";
      WMK_TRY(pack_block(P, bp, C, kHeads[s], H, (i % 2) ? 4 : 0, false, &e->stage[s][i], mode));
    }
    if (s < 4) {
      const HostTensor* t;
      const std::string dp = p + "dowsample_" + std::to_string(s) + ".conv.0.";
      WMK_TRY(get(P, dp + "weight", (size_t)2 * C * C * 16, &t));
      // [co][(kh,kw,ci)] for the im2col form; the implicit-GEMM form (16-bit plans) orders k by the space-to-depth view:
      // kernel position (kh, kw) = (2a + ph, 2b + pw) -> k = ((a*2 + b)*4 + ph*2 + pw)*C + ci (s2d_pad_kernel)
      const bool s2d_order = mode != 0 && implicit_down() && !(mode == 2 && ext_down_wsplit());
      std::vector<float> g((size_t)2 * C * 16 * C);
      for (int co = 0; co < 2 * C; ++co)
        for (int ci = 0; ci < C; ++ci)
          for (int tap = 0; tap < 16; ++tap) {
            const int kh = tap >> 2, kw = tap & 3;
            const int pos = s2d_order ? (((kh >> 1) * 2 + (kw >> 1)) * 4 + (kh & 1) * 2 + (kw & 1)) : tap;
            g[((size_t)co * 16 + pos) * C + ci] = t->data[((size_t)co * C + ci) * 16 + tap];
          }
      // WMK_EXT_DOWN_WSPLIT=1 (experiment, off): fp16 im2col rows x (hi + lo) fp16 weights for the precise extractor's
      // downsample conv - 3.9 ms faster per step, but the fp16 rounding of the residual stream moves the logits by up to
      // 1.24e-4 (> the 1e-4 margin: test_mixed_extractor_bits_match_oracle_config2_shape fails), so split rows stay
      WMK_TRY(upload_op(P, g, &e->down_w[s], (mode == 2 && ext_down_wsplit()) ? 4 : mode, 16 * C));
      WMK_TRY(get_f32(P, dp + "bias", 2 * (size_t)C, &e->down_b[s]));
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------ workspace
template <typename T>
int ws_alloc(wmk_plan* P, T** p, size_t bytes) {
  void* d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess) { set_error("workspace cudaMalloc of %zu bytes failed", bytes); return WMK_ERR_ALLOC; }
  P->ws_allocs.push_back(d);
  P->ws_bytes += bytes;
  *p = reinterpret_cast<T*>(d);
  return 0;
}

int ensure_workspace(wmk_plan* P) {
  if (P->ws_ready) return 0;
  const size_t n = (size_t)P->chunk, os = P->op_size();
  for (int s = 0; s < 5; ++s) WMK_TRY(ws_alloc(P, &P->E[s], n * (16384 >> (2 * s)) * (32 << s) * 4));
  for (int s = 0; s < 4; ++s) WMK_TRY(ws_alloc(P, &P->D[s], n * (256 << (2 * s)) * (512 >> s) * 4));
  const size_t tokC = 16384 * 64;           // widest stage: decoder stage 3 (C = 64 at 128x128)
  WMK_TRY(ws_alloc(P, &P->bufA, n * tokC * os));
  WMK_TRY(ws_alloc(P, &P->bufQKV, n * tokC * 3 * os));
  WMK_TRY(ws_alloc(P, &P->bufO, n * tokC * os));
  WMK_TRY(ws_alloc(P, &P->bufH1, n * tokC * 4 * os));
  WMK_TRY(ws_alloc(P, &P->bufH2, n * tokC * 4 * os));
  WMK_TRY(ws_alloc(P, &P->feat, n * 256 * 4));
  WMK_TRY(ws_alloc(P, &P->pool, n * 256 * 4));
  WMK_TRY(ws_alloc(P, &P->headout, n * 256 * 4));
  WMK_TRY(ws_alloc(P, &P->ybuf, n * 32768 * 4));
  WMK_TRY(ws_alloc(P, &P->wavebuf, n * 8002 * 4));
  WMK_TRY(ws_alloc(P, &P->rt, n * 32768 * 4));
  WMK_TRY(ws_alloc(P, &P->rt2, n * 65536 * 4));
  P->ws_ready = true;
  return 0;
}

int tap(wmk_plan* P, const std::string& name, const float* src, size_t n, cudaStream_t st) {
  if (!P->taps_on) return 0;
  auto& slot = P->taps[name];
  if (slot.first && slot.second != n) { cudaFree(slot.first); slot.first = nullptr; }
  if (!slot.first) WMK_CHECK_CUDA(cudaMalloc(&slot.first, n * 4));
  slot.second = n;
  WMK_CHECK_CUDA(cudaMemcpyAsync(slot.first, src, n * 4, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// ------------------------------------------------------------------------------------ execution
int gemm(int mode, const GemmArgs& g, cudaStream_t st) {
  return mode != 0 ? gemm_bf16_tcgen05(g, st) : gemm_fp32_simt(g, st);
}

// One LeWin block (uformerWM/model.py:937-1019) on the residual stream x.  16-bit modes, C <= 128: the two
// LayerNorms are fused into the epilogues of the dense layers that produce their input - norm2 into the
// attention projection, and the NEXT block's norm1 (+ modulator) into this block's linear2 - so the fp32
// stream is not re-read; `ln1_ready` says the previous block already left LN1(x) in bufA, `next` is the
// following block of the same stage (nullptr for the last one).
//
// Precise mode (OpT = SplitBf16, the WMK_PREC_MIXED extractor; the precision of every tensor follows the measured
// sensitivity of the logits, tools/precision_study.py - WEIGHT rounding and the A operands of the attention
// projections dominate, the 16-bit rounding of q / k / v / P and of the LeFF activations costs ~3e-5):
//   LN1 out (bufA)     split-bf16 rows [hi | lo]          -> QKV projection  hi*hi + lo*hi + hi*lo   (3 MMAs)
//   q | k | v (bufQKV) fp16                               -> the fp16 attention kernel
//   attention out      split-bf16 rows                    -> output projection, 3 MMAs, fp32 residual stream
//   LN2 out (bufA)     fp16                               -> linear1: fp16 x (hi + lo fp16 weights), 2 MMAs, erf-form GELU
//   hidden H1 / H2     fp16, depthwise conv in fp32 + erf-form GELU -> linear2: fp16 x (hi + lo) weights, 2 MMAs
template <typename OpT>
int run_block(wmk_plan* P, const BlockW& w, float* x, int n, cudaStream_t st, bool ln1_ready = false,
              const BlockW* next = nullptr) {
  constexpr int MODE = OpMode<OpT>::v;
  constexpr bool P16 = OpPlain16<OpT>::v;          // plain 16-bit operands: bf16 (MODE 1) or fp16 (MODE 3)
  constexpr bool PRECISE = MODE == 2;
  constexpr int F16 = MODE == 3 || PRECISE;        // 16-bit tensors of this mode are fp16
  const int C = w.C, H = w.H;
  const int M = n * H * H;
  const int ob = P16 || PRECISE;                   // QKV / hidden tensors are 16-bit
  static const int fuse_min_c = getenv("WMK_FUSE_LN_MINC") ? atoi(getenv("WMK_FUSE_LN_MINC")) : 32;
  const bool fuse_ln = (P16 || PRECISE) && C <= 128 && C >= fuse_min_c;
  // size of one LayerNorm-output element in bufA (profile accounting)
  const double ln1_bytes = PRECISE ? 4 : sizeof(OpT), ln2_bytes = PRECISE ? 2 : sizeof(OpT);
  if (!(ln1_ready && fuse_ln)) {
    ProfScope prof(FAM_LAYERNORM, (double)M * C * (4 + ln1_bytes), st);
    launch_layernorm<OpT>(x, reinterpret_cast<OpT*>(P->bufA), w.ln1_w, w.ln1_b, w.mod, M, C, H, w.shift, st);
    WMK_CHECK_LAUNCH("layernorm_kernel");
  }
  GemmArgs g;
  static const int fused_attn = getenv("WMK_FUSED_ATTN") ? atoi(getenv("WMK_FUSED_ATTN")) : 1;
  const bool attn_fused = fused_attn && MODE == 3 && C <= 128 && H >= 16 && w.w_qkv_heads;
  if (attn_fused) {
    // q|k|v projection + window attention in ONE tcgen05 kernel (attn_block.cu): q, k, v never reach HBM
    WMK_TRY(attn_block(P->bufA, w.w_qkv_heads, w.b_qkv_heads, w.bias_quad, reinterpret_cast<uint16_t*>(P->bufO), n, H, C, w.shift, st));
  } else {
  g.A = P->bufA; g.W = w.w_qkv; g.bias = w.b_qkv; g.C = P->bufQKV; g.M = M; g.N = 3 * C; g.K = C; g.ldc = 3 * C;
  g.epi = EPI_BIAS; g.out_bf16 = ob; g.split = PRECISE; g.f16 = F16;
  WMK_TRY(gemm(MODE, g, st));
  }
  if (!attn_fused) {
    ProfScope prof(FAM_ATTENTION, 256.0 * C * M, st);
    const int n_windows = n * (H / 8) * (H / 8);
    if constexpr (P16 || PRECISE) {
      static const int att_nst = getenv("WMK_ATT_STAGES") ? atoi(getenv("WMK_ATT_STAGES")) : 2;   // 3 (two windows ahead, 4 CTAs per SM) measured 2 % slower: not bound by bytes in flight
      const int resident = att_nst == 3 ? 4 : 5;          // CTAs per SM (shared memory: 56 KB / 41 KB per CTA)
      int per_head = (148 * resident) / w.heads;          // CTAs per head
      if (per_head > n_windows) per_head = n_windows;
      if (per_head < 1) per_head = 1;
      const size_t att_smem = (size_t)(att_nst == 3 ? 3 : 2) * 3 * ATT_TILE * sizeof(uint16_t);
      if (att_nst == 3) {
        auto kern = window_attention_mma_kernel<F16 != 0, PRECISE, 3>;
        static bool attr = false;
        if (!attr) { WMK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem)); attr = true; }
        kern<<<per_head * w.heads, 128, att_smem, st>>>(
            reinterpret_cast<const uint16_t*>(P->bufQKV), reinterpret_cast<uint16_t*>(P->bufO), w.attn_bias, C, H, w.shift, n_windows);
      } else {
        window_attention_mma_kernel<F16 != 0, PRECISE, 2><<<per_head * w.heads, 128, att_smem, st>>>(
            reinterpret_cast<const uint16_t*>(P->bufQKV), reinterpret_cast<uint16_t*>(P->bufO), w.attn_bias, C, H, w.shift, n_windows);
      }
    } else {
      window_attention_kernel<float><<<dim3(n_windows, w.heads), 128, 0, st>>>(
          reinterpret_cast<const float*>(P->bufQKV), reinterpret_cast<float*>(P->bufO), w.attn_bias, C, H, w.shift);
    }
    WMK_CHECK_LAUNCH("window_attention_kernel");
  }
  g = GemmArgs();
  g.A = P->bufO; g.W = w.w_proj; g.bias = w.b_proj; g.resid = x; g.C = x; g.M = M; g.N = C; g.K = C; g.ldc = C;
  g.epi = EPI_BIAS_RESID; g.out_bf16 = 0; g.split = PRECISE; g.f16 = F16;
  if (fuse_ln) { g.ln_out = P->bufA; g.ln_gamma = w.ln2_w; g.ln_beta = w.ln2_b; }   // norm2 (model.py:1017), plain 16-bit rows
  WMK_TRY(gemm(MODE, g, st));
  if (!fuse_ln) {
    ProfScope prof(FAM_LAYERNORM, (double)M * C * (4 + ln2_bytes), st);
    if constexpr (PRECISE) launch_layernorm<__half>(x, reinterpret_cast<__half*>(P->bufA), w.ln2_w, w.ln2_b, nullptr, M, C, H, 0, st);
    else launch_layernorm<OpT>(x, reinterpret_cast<OpT*>(P->bufA), w.ln2_w, w.ln2_b, nullptr, M, C, H, 0, st);
    WMK_CHECK_LAUNCH("layernorm_kernel");
  }
  if constexpr (P16 || PRECISE) {
    // the whole LeFF (linear1 -> GELU -> depthwise 3x3 -> GELU -> linear2 + residual) in ONE tcgen05 kernel: the 4C-wide
    // hidden tensor stays in shared memory (leff_block.cu); C <= 128, 16 x 8 pixel tiles
    static const int fused = getenv("WMK_FUSED_LEFF") ? atoi(getenv("WMK_FUSED_LEFF")) : 0;
    if (fused && F16 && C <= 128 && H >= 16) {
      WMK_TRY(leff_block(P->bufA, w.w_l1, w.w_l2, w.b_l1, w.dw16, PRECISE ? w.dw_b : w.dw_bh, w.b_l2, x, n, H, C, PRECISE, st));
      if (fuse_ln && next) {               // the next block expects its norm1 in bufA
        ProfScope prof(FAM_LAYERNORM, (double)M * C * (4 + ln1_bytes), st);
        launch_layernorm<OpT>(x, reinterpret_cast<OpT*>(P->bufA), next->ln1_w, next->ln1_b, next->mod, M, C, H, next->shift, st);
        WMK_CHECK_LAUNCH("layernorm_kernel");
      }
      return 0;
    }
  }
  g = GemmArgs();
  g.A = P->bufA; g.W = w.w_l1; g.bias = w.b_l1; g.C = P->bufH1; g.M = M; g.N = 4 * C; g.K = C; g.ldc = 4 * C;
  g.epi = EPI_BIAS_GELU; g.out_bf16 = ob; g.gelu_half = P16;      // plain 16-bit plans carry W1 / 2, b1 / 2 (pack_block)
  g.wsplit = PRECISE; g.gelu_exact = PRECISE; g.f16 = F16;
  WMK_TRY(gemm(MODE, g, st));
  {
    ProfScope prof(FAM_DWCONV, 8.0 * M * C * (ob ? 2 : 4), st);
    if constexpr (P16) {
      WMK_TRY(dwconv3x3_gelu_op16(P->bufH1, P->bufH2, w.dw_wh, w.dw_bh, n, H, 4 * C, F16, st));
    } else if constexpr (PRECISE) {
      WMK_TRY(dwconv3x3_gelu_op16(P->bufH1, P->bufH2, w.dw_w, w.dw_b, n, H, 4 * C, 2, st));     // fp16 tensors, erf-form GELU
    } else {
      dwconv3x3_gelu_kernel<float><<<n * (H / 8) * (H / 8) * ((4 * C) / 64), 128, 0, st>>>(
          reinterpret_cast<const float*>(P->bufH1), reinterpret_cast<float*>(P->bufH2), w.dw_w, w.dw_b, n, H, 4 * C);
      WMK_CHECK_LAUNCH("dwconv3x3_gelu_kernel");
    }
  }
  g = GemmArgs();
  g.A = P->bufH2; g.W = w.w_l2; g.bias = w.b_l2; g.resid = x; g.C = x; g.M = M; g.N = C; g.K = 4 * C; g.ldc = C;
  g.epi = EPI_BIAS_RESID; g.out_bf16 = 0; g.wsplit = PRECISE; g.f16 = F16;
  if (fuse_ln && next) {                                                           // the next block's norm1 + modulator
    g.ln_out = P->bufA; g.ln_gamma = next->ln1_w; g.ln_beta = next->ln1_b; g.ln_mod = next->mod; g.ln_H = H; g.ln_shift = next->shift;
    g.ln_split = PRECISE;                                                          // ... as split rows for the next QKV projection
  }
  WMK_TRY(gemm(MODE, g, st));
  return 0;
}

// all blocks of one stage
template <typename OpT>
int run_stage(wmk_plan* P, const std::vector<BlockW>& blocks, float* x, int n, cudaStream_t st, bool first_ln_ready = false) {
  for (size_t i = 0; i < blocks.size(); ++i)
    WMK_TRY(run_block<OpT>(P, blocks[i], x, n, st, i > 0 || first_ln_ready,
                           i + 1 < blocks.size() ? &blocks[i + 1] : nullptr));
  return 0;
}

// Encoder / EncoderTransformerWM stages (model.py:1381-1394, 1569-1579): x NCHW -> E[0..4].
template <typename OpT>
int run_encoder(wmk_plan* P, const EncW& e, const float* x_nchw, int n, const char* tag, cudaStream_t st) {
  bool stage0_ln_ready = false;
  static const int split_ln0 = getenv("WMK_SPLIT_FUSE_LN") ? atoi(getenv("WMK_SPLIT_FUSE_LN")) : 1;
  {
    ProfScope prof(FAM_SMALL, (double)n * 16384 * (8 + 128), st);
    static const int fuse_ln0 = getenv("WMK_FUSE_FIRST_LN") ? atoi(getenv("WMK_FUSE_FIRST_LN")) : 1;
    const bool ln0 = (OpPlain16<OpT>::v || (OpMode<OpT>::v == 2 && split_ln0)) && fuse_ln0 && e.in_proj_has_ln;       // norm1 of stage 0's block in the same pass
    input_proj_kernel<<<cdiv((size_t)n * 16384, 128), 128, 0, st>>>(x_nchw, P->E[0], e.in_proj, n,
                                                                   ln0 ? reinterpret_cast<uint16_t*>(P->bufA) : nullptr,
                                                                   OpMode<OpT>::v == 3 ? 1 : OpMode<OpT>::v == 2 ? 2 : 0);
    WMK_CHECK_LAUNCH("input_proj_kernel");
    stage0_ln_ready = ln0;
  }
  const std::string t(tag);
  if (t == "enc") WMK_TRY(tap(P, "emb.inproj", P->E[0], (size_t)n * 16384 * 32, st));
  bool ln_ready = stage0_ln_ready;          // norm1 of the stage's first block already left in bufA by the producer of E[s]
  for (int s = 0; s < 5; ++s) {
    const int C = 32 << s, H = 128 >> s;
    WMK_TRY(run_stage<OpT>(P, e.stage[s], P->E[s], n, st, ln_ready));
    ln_ready = false;
    WMK_TRY(tap(P, t + ".conv" + std::to_string(s), P->E[s], (size_t)n * H * H * C, st));
    if (s == 4) break;
    const int Ho = H / 2;
    OpT* col = reinterpret_cast<OpT*>(P->bufH1);
    const size_t total = (size_t)n * Ho * Ho * 4 * (C / 8);
    const bool down_wsplit = OpMode<OpT>::v == 2 && ext_down_wsplit();      // fp16 rows x (hi + lo) weights
    const bool s2d = OpMode<OpT>::v != 0 && implicit_down() && !down_wsplit;   // implicit GEMM: every element written once
    if (s2d) {
      const size_t cells = (size_t)n * (Ho + 1) * (Ho + 1);
      ProfScope prof_l(FAM_LAYOUT, (double)n * H * H * C * 4 + (double)cells * 4 * C * sizeof(OpT), st);
      s2d_pad_kernel<OpT><<<cdiv(cells * 4 * (C / 8), 256), 256, 0, st>>>(P->E[s], col, n, H, C);
      WMK_CHECK_LAUNCH("s2d_pad_kernel");
    } else {
      ProfScope prof_l(FAM_LAYOUT, (double)n * H * H * C * 4 + (double)n * Ho * Ho * 16 * C * (down_wsplit ? 2 : sizeof(OpT)), st);
      if (down_wsplit) im2col_4x4s2_kernel<__half><<<cdiv(total, 256), 256, 0, st>>>(P->E[s], reinterpret_cast<__half*>(col), n, H, C);
      else im2col_4x4s2_kernel<OpT><<<cdiv(total, 256), 256, 0, st>>>(P->E[s], col, n, H, C);
      WMK_CHECK_LAUNCH("im2col_4x4s2_kernel");
    }
    GemmArgs g;
    g.A = col; g.W = e.down_w[s]; g.bias = e.down_b[s]; g.C = P->E[s + 1]; g.M = n * Ho * Ho; g.N = 2 * C;
    g.K = 16 * C; g.ldc = 2 * C; g.epi = EPI_BIAS; g.out_bf16 = 0; g.split = OpMode<OpT>::v == 2 && !down_wsplit; g.wsplit = down_wsplit;
    g.f16 = OpMode<OpT>::v == 3;
    if (s2d) { g.dn_Ho = Ho; g.dn_B = n; }
    static const int fuse_first_ln = getenv("WMK_FUSE_FIRST_LN") ? atoi(getenv("WMK_FUSE_FIRST_LN")) : 1;
    if ((OpPlain16<OpT>::v || (OpMode<OpT>::v == 2 && split_ln0)) && fuse_first_ln && 2 * C <= 128) {
      // the next stage's first norm1 rides on the downsample conv's epilogue (encoder blocks carry no modulator)
      const BlockW& nb = e.stage[s + 1][0];
      g.ln_out = P->bufA; g.ln_gamma = nb.ln1_w; g.ln_beta = nb.ln1_b; g.ln_mod = nb.mod; g.ln_H = Ho; g.ln_shift = nb.shift;
      g.ln_split = OpMode<OpT>::v == 2;
      ln_ready = true;
    }
    WMK_TRY(gemm(OpMode<OpT>::v, g, st));
    WMK_TRY(tap(P, t + ".pool" + std::to_string(s), P->E[s + 1], (size_t)n * Ho * Ho * 2 * C, st));
  }
  return 0;
}

template <typename OpT>
int run_extract(wmk_plan* P, const float* y, int n, float* wm, float* logits, cudaStream_t st) {
  WMK_TRY(run_encoder<OpT>(P, P->ext, y, n, "ext", st));
  extract_head_kernel<<<cdiv((size_t)n * 256, 256), 256, 0, st>>>(P->E[4], P->headout, P->head_w, P->head_b, n);
  WMK_CHECK_LAUNCH("extract_head_kernel");
  WMK_TRY(tap(P, "ext.feat", P->headout, (size_t)n * 256, st));
  wm_decode_kernel<<<n, 256, 0, st>>>(P->headout, nullptr, wm, logits, P->codec_t1w, P->codec_t1b, P->codec_t2w,
                                      P->codec_t2b);
  WMK_CHECK_LAUNCH("wm_decode_kernel");
  return 0;
}

// the extractor in the plan's extractor mode (WMK_PREC_MIXED: split-bf16 whatever the embedder runs in)
int run_extract_any(wmk_plan* P, const float* y, int n, float* wm, float* logits, cudaStream_t st) {
  switch (P->extract_mode()) {
    case 0: return run_extract<float>(P, y, n, wm, logits, st);
    case 1: return run_extract<__nv_bfloat16>(P, y, n, wm, logits, st);
    case 3: return run_extract<__half>(P, y, n, wm, logits, st);
    default: return run_extract<SplitBf16>(P, y, n, wm, logits, st);
  }
}

template <typename OpT>
int run_forward(wmk_plan* P, const float* x, const float* msg, MsgMap mm, int clip0, int n, float* stft_new, float* noise,
                float* y_out, float* wm_pred, float* wm, float* wm_logits, cudaStream_t st) {
  wm_encode_kernel<<<n, 256, 0, st>>>(msg, mm, clip0, P->feat, P->codec_c1w, P->codec_c1b, P->codec_c2w, P->codec_c2b);
  WMK_CHECK_LAUNCH("wm_encode_kernel");
  WMK_TRY(run_encoder<OpT>(P, P->enc, x, n, "enc", st));
  if (wm_pred) {
    bottleneck_maxpool_kernel<<<cdiv((size_t)n * 256, 256), 256, 0, st>>>(P->E[4], P->pool, n);
    WMK_CHECK_LAUNCH("bottleneck_maxpool_kernel");
    wm_decode_kernel<<<n, 256, 0, st>>>(P->feat, P->pool, wm_pred, nullptr, P->codec_t1w, P->codec_t1b, P->codec_t2w,
                                        P->codec_t2b);
    WMK_CHECK_LAUNCH("wm_decode_kernel");
  }
  // decoder (model.py:1221-1240)
  OpT* A = reinterpret_cast<OpT*>(P->bufH1);
  bottleneck_concat_kernel<OpT><<<cdiv((size_t)n * 65536, 256), 256, 0, st>>>(P->feat, P->E[4], A, n);
  WMK_CHECK_LAUNCH("bottleneck_concat_kernel");
  for (int s = 0; s < 4; ++s) {
    const int Hin = 8 << s, Cin = (s == 0) ? 1024 : (1024 >> s), Cout = 256 >> s;
    const int Hout = 2 * Hin, Cd = 2 * Cout;
    if (s > 0) {
      const size_t rows = (size_t)n * Hin * Hin;
      ProfScope prof_cp(FAM_LAYOUT, (double)rows * Cin * 6, st);
      copy_cols_kernel<OpT><<<cdiv(rows * (Cin / 4), 256), 256, 0, st>>>(P->D[s - 1], A, rows, Cin, Cin, 0);
      WMK_CHECK_LAUNCH("copy_cols_kernel");
    }
    GemmArgs g;
    g.A = A; g.W = P->up_w[s]; g.bias = P->up_b[s]; g.C = P->D[s]; g.M = n * Hin * Hin; g.N = 4 * Cout; g.K = Cin;
    g.ldc = Cd; g.epi = EPI_UPSAMPLE; g.out_bf16 = 0; g.up_h = Hin; g.up_w = Hin; g.up_cout = Cout; g.f16 = OpMode<OpT>::v == 3;
    WMK_TRY(gemm(OpMode<OpT>::v, g, st));
    {
      const size_t rows = (size_t)n * Hout * Hout;
      ProfScope prof_cp(FAM_LAYOUT, (double)rows * Cout * 8, st);
      copy_cols_kernel<float><<<cdiv(rows * (Cout / 4), 256), 256, 0, st>>>(P->E[3 - s], P->D[s], rows, Cout, Cd, Cout);
      WMK_CHECK_LAUNCH("copy_cols_kernel");
    }
    WMK_TRY(run_stage<OpT>(P, P->dec[s], P->D[s], n, st));
    WMK_TRY(tap(P, "dec.deconv" + std::to_string(s), P->D[s], (size_t)n * Hout * Hout * Cd, st));
  }
  float* y = y_out ? y_out : P->ybuf;
  {
    ProfScope prof(FAM_SMALL, (double)n * 16384 * (256 + 24), st);
    output_proj_kernel<<<n * (128 / OP_TH) * (128 / OP_TW), OP_THREADS, 0, st>>>(P->D[3], x, noise, y, P->out_proj, n);
    WMK_CHECK_LAUNCH("output_proj_kernel");
  }
  if (stft_new) {
    // in-graph ISTFT -> STFT projection + stft_layer (model.py:2458-2465)
    WMK_TRY(istft_clips(y, n, 1, 128, P->wavebuf, 8002, st));
    WMK_TRY(tap(P, "emb.wave", P->wavebuf, (size_t)n * 8002, st));
    WMK_TRY(stft_clips(P->wavebuf, n, 8002, P->rt, 1, st));
    WMK_TRY(tap(P, "emb.roundtrip", P->rt, (size_t)n * 32768, st));
    {
      ProfScope prof_sl(FAM_SMALL, (double)n * 16384 * 48, st);
      conv3x3_nchw_kernel<2, 4, true><<<cdiv((size_t)n * 16384, 256), 256, 0, st>>>(P->rt, P->rt2, P->sl0_w, P->sl0_b, n);
      WMK_CHECK_LAUNCH("conv3x3_nchw_kernel<2,4>");
      conv3x3_nchw_kernel<4, 2, false><<<cdiv((size_t)n * 16384, 256), 256, 0, st>>>(P->rt2, stft_new, P->sl2_w, P->sl2_b, n);
      WMK_CHECK_LAUNCH("conv3x3_nchw_kernel<4,2>");
    }
  }
  if (wm || wm_logits) WMK_TRY(run_extract_any(P, y, n, wm, wm_logits, st));    // model.py:2508-2509 reads y
  return 0;
}

int check_ready(wmk_plan* P) {
  if (!P) { set_error("null plan"); return WMK_ERR_ARG; }
  if (!P->finalized) { set_error("plan is not finalized"); return WMK_ERR_STATE; }
  WMK_CHECK_CUDA(cudaSetDevice(P->device));
  return ensure_workspace(P);
}

}  // namespace
}  // namespace wmk

// ------------------------------------------------------------------------------------ C ABI
extern "C" int wmk_uformer_plan_create(int precision, wmk_plan** out) {
  WMK_REQUIRE(out, "plan_create: null out");
  WMK_REQUIRE(precision == WMK_PREC_FP32 || precision == WMK_PREC_BF16 || precision == WMK_PREC_MIXED || precision == WMK_PREC_F16,
              "plan_create: unknown precision %d", precision);
  int dev = 0;
  WMK_CHECK_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  WMK_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (precision != WMK_PREC_FP32 && prop.major != 10) {
    set_error("bf16 / mixed precision needs an sm_100 device (tcgen05); found sm_%d%d", prop.major, prop.minor);
    return WMK_ERR_UNSUPPORTED;
  }
  wmk_plan* P = new wmk_plan();
  P->precision = precision;
  P->device = dev;
  *out = P;
  return 0;
}

extern "C" int wmk_plan_destroy(wmk_plan* P) {
  if (!P) return 0;
  for (void* p : P->allocs) cudaFree(p);
  for (void* p : P->ws_allocs) cudaFree(p);
  for (auto& kv : P->taps) cudaFree(kv.second.first);
  delete P;
  return 0;
}

extern "C" int wmk_plan_set_tensor(wmk_plan* P, const char* name, const float* data_host, const int64_t* shape, int ndim) {
  WMK_REQUIRE(P && name && data_host && shape && ndim >= 0 && ndim <= 8, "set_tensor: bad arguments");
  WMK_REQUIRE(!P->finalized, "set_tensor: plan already finalized");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.data.assign(data_host, data_host + n);
  P->host[name] = std::move(t);
  return 0;
}

extern "C" int wmk_plan_set_chunk(wmk_plan* P, int clips_per_pass) {
  WMK_REQUIRE(P, "set_chunk: null plan");
  WMK_REQUIRE(!P->ws_ready, "set_chunk: workspace already allocated");
  if (clips_per_pass > 0) P->chunk = clips_per_pass;
  return 0;
}

extern "C" size_t wmk_plan_workspace_bytes(const wmk_plan* P) { return P ? P->ws_bytes : 0; }

extern "C" int wmk_plan_finalize(wmk_plan* P) {
  WMK_REQUIRE(P, "finalize: null plan");
  if (P->finalized) return 0;
  WMK_CHECK_CUDA(cudaSetDevice(P->device));
  WMK_TRY(pack_encoder(P, "encoder.", "input_proj.", &P->enc, P->embed_mode()));
  WMK_TRY(pack_encoder(P, "decoder_wm.", "decoder_wm.input_proj.", &P->ext, P->extract_mode()));
  for (int s = 0; s < 4; ++s) {
    const int Cin = (s == 0) ? 1024 : (1024 >> s), Cout = 256 >> s, Cd = 2 * Cout, H = 16 << s;
    const std::string up = "decoder.upsample_" + std::to_string(s) + ".deconv.0.";
    const HostTensor *t, *tb;
    WMK_TRY(get(P, up + "weight", (size_t)Cin * Cout * 4, &t));
    WMK_TRY(get(P, up + "bias", Cout, &tb));
    std::vector<float> g((size_t)4 * Cout * Cin), b4((size_t)4 * Cout);      // [(i,j,co)][ci]
    for (int ci = 0; ci < Cin; ++ci)
      for (int co = 0; co < Cout; ++co)
        for (int ij = 0; ij < 4; ++ij) g[((size_t)ij * Cout + co) * Cin + ci] = t->data[((size_t)ci * Cout + co) * 4 + ij];
    for (int ij = 0; ij < 4; ++ij)
      for (int co = 0; co < Cout; ++co) b4[(size_t)ij * Cout + co] = tb->data[co];
    WMK_TRY(upload_op(P, g, &P->up_w[s], P->embed_mode(), Cin));
    WMK_TRY(upload_f32(P, b4, &P->up_b[s]));
    P->dec[s].resize(kDepths[5 + s]);
    for (int i = 0; i < kDepths[5 + s]; ++i) {
      const std::string bp = "decoder.decoderlayer_" + std::to_string(s) + ".blocks." + std::to_string(i) + ".";
      WMK_TRY(pack_block(P, bp, Cd, kHeads[5 + s], H, (i % 2) ? 4 : 0, true, &P->dec[s][i], P->embed_mode()));
    }
  }
  {
    const HostTensor *tw, *tb;
    WMK_TRY(get(P, "output_proj.proj.0.weight", 2 * 64 * 9, &tw));   // [o][c][tap] -> [c][tap*2 + o]
    WMK_TRY(get(P, "output_proj.proj.0.bias", 2, &tb));
    for (int o = 0; o < 2; ++o)
      for (int c = 0; c < 64; ++c)
        for (int t = 0; t < 9; ++t) P->out_proj.w[c * 18 + t * 2 + o] = tw->data[(o * 64 + c) * 9 + t];
    P->out_proj.b[0] = tb->data[0];
    P->out_proj.b[1] = tb->data[1];
  }
  WMK_TRY(get_f32(P, "encoder_wm.conv1.weight", 144, &P->codec_c1w));
  WMK_TRY(get_f32(P, "encoder_wm.conv1.bias", 16, &P->codec_c1b));
  WMK_TRY(get_f32(P, "encoder_wm.conv2.weight", 576, &P->codec_c2w));
  WMK_TRY(get_f32(P, "encoder_wm.conv2.bias", 4, &P->codec_c2b));
  WMK_TRY(get_f32(P, "encoder_wm.t_conv1.weight", 256, &P->codec_t1w));
  WMK_TRY(get_f32(P, "encoder_wm.t_conv1.bias", 16, &P->codec_t1b));
  WMK_TRY(get_f32(P, "encoder_wm.t_conv2.weight", 64, &P->codec_t2w));
  WMK_TRY(get_f32(P, "encoder_wm.t_conv2.bias", 1, &P->codec_t2b));
  WMK_TRY(get_f32(P, "decoder_wm.conv2.weight", 64, &P->head_w));
  WMK_TRY(get_f32(P, "decoder_wm.conv2.bias", 1, &P->head_b));
  WMK_TRY(get_f32(P, "stft_layer.0.weight", 72, &P->sl0_w));
  WMK_TRY(get_f32(P, "stft_layer.0.bias", 4, &P->sl0_b));
  WMK_TRY(get_f32(P, "stft_layer.2.weight", 72, &P->sl2_w));
  WMK_TRY(get_f32(P, "stft_layer.2.bias", 2, &P->sl2_b));
  P->host.clear();
  P->finalized = true;
  return 0;
}

static int forward_mapped(wmk_plan* P, const float* x, const float* msg, MsgMap mm, int B, float* stft_new, float* noise,
                          float* y, float* wm_pred, float* wm, float* wm_logits, void* stream) {
  WMK_TRY(check_ready(P));
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += P->chunk) {
    const int n = B - b0 < P->chunk ? B - b0 : P->chunk;
    const size_t o = (size_t)b0;
    auto off = [&](float* p, size_t per) { return p ? p + o * per : nullptr; };
    int s;
    if (P->embed_mode() == 1)
      s = run_forward<__nv_bfloat16>(P, x + o * 32768, msg, mm, b0, n, off(stft_new, 32768), off(noise, 32768), off(y, 32768),
                                     off(wm_pred, 1024), off(wm, 1024), off(wm_logits, 1024), st);
    else if (P->embed_mode() == 3)
      s = run_forward<__half>(P, x + o * 32768, msg, mm, b0, n, off(stft_new, 32768), off(noise, 32768), off(y, 32768),
                              off(wm_pred, 1024), off(wm, 1024), off(wm_logits, 1024), st);
    else
      s = run_forward<float>(P, x + o * 32768, msg, mm, b0, n, off(stft_new, 32768), off(noise, 32768), off(y, 32768),
                             off(wm_pred, 1024), off(wm, 1024), off(wm_logits, 1024), st);
    if (s) return s;
  }
  return 0;
}

extern "C" int wmk_uformer_forward(wmk_plan* P, const float* x, const float* msg, int msg_stride, int B, float* stft_new,
                                   float* noise, float* y, float* wm_pred, float* wm, float* wm_logits, void* stream) {
  WMK_REQUIRE(x && msg && B > 0 && (msg_stride == 0 || msg_stride == 1024), "forward: bad arguments");
  const MsgMap mm = msg_stride == 0 ? MsgMap{0x7fffffff, 1} : MsgMap{1, 1};
  return forward_mapped(P, x, msg, mm, B, stft_new, noise, y, wm_pred, wm, wm_logits, stream);
}

extern "C" int wmk_uformer_forward_mapped(wmk_plan* P, const float* x, const float* msg, int clips_per_utt, int msgs_per_utt,
                                          int B, float* stft_new, float* noise, float* y, float* wm_pred, float* wm,
                                          float* wm_logits, void* stream) {
  WMK_REQUIRE(x && msg && B > 0 && clips_per_utt > 0 && msgs_per_utt > 0, "forward_mapped: bad arguments");
  return forward_mapped(P, x, msg, MsgMap{clips_per_utt, msgs_per_utt}, B, stft_new, noise, y, wm_pred, wm, wm_logits, stream);
}

extern "C" int wmk_uformer_extract(wmk_plan* P, const float* y, int B, float* wm, float* wm_logits, void* stream) {
  WMK_TRY(check_ready(P));
  WMK_REQUIRE(y && B > 0 && (wm || wm_logits), "extract: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += P->chunk) {
    const int n = B - b0 < P->chunk ? B - b0 : P->chunk;
    const size_t o = (size_t)b0;
    const int s = run_extract_any(P, y + o * 32768, n, wm ? wm + o * 1024 : nullptr,
                                  wm_logits ? wm_logits + o * 1024 : nullptr, st);
    if (s) return s;
  }
  return 0;
}

extern "C" int wmk_uformer_autoencode(wmk_plan* P, const float* msg, int msg_stride, int B, float* wm_pred, void* stream) {
  WMK_TRY(check_ready(P));
  WMK_REQUIRE(msg && wm_pred && B > 0 && (msg_stride == 0 || msg_stride == 1024), "autoencode: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  for (int b0 = 0; b0 < B; b0 += P->chunk) {
    const int n = B - b0 < P->chunk ? B - b0 : P->chunk;
    wm_encode_kernel<<<n, 256, 0, st>>>(msg, msg_stride == 0 ? MsgMap{0x7fffffff, 1} : MsgMap{1, 1}, b0, P->feat, P->codec_c1w,
                                        P->codec_c1b, P->codec_c2w, P->codec_c2b);
    WMK_CHECK_LAUNCH("wm_encode_kernel");
    wm_decode_kernel<<<n, 256, 0, st>>>(P->feat, nullptr, wm_pred + (size_t)b0 * 1024, nullptr, P->codec_t1w, P->codec_t1b,
                                        P->codec_t2w, P->codec_t2b);
    WMK_CHECK_LAUNCH("wm_decode_kernel");
  }
  return 0;
}

extern "C" int wmk_plan_enable_taps(wmk_plan* P, int enable) {
  WMK_REQUIRE(P, "enable_taps: null plan");
  P->taps_on = enable != 0;
  return 0;
}

extern "C" int wmk_plan_get_tap(wmk_plan* P, const char* name, float* out, size_t capacity, size_t* n_out) {
  WMK_REQUIRE(P && name && n_out, "get_tap: bad arguments");
  auto it = P->taps.find(name);
  if (it == P->taps.end()) { set_error("get_tap: no tap named '%s'", name); return WMK_ERR_ARG; }
  *n_out = it->second.second;
  if (out) {
    WMK_REQUIRE(capacity >= it->second.second, "get_tap: capacity %zu < %zu", capacity, it->second.second);
    WMK_CHECK_CUDA(cudaMemcpy(out, it->second.first, it->second.second * 4, cudaMemcpyDeviceToDevice));
  }
  return 0;
}

static __global__ void widen_kernel(const uint16_t* in, float* out, size_t n, bool f16) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = f16 ? __half2float(reinterpret_cast<const __half*>(in)[i]) : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(in)[i]);
}

// fp32 [rows][K] -> split-bf16 rows.  form 0: [hi(K) | lo(K)]; form 1 (K = 32 weight rows): [hi | hi | lo | 0].
static __global__ void split_rows_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t rows, int K, int form) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (size_t)K) return;
  const size_t r = i / K;
  const int k = (int)(i - r * K);
  const float v = src[i];
  const __nv_bfloat16 hi = __float2bfloat16(v);
  const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
  if (form == 1) {
    __nv_bfloat16* d = dst + r * 128;
    d[k] = hi; d[32 + k] = hi; d[64 + k] = lo; d[96 + k] = __float2bfloat16(0.f);
  } else {
    __nv_bfloat16* d = dst + r * 2 * (size_t)K;
    d[k] = hi; d[K + k] = lo;
  }
}

// fp32 [rows][K] -> fp16 rows [hi(K) | lo(K)] (the W-only split weight form)
static __global__ void wsplit_rows_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t rows, int K) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (size_t)K) return;
  const size_t r = i / K;
  const int k = (int)(i - r * K);
  const float v = fminf(fmaxf(src[i], -65504.f), 65504.f);
  const __half hi = __float2half_rn(v);
  dst[r * 2 * (size_t)K + k] = hi;
  dst[r * 2 * (size_t)K + K + k] = __float2half_rn(v - __half2float(hi));
}

static __global__ void scale_kernel(const float* __restrict__ in, float* __restrict__ out, size_t n, float s) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i] * s;
}

// Stand-alone LeFF block for the unit tests: x[M][C] += Linear2(GELU(dwconv3x3(GELU(Linear1(A))))) on n images of H x H
// tokens, all tensors fp32 on the device (A = the LayerNorm-2 output; W1 [4C][C], b1 [4C], dw_w [9][4C] tap-major,
// dw_b [4C], W2 [C][4C], b2 [C]); operands are converted on the fly.  precise = 0: fp16 operands, tanh-form GELU;
// precise = 1: fp16 activations x (hi + lo) fp16 weights, erf-form GELU (the WMK_PREC_MIXED extractor).
extern "C" int wmk_leff_block_f32(const float* A, const float* W1, const float* b1, const float* dw_w, const float* dw_b,
                                  const float* W2, const float* b2, float* x, int n, int H, int C, int precise, void* stream) {
  WMK_REQUIRE(A && W1 && b1 && dw_w && dw_b && W2 && b2 && x && n > 0, "leff_block: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t M = (size_t)n * H * H, K4 = 4 * (size_t)C;
  __half *a16 = nullptr, *w1 = nullptr, *w2 = nullptr;
  float *b1s = nullptr, *dws = nullptr, *dbs = nullptr, *w1s = nullptr;
  __half* dw16 = nullptr;
  const int wt = precise ? 2 : 1;
  WMK_CHECK_CUDA(cudaMallocAsync(&a16, M * C * 2, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&w1, K4 * C * 2 * wt, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&w2, K4 * C * 2 * wt, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&b1s, K4 * 4, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&dws, 9 * K4 * 4, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&dbs, K4 * 4, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&w1s, K4 * C * 4, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&dw16, 9 * K4 * 2 * wt, st));
  copy_cols_kernel<__half><<<cdiv(M * (C / 4), 256), 256, 0, st>>>(A, a16, M, C, C, 0);
  WMK_CHECK_LAUNCH("copy_cols_kernel");
  const float sc = precise ? 1.0f : 0.5f;            // plain plans fold the GELU's 0.5 into W1 / b1 / dw (pack_block)
  scale_kernel<<<cdiv(K4 * C, 256), 256, 0, st>>>(W1, w1s, K4 * C, sc);
  scale_kernel<<<cdiv(K4, 256), 256, 0, st>>>(b1, b1s, K4, sc);
  scale_kernel<<<cdiv(9 * K4, 256), 256, 0, st>>>(dw_w, dws, 9 * K4, sc);
  scale_kernel<<<cdiv(K4, 256), 256, 0, st>>>(dw_b, dbs, K4, sc);
  count_launch(4);
  if (precise) {
    wsplit_rows_kernel<<<cdiv(K4 * C, 256), 256, 0, st>>>(w1s, w1, K4, C);
    wsplit_rows_kernel<<<cdiv(K4 * C, 256), 256, 0, st>>>(W2, w2, (size_t)C, (int)K4);
  } else {
    copy_cols_kernel<__half><<<cdiv(K4 * (C / 4), 256), 256, 0, st>>>(w1s, w1, K4, C, C, 0);
    copy_cols_kernel<__half><<<cdiv((size_t)C * (K4 / 4), 256), 256, 0, st>>>(W2, w2, (size_t)C, (int)K4, (int)K4, 0);
  }
  count_launch(2);
  if (precise) {      // the 9 x 4C depthwise weights as ONE row of (hi | lo) = the two sets
    wsplit_rows_kernel<<<cdiv(9 * K4, 256), 256, 0, st>>>(dws, dw16, 1, (int)(9 * K4));
  } else {
    copy_cols_kernel<__half><<<cdiv(9 * K4 / 4, 256), 256, 0, st>>>(dws, dw16, 1, (int)(9 * K4), (int)(9 * K4), 0);
  }
  count_launch(1);
  const int s = leff_block(a16, w1, w2, b1s, reinterpret_cast<const uint16_t*>(dw16), dbs, b2, x, n, H, C, precise, st);
  cudaFreeAsync(dw16, st);
  cudaFreeAsync(a16, st); cudaFreeAsync(w1, st); cudaFreeAsync(w2, st); cudaFreeAsync(b1s, st);
  cudaFreeAsync(dws, st); cudaFreeAsync(dbs, st); cudaFreeAsync(w1s, st);
  return s;
}

// Stand-alone fused q|k|v projection + window attention for the unit tests (attn_block.cu): A [n*H*H][C] fp32 on the device
// (the LayerNorm-1 output), out [n*H*H][C] fp32 on the device; the block's reference tensors on the HOST: Wq [C][C], bq [C],
// Wkv [2C][C], bkv [2C], relative_position_bias_table [225][heads].  No output projection.
extern "C" int wmk_window_attention_f32(const float* A, const float* Wq_host, const float* bq_host, const float* Wkv_host,
                                        const float* bkv_host, const float* table_host, float* out, int n, int H, int C, int shift,
                                        void* stream) {
  WMK_REQUIRE(A && Wq_host && bq_host && Wkv_host && bkv_host && table_host && out && n > 0, "window_attention: bad arguments");
  WMK_REQUIRE(C == 32 || C == 64 || C == 128, "window_attention: covers C in {32,64,128}, got %d", C);
  cudaStream_t st = (cudaStream_t)stream;
  const int heads = C / 32;
  const size_t M = (size_t)n * H * H;
  std::vector<float> wqkv, bqkv, bias, wh, bh, bquad;
  build_qkv_operands(Wq_host, bq_host, Wkv_host, bkv_host, table_host, C, heads, true, wqkv, bqkv, bias);
  pack_attn_heads(wqkv, bqkv, bias, C, heads, wh, bh, bquad);
  std::vector<__half> wh16(wh.size()), bq16(bquad.size());
  for (size_t i = 0; i < wh.size(); ++i) wh16[i] = __float2half_rn(wh[i]);
  for (size_t i = 0; i < bquad.size(); ++i) bq16[i] = __float2half_rn(bquad[i]);
  __half *a16 = nullptr, *o16 = nullptr, *w16 = nullptr, *b16 = nullptr;
  float* bh_d = nullptr;
  WMK_CHECK_CUDA(cudaMalloc(&a16, M * C * 2));
  WMK_CHECK_CUDA(cudaMalloc(&o16, M * C * 2));
  WMK_CHECK_CUDA(cudaMalloc(&w16, wh16.size() * 2));
  WMK_CHECK_CUDA(cudaMalloc(&b16, bq16.size() * 2));
  WMK_CHECK_CUDA(cudaMalloc(&bh_d, bh.size() * 4));
  WMK_CHECK_CUDA(cudaMemcpyAsync(w16, wh16.data(), wh16.size() * 2, cudaMemcpyHostToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(b16, bq16.data(), bq16.size() * 2, cudaMemcpyHostToDevice, st));
  WMK_CHECK_CUDA(cudaMemcpyAsync(bh_d, bh.data(), bh.size() * 4, cudaMemcpyHostToDevice, st));
  copy_cols_kernel<__half><<<cdiv(M * (C / 4), 256), 256, 0, st>>>(A, a16, M, C, C, 0);
  WMK_CHECK_LAUNCH("copy_cols_kernel");
  int s = attn_block(a16, w16, bh_d, reinterpret_cast<const uint16_t*>(b16), reinterpret_cast<uint16_t*>(o16), n, H, C, shift, st);
  if (s == 0) {
    widen_kernel<<<cdiv(M * C, 256), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(o16), out, M * C, true);
    count_launch();
  }
  cudaStreamSynchronize(st);                      // the host staging vectors die with this frame
  cudaFree(a16); cudaFree(o16); cudaFree(w16); cudaFree(b16); cudaFree(bh_d);
  return s;
}

extern "C" int wmk_linear_f32(const float* A, const float* W, const float* bias, float* C, int M, int N, int K,
                              int precision, int gelu, void* stream) {
  WMK_REQUIRE(A && W && C && M > 0 && N > 0 && K > 0, "linear: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  GemmArgs g;
  g.bias = bias; g.C = C; g.M = M; g.N = N; g.K = K; g.ldc = N; g.epi = gelu ? EPI_BIAS_GELU : EPI_BIAS; g.out_bf16 = 0;
  if (precision == WMK_PREC_FP32) {
    g.A = A; g.W = W;
    return gemm_fp32_simt(g, st);
  }
  if (precision == WMK_PREC_MIXED) {       // split-bf16 operands: three tcgen05 MMAs per product, fp32 output
    WMK_REQUIRE(K == 32 || K % 64 == 0, "linear(split): K must be 32 or a multiple of 64, got %d", K);
    const int ldw = K == 32 ? 128 : 2 * K;
    __nv_bfloat16 *as = nullptr, *wsp = nullptr;
    WMK_CHECK_CUDA(cudaMallocAsync(&as, (size_t)M * 2 * K * 2, st));
    WMK_CHECK_CUDA(cudaMallocAsync(&wsp, (size_t)N * ldw * 2, st));
    split_rows_kernel<<<cdiv((size_t)M * K, 256), 256, 0, st>>>(A, as, (size_t)M, K, 0);
    WMK_CHECK_LAUNCH("split_rows_kernel");
    split_rows_kernel<<<cdiv((size_t)N * K, 256), 256, 0, st>>>(W, wsp, (size_t)N, K, K == 32 ? 1 : 0);
    WMK_CHECK_LAUNCH("split_rows_kernel");
    g.A = as; g.W = wsp; g.split = 1; g.gelu_exact = 1;
    const int s = gemm_bf16_tcgen05(g, st);
    cudaFreeAsync(as, st);
    cudaFreeAsync(wsp, st);
    return s;
  }
  WMK_REQUIRE(precision == WMK_PREC_BF16 || precision == WMK_PREC_F16 || precision == WMK_LINEAR_WSPLIT,
              "linear: unknown precision %d", precision);
  const bool f16 = precision != WMK_PREC_BF16;
  if (precision == WMK_LINEAR_WSPLIT) {    // fp16 A x (hi + lo) fp16 W: two tcgen05 MMAs per product
    WMK_REQUIRE(K == 32 || K % 64 == 0, "linear(wsplit): K must be 32 or a multiple of 64, got %d", K);
    __half *ah = nullptr, *wh = nullptr;
    WMK_CHECK_CUDA(cudaMallocAsync(&ah, (size_t)M * K * 2, st));
    WMK_CHECK_CUDA(cudaMallocAsync(&wh, (size_t)N * 2 * K * 2, st));
    copy_cols_kernel<__half><<<cdiv((size_t)M * (K / 4), 256), 256, 0, st>>>(A, ah, (size_t)M, K, K, 0);
    WMK_CHECK_LAUNCH("copy_cols_kernel");
    wsplit_rows_kernel<<<cdiv((size_t)N * K, 256), 256, 0, st>>>(W, wh, (size_t)N, K);
    WMK_CHECK_LAUNCH("wsplit_rows_kernel");
    g.A = ah; g.W = wh; g.wsplit = 1; g.gelu_exact = 1;
    uint16_t* c16w = nullptr;
    if (gelu) {
      WMK_CHECK_CUDA(cudaMallocAsync(&c16w, (size_t)M * N * 2, st));
      g.C = c16w; g.out_bf16 = 1;
    }
    const int s = gemm_bf16_tcgen05(g, st);
    if (gelu && s == 0) {
      widen_kernel<<<cdiv((size_t)M * N, 256), 256, 0, st>>>(c16w, C, (size_t)M * N, true);
      count_launch();
    }
    if (c16w) cudaFreeAsync(c16w, st);
    cudaFreeAsync(ah, st);
    cudaFreeAsync(wh, st);
    return s;
  }
  uint16_t *a16 = nullptr, *w16 = nullptr;
  WMK_CHECK_CUDA(cudaMallocAsync(&a16, (size_t)M * K * 2, st));
  WMK_CHECK_CUDA(cudaMallocAsync(&w16, (size_t)N * K * 2, st));
  if (f16) {
    copy_cols_kernel<__half><<<cdiv((size_t)M * (K / 4), 256), 256, 0, st>>>(A, reinterpret_cast<__half*>(a16), (size_t)M, K, K, 0);
    WMK_CHECK_LAUNCH("copy_cols_kernel");
    copy_cols_kernel<__half><<<cdiv((size_t)N * (K / 4), 256), 256, 0, st>>>(W, reinterpret_cast<__half*>(w16), (size_t)N, K, K, 0);
    WMK_CHECK_LAUNCH("copy_cols_kernel");
  } else {
    copy_cols_kernel<__nv_bfloat16><<<cdiv((size_t)M * (K / 4), 256), 256, 0, st>>>(A, reinterpret_cast<__nv_bfloat16*>(a16), (size_t)M, K, K, 0);
    WMK_CHECK_LAUNCH("copy_cols_kernel");
    copy_cols_kernel<__nv_bfloat16><<<cdiv((size_t)N * (K / 4), 256), 256, 0, st>>>(W, reinterpret_cast<__nv_bfloat16*>(w16), (size_t)N, K, K, 0);
    WMK_CHECK_LAUNCH("copy_cols_kernel");
  }
  g.A = a16; g.W = w16; g.f16 = f16;
  uint16_t* c16 = nullptr;
  if (gelu) {        // the GELU epilogue stores 16-bit (as inside the model): widen afterwards
    WMK_CHECK_CUDA(cudaMallocAsync(&c16, (size_t)M * N * 2, st));
    g.C = c16; g.out_bf16 = 1;
  }
  int s = gemm_bf16_tcgen05(g, st);
  if (gelu && s == 0) {
    widen_kernel<<<cdiv((size_t)M * N, 256), 256, 0, st>>>(c16, C, (size_t)M * N, f16);
    count_launch();
  }
  if (c16) cudaFreeAsync(c16, st);
  cudaFreeAsync(a16, st);
  cudaFreeAsync(w16, st);
  return s;
}

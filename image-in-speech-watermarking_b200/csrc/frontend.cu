// STFT / ISTFT front end (n_fft 255, hop 63, rectangular window, centre reflect pad 127,
// 128 one-sided bins) fused with the reference's clip split / concatenation.
// Replaces torch.stft / torch.istft at uformerWM/audio_test.py:315-316,598-600,677-678 and
// uformerWM/model.py:2458,2463, plus the pad / slice / permute of audio_test.py:319-343,681-688.
//
// v2 formulation (HBM bound by design): one CTA = one tile of 32 consecutive frames, lane = frame.
// The 255-point transform is the twiddle-free 15 x 17 prime-factor FFT of dft255.cuh with
// shared-memory staged butterflies: stage A (17 warps, one per residue n2) -> stage B (8 warps,
// one per k1, storing straight to global memory).  Every waveform sample is read from HBM
// once per CTA (the 75% frame overlap is served from shared memory, 9% halo re-read hits L2) and
// every spectrogram value is written / read once, directly in the (2,128,128) clip layout the
// model consumes, as 128-byte row segments.  Algorithmic traffic: 1276 B per frame.
#include <math.h>
#include <stdlib.h>
#include <mutex>

#include "dft255.cuh"
#include "uformer_kernels.cuh"

namespace wmk {

namespace tc { int num_sms(); }

namespace {

constexpr int NFFT = 255, HOP = 63, PAD = 127, BINS = 128;
using dft255::FT;
constexpr int kFrontThreads = 17 * 32;
constexpr int kTileSamples = HOP * (FT - 1) + NFFT;      // 2208 samples feed the 32 frames of a tile
constexpr int kSampFloats = kTileSamples + 8;            // + up to 3 floats of alignment slack, 16-byte multiple

// 1 = two warps per 17-point transform (k1, half): balances the 17-warp CTA across its phases but re-does the
// butterflies; measured slower (both kernels are instruction-issue bound: STFT 3.94 -> 3.64 TB/s), so off.
#ifndef STFT_SPLIT
#define STFT_SPLIT 0
#endif
#ifndef ISTFT_SPLIT
#define ISTFT_SPLIT 0
#endif

__constant__ dft255::Tables c_tab;

int ensure_tables() {
  static bool done[64] = {};
  static std::mutex mu;
  int dev = 0;
  WMK_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (!done[dev]) {
    dft255::Tables t;
    dft255::build_tables(&t);
    WMK_CHECK_CUDA(cudaMemcpyToSymbol(c_tab, &t, sizeof(t)));
    done[dev] = true;
  }
  return 0;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// grid (ceil(4 n_clips / G), B); 544 threads = 17 warps; a CTA walks G consecutive 32-frame tiles of
// one utterance and prefetches the next tile's samples with cp.async while it transforms the
// current one.
//   load : 2208 samples -> smem (16-byte cp.async on interior tiles; reflect padding on edge tiles)
//   A    : warp = residue n2 (17 warps): 15-point real DFTs                -> SA[k1][n2][f]
//   B    : warp = k1 (8 warps): 17-point complex DFTs, results stored straight to the clip rows
//          (lane = frame, so every store instruction writes one 128-byte row segment)
struct StftTile {
  const float* wv; int L, q0, mis, n4; bool fast;
  __device__ StftTile(const float* wv_, int L_, int f0) : wv(wv_), L(L_) {
    q0 = HOP * f0 - PAD;                            // first sample of the tile (centre=True: 127 of padding)
    mis = (int)(((uintptr_t)(wv + q0) >> 2) & 3);   // floats past a 16-byte boundary
    n4 = (kTileSamples + mis + 3) >> 2;
    fast = q0 - mis >= 0 && q0 - mis + 4 * n4 <= L;
  }
  __device__ void prefetch(float* buf, int tid) const {
    const float4* src = reinterpret_cast<const float4*>(wv + q0 - mis);
    for (int i = tid; i < n4; i += kFrontThreads) cp_async16(reinterpret_cast<float4*>(buf) + i, src + i);
  }
  __device__ void load_edge(float* buf, int tid) const {
    float r[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = tid + j * kFrontThreads;
      int q = q0 + i;
      if (q < 0) q = -q;                            // reflect padding
      if (q >= L) q = 2 * (L - 1) - q;
      r[j] = (i < kTileSamples && q >= 0 && q < L) ? wv[q] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int i = tid + j * kFrontThreads;
      if (i < kTileSamples) buf[i] = r[j];
    }
  }
};

template <bool ALL_LIVE, int HALF>
__device__ __forceinline__ void stft_stage_b(const float2* SA, int k1, int lane, float* orow, float live) {
  dft255::fwd_stage_b<HALF>(SA, c_tab.fwd, k1, lane, [&](int bin, float re, float im) {
    float* o = orow + bin * 128;
    o[0] = ALL_LIVE ? re : re * live;
    o[BINS * 128] = ALL_LIVE ? im : im * live;
  });
}

__global__ void __launch_bounds__(kFrontThreads, 2)
stft_clips_kernel(const float* __restrict__ wave, int L, int T, float* __restrict__ clips, int n_clips, int G) {
  extern __shared__ __align__(16) float smem[];
  float2* SA = reinterpret_cast<float2*>(smem + 2 * kSampFloats);        // [8][17][32]
  const int b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* wv = wave + (size_t)b * L;
  const int tile0 = blockIdx.x * G;
  const int n_tiles = min(G, 4 * n_clips - tile0);
  bool prefetched = false;
  if (tile0 * FT < T) {
    StftTile t0(wv, L, tile0 * FT);
    if (t0.fast) { t0.prefetch(smem, tid); prefetched = true; }
  }
  for (int it = 0; it < n_tiles; ++it) {
    const int f0 = (tile0 + it) * FT;
    float* orow = clips + ((size_t)b * n_clips + (f0 >> 7)) * 2 * BINS * 128 + (f0 & 127) + lane;
    if (f0 >= T) {                                  // padding frames of the last clip (audio_test.py:319-320)
      for (int r = warp; r < 2 * BINS; r += 17) orow[(size_t)r * 128] = 0.f;
      continue;
    }
    float* buf = smem + (it & 1) * kSampFloats;
    const StftTile cur(wv, L, f0);
    if (prefetched) cp_async_wait_all();
    else cur.load_edge(buf, tid);
    __syncthreads();                                // samples visible; everyone is done with SA of the last tile
    prefetched = false;
    if (it + 1 < n_tiles && f0 + FT < T) {
      const StftTile nxt(wv, L, f0 + FT);
      if (nxt.fast) { nxt.prefetch(smem + ((it + 1) & 1) * kSampFloats, tid); prefetched = true; }
    }
    dft255::fwd_stage_a(buf + (cur.fast ? cur.mis : 0), SA, warp, lane);
    __syncthreads();
    if (warp < 8) {
      if (f0 + FT <= T) stft_stage_b<true, STFT_SPLIT ? 0 : -1>(SA, warp, lane, orow, 1.f);
      else stft_stage_b<false, STFT_SPLIT ? 0 : -1>(SA, warp, lane, orow, f0 + lane < T ? 1.f : 0.f);
    } else if (STFT_SPLIT && warp < 16) {
      if (f0 + FT <= T) stft_stage_b<true, 1>(SA, warp - 8, lane, orow, 1.f);
      else stft_stage_b<false, 1>(SA, warp - 8, lane, orow, f0 + lane < T ? 1.f : 0.f);
    }
  }
}

// One CTA reconstructs a run of hops of the padded overlap-add buffer in G steps of 32 frames.  A hop needs the four
// (five) frames before it: the FIRST step of a CTA recomputes 4 halo frames (28 hops out of 32 frames, no atomics), every
// later step takes them from the CARRY - the time-domain rows of the previous step's last 4 frames, kept in front of
// the frame buffer - so it turns 32 frames into 32 hops (the halo recompute was 14 % of all frame work at G = 1);
// the next spectrum tile is prefetched with cp.async.
//   load : the 256 x 32 spectrum tile -> XS (lane = frame)
//   B'   : warp = k1 (8 warps): inverse 17-point DFTs        -> ZS[k1][n2][f]
//   A'   : warp = n2 (17 warps): complex-to-real 15-point inverse DFTs -> FR[f][n] (over XS)
//   OLA  : each output sample sums the 4-5 frames that cover it (rows -4 .. -1 of FR = the carry)
constexpr int IFT = FT - 4;
constexpr int kXsFloats = 2 * BINS * FT;
constexpr int kCarryFloats = 1024;                 // 4 rows x 255 floats in front of each frame buffer (16-byte multiple)
constexpr int kIstftBuf = kCarryFloats + kXsFloats;

__device__ __forceinline__ void istft_prefetch(const float* __restrict__ clips_b, int fbase, float* XS, int tid) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tid + j * kFrontThreads;          // 16-byte chunk: row c>>3, frames fbase + 4 (c&7) ..
    if (c < 2 * BINS * 8) {
      const int row = c >> 3, t = fbase + 4 * (c & 7);
      cp_async16(XS + row * FT + 4 * (c & 7), clips_b + ((size_t)(t >> 7) * 2 * BINS + row) * 128 + (t & 127));
    }
  }
}

__global__ void __launch_bounds__(kFrontThreads, 2)
istft_clips_kernel(const float* __restrict__ clips, int n_clips, int T, float* __restrict__ wave, int length, int G,
                   int need_hops) {
  extern __shared__ __align__(16) float smem[];
  float2* ZS = reinterpret_cast<float2*>(smem + 2 * kIstftBuf);         // [8][17][32]
  const int b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* clips_b = clips + (size_t)b * n_clips * 2 * BINS * 128;
  float* wv = wave + (size_t)b * length;
  const int h0 = blockIdx.x * (IFT + FT * (G - 1));                     // first hop of this CTA
  const int hops_left = need_hops - h0;
  int n_it = hops_left <= IFT ? 1 : 1 + (hops_left - IFT + FT - 1) / FT;
  if (n_it > G) n_it = G;
  auto fbase_of = [&](int it) { return it == 0 ? h0 - 4 : h0 + IFT + FT * (it - 1); };   // first frame the step loads
  auto is_interior = [&](int fb) { return fb >= 0 && fb + FT - 1 <= T - 1; };
  bool prefetched = false;
  if (is_interior(fbase_of(0))) { istft_prefetch(clips_b, fbase_of(0), smem + kCarryFloats, tid); prefetched = true; }
  for (int it = 0; it < n_it; ++it) {
    const int fbase = fbase_of(it);
    const int row0 = it == 0 ? 4 : 0, n_hops = FT - row0;               // FR row of the first output hop; hops written
    float* XS = smem + kCarryFloats + (it & 1) * kIstftBuf;             // [256][32], later FR [32][255]; XS[-1020 .. -1] = carry
    if (prefetched) {
      cp_async_wait_all();
    } else {
      const int t = fbase + lane;
      const bool tv = t >= 0 && t < T;
      const int tc = tv ? t : 0;                    // out-of-range frames read a valid address and are zeroed
      const float* src = clips_b + (size_t)(tc >> 7) * 2 * BINS * 128 + (tc & 127) + warp * 128;
      float r[16];
#pragma unroll
      for (int j = 0; j < 15; ++j) r[j] = __ldg(src + j * 17 * 128);     // rows warp + 17 j <= 254
      r[15] = warp == 0 ? __ldg(src + 255 * 128) : 0.f;                   // row 255
      float* dst = XS + warp * FT + lane;
#pragma unroll
      for (int j = 0; j < 15; ++j) dst[j * 17 * FT] = tv ? r[j] : 0.f;
      if (warp == 0) dst[255 * FT] = tv ? r[15] : 0.f;
    }
    __syncthreads();                                // tile (and carry) visible; the other buffer (FR of the last step) is free
    prefetched = false;
    if (it + 1 < n_it && is_interior(fbase_of(it + 1))) {
      istft_prefetch(clips_b, fbase_of(it + 1), smem + kCarryFloats + ((it + 1) & 1) * kIstftBuf, tid);
      prefetched = true;
    }
    if (warp < 8) dft255::inv_stage_b<ISTFT_SPLIT ? 0 : -1>(XS, c_tab.inv, ZS, warp, lane);
    else if (ISTFT_SPLIT && warp < 16) dft255::inv_stage_b<1>(XS, c_tab.inv, ZS, warp - 8, lane);
    __syncthreads();
    float* FR = XS;
    dft255::inv_stage_a(ZS, FR, warp, lane);
    __syncthreads();
    if (it + 1 < n_it) {                            // the next step's carry: rows 28 .. 31 of this step
      float* cdst = smem + kCarryFloats + ((it + 1) & 1) * kIstftBuf - 4 * NFFT;
      for (int i = tid; i < 4 * NFFT; i += kFrontThreads) cdst[i] = FR[IFT * NFFT + i];
    }
    const int hop0 = fbase + row0;                  // first hop written by this step
    if (is_interior(fbase) && PAD + length >= HOP * (fbase + FT)) {
      // 8 groups of 63 threads: thread (g, r) owns offset r of hops g, g+8, g+16, g+24.  Sample p = 63 h' + r
      // is covered by frames h', h'-1, .., h'-3 (and h'-4 when r <= 2) at offsets r, r+63, ..: rows 192 apart.
      if (tid < 8 * HOP) {
        const int g = tid / HOP, r = tid - g * HOP;
        const bool five = r <= 2;
        const float inv = five ? 0.2f : 0.25f;
        const float* fr = FR + (g + row0) * NFFT + r;
        float* o = wv + HOP * (hop0 + g) + r - PAD;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (g + 8 * j < n_hops) {
            const float* q = fr + j * 8 * NFFT;
            float s = (q[0] + q[-192]) + (q[-384] + q[-576]);
            if (five) s += q[-768];
            o[j * 8 * HOP] = s * inv;
          }
        }
      }
    } else {
      for (int i = tid; i < HOP * n_hops; i += kFrontThreads) {
        const int p = HOP * hop0 + i;
        const int jn = p - PAD;
        if (jn < 0 || jn >= length) continue;
        int t_hi = p / HOP;
        if (t_hi > T - 1) t_hi = T - 1;
        int t_lo = (p - (NFFT - 1) + HOP - 1) / HOP;
        if (p - (NFFT - 1) < 0) t_lo = 0;
        float s = 0.f;
        for (int t = t_lo; t <= t_hi; ++t) s += FR[(t - fbase) * NFFT + (p - HOP * t)];
        const int cnt = t_hi - t_lo + 1;
        wv[jn] = cnt > 0 ? s / (float)cnt : 0.f;
      }
    }
  }
}

// tiles (steps) per CTA: 1 until the grid is several waves deep, then up to 4 (prefetch pays once
// there are enough CTAs to keep 2 x 148 resident ones busy)
int tiles_per_cta(long long tiles_total) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const long long per = tiles_total / (2LL * sms * 4);
  return per >= 4 ? 4 : per >= 2 ? 2 : 1;
}

}  // namespace

int stft_clips(const float* wave, int B, int L, float* clips, int n_clips, cudaStream_t st) {
  WMK_REQUIRE(wave && clips && B > 0 && L > PAD && n_clips > 0, "stft: bad arguments (B=%d L=%d n_clips=%d)", B, L, n_clips);
  const int T = 1 + (L - 1) / HOP;
  WMK_REQUIRE(n_clips * 128 >= T, "stft: n_clips=%d cannot hold %d frames", n_clips, T);
  WMK_TRY(ensure_tables());
  const size_t smem = 2 * kSampFloats * sizeof(float) + dft255::SA_FLOAT2 * sizeof(float2);
  static bool attr = false;
  if (!attr) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(stft_clips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  ProfScope prof(FAM_STFT, 1276.0 * T * B, st);
  const int tiles = n_clips * (128 / FT);
  const int G = tiles_per_cta((long long)tiles * B);
  dim3 grid(cdiv(tiles, G), B);
  stft_clips_kernel<<<grid, kFrontThreads, smem, st>>>(wave, L, T, clips, n_clips, G);
  WMK_CHECK_LAUNCH("stft_clips_kernel");
  return 0;
}

int istft_clips(const float* clips, int B, int n_clips, int T, float* wave, int length, cudaStream_t st) {
  WMK_REQUIRE(clips && wave && B > 0 && T > 0 && n_clips * 128 >= T, "istft: bad arguments (B=%d T=%d n_clips=%d)", B, T, n_clips);
  if (length <= 0) length = HOP * (T - 1) + 1;
  WMK_TRY(ensure_tables());
  const int total = NFFT + HOP * (T - 1);
  int need = PAD + length;
  if (need < total) need = total;
  const size_t smem = 2 * kIstftBuf * sizeof(float) + dft255::SA_FLOAT2 * sizeof(float2);
  static bool attr = false;
  if (!attr) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(istft_clips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  ProfScope prof(FAM_ISTFT, 1276.0 * T * B, st);
  const int need_hops = cdiv(need, HOP);
  static const int g_max = getenv("WMK_ISTFT_G") ? atoi(getenv("WMK_ISTFT_G")) : 8;
  const long long per = (long long)cdiv(need_hops, IFT) * B / (2LL * tc::num_sms() * 4);    // steps per resident CTA slot and wave
  int G = per >= 8 ? 8 : per >= 4 ? 4 : per >= 2 ? 2 : 1;
  if (G > g_max) G = g_max;
  dim3 grid(cdiv(need_hops, IFT + FT * (G - 1)), B);
  istft_clips_kernel<<<grid, kFrontThreads, smem, st>>>(clips, n_clips, T, wave, length, G, need_hops);
  WMK_CHECK_LAUNCH("istft_clips_kernel");
  return 0;
}

}  // namespace wmk

extern "C" int wmk_stft_num_frames(int L) { return L > 0 ? 1 + (L - 1) / 63 : 0; }

extern "C" int wmk_stft_clips_f32(const float* wave, int B, int L, float* clips, int n_clips, void* stream) {
  return wmk::stft_clips(wave, B, L, clips, n_clips, (cudaStream_t)stream);
}

extern "C" int wmk_istft_clips_f32(const float* clips, int B, int n_clips, int T, float* wave, int length,
                                   void* stream) {
  return wmk::istft_clips(clips, B, n_clips, T, wave, length, (cudaStream_t)stream);
}

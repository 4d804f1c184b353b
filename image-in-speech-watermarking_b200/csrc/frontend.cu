// STFT / ISTFT front end (n_fft 255, hop 63, rectangular window, centre reflect pad 127,
// 128 one-sided bins) fused with the reference's clip split / concatenation.
// Replaces torch.stft / torch.istft at uformerWM/audio_test.py:315-316,598-600,677-678 and
// uformerWM/model.py:2458,2463, plus the pad / slice / permute of audio_test.py:319-343,681-688.
//
// v2 formulation (HBM bound by design): one CTA = one tile of 32 consecutive frames, lane = frame.
// The 255-point transform is the twiddle-free 15 x 17 prime-factor FFT of dft255.cuh with
// shared-memory staged butterflies: stage A (17 warps, one per residue n2) -> stage B (16 warps,
// one per (half, k1)) -> stage C (bins, warp uniform).  Every waveform sample is read from HBM
// once per CTA (the 75% frame overlap is served from shared memory, 9% halo re-read hits L2) and
// every spectrogram value is written / read once, directly in the (2,128,128) clip layout the
// model consumes, as 128-byte row segments.  Algorithmic traffic: 1276 B per frame.
#include <math.h>
#include <mutex>

#include "dft255.cuh"
#include "uformer_kernels.cuh"

namespace wmk {

namespace {

constexpr int NFFT = 255, HOP = 63, PAD = 127, BINS = 128;
using dft255::FT;
constexpr int kFrontThreads = 17 * 32;
constexpr int kSampFloats = HOP * (FT - 1) + NFFT + 1;   // 2209 (kept even for float2 alignment after it)

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

// grid (frame tiles of 32 = 4 * n_clips, B); 544 threads = 17 warps.
__global__ void __launch_bounds__(kFrontThreads, 2)
stft_clips_kernel(const float* __restrict__ wave, int L, int T, float* __restrict__ clips, int n_clips) {
  extern __shared__ __align__(16) float smem[];
  float* samp = smem;                                                   // [2210]
  float2* SA = reinterpret_cast<float2*>(smem + kSampFloats + 1);        // [8][17][32]
  float2* R = SA + dft255::SA_FLOAT2;                                    // [2][8][9][32]
  const int b = blockIdx.y, f0 = blockIdx.x * FT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* orow = clips + ((size_t)b * n_clips + (f0 >> 7)) * 2 * BINS * 128 + (f0 & 127) + lane;
  if (f0 >= T) {                                    // padding frames of the last clip (audio_test.py:319-320)
    for (int r = warp; r < 2 * BINS; r += 17) orow[(size_t)r * 128] = 0.f;
    return;
  }
  const float* wv = wave + (size_t)b * L;
  for (int i = tid; i < HOP * (FT - 1) + NFFT; i += kFrontThreads) {
    int q = HOP * f0 + i - PAD;
    if (q < 0) q = -q;                              // centre=True reflect padding
    if (q >= L) q = 2 * (L - 1) - q;
    samp[i] = (q >= 0 && q < L) ? wv[q] : 0.f;
  }
  __syncthreads();
  dft255::fwd_stage_a(samp, SA, warp, lane);
  __syncthreads();
  if (warp < 8) dft255::fwd_stage_b<0>(SA, R, warp, lane);
  else if (warp < 16) dft255::fwd_stage_b<1>(SA, R, warp - 8, lane);
  __syncthreads();
  const bool live = f0 + lane < T;
  for (int bin = warp; bin < BINS; bin += 17) {
    const float2 X = dft255::fwd_stage_c(R, c_tab.fwd[bin], lane);
    orow[(size_t)bin * 128] = live ? X.x : 0.f;
    orow[(size_t)(BINS + bin) * 128] = live ? X.y : 0.f;
  }
}

// One CTA reconstructs 28 hops (1764 samples) of the padded overlap-add buffer from 32 frames
// (4 halo frames recomputed instead of atomics), divides by the overlap count and trims.
constexpr int IFT = FT - 4;
__global__ void __launch_bounds__(kFrontThreads, 2)
istft_clips_kernel(const float* __restrict__ clips, int n_clips, int T, float* __restrict__ wave, int length) {
  extern __shared__ __align__(16) float smem[];
  float* XS = smem;                                                     // [256][32], later FR [32][255]
  float2* R = reinterpret_cast<float2*>(smem + 2 * BINS * FT);          // [2][8][9][32]
  const int b = blockIdx.y;
  const int fbase = blockIdx.x * IFT - 4;           // first (halo) frame of this CTA
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const int t = fbase + lane;
    const bool tv = t >= 0 && t < T;
    const float* src = clips + ((size_t)b * n_clips + (tv ? (t >> 7) : 0)) * 2 * BINS * 128 + (t & 127);
    for (int r = warp; r < 2 * BINS; r += 17) XS[r * FT + lane] = tv ? src[(size_t)r * 128] : 0.f;
  }
  __syncthreads();
  if (warp < 8) dft255::inv_stage_b<0>(XS, c_tab.inv, R, warp, lane);
  else if (warp < 16) dft255::inv_stage_b<1>(XS, c_tab.inv, R, warp - 8, lane);
  __syncthreads();
  float* FR = XS;
  dft255::inv_stage_a(R, FR, warp, lane);
  __syncthreads();
  const int p0 = HOP * (fbase + 4);
  float* wv = wave + (size_t)b * length;
  for (int i = tid; i < HOP * IFT; i += kFrontThreads) {
    const int p = p0 + i;
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

}  // namespace

int stft_clips(const float* wave, int B, int L, float* clips, int n_clips, cudaStream_t st) {
  WMK_REQUIRE(wave && clips && B > 0 && L > PAD && n_clips > 0, "stft: bad arguments (B=%d L=%d n_clips=%d)", B, L, n_clips);
  const int T = 1 + (L - 1) / HOP;
  WMK_REQUIRE(n_clips * 128 >= T, "stft: n_clips=%d cannot hold %d frames", n_clips, T);
  WMK_TRY(ensure_tables());
  const size_t smem = (kSampFloats + 1) * sizeof(float) + (dft255::SA_FLOAT2 + dft255::R_FLOAT2) * sizeof(float2);
  static bool attr = false;
  if (!attr) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(stft_clips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  ProfScope prof(FAM_STFT, 1276.0 * T * B, st);
  dim3 grid(n_clips * (128 / FT), B);
  stft_clips_kernel<<<grid, kFrontThreads, smem, st>>>(wave, L, T, clips, n_clips);
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
  const size_t smem = 2 * BINS * FT * sizeof(float) + dft255::R_FLOAT2 * sizeof(float2);
  static bool attr = false;
  if (!attr) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(istft_clips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  ProfScope prof(FAM_ISTFT, 1276.0 * T * B, st);
  dim3 grid(cdiv(need, HOP * IFT), B);
  istft_clips_kernel<<<grid, kFrontThreads, smem, st>>>(clips, n_clips, T, wave, length);
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

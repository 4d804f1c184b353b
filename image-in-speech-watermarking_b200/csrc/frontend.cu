// STFT / ISTFT front end (n_fft 255, hop 63, rectangular window, centre reflect pad 127,
// 128 one-sided bins) fused with the reference's clip split / concatenation.
// Replaces torch.stft / torch.istft at uformerWM/audio_test.py:315-316,598-600,677-678 and
// uformerWM/model.py:2458,2463, plus the pad / slice / permute of audio_test.py:319-343,681-688.
//
// v1 formulation: the 255-point real DFT of 64 overlapping frames is a small fp32 matrix product
// against a constant twiddle matrix that lives in L2.  Every waveform sample is read from HBM once
// per CTA (the 75% frame overlap is served from shared memory) and every spectrogram value is
// written / read once, in the clip layout the model consumes, as 256-byte rows.
#include <math.h>
#include <mutex>
#include <vector>

#include "uformer_kernels.cuh"

namespace wmk {

namespace {

constexpr int NFFT = 255, HOP = 63, PAD = 127, BINS = 128;

struct Twiddles {
  float* fwd = nullptr;   // [256 n][256 j]: j<128 cos(2 pi j n/255); j>=128 -sin(2 pi (j-128) n/255); row 255 = 0
  float* inv = nullptr;   // [256 j][256 n]: irfft weights / 255 (imag of DC ignored); column 255 = 0
};

int get_twiddles(Twiddles* out) {
  static Twiddles tw[64];
  static std::mutex mu;
  int dev = 0;
  WMK_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (!tw[dev].fwd) {
    std::vector<float> f(256 * 256, 0.f), iv(256 * 256, 0.f);
    for (int n = 0; n < NFFT; ++n)
      for (int k = 0; k < BINS; ++k) {
        const double ang = 2.0 * M_PI * (double)((k * n) % NFFT) / NFFT;
        f[n * 256 + k] = (float)cos(ang);
        f[n * 256 + 128 + k] = (float)(-sin(ang));
        const double wk = (k == 0 ? 1.0 : 2.0) / NFFT;
        iv[k * 256 + n] = (float)(wk * cos(ang));
        iv[(128 + k) * 256 + n] = (float)(k == 0 ? 0.0 : -wk * sin(ang));
      }
    float *df = nullptr, *di = nullptr;
    WMK_CHECK_CUDA(cudaMalloc(&df, f.size() * 4));
    WMK_CHECK_CUDA(cudaMalloc(&di, iv.size() * 4));
    WMK_CHECK_CUDA(cudaMemcpy(df, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    WMK_CHECK_CUDA(cudaMemcpy(di, iv.data(), iv.size() * 4, cudaMemcpyHostToDevice));
    tw[dev].fwd = df;
    tw[dev].inv = di;
  }
  *out = tw[dev];
  return 0;
}

// grid (frame tiles of 64, 4 column tiles of 64, B); 256 threads, 4x4 outputs per thread.
__global__ void __launch_bounds__(256)
stft_clips_kernel(const float* __restrict__ wave, int L, int T, float* __restrict__ clips, int n_clips,
                  const float* __restrict__ tw) {
  __shared__ __align__(16) float samp[63 * 63 + 256 + 8];
  __shared__ __align__(16) float Bs[16][68];
  __shared__ float Cs[64][65];
  const int b = blockIdx.z, j0 = blockIdx.y * 64, f0 = blockIdx.x * 64;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][4] = {};
  if (f0 < T) {
    const float* wv = wave + (size_t)b * L;
    for (int i = tid; i < 63 * 63 + 256; i += 256) {
      int q = HOP * f0 + i - PAD;
      if (q < 0) q = -q;
      if (q >= L) q = 2 * (L - 1) - q;
      samp[i] = (q >= 0 && q < L) ? wv[q] : 0.f;
    }
    for (int k0 = 0; k0 < 256; k0 += 16) {
      __syncthreads();
      for (int e = tid; e < 16 * 64; e += 256) Bs[e >> 6][e & 63] = __ldg(tw + (size_t)(k0 + (e >> 6)) * 256 + j0 + (e & 63));
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < 16; ++kk) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = samp[HOP * (ty * 4 + i) + k0 + kk];
        const float4 w4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float wv4[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], wv4[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  for (int e = tid; e < 64 * 64; e += 256) {
    const int col = e >> 6, row = e & 63;
    const int frame = f0 + row, j = j0 + col;
    const int clip = frame >> 7;
    if (clip >= n_clips) continue;
    const int reim = j >> 7, bin = j & 127;
    clips[((((size_t)b * n_clips + clip) * 2 + reim) * 128 + bin) * 128 + (frame & 127)] =
        frame < T ? Cs[row][col] : 0.f;
  }
}

// One CTA reconstructs 60 hops (3780 samples) of the padded overlap-add buffer from 64 frames
// (4 halo frames recomputed instead of atomics), divides by the overlap count and trims.
constexpr int IFT = 60;
__global__ void __launch_bounds__(256)
istft_clips_kernel(const float* __restrict__ clips, int n_clips, int T, float* __restrict__ wave, int length,
                   const float* __restrict__ iw) {
  extern __shared__ float sm[];
  float* As = sm;                    // [16][64]
  float* Bs = sm + 16 * 64;          // [16][256]
  float* Fs = Bs + 16 * 256;         // [64][257]
  const int b = blockIdx.y;
  const int fbase = blockIdx.x * IFT - 4;   // first (halo) frame of this CTA
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < 256; k0 += 16) {
    __syncthreads();
    for (int e = tid; e < 16 * 64; e += 256) {
      const int kk = e >> 6, m = e & 63;
      const int t = fbase + m, j = k0 + kk;
      float v = 0.f;
      if (t >= 0 && t < T)
        v = clips[((((size_t)b * n_clips + (t >> 7)) * 2 + (j >> 7)) * 128 + (j & 127)) * 128 + (t & 127)];
      As[kk * 64 + m] = v;
    }
    for (int e = tid; e < 16 * 256; e += 256) Bs[e] = __ldg(iw + (size_t)k0 * 256 + e);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk * 64 + ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      float w[16];
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 t4 = *reinterpret_cast<const float4*>(&Bs[kk * 256 + j4 * 64 + tx * 4]);
        w[j4 * 4] = t4.x; w[j4 * 4 + 1] = t4.y; w[j4 * 4 + 2] = t4.z; w[j4 * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4)
#pragma unroll
      for (int j = 0; j < 4; ++j) Fs[(ty * 4 + i) * 257 + j4 * 64 + tx * 4 + j] = acc[i][j4 * 4 + j];
  __syncthreads();
  const int p0 = HOP * (fbase + 4);
  float* wv = wave + (size_t)b * length;
  for (int i = tid; i < HOP * IFT; i += 256) {
    const int p = p0 + i;
    const int jn = p - PAD;
    if (jn < 0 || jn >= length) continue;
    int t_hi = p / HOP;
    if (t_hi > T - 1) t_hi = T - 1;
    int t_lo = (p - (NFFT - 1) + HOP - 1) / HOP;
    if (p - (NFFT - 1) < 0) t_lo = 0;
    float s = 0.f;
    for (int t = t_lo; t <= t_hi; ++t) s += Fs[(t - fbase) * 257 + (p - HOP * t)];
    const int cnt = t_hi - t_lo + 1;
    wv[jn] = cnt > 0 ? s / (float)cnt : 0.f;
  }
}

}  // namespace

int stft_clips(const float* wave, int B, int L, float* clips, int n_clips, cudaStream_t st) {
  WMK_REQUIRE(wave && clips && B > 0 && L > PAD && n_clips > 0, "stft: bad arguments (B=%d L=%d n_clips=%d)", B, L, n_clips);
  const int T = 1 + (L - 1) / HOP;
  WMK_REQUIRE(n_clips * 128 >= T, "stft: n_clips=%d cannot hold %d frames", n_clips, T);
  Twiddles tw;
  WMK_TRY(get_twiddles(&tw));
  ProfScope prof(FAM_STFT, 1276.0 * T * B, st);
  dim3 grid(n_clips * 2, 4, B);
  stft_clips_kernel<<<grid, 256, 0, st>>>(wave, L, T, clips, n_clips, tw.fwd);
  WMK_CHECK_LAUNCH("stft_clips_kernel");
  return 0;
}

int istft_clips(const float* clips, int B, int n_clips, int T, float* wave, int length, cudaStream_t st) {
  WMK_REQUIRE(clips && wave && B > 0 && T > 0 && n_clips * 128 >= T, "istft: bad arguments (B=%d T=%d n_clips=%d)", B, T, n_clips);
  if (length <= 0) length = HOP * (T - 1) + 1;
  Twiddles tw;
  WMK_TRY(get_twiddles(&tw));
  const int total = NFFT + HOP * (T - 1);
  int need = PAD + length;
  if (need < total) need = total;
  const size_t smem = (16 * 64 + 16 * 256 + 64 * 257) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(istft_clips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  ProfScope prof(FAM_ISTFT, 1276.0 * T * B, st);
  dim3 grid(cdiv(need, HOP * IFT), B);
  istft_clips_kernel<<<grid, 256, smem, st>>>(clips, n_clips, T, wave, length, tw.inv);
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

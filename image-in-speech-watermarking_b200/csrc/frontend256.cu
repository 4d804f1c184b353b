// Training-time analysis STFT of the reference's dataset front end
// (uformerWM/audio_test.py:465-491, SpeechDataTrain.prepare_data):
//     torch.stft(x, n_fft=256, hop_length=128, win_length=256)   rectangular window, centre reflect pad 128
//     drop the Nyquist row (129 -> 128 bins), zero-pad the frames to a multiple of 128, cut 128-frame clips
// written straight into the (2,128,128) clip layout the model consumes, plus the global min / max of
// normalize_batch (audio_test.py:33-55).
//
// HBM-bound (1536 B per frame: 128 new samples in, 128 complex bins out).  One CTA = 32 consecutive frames
// of one utterance, lane = frame, so every shared-memory access is conflict free and every global store
// is a 128-byte row segment of the clip.  256 = 16 x 16 Cooley-Tukey: pass A (per n2: 16-point DFT over
// the stride-16 samples; real input: only k1 = 0..8 go to shared memory, 54 KB per CTA, 4 CTAs per SM) ->
// pass B (per k1: W256^(n2 k1) twiddle, 16-point DFT over n2) -> registers -> global; each 16-point DFT is two radix-4 stages in registers.  The frame overlap
// (hop 128 = 0 mod 32 banks) is broken by skewing the staged waveform by one word per hop.
#include <math.h>

#include "dft256.cuh"
#include "tc_ptx.cuh"

namespace wmk {

namespace {

constexpr int F256_TILE = 32;                      // frames per CTA
constexpr int F256_THREADS = 256;
constexpr int F256_SPAN = (F256_TILE + 1) * 128;   // samples covered by 32 frames
constexpr int F256_SPAN_WORDS = F256_SPAN + F256_TILE + 2;

__constant__ float2 c_tw256[256];                  // e^{-2 pi i m / 256}

using namespace dft256;

__global__ void __launch_bounds__(F256_THREADS)
stft256_clips_kernel(const float* __restrict__ wave, float* __restrict__ clips, int L, int T, int n_clips) {
  extern __shared__ float smem256[];
  float* span = smem256;                                 // skewed waveform span
  float* wre = smem256 + F256_SPAN_WORDS;                // [16 n2][9 k1][32]: Hermitian half of pass A
  float* wim = wre + 144 * F256_TILE;
  const int tiles_per_utt = n_clips * (128 / F256_TILE);
  const int b = blockIdx.x / tiles_per_utt;
  const int tile = blockIdx.x - b * tiles_per_utt;
  const int t0 = tile * F256_TILE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* x = wave + (size_t)b * L;

  // stage samples 128 t0 - 128 .. 128 (t0 + 32) + 127 with torch's reflect padding (centre = True)
  const int g0 = 128 * t0 - 128;
  {
    constexpr int NLD = (F256_SPAN + F256_THREADS - 1) / F256_THREADS;     // 17 independent loads in flight
    float r[NLD];
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * F256_THREADS;
      int g = g0 + i;
      if (g < 0) g = -g;
      if (g >= L) g = 2 * (L - 1) - g;
      r[u] = (i < F256_SPAN && g >= 0 && g < L) ? __ldg(x + g) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < NLD; ++u) {
      const int i = tid + u * F256_THREADS;
      if (i < F256_SPAN) span[i + (i >> 7)] = r[u];
    }
  }
  __syncthreads();

  // pass A: n2 = 2 warp, 2 warp + 1: DFT16 over n1 of x[16 n1 + n2], times W256^(n2 k1) -> work[16 n2 + k1]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n2 = 2 * warp + h;
    float xs[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int n = 16 * n1 + n2;                        // sample of frame `lane`: span index 128 lane + n, skewed
      xs[n1] = span[129 * lane + n + (n >> 7)];
    }
    c32 v[16];
    pass_a(xs, v);
#pragma unroll
    for (int k1 = 0; k1 <= 8; ++k1) {
      wre[(9 * n2 + k1) * F256_TILE + lane] = v[k1].re;
      wim[(9 * n2 + k1) * F256_TILE + lane] = v[k1].im;
    }
  }
  __syncthreads();

  // pass B: k1 = 2 warp, 2 warp + 1: DFT16 over n2 -> X[k1 + 16 k2]; bins 0..127 = k2 < 8
  const int t = t0 + lane;
  const bool live = t < T;
  const int clip = t0 >> 7, tin = (t0 & 127) + lane;
  float* ore = clips + ((size_t)(b * n_clips + clip) * 2) * 16384 + tin;
  float* oim = ore + 16384;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k1 = 2 * warp + h;
    const int ks = pass_b_src(k1);
    c32 v[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
      v[n2].re = wre[(9 * n2 + ks) * F256_TILE + lane];
      v[n2].im = wim[(9 * n2 + ks) * F256_TILE + lane];
    }
    pass_b(v, k1, c_tw256);
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) {
      const int k = k1 + 16 * k2;
      ore[(size_t)k * 128] = live ? v[k2].re : 0.f;
      oim[(size_t)k * 128] = live ? v[k2].im : 0.f;
    }
  }
}

// global min / max of a tensor (normalize_batch, audio_test.py:35-37): order-preserving integer keys
__device__ __forceinline__ int f2key(float f) { const int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void minmax_init_kernel(int* keys) { keys[0] = 0x7fffffff; keys[1] = (int)0x80000000; }

__global__ void __launch_bounds__(256)
minmax_kernel(const float* __restrict__ x, size_t n, int* keys) {
  float lo = INFINITY, hi = -INFINITY;
  const size_t n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
    hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - 4 * n4)) {
    const float v = x[4 * n4 + threadIdx.x];
    lo = fminf(lo, v); hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(keys, f2key(lo));
    atomicMax(keys + 1, f2key(hi));
  }
}

__global__ void minmax_finish_kernel(const int* keys, float* out) { out[0] = key2f(keys[0]); out[1] = key2f(keys[1]); }

int init_tw256() {
  static int done = 0;
  static int status = 0;
  if (done) return status;
  float2 h[256];
  for (int m = 0; m < 256; ++m) {
    const double a = -2.0 * 3.14159265358979323846 * m / 256.0;
    h[m] = make_float2((float)cos(a), (float)sin(a));
  }
  if (cudaMemcpyToSymbol(c_tw256, h, sizeof(h)) != cudaSuccess) {
    set_error("stft256: twiddle upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    status = WMK_ERR_CUDA;
  }
  done = 1;
  return status;
}

}  // namespace

int stft256_clips(const float* wave, int B, int L, float* clips, int n_clips, cudaStream_t st) {
  WMK_REQUIRE(wave && clips, "stft256: null buffer");
  WMK_REQUIRE(B > 0 && L > 128, "stft256: need B > 0 and L > 128 (reflect padding), got B=%d L=%d", B, L);
  const int T = 1 + L / 128;
  WMK_REQUIRE(n_clips * 128 >= T, "stft256: %d clips cannot hold %d frames", n_clips, T);
  WMK_REQUIRE((long long)B * n_clips * 4 < (1LL << 31), "stft256: too many tiles");
  WMK_TRY(init_tw256());
  const size_t smem = (size_t)(F256_SPAN_WORDS + 2 * 144 * F256_TILE) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    WMK_CHECK_CUDA(cudaFuncSetAttribute(stft256_clips_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  ProfScope prof(FAM_STFT, (double)B * T * 1536.0, st);
  stft256_clips_kernel<<<B * n_clips * (128 / F256_TILE), F256_THREADS, smem, st>>>(wave, clips, L, T, n_clips);
  WMK_CHECK_LAUNCH("stft256_clips_kernel");
  return 0;
}

int minmax_f32(const float* x, size_t n, float* out2, int* scratch2, cudaStream_t st) {
  WMK_REQUIRE(x && out2 && scratch2 && n > 0, "minmax: bad arguments");
  WMK_REQUIRE(((uintptr_t)x & 15) == 0, "minmax: input must be 16-byte aligned");
  minmax_init_kernel<<<1, 1, 0, st>>>(scratch2);
  size_t blocks = (n / 4 + 255) / 256;
  const size_t cap = (size_t)tc::num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  minmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, scratch2);
  minmax_finish_kernel<<<1, 1, 0, st>>>(scratch2, out2);
  WMK_CHECK_LAUNCH("minmax_kernel");
  return 0;
}

}  // namespace wmk

extern "C" int wmk_stft256_num_frames(int L) { return L > 0 ? 1 + L / 128 : 0; }

extern "C" int wmk_stft256_clips_f32(const float* wave, int B, int L, float* clips, int n_clips, void* stream) {
  return wmk::stft256_clips(wave, B, L, clips, n_clips, (cudaStream_t)stream);
}

extern "C" int wmk_minmax_f32(const float* x, size_t n, float* out2, void* scratch8, void* stream) {
  return wmk::minmax_f32(x, n, out2, reinterpret_cast<int*>(scratch8), (cudaStream_t)stream);
}

"""STFT / ISTFT front end of the uformer trainer / evaluator, on the GPU.

Drop-in for the `torch.stft(x, n_fft=255)` / `torch.istft(spec, n_fft=255, length=L)` call
sites of the reference (`uformerWM/audio_test.py:315-316,598-600,677-678`,
`uformerWM/model.py:2458,2463`; the model they serve is named by
`uformerWM/audio_uformer_stft.py:452`) fused with the 128-frame clip split
(`uformerWM/audio_test.py:319-343`)."""
import torch

from . import _lib

N_FFT, HOP, BINS, CLIP = 255, 63, 128, 128


def num_frames(L):
    return 1 + (L - 1) // HOP


def stft_clips(wave, n_clips=None, out=None):
    """wave (B, L) float32 CUDA -> clips (B, n_clips, 2, 128, 128).  Default n_clips is the
    reference's `T // 128 + 1` (`audio_test.py:319-325`, incl. the empty clip when T % 128 == 0).
    `out`: a contiguous float32 buffer of B * n_clips clips to write into (no allocation)."""
    lib = _lib.load()
    if wave.dim() == 1:
        wave = wave[None]
    wave = wave.contiguous().float()
    B, L = wave.shape
    T = num_frames(L)
    if n_clips is None:
        n_clips = T // CLIP + 1
    if out is None:
        out = torch.empty((B, n_clips, 2, BINS, CLIP), device=wave.device, dtype=torch.float32)
    elif out.numel() != B * n_clips * 2 * BINS * CLIP or not out.is_contiguous() or out.dtype != torch.float32:
        raise ValueError("stft_clips: `out` must be a contiguous float32 buffer of %d clips" % (B * n_clips))
    _lib.check(lib.wmk_stft_clips_f32(_lib.ptr(wave), B, L, _lib.ptr(out), n_clips, _lib.stream_ptr()))
    return out.view(B, n_clips, 2, BINS, CLIP)


def istft_clips(clips, T, length=None):
    """clips (B, n_clips, 2, 128, 128) -> wave (B, length): `torch.istft(n_fft=255, length=...)` of
    the first T frames of the concatenated clips (`audio_test.py:595-600`)."""
    lib = _lib.load()
    clips = clips.contiguous().float()
    B, n_clips = clips.shape[0], clips.shape[1]
    if length is None:
        length = HOP * (T - 1) + 1
    out = torch.empty((B, length), device=clips.device, dtype=torch.float32)
    _lib.check(lib.wmk_istft_clips_f32(_lib.ptr(clips), B, n_clips, T, _lib.ptr(out), length, _lib.stream_ptr()))
    return out


def stft(x, n_fft=N_FFT, return_complex=False):
    """`torch.stft(x, n_fft=255)` legacy layout: (B, L) -> (B, 128, T, 2) (or complex)."""
    if n_fft != N_FFT:
        raise ValueError("the CUDA front end implements n_fft=255 (hop 63) only")
    squeeze = x.dim() == 1
    xs = x[None] if squeeze else x
    T = num_frames(xs.shape[-1])
    c = stft_clips(xs, (T + CLIP - 1) // CLIP)                      # (B, nc, 2, F, 128)
    s = c.permute(0, 3, 1, 4, 2).reshape(c.shape[0], BINS, -1, 2)[:, :, :T, :].contiguous()
    if squeeze:
        s = s[0]
    return torch.view_as_complex(s) if return_complex else s


def istft(spec, n_fft=N_FFT, length=None, return_complex=False):
    """`torch.istft(spec, n_fft=255[, length])` on the legacy (…, 128, T, 2) layout (or complex)."""
    if n_fft != N_FFT:
        raise ValueError("the CUDA front end implements n_fft=255 (hop 63) only")
    if torch.is_complex(spec):
        spec = torch.view_as_real(spec)
    squeeze = spec.dim() == 3
    s = spec[None] if squeeze else spec
    B, F, T, _ = s.shape
    nc = (T + CLIP - 1) // CLIP
    pad = nc * CLIP - T
    s = torch.nn.functional.pad(s.float(), (0, 0, 0, pad))
    clips = s.reshape(B, F, nc, CLIP, 2).permute(0, 2, 4, 1, 3).contiguous()
    w = istft_clips(clips, T, length)
    return w[0] if squeeze else w


# ---------------------------------------------------------------------------------------------
# training-time analysis (`SpeechDataTrain.prepare_data`, `uformerWM/audio_test.py:465-491`)
N_FFT_TRAIN, HOP_TRAIN = 256, 128


def num_frames_train(L):
    """frames of `torch.stft(x, n_fft=256, hop_length=128, win_length=256)`: 1 + L // 128."""
    return 1 + L // HOP_TRAIN


def stft256_clips(wave, n_clips=None):
    """wave (B, L) float32 CUDA -> clips (B, n_clips, 2, 128, 128) of the training-time STFT: n_fft 256,
    hop 128, rectangular window, centre reflect padding, Nyquist row dropped (`audio_test.py:465-469`),
    frames zero-padded and cut into 128-frame clips (`:475-487`).  Default n_clips is the reference's
    `T // 128 + 1` (its `len_pad = 128 - T % 128` appends a whole empty clip when T % 128 == 0)."""
    lib = _lib.load()
    if wave.dim() == 1:
        wave = wave[None]
    wave = wave.contiguous().float()
    B, L = wave.shape
    T = num_frames_train(L)
    if n_clips is None:
        n_clips = T // CLIP + 1
    out = torch.empty((B, n_clips, 2, BINS, CLIP), device=wave.device, dtype=torch.float32)
    _lib.check(lib.wmk_stft256_clips_f32(_lib.ptr(wave), B, L, _lib.ptr(out), n_clips, _lib.stream_ptr()))
    return out


_TAPS = {}


def resample_poly(wave, up, down):
    """`scipy.signal.resample_poly(wave, up, down)` (Kaiser-5 windowed-sinc FIR of 20 max(up, down) + 1 taps) on the
    GPU: wave (B, L) or (L,) CUDA fp32 -> (B, ceil(L up / down)).  The tap design is host-side scalar work, cached
    per (up, down) on the device; corpora that are not 16 kHz go through this before the STFT."""
    import math
    import numpy as np
    from scipy import signal
    g = math.gcd(int(up), int(down))
    up, down = int(up) // g, int(down) // g
    w = wave.float().contiguous()
    w = w[None] if w.dim() == 1 else w
    if up == down:
        return w
    key = (up, down, str(w.device))
    if key not in _TAPS:
        mx = max(up, down)
        h = signal.firwin(2 * 10 * mx + 1, 1.0 / mx, window=("kaiser", 5.0)) * up
        _TAPS[key] = torch.from_numpy(np.ascontiguousarray(h, np.float32)).to(w.device)
    h = _TAPS[key]
    B, L = w.shape
    Lout = -(-L * up // down)
    out = torch.empty((B, Lout), device=w.device, dtype=torch.float32)
    _lib.check(_lib.load().wmk_resample_poly_f32(_lib.ptr(w), _lib.ptr(out), B, L, Lout, up, down, _lib.ptr(h), h.numel(),
                                                 _lib.stream_ptr()))
    return out


def load_utterance(path, target_sr=16000):
    """WAV file -> (1, L) CUDA waveform at `target_sr`: header parsed on the host, samples decoded, mixed down to
    mono and resampled on the device - the loader step in front of `stft_clips` (`uformerWM/audio_test.py:299-316`
    gets this from torchaudio's dataset classes)."""
    from . import wavio
    x, sr = wavio.read_wav_cuda(path)
    x = x.mean(0, keepdim=True) if x.shape[0] > 1 else x
    return resample_poly(x, target_sr, sr) if sr != target_sr else x


def minmax(x):
    """(min, max) of a CUDA float32 tensor as a 2-element device tensor (`normalize_batch`, `audio_test.py:35-37`)."""
    lib = _lib.load()
    x = x.contiguous().float()
    out = torch.empty(2, device=x.device, dtype=torch.float32)
    scratch = torch.empty(2, device=x.device, dtype=torch.int32)
    _lib.check(lib.wmk_minmax_f32(_lib.ptr(x), x.numel(), _lib.ptr(out), _lib.ptr(scratch), _lib.stream_ptr()))
    return out


"""numpy float64 restatement of the signal front end, waveform attacks and metrics.

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
import math

import numpy as np
from scipy import signal as sps

N_FFT = 255
HOP = 63
PAD = 127
BINS = 128
CLIP = 128


# --------------------------------------------------------------------------- STFT / ISTFT
def stft(x, n_fft=N_FFT):
    """`torch.stft(x, n_fft=255)` defaults, call sites `uformerWM/audio_test.py:315-316,677-678`,
    `uformerWM/model.py:2463`: hop n_fft//4, rectangular window, centre reflect pad n_fft//2,
    one-sided, unnormalised.  x: (..., L) -> (..., bins, T, 2)."""
    x = np.asarray(x, dtype=np.float64)
    hop = n_fft // 4
    pad = n_fft // 2
    xp = np.pad(x, [(0, 0)] * (x.ndim - 1) + [(pad, pad)], mode="reflect")
    T = 1 + (x.shape[-1] + 2 * pad - n_fft) // hop
    idx = hop * np.arange(T)[:, None] + np.arange(n_fft)[None, :]
    frames = xp[..., idx]                                   # (..., T, n_fft)
    X = np.fft.rfft(frames, n=n_fft, axis=-1)               # (..., T, bins)
    X = np.swapaxes(X, -1, -2)
    return np.stack([X.real, X.imag], axis=-1)


def istft(spec, n_fft=N_FFT, length=None):
    """`torch.istft(spec, n_fft=255[, length])`, call sites `uformerWM/audio_test.py:598-600`,
    `uformerWM/model.py:2458`.  spec: (..., bins, T, 2) -> (..., L)."""
    spec = np.asarray(spec, dtype=np.float64)
    hop = n_fft // 4
    pad = n_fft // 2
    X = spec[..., 0] + 1j * spec[..., 1]
    T = X.shape[-1]
    fr = np.fft.irfft(np.swapaxes(X, -1, -2), n=n_fft, axis=-1)    # (..., T, n_fft)
    total = n_fft + hop * (T - 1)
    y = np.zeros(X.shape[:-2] + (total,))
    env = np.zeros(total)
    for t in range(T):
        y[..., hop * t:hop * t + n_fft] += fr[..., t, :]
        env[hop * t:hop * t + n_fft] += 1.0
    end = total - pad if length is None else pad + length
    y = y[..., pad:min(end, total)]
    env = env[pad:min(end, total)]
    y = y / env
    if length is not None and y.shape[-1] < length:
        y = np.pad(y, [(0, 0)] * (y.ndim - 1) + [(0, length - y.shape[-1])])
    return y


def stft_train(x):
    """Training-time analysis `torch.stft(x, n_fft=256, hop_length=128, win_length=256)` with the Nyquist row
    dropped (`uformerWM/audio_test.py:465-469`): rectangular window, centre reflect pad 128, T = 1 + L // 128.
    x: (L,) -> (128, T, 2) float64."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    xp = np.pad(x, (128, 128), mode="reflect")
    T = 1 + x.shape[0] // 128
    idx = 128 * np.arange(T)[:, None] + np.arange(256)[None, :]
    X = np.fft.rfft(xp[idx], n=256, axis=-1)[:, :128].T          # (128, T)
    return np.stack([X.real, X.imag], axis=-1)


def prepare_data_train(waves, audio_scale="0"):
    """`SpeechDataTrain.prepare_data` + `normalize_batch` (`uformerWM/audio_test.py:33-55,439-502`): per utterance
    stft_train, zero-pad by `128 - T % 128` frames (a whole empty clip when T % 128 == 0), cut 128-frame clips;
    then the dataset-wide scaling.  Returns (data (N, 2, 128, 128) float64 in the (re/im, bin, frame) layout the
    models consume, min, max) - min = max = 0 when len(audio_scale) <= 1 (`:489-499`).  The 'a-b' branch uses
    the GLOBAL min / max (the reference's `view(c,1,1,1,1)` of those scalars only runs for a one-clip dataset)."""
    clips = []
    for w in waves:
        s = stft_train(w)
        T = s.shape[1]
        s = np.pad(s, ((0, 0), (0, 128 - T % 128), (0, 0)))
        for j in range(s.shape[1] // 128):
            clips.append(np.transpose(s[:, 128 * j:128 * (j + 1), :], (2, 0, 1)))
    data = np.stack(clips)
    a = str(audio_scale)
    if len(a) <= 1:
        return data, 0, 0
    mn, mx = data.min(), data.max()
    if "-" not in a:
        return data * float(a), mn, mx
    lo, hi = (float(v) for v in a.split("-"))
    return (data - mn) / (mx - mn) * (hi - lo) + lo, mn, mx


def clip_spectrogram(spec):
    """`SpeechDataTest.prepare_data` `uformerWM/audio_test.py:319-347`.
    spec (1,bins,T,2) -> list of (2,128,128) clips, len_last_clip.  Keeps quirk B-6:
    a full empty clip is appended when T % 128 == 0."""
    T = spec.shape[2]
    len_pad = CLIP - T % CLIP
    s2 = np.pad(spec, [(0, 0), (0, 0), (0, len_pad), (0, 0)])
    clips = [np.transpose(s2[:, :, CLIP * j:CLIP * (j + 1), :], (0, 3, 1, 2))[0]
             for j in range(s2.shape[2] // CLIP)]
    return clips, T % CLIP


def attacked_clips(audio_att, n_fft=N_FFT):
    """`reconstruct_audio` `uformerWM/audio_test.py:677-688`.  Keeps quirk B-7: the pad length
    is computed from the re/im axis (size 2) of the 3-D attacked spectrogram -> always 126."""
    feat = stft(audio_att, n_fft)                           # (bins, T, 2)
    len_pad = 128 - feat.shape[2] % 128                     # == 126
    feat = np.pad(feat, [(0, 0), (0, len_pad), (0, 0)])
    feat = np.transpose(feat, (2, 0, 1))[None]              # (1,2,bins,T')
    return [feat[:, :, :, 128 * j:128 * (j + 1)].astype(np.float32)
            for j in range(feat.shape[3] // 128)]


# --------------------------------------------------------------------------- attacks
def low_pass_filter(x, Fs=16000, low_pass_parameter=8000):
    """`uformerWM/audio_attack.py:21-30`."""
    wn = 2 * low_pass_parameter / (Fs * 2)
    b, a = sps.butter(8, wn, "lowpass")
    return sps.filtfilt(b, a, x)


def echo_addition(x, Fs=16000, td=0.5, AA=0.2):
    """`uformerWM/audio_attack.py:33-52`."""
    d = int(td * Fs)
    echo = np.append(np.zeros([d, 1]), AA * x[0:int(len(x) - td * Fs)])
    return x + echo


def amplitude_scaling(x, factor=0.8):
    """`uformerWM/audio_attack.py:55-58`."""
    return x * float(factor)


def closed_loop(x):
    """`uformerWM/audio_attack.py:67-69`."""
    return x


def awgn(x, snr=15, noise_unit=None, rng=None):
    """`uformerWM/audio_attack.py:99-125`.  ``noise_unit`` injects the N(0,1) draws
    (np.random.normal(0, s, shape) == s * N(0,1)) so the CUDA path can be compared."""
    p = np.mean(x ** 2)
    p_db = 10 * np.log10(p)
    n_db = p_db - snr
    n_p = 10 ** (n_db / 10)
    if noise_unit is None:
        noise_unit = (rng or np.random).standard_normal(x.shape)
    return x + np.sqrt(n_p) * noise_unit


def jittering_2(x, jit_ratio=1000, indices=None, rng=None):
    """`uformerWM/audio_attack.py:176-193` (zeroes samples; the reference does it in place)."""
    import random
    if indices is None:
        r = rng or random
        indices = [r.randint(0, len(x) - 1) for _ in range(jit_ratio)]
    y = np.array(x, copy=True)
    y[np.asarray(indices, dtype=np.int64)] = 0
    return y


def jittering(x, jit_ratio=1000, indices=None, rng=None):
    """`uformerWM/audio_attack.py:156-173`: np.delete of `jit_ratio` random samples (inclusive upper bound
    `len(x)` as in the reference, so an out-of-bounds draw raises IndexError exactly like the reference)."""
    import random
    if indices is None:
        r = rng or random
        indices = [r.randint(0, len(x)) for _ in range(jit_ratio)]
    return np.delete(np.asarray(x), np.asarray(indices, dtype=np.int64))


def requantization(x):
    """`uformerWM/audio_attack.py:85-96`: `sf.write(..., subtype='PCM_U8')` + `sf.read`.
    PARITY UNPINNED (neither python-soundfile nor libsndfile is in the reference tree or installed here):
    restated from their sources - soundfile opens every file with SFC_SET_CLIPPING = TRUE, so libsndfile's
    pcm.c converts with `f2uc_clip_array` / `d2uc_clip_array`: u = (lrint(x * 2^31) >> 24) + 128, saturating to
    255 for x * 2^31 >= 2^31 - 1 and to 0 for x <= -1; the read (`uc2f_array`) is (u - 128) / 128.
    Net effect: floor(x * 128) / 128 clipped to [-1, 127/128]."""
    sv = np.asarray(x, np.float64) * 2147483648.0
    q = np.floor_divide(np.rint(np.clip(sv, -2147483648.0, 2147483647.0)).astype(np.int64), 1 << 24) + 128
    u = np.where(sv >= 2147483647.0, 255, np.where(sv <= -2147483648.0, 0, q))
    return (u.astype(np.float64) - 128.0) / 128.0


def resampling(x):
    """`uformerWM/audio_attack.py:71-83`: librosa 16k -> 8k -> 16k.
    PARITY UNPINNED (librosa not available; default res_type differs by version): restated
    as scipy polyphase resampling (Kaiser-windowed FIR, `scipy.signal.resample_poly`)."""
    d = sps.resample_poly(x, 1, 2)
    return sps.resample_poly(d, 2, 1)[:len(x)]


def apply_attack(x, attack, draws=None):
    """Attack dispatch grammar `uformerWM/audio_test.py:631-660` ('name-p1[-p2]'); '+' chains
    several attacks on the same waveform (BASELINE config 2: 'awgn-20+low_pass')."""
    if "+" in attack:
        for one in attack.split("+"):
            x = apply_attack(x, one, draws)
        return x
    p = attack.split("-")
    draws = draws or {}
    if p[0] == "echo_addition":
        return echo_addition(x)
    if p[0] == "amplitude_scaling":
        return amplitude_scaling(x, factor=float(p[1]))
    if p[0] == "low_pass":
        return low_pass_filter(x)
    if p[0] == "closed_loop":
        return closed_loop(x)
    if p[0] == "awgn":
        return awgn(x, snr=float(p[1]), noise_unit=draws.get("awgn"))
    if p[0] == "resampling":
        return resampling(x)
    if p[0] == "requantization":
        return requantization(x)
    if p[0] == "jittering_2":
        return jittering_2(x, int(p[1]), indices=draws.get("jitter"))
    if p[0] == "jittering":
        return jittering(x, indices=draws.get("jitter_delete"))
    raise ValueError("attack %r is outside the hot-path scope (needs third-party codecs)" % attack)


# --------------------------------------------------------------------------- metrics
def signaltonoise(a, axis=0, ddof=0):
    """`uformerWM/evaluate.py:133-137`, `uformerWM/audio_test.py:522-526`."""
    a = np.asanyarray(a)
    m = a.mean(axis)
    sd = a.std(axis=axis, ddof=ddof)
    return 20 * np.log10(abs(np.where(sd == 0, 0, m / sd)))


def cal_snr(audio_ori, audio_recon):
    """`uformerWM/evaluate.py:139-144`."""
    n = min(len(audio_ori), len(audio_recon))
    ps = np.sum(np.square(audio_ori[:n]))
    pn = np.sum(np.square(audio_ori[:n] - audio_recon[:n]))
    return 10 * np.log10(ps / pn)


def SNR_singlech(S, SN):
    """`uformerWM/evaluate.py:83-90`."""
    S = S - np.mean(S)
    S = S / np.max(np.abs(S))
    mean_S = np.sum(S) / len(S)
    PS = np.sum((S - mean_S) * (S - mean_S))
    PN = np.sum((S - SN) * (S - SN))
    return 10 * math.log(PS / PN, 10)


def bit_error_rate(decoded, message):
    """`hidden/test_model.py:60-64`: mean(abs(clip(round(decoded),0,1) - message)); numpy
    round is half-to-even."""
    d = np.clip(np.round(np.asarray(decoded, dtype=np.float64)), 0, 1)
    return float(np.mean(np.abs(d - np.asarray(message, dtype=np.float64))))


def mse(a, b):
    """`torch.nn.MSELoss()` as used at `uformerWM/audio_test.py:618,625,712`."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.mean((a - b) ** 2))

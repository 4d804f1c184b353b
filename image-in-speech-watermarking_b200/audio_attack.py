"""Waveform attacks of the reference (`uformerWM/audio_attack.py`) on the GPU.

Two call surfaces:
* the reference's own: ``<name>(np.ndarray 1-D[, param]) -> np.ndarray`` (float64 out, as numpy
  promotes in the reference) - host buffers, copies inside;
* the batched device API the pipeline uses: ``apply_attack(wave (B, L) CUDA, 'name-p1[-p2]')``
  with the reference's attack-id grammar (`uformerWM/audio_test.py:631-660`), extended with '+'
  to chain attacks on the same waveform (BASELINE config 2: 'awgn-20+low_pass').

Attacks that need third-party codecs / phase vocoders (aac, mp3compress, time_scaling,
pitch_scaling) are outside the hot-path scope and raise."""
import ctypes
import os

import numpy as np
import torch

from . import _lib

_BUTTER_CACHE = {}
_CALLS = [0]
_M64 = 0xFFFFFFFFFFFFFFFF


def fresh_seed():
    """A new 64-bit key for the device RNG (Philox AWGN, jitter index draws), one per attack call - the reference
    draws new `np.random` / `random` values on every call (`audio_attack.py:112-123,161-163,181-183`), so no two
    utterances, calls or ranks may share a noise realisation.  Taken from numpy's global stream (the reference's
    own source: `np.random.seed()` still makes a run reproducible) and mixed with the rank and a call counter."""
    _CALLS[0] += 1
    base = int(np.random.randint(0, 1 << 31)) | (int(np.random.randint(0, 1 << 31)) << 31)
    rank = int(os.environ.get("RANK", "0"))
    return (base ^ ((rank + 1) * 0x9E3779B97F4A7C15) ^ (_CALLS[0] * 0xD1B54A32D192ED03)) & _M64


def _butter(order=8, wn=0.5):
    """`signal.butter(8, wn, 'lowpass')` + `lfilter_zi` (`audio_attack.py:27-29`); coefficient
    design is host-side scalar work (17 numbers), done with scipy as the reference does."""
    key = (order, wn)
    if key not in _BUTTER_CACHE:
        from scipy import signal
        b, a = signal.butter(order, wn, 'lowpass')
        zi = signal.lfilter_zi(b, a)
        _BUTTER_CACHE[key] = (np.ascontiguousarray(b, np.float64), np.ascontiguousarray(a, np.float64),
                              np.ascontiguousarray(zi, np.float64))
    return _BUTTER_CACHE[key]


def _dptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _wave2d(w):
    if not w.is_cuda:
        raise _lib.WmkError("device attack API needs CUDA tensors")
    w = w.float().contiguous()
    return w[None] if w.dim() == 1 else w


# ------------------------------------------------------------------ batched device API
def awgn_(wave, snr=15.0, noise_unit=None, seed=None):
    """seed=None: a fresh key per call (`fresh_seed`); an explicit seed is for tests / reproducible benchmarks."""
    seed = fresh_seed() if seed is None else int(seed) & _M64
    w = _wave2d(wave)
    out = torch.empty_like(w)
    nu = None if noise_unit is None else noise_unit.to(w.device, torch.float32).contiguous().reshape(w.shape)
    _lib.check(_lib.load().wmk_attack_awgn_f32(_lib.ptr(w), _lib.ptr(out), w.shape[0], w.shape[1], float(snr),
                                               _lib.ptr(nu), int(seed), _lib.stream_ptr()))
    return out


def amplitude_scaling_(wave, factor=0.8):
    w = _wave2d(wave)
    out = torch.empty_like(w)
    _lib.check(_lib.load().wmk_attack_scale_f32(_lib.ptr(w), _lib.ptr(out), w.shape[0], w.shape[1], float(factor),
                                                _lib.stream_ptr()))
    return out


def echo_addition_(wave, Fs=16000, td=0.5, AA=0.2):
    w = _wave2d(wave)
    out = torch.empty_like(w)
    _lib.check(_lib.load().wmk_attack_echo_f32(_lib.ptr(w), _lib.ptr(out), w.shape[0], w.shape[1], int(td * Fs),
                                               float(AA), _lib.stream_ptr()))
    return out


def low_pass_filter_(wave, Fs=16000, low_pass_parameter=8000):
    w = _wave2d(wave)
    out = torch.empty_like(w)
    wn = 2 * low_pass_parameter / (Fs * 2)
    b, a, zi = _butter(8, wn)
    _lib.check(_lib.load().wmk_attack_lowpass_f32(_lib.ptr(w), _lib.ptr(out), w.shape[0], w.shape[1], 8, _dptr(b),
                                                  _dptr(a), _dptr(zi), _lib.stream_ptr()))
    return out


def jittering_2_(wave, jit_ratio=1000, indices=None, seed=None):
    w = _wave2d(wave).clone()
    B, L = w.shape
    if indices is None:
        g = torch.Generator(device="cpu").manual_seed(fresh_seed() if seed is None else int(seed) & _M64)
        indices = torch.randint(0, L, (B, jit_ratio), generator=g, dtype=torch.int32)
    idx = torch.as_tensor(indices, dtype=torch.int32).reshape(B, -1).to(w.device).contiguous()
    _lib.check(_lib.load().wmk_attack_jitter_zero_f32(_lib.ptr(w), B, L, _lib.ptr(idx), idx.shape[1], _lib.stream_ptr()))
    return w


def jittering_(wave, jit_ratio=1000, indices=None, seed=None):
    """`jittering` (`audio_attack.py:156-173`): np.delete of `jit_ratio` random (not necessarily distinct) samples.
    Returns (out (B, L) zero-padded, lengths: list of B ints).  Like numpy, an index >= L raises IndexError (the
    reference draws `random.randint(0, len)` inclusive, so it fails itself about 2 % of the time at 3 s)."""
    w = _wave2d(wave)
    B, L = w.shape
    if indices is None:
        g = torch.Generator(device="cpu").manual_seed(fresh_seed() if seed is None else int(seed) & _M64)
        indices = torch.randint(0, L, (B, jit_ratio), generator=g, dtype=torch.int32)
    idx = torch.as_tensor(indices, dtype=torch.int32).reshape(B, -1)
    if int(idx.max()) >= L or int(idx.min()) < -L:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (int(idx.max()), L))
    idx = torch.where(idx < 0, idx + L, idx).to(w.device).contiguous()
    out = torch.empty_like(w)
    lens = torch.empty(B, dtype=torch.int32, device=w.device)
    _lib.check(_lib.load().wmk_attack_jitter_delete_f32(_lib.ptr(w), _lib.ptr(out), B, L, _lib.ptr(idx), idx.shape[1],
                                                        _lib.ptr(lens), _lib.stream_ptr()))
    return out, lens.tolist()


def requantization_(wave):
    w = _wave2d(wave)
    out = torch.empty_like(w)
    _lib.check(_lib.load().wmk_attack_requant8_f32(_lib.ptr(w), _lib.ptr(out), w.shape[0], w.shape[1], _lib.stream_ptr()))
    return out


def resampling_(wave):
    from scipy import signal
    w = _wave2d(wave)
    out = torch.empty_like(w)
    h = np.ascontiguousarray(signal.firwin(41, 0.5, window=('kaiser', 5.0)), np.float64)
    _lib.check(_lib.load().wmk_attack_resample2_f32(_lib.ptr(w), _lib.ptr(out), w.shape[0], w.shape[1], _dptr(h), 41,
                                                    _lib.stream_ptr()))
    return out


def apply_attack(wave, attack, draws=None, seed=None):
    """Device-resident attack dispatch; `attack` follows `uformerWM/audio_test.py:631-660`,
    '+' chains several attacks.  Random attacks draw a fresh key per call (`fresh_seed`) unless `seed` is given
    (tests, reproducible benchmark steps); stage i of a chain uses seed + i."""
    draws = draws or {}
    w = _wave2d(wave)
    base_seed = seed
    for stage, one in enumerate(attack.split("+")):
        seed = None if base_seed is None else (int(base_seed) + stage) & _M64
        p = one.split("-")
        if p[0] == "echo_addition":
            w = echo_addition_(w)
        elif p[0] == "amplitude_scaling":
            w = amplitude_scaling_(w, float(p[1]))
        elif p[0] == "low_pass":
            w = low_pass_filter_(w)
        elif p[0] == "closed_loop":
            pass
        elif p[0] == "awgn":
            w = awgn_(w, float(p[1]), draws.get("awgn"), seed)
        elif p[0] == "resampling":
            w = resampling_(w)
        elif p[0] == "requantization":
            w = requantization_(w)
        elif p[0] == "jittering_2":
            w = jittering_2_(w, int(p[1]), draws.get("jitter"), seed)
        elif p[0] == "jittering":
            # sample deletion shortens every utterance by its own number of distinct indices: one utterance at a
            # time (the reference driver's batch size), so that the result stays a dense (1, L') waveform
            if w.shape[0] != 1:
                raise ValueError("attack 'jittering' (sample deletion) yields ragged lengths: run it one utterance at a time")
            w, lens = jittering_(w, 1000, draws.get("jitter_delete"), seed)
            w = w[:, :lens[0]].contiguous()
        else:
            raise ValueError("attack %r is outside the hot-path scope (needs third-party codecs)" % one)
    return w


# ------------------------------------------------------------------ reference call surface
def _np_call(fn, x, *a, **k):
    t = torch.as_tensor(np.ascontiguousarray(x), dtype=torch.float32).cuda()
    return fn(t, *a, **k)[0].double().cpu().numpy()


def low_pass_filter(S_watermarked, Fs=16000, low_pass_parameter=8000):
    return _np_call(low_pass_filter_, S_watermarked, Fs, low_pass_parameter)


def echo_addition(S_watermarked, Fs=16000, td=0.5, AA=0.2):
    return _np_call(echo_addition_, S_watermarked, Fs, td, AA)


def amplitude_scaling(S_watermarked, factor=0.8):
    return _np_call(amplitude_scaling_, S_watermarked, factor)


def closed_loop(S_watermarked):
    return S_watermarked


def resampling(S_watermarked, fs=16000):
    return _np_call(resampling_, S_watermarked)


def requantization(S_watermarked, quantization_bits=8):
    return _np_call(requantization_, S_watermarked)


def awgn(signal, snr=15):
    unit = torch.from_numpy(np.random.normal(0, 1.0, np.shape(signal)))     # reference RNG stream
    return _np_call(awgn_, signal, snr, unit)


def jittering(S_watermarked, jit_ratio=1000):
    import random
    idx = [random.randint(0, len(S_watermarked)) for _ in range(jit_ratio)]      # the reference's inclusive bound
    t = torch.as_tensor(np.ascontiguousarray(S_watermarked), dtype=torch.float32).cuda()
    out, lens = jittering_(t, jit_ratio, np.asarray(idx)[None])
    return out[0, :lens[0]].double().cpu().numpy()


def jittering_2(S_watermarked, jit_ratio=1000):
    import random
    idx = [random.randint(0, len(S_watermarked) - 1) for _ in range(jit_ratio)]
    return _np_call(jittering_2_, S_watermarked, jit_ratio, np.asarray(idx)[None])

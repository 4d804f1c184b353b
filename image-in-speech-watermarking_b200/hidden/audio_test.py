"""HiDDeN flavour of the hot path on STFT magnitudes (reference `hidden/audio_test.py:397-630`;
BASELINE config 3).  The reference's `Hidden` / `EncoderDecoder` / `Encoder` classes are absent
from its tree, so the part that exists - magnitude view of the clips, one
randomly chosen noise layer per batch (`noise_layers/noiser.py:29-31`), the modified `Decoder`
(`hidden/model/decoder.py:12-40`) and the bit-error rate of `hidden/test_model.py:57-64` - is what
runs here, all on libwmk kernels."""
import torch

from .. import _lib
from .. import audio_uformer_stft as FE
from .. import evaluate as EV

AUDIO_SCALE = 0.025          # the constant x0.025 / x40 scaling of `hidden/audio_test.py:45,218,430,548`


def magphase_split(clips, want_phase=True):
    """clips (n,2,F,T) re/im CUDA -> mag (n,1,F,T), phase (n,1,F,T) (or None)."""
    lib = _lib.load()
    c = clips.contiguous().float()
    if not c.is_cuda:
        raise _lib.WmkError("magphase_split has no CPU implementation")
    n, _, F_, T = c.shape
    mag = torch.empty((n, 1, F_, T), device=c.device, dtype=torch.float32)
    ph = torch.empty_like(mag) if want_phase else None
    _lib.check(lib.wmk_magphase_split_f32(_lib.ptr(c), _lib.ptr(mag), _lib.ptr(ph), n, F_ * T, _lib.stream_ptr()))
    return mag, ph


def magphase_merge(mag, phase):
    """mag, phase (n,1,F,T) -> clips (n,2,F,T)."""
    lib = _lib.load()
    m, p = mag.contiguous().float(), phase.contiguous().float()
    n, _, F_, T = m.shape
    out = torch.empty((n, 2, F_, T), device=m.device, dtype=torch.float32)
    _lib.check(lib.wmk_magphase_merge_f32(_lib.ptr(m), _lib.ptr(p), _lib.ptr(out), n, F_ * T, _lib.stream_ptr()))
    return out


def attack_and_decode(waves, messages, decoder, noiser=None, audio_scale=AUDIO_SCALE):
    """waves (B,L) CUDA -> STFT clips -> magnitudes x audio_scale -> noiser([noised, cover]) ->
    decoder -> (B*nc, 1, H/4, W/4) images and the per-clip statistics {bit errors, sum sq err}
    against messages (B or 1, 1, 32, 32) (only when the noise layer kept the 128x128 geometry)."""
    waves = waves.float().contiguous()
    B, L = waves.shape
    T = FE.num_frames(L)
    nc = (T + 127) // 128
    clips = FE.stft_clips(waves, nc).reshape(B * nc, 2, 128, 128)
    mag, _ = magphase_split(clips, want_phase=False)
    mag = mag * audio_scale
    noised = noiser([mag.clone(), mag])[0] if noiser is not None else mag
    decoded = decoder(noised)
    stats = None
    if decoded.shape[-2:] == (32, 32):
        msg = messages.float().reshape(-1, 1, 32, 32)
        if msg.shape[0] == B:
            msg = msg[:, None].expand(B, nc, 1, 32, 32).reshape(B * nc, 1, 32, 32)
        stats = EV.wm_stats(decoded, msg)
    return decoded, stats

"""Minimal RIFF/WAVE reader / writer for the evaluator's audio dumps.

The reference writes the original, watermarked and attacked waveforms with
`torchaudio.save(path, tensor, sr)` (`uformerWM/evaluate.py:240-247`); for float32 tensors that is a 32-bit
IEEE-float WAV (format tag 3), which is what `write_wav` produces.  `read_wav` also understands 16-bit and
8-bit PCM (format tag 1) with libsndfile's scaling (x / 32768, (x - 128) / 128).  Host-side format code only:
no third-party package (torchaudio's backends, soundfile) is needed."""
import struct

import numpy as np


def write_wav(path, wave, sample_rate=16000):
    """wave: (L,) or (channels, L) array-like / tensor of floats in [-1, 1] -> 32-bit float WAV."""
    a = wave.detach().cpu().numpy() if hasattr(wave, "detach") else np.asarray(wave)
    a = np.atleast_2d(a).astype("<f4")                                   # (channels, L)
    ch, n = a.shape
    data = np.ascontiguousarray(a.T).tobytes()                           # interleaved frames
    fmt = struct.pack("<HHIIHH", 3, ch, int(sample_rate), int(sample_rate) * ch * 4, ch * 4, 32)
    fact = struct.pack("<I", n)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"fact" + struct.pack("<I", 4) + fact + \
        b"data" + struct.pack("<I", len(data)) + data
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)


def _parse(path):
    """-> (format tag, channels, sample rate, bits, payload bytes) of a RIFF/WAVE file (host-side header walk)."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:4] != b"RIFF" or buf[8:12] != b"WAVE":
        raise ValueError("%s is not a RIFF/WAVE file" % path)
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(buf):
        tag, size = buf[pos:pos + 4], struct.unpack("<I", buf[pos + 4:pos + 8])[0]
        chunk = buf[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            fmt = struct.unpack("<HHIIHH", chunk[:16])
        elif tag == b"data":
            data = chunk
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError("%s: missing fmt / data chunk" % path)
    tag, ch, sr, _, _, bits = fmt
    if not ((tag == 3 and bits == 32) or (tag == 1 and bits in (8, 16))):
        raise ValueError("%s: unsupported WAV encoding (format %d, %d bits)" % (path, tag, bits))
    return tag, ch, sr, bits, data


def read_wav_cuda(path, device="cuda"):
    """-> (float32 CUDA tensor (channels, L), sample_rate): only the raw `data` payload crosses PCIe, the sample
    decode (int16 / 32768, (u8 - 128) / 128, float32) and the de-interleave run on the GPU (`wmk_pcm_decode_f32`)."""
    import torch
    from . import _lib
    tag, ch, sr, bits, data = _parse(path)
    raw = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(device)
    n_frames = len(data) // (ch * bits // 8)
    out = torch.empty((ch, n_frames), device=raw.device, dtype=torch.float32)
    with torch.cuda.device(raw.device):
        _lib.check(_lib.load().wmk_pcm_decode_f32(_lib.ptr(raw), bits, n_frames, ch, _lib.ptr(out), _lib.stream_ptr()))
    return out, sr


def read_wav(path):
    """-> (float32 array (channels, L), sample_rate)."""
    tag, ch, sr, bits, data = _parse(path)
    if tag == 3 and bits == 32:
        a = np.frombuffer(data, "<f4").astype(np.float32)
    elif tag == 1 and bits == 16:
        a = np.frombuffer(data, "<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 8:
        a = (np.frombuffer(data, "u1").astype(np.float32) - 128.0) / 128.0
    else:
        raise ValueError("%s: unsupported WAV encoding (format %d, %d bits)" % (path, tag, bits))
    return a.reshape(-1, ch).T.copy(), sr

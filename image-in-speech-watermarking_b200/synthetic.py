"""Synthetic workload generators: speech-like waveforms, watermark images
and deterministic random weights with the reference's parameter names.  Host-side only."""
import math
import zlib

import numpy as np
import torch

SR = 16000


def synth_speech(index, seconds, sr=SR):
    """fp32 mono 16 kHz 'voiced speech': 20 harmonics of f0 in U[90,250] Hz with 1/k amplitudes,
    a 3-6 Hz syllabic envelope, a N(0,0.005^2) floor, RMS normalised to 0.05; seed 42 + index."""
    g = torch.Generator().manual_seed(42 + int(index))
    n = int(round(seconds * sr))
    t = torch.arange(n, dtype=torch.float64) / sr
    f0 = 90.0 + 160.0 * torch.rand(1, generator=g, dtype=torch.float64).item()
    phases = 2 * math.pi * torch.rand(20, generator=g, dtype=torch.float64)
    env_f = 3.0 + 3.0 * torch.rand(1, generator=g, dtype=torch.float64).item()
    env_p = 2 * math.pi * torch.rand(1, generator=g, dtype=torch.float64).item()
    x = torch.zeros(n, dtype=torch.float64)
    for k in range(1, 21):
        x += torch.sin(2 * math.pi * k * f0 * t + phases[k - 1]) / k
    x *= 0.55 + 0.45 * torch.sin(2 * math.pi * env_f * t + env_p)
    x += 0.005 * torch.randn(n, generator=g, dtype=torch.float64)
    x *= 0.05 / x.pow(2).mean().sqrt()
    return x.to(torch.float32)


def synth_speech_batch(first_index, count, seconds):
    return torch.stack([synth_speech(first_index + i, seconds) for i in range(count)])


def synth_image_binary(index, size=32):
    """Bernoulli(0.5) {0,1} image as `BinaryWM` (`uformerWM/audio_uformer_stft.py:228-233`)."""
    g = torch.Generator().manual_seed(1234 + int(index))
    return (torch.rand(1, size, size, generator=g) > 0.5).to(torch.float32)


def synth_image_grey(index, size=64):
    """U[0,1] greyscale image min-max normalised to [0,1] (as `NormalizeBatch`,
    `uformerWM/evaluate.py:63-81`)."""
    g = torch.Generator().manual_seed(1234 + int(index))
    im = torch.rand(1, size, size, generator=g)
    return (im - im.min()) / (im.max() - im.min())


def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * int(seed)) & 0x7FFFFFFF)


def init_state_dict(schema, kind="reference", seed=0):
    """Deterministic random weights for a ``{name: (shape, kind)}`` schema.

    kind='reference': the reference's own initialisers (`uformerWM/model.py:2325-2332`, `:507`,
    PyTorch defaults for convs / Embedding).  kind='stress': every bias / LayerNorm affine /
    table non-trivial and Linear weights ~ N(0, 0.7/sqrt(fan_in)) so that every code path
    contributes visibly to the output (used by the parity tests)."""
    sd = {}
    for name, (shape, k) in schema.items():
        g = _gen(name, seed)
        if k == "index":
            c = torch.arange(8)
            co = torch.stack(torch.meshgrid([c, c], indexing="ij")).flatten(1)
            rel = (co[:, :, None] - co[:, None, :]).permute(1, 2, 0).contiguous()
            rel[:, :, 0] += 7
            rel[:, :, 1] += 7
            rel[:, :, 0] *= 15
            sd[name] = rel.sum(-1)
            continue
        t = torch.empty(shape, dtype=torch.float32)
        stress = kind == "stress"
        if k == "lin_w":
            if stress:
                t.normal_(0, 0.7 / math.sqrt(shape[1]), generator=g)
            else:
                torch.nn.init.trunc_normal_(t, std=0.02, generator=g)
        elif k == "conv_w":
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            t.uniform_(-b, b, generator=g)
        elif k == "bias":
            if stress:
                t.normal_(0, 0.1, generator=g)
            elif name.endswith("to_q.bias") or name.endswith("to_kv.bias") or name.endswith("proj.bias") \
                    or "linear1" in name or "linear2" in name:
                t.zero_()
            else:
                t.uniform_(-0.1, 0.1, generator=g)
        elif k == "ln_w":
            t.fill_(1.0)
            if stress:
                t.add_(torch.empty(shape).normal_(0, 0.1, generator=g))
        elif k == "ln_b":
            t.zero_()
            if stress:
                t.normal_(0, 0.1, generator=g)
        elif k == "table":
            if stress:
                t.normal_(0, 0.5, generator=g)
            else:
                torch.nn.init.trunc_normal_(t, std=0.02, generator=g)
        elif k == "embed":
            t.normal_(0, 0.5 if stress else 1.0, generator=g)
        else:
            raise ValueError(k)
        sd[name] = t
    return sd

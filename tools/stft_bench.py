"""Stand-alone STFT / ISTFT throughput (BASELINE metric 'STFT GB/s vs HBM peak').

    python tools/stft_bench.py [--utterances 2048] [--seconds 3] [--iters 20]

Algorithmic traffic = 1276 B per frame (63 new samples read + 128 complex bins written).
Working set (>= 2.5 GB) is far larger than the 126 MB L2, so no flush is needed between iterations."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def measure(utterances=2048, seconds=3.0, iters=20):
    import torch
    from image_in_speech_watermarking_b200 import audio_uformer_stft as FE
    L = int(16000 * seconds)
    T = FE.num_frames(L)
    g = torch.Generator(device="cuda").manual_seed(0)
    wave = torch.randn(utterances, L, device="cuda", generator=g) * 0.05
    out = {}
    clips = FE.stft_clips(wave)
    for name, fn in (("stft", lambda: FE.stft_clips(wave)), ("istft", lambda: FE.istft_clips(clips, T, L))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name] = {"ms": ms, "gbs": 1276.0 * T * utterances / (ms * 1e-3) / 1e9, "frames": T * utterances}
    # training-time analysis (n_fft 256 / hop 128): 1536 B per frame (128 samples read + 128 complex bins written)
    T2 = FE.num_frames_train(L)
    fn = lambda: FE.stft256_clips(wave)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    out["stft256"] = {"ms": ms, "gbs": 1536.0 * T2 * utterances / (ms * 1e-3) / 1e9, "frames": T2 * utterances}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--utterances", type=int, default=2048)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(measure(a.utterances, a.seconds, a.iters)))

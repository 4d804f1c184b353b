"""One-shot GPU diagnostic: runs every kernel family against the oracle and prints the errors
(never asserts) so a single gpurun call localises all problems.  Not a test, not a bench."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import warnings
warnings.filterwarnings("ignore")

from oracle import uformer as O, signal as S, pipeline as P                     # noqa: E402
from image_in_speech_watermarking_b200 import _lib, synthetic as SY            # noqa: E402
from image_in_speech_watermarking_b200 import audio_uformer_stft as FE         # noqa: E402
from image_in_speech_watermarking_b200 import audio_attack as AT, evaluate as EV, audio_test as PT   # noqa: E402
from image_in_speech_watermarking_b200.model import UformerAudio               # noqa: E402


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30)), float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def section(name, fn):
    print("=== %s" % name, flush=True)
    t = time.time()
    try:
        fn()
    except Exception:
        traceback.print_exc()
    torch.cuda.synchronize()
    print("    (%.1fs)" % (time.time() - t), flush=True)


WHICH_LINEAR = []


def linear():
    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for prec, name in ((0, "fp32"), (1, "bf16")):
        if WHICH_LINEAR and name not in WHICH_LINEAR:
            continue
        for (M, N, K) in [(128, 32, 32), (256, 96, 32), (200, 64, 64), (1024, 128, 128), (64, 512, 2048),
                          (4096, 384, 128), (3000, 256, 512), (128, 1024, 64)]:
            A = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
            b = torch.randn(N, generator=g).cuda()
            C = torch.full((M, N), float("nan"), device="cuda")
            st = lib.wmk_linear_f32(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(C), M, N, K, prec, 0, _lib.stream_ptr())
            torch.cuda.synchronize()
            if st != 0:
                print("   %s %s FAILED: %s" % (name, (M, N, K), lib.wmk_last_error().decode())); continue
            if prec == 1:
                ref = A.bfloat16().float().double() @ W.bfloat16().float().double().T + b.double()
            else:
                ref = A.double() @ W.double().T + b.double()
            print("   %s %-18s max-rel %.3e  l2-rel %.3e  nan=%d" % ((name, str((M, N, K))) + rel(C.cpu(), ref.cpu()) + (int(torch.isnan(C).sum()),)))


def frontend():
    for L in (16000, 8002, 48000, 63 * 127 + 1, 5000):
        w = torch.stack([SY.synth_speech(i, L / 16000.0)[:L] for i in range(2)])
        ref = S.stft(w.numpy())                                # (B,128,T,2)
        got = FE.stft(w.cuda()).cpu().numpy()
        print("   stft L=%d T=%d  %s" % (L, ref.shape[2], rel(got, ref)))
        back = FE.istft(torch.from_numpy(ref).float().cuda(), length=L).cpu().numpy()
        print("   istft(length=L)     %s" % (rel(back, S.istft(ref, length=L)),))
        back2 = FE.istft(torch.from_numpy(ref).float().cuda()).cpu().numpy()
        print("   istft(default)      %s" % (rel(back2, S.istft(ref)),))
    c = FE.stft_clips(w.cuda())
    print("   clips shape", tuple(c.shape), "pad frames zero:", float(c[:, -1, :, :, (ref.shape[2] % 128):].abs().max()))


def attacks():
    g = np.load(os.path.join(ROOT, "tests", "golden", "signal.npz"))
    x = torch.from_numpy(g["x"]).cuda()[None]
    print("   awgn     %s" % (rel(AT.awgn_(x, 20, torch.from_numpy(g["awgn_unit"]).float()).cpu()[0], g["awgn20"]),))
    print("   lowpass  %s" % (rel(AT.low_pass_filter_(x).cpu()[0], g["low_pass"]),))
    print("   echo     %s" % (rel(AT.echo_addition_(x).cpu()[0], g["echo"]),))
    print("   scale    %s" % (rel(AT.amplitude_scaling_(x, 0.7).cpu()[0], g["scale07"]),))
    print("   jitter   %s" % (rel(AT.jittering_2_(x, 200, g["jitter_idx"][None]).cpu()[0], g["jitter"]),))
    print("   requant  %s" % (rel(AT.requantization_(x).cpu()[0], S.requantization(g["x"].astype(np.float64))),))
    print("   resample %s" % (rel(AT.resampling_(x).cpu()[0], S.resampling(g["x"].astype(np.float64))),))
    n = AT.awgn_(x.repeat(4, 1), 20, None, seed=3)
    d = (n - x).double()
    print("   philox noise: snr %.3f dB (want 20), mean %.2e, kurt %.3f" % (
        float(10 * torch.log10(x.double().pow(2).sum() * 4 / d.pow(2).sum())), float(d.mean()),
        float((d ** 4).mean() / (d ** 2).mean() ** 2)))
    lp = torch.from_numpy(g["low_pass"]).float().cuda()[None]
    st = EV.wave_stats(x, lp)
    print("   cal_snr  got %.9f want %.9f" % (float(EV.snr_from_stats(st)[0]), float(g["cal_snr"])))
    print("   s2n      got %.9f want %.9f" % (float(EV.signaltonoise_from_stats(EV.wave_stats(x, x))[0]), float(g["signaltonoise"])))
    wm = torch.rand(3, 1, 32, 32); msg = (torch.rand(3, 1, 32, 32) > 0.5).float()
    st = EV.wm_stats(wm.cuda(), msg.cuda()).cpu().numpy()
    print("   ber      got %s want %s" % (st[:, 0] / 1024, [S.bit_error_rate(wm[i].numpy(), msg[i].numpy()) for i in range(3)]))
    print("   wm mse   got %s want %s" % (st[:, 1] / 1024, [S.mse(wm[i].numpy(), msg[i].numpy()) for i in range(3)]))


def model(prec, kind):
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_%s.npz" % kind))
    sd = SY.init_state_dict(O.state_dict_schema(), kind, int(g["seed"]))
    m = UformerAudio(precision=prec)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = torch.from_numpy(g["x"]).cuda(); msg = torch.from_numpy(g["msg"]).cuda()
    m.enable_taps(True)
    o = m.run(x, msg, want=("stft_new", "noise", "wm_pred", "wm", "wm_logits", "y"))
    torch.cuda.synchronize()
    taps = {}
    with torch.no_grad():
        ref = O.forward(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["msg"]), taps, return_logits=True)
    for name in taps:
        try:
            got = m.get_tap(name).cpu().numpy().reshape(taps[name].shape)
            print("   tap %-14s max-rel %.3e l2-rel %.3e" % ((name,) + rel(got, taps[name].numpy())))
        except Exception as e:
            print("   tap %-14s unavailable (%s)" % (name, str(e)[:60]))
    for k in ("stft_new", "noise", "wm_pred", "wm"):
        print("   out %-10s vs golden  max-rel %.3e l2-rel %.3e" % ((k,) + rel(o[k].cpu().numpy(), g[k])))
    lg = o["wm_logits"].cpu().numpy(); rl = ref[4].numpy()
    flips = (lg > 0) != (rl > 0)
    print("   logits max-abs-err %.3e; bit flips %d (of %d), max |ref logit| among flips %.3e" % (
        np.abs(lg - rl).max(), flips.sum(), flips.size, np.abs(rl[flips]).max() if flips.any() else 0.0))
    wa = m.wm_decode(torch.from_numpy(g["x_att"]).cuda()).cpu().numpy()
    print("   wm_decode(x_att) vs golden %s" % (rel(wa, g["wm_att"]),))
    m.enable_taps(False)
    return m, sd


def pipeline(m, sd):
    for name in ("pipeline_cfg1_awgn_20.npz", "pipeline_cfg1_low_pass.npz"):
        g = np.load(os.path.join(ROOT, "tests", "golden", name))
        wave = SY.synth_speech(0, 1.0)[None]
        msg = SY.synth_image_binary(0)[None]
        draws = {"awgn": torch.from_numpy(g["awgn_unit"]).float()} if g["awgn_unit"].size else None
        out = PT.reconstruct_audio(PT.prepare_data(wave), msg, m, attack=str(g["attack"]), draws=draws)
        print("   %s recon %s att %s" % (name, rel(out[1].numpy(), g["recon"]), rel(out[0], g["audio_att"])))
        print("      wms %s wms_att %s" % (rel(np.stack(out[3]), g["wms"]), rel(np.stack(out[4]), g["wms_att"])))
        print("      mse %.6e/%.6e wm_loss %.6f/%.6f wm_loss_att %.6f/%.6f snr_ori %.4f/%.4f snr_recon %.4f/%.4f" % (
            out[5], g["mse"], out[6], g["wm_loss"], out[7], g["wm_loss_att"], out[8], g["snr_ori"], out[9], g["snr_recon"]))


def batch_consistency(prec):
    sd = SY.init_state_dict(O.state_dict_schema(), "stress", 0)
    m = UformerAudio(precision=prec, clips_per_pass=3)
    m.load_state_dict(sd); m = m.cuda().eval()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(7, 2, 128, 128, generator=g).cuda(); msg = (torch.rand(7, 1, 32, 32, generator=g) > 0.5).float().cuda()
    a = m.run(x, msg, want=("stft_new", "wm_logits"))
    b = [m.run(x[i:i + 1], msg[i:i + 1], want=("stft_new", "wm_logits")) for i in range(7)]
    print("   %s chunked(3) vs single: stft_new %.3e logits %.3e" % (
        prec, float((a["stft_new"] - torch.cat([q["stft_new"] for q in b])).abs().max()),
        float((a["wm_logits"] - torch.cat([q["wm_logits"] for q in b])).abs().max())))


def speed():
    lib = _lib.load()
    sd = SY.init_state_dict(O.state_dict_schema(), "reference", 0)
    for prec, B in (("bf16", 32), ("fp32", 8)):
        m = UformerAudio(precision=prec, clips_per_pass=B)
        m.load_state_dict(sd); m = m.cuda().eval()
        x = torch.randn(B, 2, 128, 128).cuda(); msg = torch.rand(B, 1, 32, 32).cuda()
        m.run(x, msg); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); m.run(x, msg); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print("   %s forward B=%d: %.2f ms  (%.1f clips/s, %.1f TFLOP/s of 53.76 GFLOP/clip)" % (
            prec, B, ms, B / ms * 1e3, 53.76e9 * B / ms / 1e9))
        del m


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), "lib", _lib.load().wmk_version(), flush=True)
    which = sys.argv[1:] or ["linear", "frontend", "attacks", "fp32", "bf16", "pipeline", "batch", "speed"]
    if "linear_fp32" in which: WHICH_LINEAR.append("fp32")
    if "linear_bf16" in which: WHICH_LINEAR.append("bf16")
    if "linear" in which or WHICH_LINEAR: section("linear", linear)
    if "frontend" in which: section("frontend", frontend)
    if "attacks" in which: section("attacks", attacks)
    keep = {}
    if "fp32" in which:
        section("model fp32 stress", lambda: keep.update(fp32=model("fp32", "stress")))
        section("model fp32 reference-init", lambda: model("fp32", "reference"))
    if "pipeline" in which and "fp32" in keep: section("pipeline fp32", lambda: pipeline(*keep["fp32"]))
    if "bf16" in which:
        section("model bf16 stress", lambda: keep.update(bf16=model("bf16", "stress")))
        section("model bf16 reference-init", lambda: model("bf16", "reference"))
    if "pipeline" in which and "bf16" in keep: section("pipeline bf16", lambda: pipeline(*keep["bf16"]))
    if "batch" in which or "batch_fp32" in which: section("batch fp32", lambda: batch_consistency("fp32"))
    if "batch" in which or "batch_bf16" in which: section("batch bf16", lambda: batch_consistency("bf16"))
    if "speed" in which: section("speed", speed)

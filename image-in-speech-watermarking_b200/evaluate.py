"""Metrics and the report line of the reference's evaluator (`uformerWM/evaluate.py`), with the
reductions running on the GPU (`wmk_wave_stats_f64`, `wmk_wm_stats_f64`)."""
import math

import numpy as np
import torch

from . import _lib


def wave_stats(orig, test):
    """(B, L) CUDA waveforms -> (B, 6) float64 CUDA tensor
    {sum o^2, sum (o-t)^2, sum t, sum t^2, sum o, n}."""
    o = orig.float().contiguous()
    t = test.float().contiguous()
    if o.dim() == 1:
        o, t = o[None], t[None]
    L = min(o.shape[1], t.shape[1])
    if o.shape[1] != L:
        o = o[:, :L].contiguous()
    if t.shape[1] != L:
        t = t[:, :L].contiguous()
    st = torch.empty((o.shape[0], 6), device=o.device, dtype=torch.float64)
    _lib.check(_lib.load().wmk_wave_stats_f64(_lib.ptr(o), _lib.ptr(t), o.shape[0], L, _lib.ptr(st), _lib.stream_ptr()))
    return st


def wm_stats(wm, msg):
    """wm (n,1,32,32) sigmoid outputs, msg (n or 1,1,32,32) -> (n, 2) float64 {bit errors, sum sq err}."""
    w = wm.float().contiguous().reshape(-1, 1024)
    m = msg.float().contiguous().reshape(-1, 1024)
    stride = 0 if m.shape[0] == 1 and w.shape[0] > 1 else 1024
    st = torch.empty((w.shape[0], 2), device=w.device, dtype=torch.float64)
    _lib.check(_lib.load().wmk_wm_stats_f64(_lib.ptr(w), _lib.ptr(m), w.shape[0], stride, _lib.ptr(st), _lib.stream_ptr()))
    return st


def wm_stats_mapped(wm, first, step, msg, clips_per_utt, msgs_per_utt, n):
    """wm (N,1,32,32) sigmoid outputs of a batch of clips, msg (U * msgs_per_utt,1,32,32) the utterances' images:
    row i = clip c = first + i * step against image (c // clips_per_utt) * msgs_per_utt + (c % clips_per_utt) %
    msgs_per_utt -> (n, 2) float64 {bit errors, sum sq err}; no expanded message tensor."""
    w = wm.float().contiguous().reshape(-1, 1024)
    m = msg.float().contiguous().reshape(-1, 1024)
    if first + (n - 1) * step >= w.shape[0]:
        raise ValueError("clip selection runs past the %d clips" % w.shape[0])
    st = torch.empty((n, 2), device=w.device, dtype=torch.float64)
    _lib.check(_lib.load().wmk_wm_stats_mapped_f64(_lib.ptr(w), int(first), int(step), _lib.ptr(m), int(clips_per_utt),
                                                   int(msgs_per_utt), int(n), _lib.ptr(st), _lib.stream_ptr()))
    return st


def stats_finalize(st_att, st_rec, ws_clean, ws_att, n_clips_att):
    """One launch: per-utterance columns (B, 7) {snr_db, audio_mse, wm_mse_clean, wm_mse_att, bit_err_clean, bit_err_att,
    bits_att} and the additive vector (8,) of `sharding.STAT_KEYS` (what the ranks all-reduce)."""
    B = st_att.shape[0]
    stats = torch.empty((B, 7), device=st_att.device, dtype=torch.float64)
    vec = torch.empty(8, device=st_att.device, dtype=torch.float64)
    _lib.check(_lib.load().wmk_stats_finalize_f64(_lib.ptr(st_att), _lib.ptr(st_rec), _lib.ptr(ws_clean), _lib.ptr(ws_att), B,
                                                  int(n_clips_att), _lib.ptr(stats), _lib.ptr(vec), _lib.stream_ptr()))
    return stats, vec


def snr_from_stats(st):
    """`cal_snr` (`uformerWM/evaluate.py:139-144`) per utterance from wave_stats."""
    return 10.0 * torch.log10(st[:, 0] / st[:, 1])


def signaltonoise_from_stats(st):
    """`signaltonoise` (`evaluate.py:133-137`) of the *test* waveform: 20 log10 |mean/std|."""
    n = st[:, 5]
    mean = st[:, 2] / n
    var = st[:, 3] / n - mean * mean
    return 20.0 * torch.log10(torch.abs(mean / torch.sqrt(var)))


def _cuda1d(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32).reshape(1, -1).cuda()


def cal_snr(audio_ori, audio_recon):
    """`uformerWM/evaluate.py:139-144` (numpy in, python float out)."""
    n = min(len(audio_ori), len(audio_recon))
    st = wave_stats(_cuda1d(audio_ori[:n]), _cuda1d(audio_recon[:n]))
    return float(snr_from_stats(st)[0])


def SNR_singlech(S, SN):
    """`uformerWM/evaluate.py:83-90` (numpy in, python float out): S is centred and peak-normalised, then
    10 log10( sum (S - mean S)^2 / sum (S - SN)^2 ).  Mean, peak and both sums are GPU reductions
    (`wmk_wave_stats_f64`, `wmk_minmax_f32`), the normalisation is `wmk_affine_f32`."""
    from . import audio_uformer_stft as FE
    s, sn = _cuda1d(S), _cuda1d(SN)
    st = wave_stats(s, s)
    n = float(st[0, 5])
    mean = float(st[0, 4]) / n
    mn, mx = FE.minmax(s).tolist()
    peak = max(mx - mean, mean - mn)
    s2 = torch.empty_like(s)
    _lib.check(_lib.load().wmk_affine_f32(_lib.ptr(s), _lib.ptr(s2), s.numel(), 1.0 / peak, -mean / peak, _lib.stream_ptr()))
    st2 = wave_stats(s2, sn)
    ps = float(st2[0, 0]) - float(st2[0, 4]) ** 2 / n
    return 10 * math.log(ps / float(st2[0, 1]), 10)


def signaltonoise(a, axis=0, ddof=0):
    """`uformerWM/evaluate.py:133-137` for 1-D input."""
    x = _cuda1d(np.asanyarray(a).reshape(-1))
    return float(signaltonoise_from_stats(wave_stats(x, x))[0])


def bit_error_rate(decoded, message):
    """`hidden/test_model.py:60-64`."""
    st = wm_stats(torch.as_tensor(decoded).cuda(), torch.as_tensor(message).cuda())
    return float(st[:, 0].sum() / (st.shape[0] * 1024))


RESULT_LINE = ('Result on {} set, attack: {}: Total clips: {}, MSE loss {}, WM loss: {}, WM loss after attack: {}, '
               'SNR score: {}, PESQ score: {}\n')


def format_result(data_cat, attack, clips_total, mse, wm_loss, wm_loss_att, snr, pesq="N/A"):
    """The `sample_result.txt` line of `uformerWM/evaluate.py:289-291` (parsed by result_extract.py:14).
    PESQ needs the third-party pypesq and is reported as N/A."""
    return RESULT_LINE.format(data_cat, attack, clips_total, mse, wm_loss, wm_loss_att, snr, pesq)


def test(model, messages, waves, data_cat='train', result_path=None, attack=None, audio_scale='0', data_max=None,
         data_min=None, model_name='uformer', draws=None, seed=None, save_audio=False):
    """Batched counterpart of `test()` (`uformerWM/evaluate.py:174-292`): the reference loops over utterances with
    batch 1, appends python floats to lists and averages them; here all utterances (waves (B,L) CUDA - or a LIST of
    1-D CUDA waveforms of different lengths, `audio_test.embed_attack_extract_ragged` - messages (B or 1,1,32,32)) go
    through the hot path in one pass and only the per-utterance statistics vector leaves the GPU.  Appends the reference's line to `<result_path>/sample_result.txt` (if given)
    and returns (line, dict of the averaged numbers).  `save_audio` writes the reference's three WAV files per
    utterance - `audio_gen_sample/<cat>/{ori,recon,<attack>}/<i>.wav` (`evaluate.py:240-247`) - as 32-bit float WAV."""
    from . import audio_test as PT
    ragged = isinstance(waves, (list, tuple))          # a corpus of utterances of different lengths (LibriSpeech, TED-LIUM)
    if ragged:
        if model_name != 'uformer':
            raise NotImplementedError("ragged corpora: Uformer model only")
        r = PT.embed_attack_extract_ragged(list(waves), messages, model, attack or "closed_loop", draws, seed,
                                           audio_scale=audio_scale, data_min=data_min, data_max=data_max)
        n_utt = len(waves)
        clips_total = int(sum(r["n_clips"]))
    else:
        r = PT.embed_attack_extract(waves, messages, model, attack or "closed_loop", draws, seed, want_outputs=bool(save_audio),
                                    audio_scale=audio_scale, data_min=data_min, data_max=data_max, model_name=model_name)
        n_utt = int(waves.shape[0])
        clips_total = int(r["n_clips"]) * n_utt
    if save_audio:
        import os
        from . import wavio
        if not result_path:
            raise ValueError("save_audio needs result_path")
        base = os.path.join(result_path, "audio_gen_sample", data_cat)
        for sub, t in (("ori", waves), ("recon", r["recon"]), (attack or "closed_loop", r["att"])):
            os.makedirs(os.path.join(base, sub), exist_ok=True)
            for i in range(n_utt):
                wavio.write_wav(os.path.join(base, sub, "%d.wav" % i), t[i].detach().float().cpu().numpy(), 16000)
    s = r["stats"].mean(0).cpu().numpy()
    out = {"clips": clips_total, "mse": float(s[1]), "wm_loss": float(s[2]), "wm_loss_att": float(s[3]), "snr": float(s[0]),
           "ber_clean": float(r["stats"][:, 4].sum() / (1024.0 * n_utt)),
           "ber_att": float(r["stats"][:, 5].sum() / r["stats"][:, 6].sum())}
    line = format_result(data_cat, attack, clips_total, out["mse"], out["wm_loss"], out["wm_loss_att"], out["snr"])
    if result_path:
        with open('{}/sample_result.txt'.format(result_path), 'a') as f:
            f.write('\n')
            f.write(line)
    return line, out

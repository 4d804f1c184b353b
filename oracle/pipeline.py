"""CPU restatement of the embed -> attack -> extract driver.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows `reconstruct_audio`
(`uformerWM/audio_test.py:528-785`, data_mode='stft', model_name='uformer', audio_scale='0')
and the dataset front end `SpeechDataTest.prepare_data` (`uformerWM/audio_test.py:299-348`)
line by line, including quirks B-6/B-7/B-8 of SURVEY.md Appendix B, with the model replaced by
the functional oracle in ``oracle/uformer.py`` (bit-identical to the reference module).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import uformer as U
from . import signal as S


def _scale_clip(data_clip, audio_scale, data_min, data_max):
    """`uformerWM/audio_test.py:329-341` (and `:691-702`), verbatim arithmetic."""
    if '-' not in audio_scale:
        if len(audio_scale) > 1:
            data_clip = data_clip * float(audio_scale)
    else:
        min_range, max_range = audio_scale.split('-')
        min_range = float(min_range)
        max_range = float(max_range)
        data_clip = (data_clip - data_min) / (data_max - data_min)
        data_clip = data_clip * (max_range - min_range) + min_range
    return data_clip


def prepare_data(wave_1L, audio_scale='0', data_min=None, data_max=None):
    """`uformerWM/audio_test.py:314-347` for one utterance.  wave (1,L) fp32 torch.
    Returns [ (wave, sr), [clip (1,2,128,128)...], len_last_clip ]."""
    stft = torch.view_as_real(torch.stft(wave_1L, n_fft=255, return_complex=True))   # (1,128,T,2)
    len_pad = 128 - stft.shape[2] % 128
    stft_2 = F.pad(stft, (0, 0, 0, len_pad), mode="constant", value=0)
    clips = []
    for j in range(stft_2.shape[2] // 128):
        c = _scale_clip(stft_2[:, :, 128 * j:128 * (j + 1), :], audio_scale, data_min, data_max)
        clips.append(c.permute(0, 3, 1, 2).squeeze(0).unsqueeze(0))   # DataLoader(bs=1) adds the batch dim
    return [(wave_1L, 16000), clips, stft.shape[2] % 128]


def reconstruct_audio(audio_data, watermark, sd, n_fft=255, attack="closed_loop", draws=None, tiles=None,
                      audio_scale='0', data_min=None, data_max=None, model_name='uformer'):
    """`uformerWM/audio_test.py:528-785`.  Returns the reference's 10-tuple plus a dict of
    extras (logits) used by the parity tests.

    `tiles` (K,1,32,32) is the 64x64 extension of BASELINE config 4 (the reference hard-codes 32x32
    messages, `uformerWM/model.py:2388-2404`): clip j embeds / is scored against tile j mod K instead
    of `watermark`; everything else is the reference's loop unchanged."""
    clips = audio_data[1]
    if tiles is not None:
        K = tiles.shape[0]
        wm_of = lambda j: tiles[j % K][None]
    else:
        wm_of = lambda j: watermark
    len_last_clip = audio_data[2]
    preds, wm_losses, wm_losses_att, wms_decode = [], [], [], []
    logits_clean = []
    with torch.no_grad():
        for i, clip in enumerate(clips):
            if model_name == 'uformer':
                audio_clip, _, _, wm_decode, lg = U.forward(sd, clip, wm_of(i), return_logits=True)    # `:553`
            else:                                       # `:555`: sd is then a ModelA-like module (oracle/cnn.py)
                audio_clip, wm_decode = sd(clip, wm_of(i))
                lg = wm_decode
            if len(audio_scale) > 1:                    # rescale to the audio value range, `:559-571`
                if '-' not in audio_scale:
                    audio_clip = audio_clip * (1 / float(audio_scale))
                else:
                    min_range, max_range = (float(v) for v in audio_scale.split('-'))
                    audio_clip = (audio_clip - min_range) / (max_range - min_range)
                    audio_clip = audio_clip * (data_max - data_min) + data_min
            wms_decode.append(wm_decode.numpy())
            logits_clean.append(lg.numpy())
            if i != len(clips) - 1:
                preds.append(audio_clip.numpy())
            else:
                preds.append(audio_clip[:, :, :, :len_last_clip].numpy())                       # `:595`
        spec = torch.from_numpy(np.concatenate(preds, axis=3)).squeeze(0).permute(1, 2, 0)         # `:596-597`
        recon_audio = torch.istft(torch.view_as_complex(spec.contiguous()), n_fft=n_fft,
                                  length=audio_data[0][0].shape[-1], return_complex=False)        # `:598-600`
        mse_loss = torch.nn.MSELoss()(audio_data[0][0].squeeze(), recon_audio).item()             # `:618`
        wm_losses.append(torch.nn.MSELoss()(wm_of(len(clips) - 1), wm_decode).item())             # `:625` (last clip only)
        audio_att = S.apply_attack(recon_audio.numpy(), attack, draws)                            # `:631-660`
        feat = torch.view_as_real(torch.stft(torch.from_numpy(np.ascontiguousarray(audio_att)), n_fft=255,
                                             return_complex=True))                                # `:677`
        len_pad = 128 - feat.shape[2] % 128                                                       # `:681` (quirk B-7)
        feat = F.pad(feat, (0, 0, 0, len_pad), mode="constant", value=0)
        feat = feat.permute(2, 0, 1).unsqueeze(0)
        wms_att_decode, logits_att = [], []
        for j in range(feat.shape[3] // 128):
            data_clip = feat[:, :, :, 128 * j:128 * (j + 1)].float()
            if len(audio_scale) > 1:                                                              # `:691-702`
                data_clip = _scale_clip(data_clip, audio_scale, data_min, data_max)
            if model_name == 'uformer':
                wm_att, lg = U.wm_decode(sd, data_clip, return_logits=True)                       # `:706`
            else:
                wm_att = sd.decode(data_clip)                                                     # `:708`
                lg = wm_att
            wms_att_decode.append(wm_att.numpy())
            logits_att.append(lg.numpy())
            wm_losses_att.append(torch.nn.MSELoss()(wm_of(j), wm_att).item())                     # `:712`
    snr_ori = S.signaltonoise(audio_data[0][0].squeeze().numpy())
    snr_recon = S.signaltonoise(recon_audio.numpy())
    out = (audio_att, recon_audio, (watermark if tiles is None else tiles).numpy(), wms_decode, wms_att_decode, mse_loss,
           np.mean(wm_losses), np.mean(wm_losses_att), snr_ori, snr_recon)
    extras = {"logits_clean": logits_clean, "logits_att": logits_att}
    if tiles is not None:          # image-level recovery: mean of the sigmoids of the clips that carried each tile
        rec = np.stack([np.mean([w[0] for j, w in enumerate(wms_att_decode) if j % K == t], axis=0) for t in range(K)])
        extras["image_att"] = rec                                                                 # (K,1,32,32)
    return out, extras


def evaluate_utterance(wave_1L, watermark, sd, attack, draws=None):
    """One iteration of `test()` `uformerWM/evaluate.py:189-206` without PESQ / file output:
    returns dict(mse, wm_loss, wm_loss_att, snr, ber_clean, ber_att, clips)."""
    data = prepare_data(wave_1L)
    out, ex = reconstruct_audio(data, watermark, sd, attack=attack, draws=draws)
    att_audio = out[0]
    snr = S.cal_snr(wave_1L.numpy().reshape(-1), att_audio)
    msg = out[2]
    return {
        "mse": out[5], "wm_loss": float(out[6]), "wm_loss_att": float(out[7]), "snr": float(snr),
        "ber_clean": float(np.mean([S.bit_error_rate(w, msg) for w in out[3]])),
        "ber_att": float(np.mean([S.bit_error_rate(w, msg) for w in out[4]])),
        "clips": len(data[1]), "out": out, "extras": ex,
    }

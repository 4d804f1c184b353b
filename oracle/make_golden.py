"""Generate the committed golden fixtures under ``tests/golden/`` by running the UNMODIFIED
reference (``/root/reference``) in the build container.

    python -m oracle.make_golden

TEST INFRASTRUCTURE.  The reference ships no golden vectors, so parity is
pinned on outputs of the reference itself:

* ``model_<init>.npz``   - `UformerAudio.forward` / `wm_decode` (`uformerWM/model.py:2384-2511,
                           2379-2382`) of the reference nn.Module on seeded inputs with weights
                           from ``synthetic.init_state_dict`` (regenerated, not stored).
* ``pipeline_cfg1_<attack>.npz`` - the reference driver `reconstruct_audio`
                           (`uformerWM/audio_test.py:528-785`) executed unmodified (CPU proxy for
                           the hard-coded 'cuda' device) on BASELINE config 1: one 1 s utterance,
                           32x32 binary image, with the reference's own numpy attacks.
* ``signal.npz``         - reference attack / metric functions on a seeded waveform.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import shims, uformer as O, pipeline as P           # noqa: E402
from image_in_speech_watermarking_b200 import synthetic as SY   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_module(kind, seed):
    m = shims.build_reference_uformer_audio(0)
    sd = SY.init_state_dict(O.state_dict_schema(), kind, seed)
    m.load_state_dict(sd, strict=True)
    return m.eval(), sd


def model_fixture(kind, seed):
    m, sd = reference_module(kind, seed)
    wave = SY.synth_speech(0, 1.0)[None]
    data = P.prepare_data(wave)
    x = torch.cat(data[1], 0)                                   # (2,2,128,128) real STFT clips
    msg = torch.stack([SY.synth_image_binary(0), SY.synth_image_binary(1)])
    g = torch.Generator().manual_seed(99)
    x_att = x + 0.05 * torch.randn(x.shape, generator=g)
    with torch.no_grad(), shims.legacy_torch_spectral():
        stft_new, noise, wm_pred, wm = m(x, msg)
        wm_att = m.wm_decode(x_att)
    np.savez_compressed(os.path.join(OUT, "model_%s.npz" % kind), seed=seed, x=x.numpy(), msg=msg.numpy(),
                        x_att=x_att.numpy(), stft_new=stft_new.numpy(), noise=noise.numpy(),
                        wm_pred=wm_pred.numpy(), wm=wm.numpy(), wm_att=wm_att.numpy())
    print("model_%s: wm range %.4f..%.4f" % (kind, wm.min(), wm.max()))


def feature_extract_fixture(kind="stress", seed=0):
    """`UformerAudio.feature_extract` (`uformerWM/model.py:2345-2377`) of the unmodified reference on the inputs of
    `model_<kind>.npz`: wm_pred (the image codec's own reconstruction) and a subsample of y = x + noise."""
    m, sd = reference_module(kind, seed)
    wave = SY.synth_speech(0, 1.0)[None]
    x = torch.cat(P.prepare_data(wave)[1], 0)
    msg = torch.stack([SY.synth_image_binary(0), SY.synth_image_binary(1)])
    with torch.no_grad(), shims.legacy_torch_spectral():
        y, wm_pred = m.feature_extract(x, msg)
    np.savez_compressed(os.path.join(OUT, "feature_extract.npz"), kind=kind, seed=seed, wm_pred=wm_pred.numpy(),
                        y_s8=y.numpy()[:, :, ::8, ::8])
    print("feature_extract.npz", tuple(y.shape), tuple(wm_pred.shape))


def pipeline_fixture(kind, seed, attack):
    m, sd = reference_module(kind, seed)
    run = shims.reference_reconstruct_audio()
    wave = SY.synth_speech(0, 1.0)[None]
    data = P.prepare_data(wave)
    msg = SY.synth_image_binary(0)[None]
    np.random.seed(2024)
    random.seed(2024)
    state = np.random.get_state()
    with torch.no_grad():
        out = run(data, msg, m, attack=attack)
    draws = {}
    if attack.startswith("awgn"):
        np.random.set_state(state)
        draws["awgn"] = np.random.normal(0, 1.0, wave.shape[-1])
    audio_att, recon, wmk, wms, wms_att, mse, wml, wml_att, snr_o, snr_r = out
    name = "pipeline_cfg1_%s.npz" % attack.replace("-", "_")
    np.savez_compressed(os.path.join(OUT, name), seed=seed, kind=kind, attack=attack,
                        audio_att=np.asarray(audio_att), recon=recon.numpy(), wms=np.stack(wms),
                        wms_att=np.stack(wms_att), mse=mse, wm_loss=wml, wm_loss_att=wml_att,
                        snr_ori=snr_o, snr_recon=snr_r, awgn_unit=draws.get("awgn", np.zeros(0)))
    print(name, "mse %.3e wm_loss %.4f wm_loss_att %.4f" % (mse, wml, wml_att))


def signal_fixture():
    att = shims.reference_attack_functions()
    met = shims.reference_metric_functions()
    x = SY.synth_speech(3, 1.0).numpy()
    np.random.seed(7)
    st = np.random.get_state()
    a = att["awgn"](x, snr=20)
    np.random.set_state(st)
    unit = np.random.normal(0, 1.0, x.shape)
    random.seed(11)
    jit = att["jittering_2"](x.copy(), 200)
    random.seed(11)
    idx = np.array([random.randint(0, len(x) - 1) for _ in range(200)])
    lp = att["low_pass_filter"](x)
    np.savez_compressed(os.path.join(OUT, "signal.npz"), x=x, awgn20=a, awgn_unit=unit,
                        low_pass=lp, echo=att["echo_addition"](x), scale07=att["amplitude_scaling"](x, 0.7),
                        jitter=jit, jitter_idx=idx, cal_snr=met["cal_snr"](x, lp),
                        signaltonoise=met["signaltonoise"](x),
                        snr_singlech=met["SNR_singlech"](x.astype(np.float64), lp))
    print("signal.npz")


def cnn_fixture():
    """ModelA + HiDDeN Decoder of the unmodified reference, eval mode, randomised parameters and
    BatchNorm running statistics; noise layers with numpy seeds."""
    from oracle import cnn as C
    ref = shims.import_reference_model()
    ma = C.randomize_(ref.ModelA(), 11)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 2, 128, 128, generator=g)
    wm = (torch.rand(2, 1, 32, 32, generator=g) > 0.5).float()
    hid = shims.reference_hidden_modules()
    cfg = hid["options"].HiDDenConfiguration(H=128, W=128, message_length=30, encoder_blocks=4, encoder_channels=64,
                                             decoder_blocks=7, decoder_channels=64, use_discriminator=True,
                                             use_vgg=False, discriminator_blocks=3, discriminator_channels=64,
                                             decoder_loss=1, encoder_loss=0.7, adversarial_loss=1e-3)
    dec = C.randomize_(hid["decoder"].Decoder(cfg), 12)
    xd = torch.randn(2, 1, 128, 128, generator=g)
    with torch.no_grad():
        enc, ext = ma(x, wm)
        dout = dec(xd)
    out = dict(x=x.numpy(), wm=wm.numpy(), modelA_encoded=enc.numpy(), modelA_extracted=ext.numpy(),
               modelA_decode_x=ma.decode(x).detach().numpy(), dec_x=xd.numpy(), dec_out=dout.numpy())
    # noise layers: [noised, cover] on (2,2,64,64) tensors
    noised = torch.randn(2, 2, 64, 64, generator=g)
    cover = torch.randn(2, 2, 64, 64, generator=g)
    out.update(noised=noised.numpy(), cover=cover.numpy())
    layers = {"crop": hid["crop"].Crop((0.4, 0.55), (0.4, 0.55)), "cropout": hid["cropout"].Cropout((0.25, 0.35), (0.25, 0.35)),
              "dropout": hid["dropout"].Dropout((0.25, 0.35)), "resize": hid["resize"].Resize((0.4, 0.6)),
              "quant": hid["quantization"].Quantization(torch.device("cpu")), "identity": hid["identity"].Identity()}
    for i, (name, layer) in enumerate(layers.items()):
        np.random.seed(100 + i)
        with torch.no_grad():
            r = layer([noised.clone(), cover.clone()])
        out["noise_" + name] = r[0].numpy().astype(np.float32)
    # JpegCompression is hard-wired to 3 channels (jpeg_compression.py:53-55): (2,3,36,44) exercises the padding
    rgb = torch.rand(2, 3, 36, 44, generator=g)
    with torch.no_grad():
        jr = hid["jpeg_compression"].JpegCompression(torch.device("cpu"))([rgb.clone(), rgb.clone()])
    out.update(jpeg_in=rgb.numpy(), noise_jpeg=jr[0].numpy().astype(np.float32))
    np.savez_compressed(os.path.join(OUT, "cnn.npz"), **out)
    print("cnn.npz", {k: v.shape for k, v in out.items() if k.startswith("noise_")})


def modelA_train_fixture():
    """One training step of the UNMODIFIED reference ModelA (train mode, nn.Dropout drawn from a seeded torch
    RNG; the keep mask is recovered from the Dropout module's input/output with a hook - where the input is 0
    the mask is irrelevant and recorded as 1).  Saves inputs, mask, outputs, losses, all gradients (flat, in
    named_parameters order) and the updated BatchNorm running statistics."""
    from oracle import cnn as C
    ref = shims.import_reference_model()
    m = C.randomize_(ref.ModelA(), 11)
    m.train()
    g = torch.Generator().manual_seed(31)
    x = torch.rand(4, 2, 128, 128, generator=g)            # the sigmoid output regresses onto [0,1] spectrograms
    wm = (torch.rand(4, 1, 32, 32, generator=g) > 0.5).float()
    rec = {}

    def hook(mod, inp, out):
        rec["mask"] = ((out != 0) | (inp[0] == 0)).float()
    m.embedder_decoder[3].register_forward_hook(hook)
    torch.manual_seed(5)
    enc, ext = m(x, wm)
    loss1 = torch.nn.MSELoss()(x, enc)                     # train_modelA.py:435
    loss2 = torch.nn.MSELoss()(ext, wm)                    # :445
    (loss1 + loss2).backward()
    names = [n for n, _ in m.named_parameters()]
    flat = torch.cat([p.grad.reshape(-1) for _, p in m.named_parameters()])
    out = dict(x=x.numpy(), wm=wm.numpy(), keep_mask=rec["mask"].numpy().astype(np.uint8), encoded=enc.detach().numpy(),
               extracted=ext.detach().numpy(), loss1=np.float64(loss1.item()), loss2=np.float64(loss2.item()),
               grads_flat=flat.numpy(), names=np.array(names))
    for k, v in m.state_dict().items():
        if "running" in k:
            out["bn." + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "modelA_train.npz"), **out)
    print("modelA_train.npz", float(loss1), float(loss2), flat.shape)


def jitter_delete_fixture():
    """`jittering` (sample deletion, `uformerWM/audio_attack.py:156-173`) executed unmodified; the python RNG is
    re-seeded until the reference's inclusive `randint(0, len)` stays in bounds (np.delete raises otherwise)."""
    att = shims.reference_attack_functions()
    x = SY.synth_speech(3, 1.0).numpy()
    seed = 11
    while True:
        random.seed(seed)
        idx = np.array([random.randint(0, len(x)) for _ in range(1000)])
        if idx.max() < len(x):
            break
        seed += 1
    random.seed(seed)
    y = att["jittering"](x.copy())
    np.savez_compressed(os.path.join(OUT, "jitter_delete.npz"), seed=seed, idx=idx, out=y)
    print("jitter_delete.npz", seed, len(x), len(y), len(np.unique(idx)))


def train_frontend_fixture():
    """`SpeechDataTrain.prepare_data` (unmodified reference, `uformerWM/audio_test.py:439-502`) on three seeded
    utterances: 16 000 samples (T = 126: one clip), 16 300 (T = 128: the reference appends an empty clip),
    24 000 (T = 188); audio_scale '0' (list of clips), '10' (x 10 + global min / max) and, on the single-clip
    utterance alone, '0-1' (the reference's min-max branch only runs when the dataset holds ONE clip:
    `min_values.view(c, 1, 1, 1, 1)` of a scalar, `audio_test.py:49-50`)."""
    run = shims.reference_prepare_data_train()
    lens = [16000, 16300, 24000]
    waves = [SY.synth_speech(20 + i, 1.5)[:L].reshape(1, L).clone() for i, L in enumerate(lens)]
    d0, _, _ = run(waves, "0")
    d0 = torch.stack(d0)                                     # (5, 1, 128, 128, 2)
    d10, mn10, mx10 = run(waves, "10")
    d01, mn01, mx01 = run(waves[:1], "0-1")
    np.savez_compressed(os.path.join(OUT, "train_frontend.npz"), lens=np.array(lens),
                        wave0=waves[0].numpy(), wave1=waves[1].numpy(), wave2=waves[2].numpy(),
                        data0=d0.numpy(), data10_s8=d10.numpy()[:, :, ::8, ::8], min10=float(mn10), max10=float(mx10),
                        data01=d01.numpy(), min01=float(mn01.reshape(-1)[0]), max01=float(mx01.reshape(-1)[0]))
    print("train_frontend.npz", tuple(d0.shape), tuple(d10.shape), tuple(d01.shape), float(mn10), float(mx10))


def uformer_train_fixture():
    """One training step of the UNMODIFIED reference `UformerAudio` in train mode (`audio_uformer_stft.py:452-482`: forward
    with stochastic depth - the reference's default drop_path_rate 0.1 -, the four losses, backward) on two clips.  The
    DropPath factors each block drew (mask / keep per sample, attention branch then MLP branch) are recovered with
    forward hooks on the DropPath modules so that the CUDA path can replay them.  275 MB of gradients are not committed:
    every tensor of <= 4096 elements is stored whole, larger ones as their L2 norm, their sum, 256 elements at seeded
    positions and a projection on a seeded +-1 vector."""
    m, sd = reference_module("stress", 0)
    m.train()
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 2, 128, 128, generator=g) * 0.5
    msg = (torch.rand(2, 1, 32, 32, generator=g) > 0.5).float()
    scales = {}

    def make_hook(name):
        def hook(mod, inp, out):
            a, b = inp[0].detach(), out.detach()
            n = a.shape[0]
            idx = a.reshape(n, -1).abs().argmax(1)                 # a non-zero element of every sample
            num, den = b.reshape(n, -1)[torch.arange(n), idx], a.reshape(n, -1)[torch.arange(n), idx]
            scales.setdefault(name, []).append((num / den).numpy().astype(np.float32))
        return hook
    for name, mod in m.named_modules():
        if type(mod).__name__ == "_DropPath" and mod.drop_prob > 0:
            mod.register_forward_hook(make_hook(name[:-len("drop_path")]))
    torch.manual_seed(13)
    with shims.legacy_torch_spectral():
        audio, audio_noise, wm_gen, wm_decode = m(x, msg)
    mse = torch.nn.MSELoss()
    loss1 = mse(audio, x)                                          # audio_uformer_stft.py:463
    noise_norm = torch.norm(audio_noise) / audio_noise.shape[0]    # :471
    loss4 = mse(noise_norm, torch.ones_like(noise_norm))
    loss2 = mse(wm_gen, msg)                                       # :474
    loss3 = mse(wm_decode, msg)                                    # :476
    (loss1 + loss2 + loss3 + loss4).backward()
    out = dict(x=x.numpy(), msg=msg.numpy(), losses=np.array([loss1.item(), loss2.item(), loss3.item(), loss4.item()]))
    for k, v in scales.items():
        assert len(v) == 2, (k, len(v))
        out["drop." + k] = np.stack(v)                            # (2 branches, n)
    rng = np.random.default_rng(2024)
    names = []
    for k, p_ in m.named_parameters():
        if p_.grad is None:
            continue
        names.append(k)
        gflat = p_.grad.reshape(-1).double().numpy()
        if gflat.size <= 4096:
            out["g." + k] = gflat.astype(np.float32)
        else:
            pos = rng.integers(0, gflat.size, 256)
            sign = rng.integers(0, 2, gflat.size) * 2.0 - 1.0
            out["s." + k] = np.concatenate([[np.sqrt((gflat ** 2).sum()), gflat.sum(), (gflat * sign).sum()], gflat[pos]])
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "uformer_train.npz"), **out)
    print("uformer_train.npz", out["losses"], len(names), "tensors,", len(scales), "DropPath modules")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1 and sys.argv[1] == "train_frontend":
        train_frontend_fixture()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "feature_extract":
        feature_extract_fixture()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "jitter_delete":
        jitter_delete_fixture()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "uformer_train":
        uformer_train_fixture()
        return
    signal_fixture()
    jitter_delete_fixture()
    train_frontend_fixture()
    cnn_fixture()
    modelA_train_fixture()
    uformer_train_fixture()
    model_fixture("stress", 0)
    model_fixture("reference", 0)
    feature_extract_fixture()
    pipeline_fixture("stress", 0, "awgn-20")
    pipeline_fixture("stress", 0, "low_pass")


if __name__ == "__main__":
    main()

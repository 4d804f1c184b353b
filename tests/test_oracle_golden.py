"""The CPU oracle against the committed golden fixtures (outputs of the unmodified reference,
made by ``oracle/make_golden.py``).  Runs everywhere; no GPU, no /root/reference."""
import numpy as np
import torch

from oracle import uformer as O, signal as S, pipeline as P
from image_in_speech_watermarking_b200 import synthetic as SY


def _t(a):
    return torch.from_numpy(a)


def test_model_forward_matches_reference(golden, weights):
    for kind in ("stress", "reference"):
        g = golden("model_%s.npz" % kind)
        sd = weights(kind, int(g["seed"]))
        with torch.no_grad():
            s, n, wp, wm = O.forward(sd, _t(g["x"]), _t(g["msg"]))
            wa = O.wm_decode(sd, _t(g["x_att"]))
        # same torch build => bit-exact; 1e-5 leaves room for a different CPU's GEMM blocking
        for got, ref in ((s, "stft_new"), (n, "noise"), (wp, "wm_pred"), (wm, "wm"), (wa, "wm_att")):
            np.testing.assert_allclose(got.numpy(), g[ref], rtol=0, atol=2e-5, err_msg=kind + ":" + ref)


def test_feature_extract_matches_reference(golden, weights):
    """oracle `feature_extract` vs the unmodified reference (`uformerWM/model.py:2345-2377`): wm_pred is the image
    codec's own reconstruction, not forward()'s bottleneck-added one."""
    g = golden("feature_extract.npz")
    m = golden("model_%s.npz" % str(g["kind"]))
    sd = weights(str(g["kind"]), int(g["seed"]))
    with torch.no_grad():
        y, wp = O.feature_extract(sd, _t(m["x"]), _t(m["msg"]))
    np.testing.assert_allclose(wp.numpy(), g["wm_pred"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(y.numpy()[:, :, ::8, ::8], g["y_s8"], rtol=0, atol=2e-5)
    assert np.abs(g["wm_pred"] - m["wm_pred"]).max() > 1e-3        # the two wm_pred definitions really differ


def test_signal_functions_match_reference(golden):
    g = golden("signal.npz")
    x = g["x"]
    np.testing.assert_allclose(S.awgn(x, 20, noise_unit=g["awgn_unit"]), g["awgn20"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(S.low_pass_filter(x), g["low_pass"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(S.echo_addition(x), g["echo"], rtol=0, atol=0)
    np.testing.assert_allclose(S.amplitude_scaling(x, 0.7), g["scale07"], rtol=0, atol=0)
    np.testing.assert_allclose(S.jittering_2(x, 200, indices=g["jitter_idx"]), g["jitter"], rtol=0, atol=0)
    assert abs(S.cal_snr(x, g["low_pass"]) - float(g["cal_snr"])) < 1e-9
    assert abs(S.signaltonoise(x) - float(g["signaltonoise"])) < 1e-9
    assert abs(S.SNR_singlech(x.astype(np.float64), g["low_pass"]) - float(g["snr_singlech"])) < 1e-9


def test_requantization_restates_the_soundfile_clipping_path():
    """PCM_U8 round trip as python-soundfile + libsndfile do it (SFC_SET_CLIPPING on: `f2uc_clip_array`):
    floor(x * 128) / 128 saturating to [-1, 127/128].  PARITY UNPINNED: no soundfile here to generate a golden."""
    x = np.array([0.0, 0.5, -0.5, 1.0, -1.0, 1.5, -1.5, 0.003, -0.003, 127.0 / 128, 0.999999, -0.999999, 1.0 / 256])
    want = np.array([0.0, 0.5, -0.5, 127.0 / 128, -1.0, 127.0 / 128, -1.0, 0.0, -1.0 / 128, 127.0 / 128, 127.0 / 128,
                     -1.0, 0.0])
    np.testing.assert_array_equal(S.requantization(x), want)
    r = np.random.default_rng(0).uniform(-0.9, 0.9, 10000)
    np.testing.assert_array_equal(S.requantization(r), np.floor(r * 128.0) / 128.0)


def test_jitter_delete_matches_reference(golden):
    """oracle `jittering` (sample deletion) vs the unmodified reference run (`audio_attack.py:156-173`)."""
    g = golden("jitter_delete.npz")
    x = golden("signal.npz")["x"]
    y = S.jittering(x, indices=g["idx"])
    assert y.shape == g["out"].shape == (len(x) - len(np.unique(g["idx"])),)
    np.testing.assert_array_equal(y, g["out"])
    with np.testing.assert_raises(IndexError):
        S.jittering(x, indices=[len(x)])                 # the reference's inclusive randint bound


def test_numpy_stft_istft_match_torch():
    x = SY.synth_speech(5, 1.0)
    ref = torch.view_as_real(torch.stft(x.double()[None], 255, return_complex=True)).numpy()
    got = S.stft(x.numpy()[None])
    assert got.shape == (1, 128, 254, 2)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    z = torch.view_as_complex(torch.from_numpy(ref))
    np.testing.assert_allclose(S.istft(ref, length=16000), torch.istft(z, 255, length=16000).numpy(), atol=1e-13)
    np.testing.assert_allclose(S.istft(ref), torch.istft(z, 255).numpy(), atol=1e-13)


def test_train_frontend_matches_reference(golden):
    """oracle `prepare_data_train` vs the unmodified `SpeechDataTrain.prepare_data` (`audio_test.py:439-502`)."""
    g = golden("train_frontend.npz")
    waves = [g["wave%d" % i].reshape(-1) for i in range(3)]
    ref0 = np.transpose(g["data0"][:, 0], (0, 3, 1, 2))              # (N,1,128,128,2) -> (N,2,128,128)
    d0, mn, mx = S.prepare_data_train(waves, "0")
    assert d0.shape == ref0.shape == (5, 2, 128, 128) and mn == 0 and mx == 0
    scale = np.abs(ref0).max()
    assert np.abs(d0 - ref0).max() < 2e-6 * scale                     # reference is fp32 (torch.stft)
    assert np.abs(d0[2]).max() == 0.0                                 # T = 128: the appended clip is empty
    d10, mn10, mx10 = S.prepare_data_train(waves, "10")
    assert abs(mn10 - float(g["min10"])) < 2e-6 * scale and abs(mx10 - float(g["max10"])) < 2e-6 * scale
    ref10 = np.transpose(g["data10_s8"][:, 0], (0, 3, 1, 2))
    assert np.abs(d10[:, :, ::8, ::8] - ref10).max() < 2e-5 * scale
    d01, mn01, mx01 = S.prepare_data_train(waves[:1], "0-1")
    ref01 = np.transpose(g["data01"][:, 0], (0, 3, 1, 2))
    assert np.abs(d01 - ref01).max() < 2e-6
    assert abs(mn01 - float(g["min01"])) < 2e-6 * scale and abs(mx01 - float(g["max01"])) < 2e-6 * scale
    # the numpy restatement against torch.stft itself (float64)
    x = SY.synth_speech(9, 0.7)
    ref = torch.view_as_real(torch.stft(x.double(), 256, hop_length=128, win_length=256, return_complex=True))[:-1]
    np.testing.assert_allclose(S.stft_train(x.numpy()), ref.numpy(), rtol=0, atol=1e-11)


def test_pipeline_matches_reference_driver(golden, weights):
    for name in ("pipeline_cfg1_awgn_20.npz", "pipeline_cfg1_low_pass.npz"):
        g = golden(name)
        sd = weights(str(g["kind"]), int(g["seed"]))
        wave = SY.synth_speech(0, 1.0)[None]
        msg = SY.synth_image_binary(0)[None]
        draws = {"awgn": g["awgn_unit"]} if g["awgn_unit"].size else None
        out, _ = P.reconstruct_audio(P.prepare_data(wave), msg, sd, attack=str(g["attack"]), draws=draws)
        np.testing.assert_allclose(out[1].numpy(), g["recon"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(out[0], g["audio_att"], rtol=0, atol=1e-6)
        np.testing.assert_allclose(np.stack(out[3]), g["wms"], rtol=0, atol=1e-5)
        np.testing.assert_allclose(np.stack(out[4]), g["wms_att"], rtol=0, atol=1e-5)
        assert abs(out[5] - float(g["mse"])) < 1e-7
        assert abs(out[6] - float(g["wm_loss"])) < 1e-6
        assert abs(out[7] - float(g["wm_loss_att"])) < 1e-6
        assert abs(out[8] - float(g["snr_ori"])) < 1e-6
        assert abs(out[9] - float(g["snr_recon"])) < 1e-4


def test_edge_cases_clip_quirks():
    # quirk B-6: T % 128 == 0 appends an empty clip and len_last_clip == 0
    L = 63 * 127 + 1                                # T = 1 + (L-1)//63 = 128
    spec = S.stft(np.zeros((1, L)))
    clips, last = S.clip_spectrogram(spec)
    assert spec.shape[2] == 128 and len(clips) == 2 and last == 0
    # quirk B-7: attacked clips = (T + 126) // 128
    for L in (16000, 63 * 127, 63 * 127 + 1, 63 * 128, 48000):
        T = 1 + (L - 1) // 63
        assert len(S.attacked_clips(np.zeros(L))) == (T + 126) // 128
    # BER: numpy round is half-to-even
    assert S.bit_error_rate(np.array([0.5, 1.5, 0.49, 0.51]), np.array([0, 1, 0, 1])) == 0.0


def test_oracle_training_step_matches_reference_golden(golden, weights):
    """The oracle's forward (fp32) + torch autograd with the DropPath factors of `uformer_train.npz` replayed reproduces the
    training step of the UNMODIFIED reference in train mode: the four losses of `audio_uformer_stft.py:463-482` and the
    stored gradients (whole small tensors, norms and samples of the large ones) - the oracle the CUDA training kernels are
    checked against is pinned on the reference itself."""
    g = golden("uformer_train.npz")
    sd32 = weights("stress")
    names = [str(k) for k in g["names"]]
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in sd32.items()}
    drops = {k[5:]: g[k] for k in g if k.startswith("drop.")}
    x, msg = torch.from_numpy(g["x"]), torch.from_numpy(g["msg"])
    with O.drop_scales(drops):
        s, noise, wm_pred, wm = O.forward(sd, x, msg)
    mse = torch.nn.functional.mse_loss
    nn_ = torch.norm(noise) / noise.shape[0]
    losses = [mse(s, x), mse(wm_pred, msg), mse(wm, msg), mse(nn_, torch.ones_like(nn_))]
    sum(losses).backward()
    for a, b in zip(losses, g["losses"]):
        assert abs(float(a) - float(b)) <= 2e-6 * abs(float(b)), (float(a), float(b))
    rng = np.random.default_rng(2024)
    for k in names:
        got = sd[k].grad.reshape(-1).double().numpy()
        if got.size <= 4096:
            ref = g["g." + k].astype(np.float64)
            assert np.abs(got - ref).max() <= 2e-4 * np.abs(ref).max() + 1e-12, k
        else:
            pos = rng.integers(0, got.size, 256)
            rng.integers(0, 2, got.size)                                   # keeps the generator in step with make_golden
            ref = g["s." + k]
            assert abs(np.sqrt((got ** 2).sum()) - ref[0]) <= 2e-4 * ref[0], k
            assert np.abs(got[pos] - ref[3:]).max() <= 2e-4 * np.abs(ref[3:]).max() + 1e-12, k

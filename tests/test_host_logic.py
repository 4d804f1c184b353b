"""CPU tests of host-side logic that needs no GPU: 64x64 image tiling (BASELINE config 4),
statistics vector layout, attack-string grammar."""
import numpy as np
import pytest
import torch

from image_in_speech_watermarking_b200 import audio_test as PT, sharding as SH


def test_tile_untile_roundtrip_and_order():
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 1, 64, 64, generator=g)
    t = PT.tile_image(img)
    assert t.shape == (3, 4, 1, 32, 32)
    assert torch.equal(t[:, 0, 0], img[:, 0, :32, :32]) and torch.equal(t[:, 1, 0], img[:, 0, :32, 32:])
    assert torch.equal(t[:, 2, 0], img[:, 0, 32:, :32]) and torch.equal(t[:, 3, 0], img[:, 0, 32:, 32:])
    assert torch.equal(PT.untile_image(t, 64, 64), img)
    with pytest.raises(ValueError):
        PT.tile_image(torch.zeros(1, 1, 48, 64))


def test_recover_tiled_averages_clips_of_the_same_tile():
    g = torch.Generator().manual_seed(1)
    wm = torch.rand(2, 6, 1, 32, 32, generator=g)                 # 6 clips, 4 tiles: tiles 0,1 seen twice
    rec = PT.recover_tiled(wm, 4)
    assert torch.allclose(rec[:, 0], (wm[:, 0] + wm[:, 4]) / 2) and torch.allclose(rec[:, 1], (wm[:, 1] + wm[:, 5]) / 2)
    assert torch.allclose(rec[:, 2], wm[:, 2]) and torch.allclose(rec[:, 3], wm[:, 3])
    with pytest.raises(ValueError):
        PT.recover_tiled(wm[:, :3], 4)


def test_stats_vector_and_summary():
    st = torch.tensor([[10.0, 1e-4, 0.2, 0.3, 500.0, 3000.0, 6144.0],
                       [12.0, 3e-4, 0.1, 0.2, 520.0, 3100.0, 6144.0]], dtype=torch.float64)
    v = SH.stats_vector(st)
    assert v.shape == (8,) and len(SH.STAT_KEYS) == 8
    s = SH.summarize(SH.allreduce_stats(v))
    assert abs(s["ber_clean"] - 1020.0 / 2048.0) < 1e-12 and abs(s["ber_attacked"] - 6100.0 / 12288.0) < 1e-12
    assert abs(s["mean_snr_db"] - 11.0) < 1e-12 and s["utterances"] == 2
    assert list(SH.shard_range(10, 1, 4)) == [3, 4, 5] and list(SH.shard_range(10, 3, 4)) == [9]


def test_result_line_roundtrips_through_result_extract(tmp_path):
    from image_in_speech_watermarking_b200 import evaluate as EV, result_extract as RX
    text = "\n" + EV.format_result("train", "awgn-20", 30, 1.5e-5, 0.21, 0.24, 19.7) + \
           "\n" + EV.format_result("test", "low_pass", 12, 2.5e-5, 0.20, 0.26, 21.3, pesq=3.1)
    rows = RX.process_data_to_csv(text, str(tmp_path / "results.csv"))
    assert [r["Attack"] for r in rows] == ["awgn-20", "low_pass"] and rows[0]["Total Clips"] == 30
    assert rows[0]["PESQ Score"] == "" and rows[1]["PESQ Score"] == 3.1 and abs(rows[1]["SNR Score"] - 21.3) < 1e-12
    head = open(tmp_path / "results.csv").read().splitlines()[0]
    assert head == "Set,Attack,Total Clips,MSE Loss,WM Loss,WM Loss After Attack,SNR Score,PESQ Score"


def test_wav_roundtrip(tmp_path):
    """float32 WAV dumps of the evaluator (`evaluate.py:240-247`): write -> read is exact; PCM16 / PCM8 files decode
    with libsndfile's scaling."""
    import struct
    import numpy as np
    from image_in_speech_watermarking_b200 import wavio
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(16001) * 0.1).astype(np.float32)
    p = str(tmp_path / "a.wav")
    wavio.write_wav(p, x, 16000)
    y, sr = wavio.read_wav(p)
    assert sr == 16000 and y.shape == (1, 16001) and np.array_equal(y[0], x)
    st = np.stack([x[:100], -x[:100]])
    wavio.write_wav(p, st, 8000)
    y, sr = wavio.read_wav(p)
    assert sr == 8000 and np.array_equal(y, st)
    pcm = (np.arange(-5, 5) * 1000).astype("<i2")
    body = b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 16000, 32000, 2, 16) + b"data" + struct.pack("<I", pcm.nbytes) + pcm.tobytes()
    q = str(tmp_path / "b.wav")
    with open(q, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
    y, sr = wavio.read_wav(q)
    assert np.allclose(y[0], pcm.astype(np.float32) / 32768.0)



def test_attack_seeds_are_fresh_per_call_and_per_rank(monkeypatch):
    """random attacks must not replay one noise realisation: a new key per call, distinct across ranks,
    reproducible under np.random.seed (the reference's own source of randomness)."""
    import numpy as np
    from image_in_speech_watermarking_b200 import audio_attack as AT
    np.random.seed(5)
    a = [AT.fresh_seed() for _ in range(4)]
    assert len(set(a)) == 4 and all(0 <= s < 2 ** 64 for s in a)
    np.random.seed(5)
    AT._CALLS[0] -= 4
    assert [AT.fresh_seed() for _ in range(4)] == a                      # reproducible run
    np.random.seed(5)
    AT._CALLS[0] -= 4
    monkeypatch.setenv("RANK", "3")
    b = [AT.fresh_seed() for _ in range(4)]
    assert not set(a) & set(b)                                           # another rank, other keys


def test_fixed_batch_sharding_covers_every_utterance_once():
    """bench.py --config 4 (strong scaling): 64 utterances over 1 / 2 / 4 / 8 ranks, and ragged cases."""
    from image_in_speech_watermarking_b200 import sharding as SH
    for n, world in ((64, 1), (64, 2), (64, 4), (64, 8), (10, 4), (3, 8)):
        got = [i for r in range(world) for i in SH.shard_range(n, r, world)]
        assert got == list(range(n))
    assert len(SH.shard_range(64, 5, 8)) == 8


def test_flat_adam_layout_and_gradient_gather():
    """`cnn_train.FlatAdam` host logic (no kernel call): every parameter starts on a 64-float boundary of the flat buffer and
    becomes a view of it; `gather_grads` writes the gradients at those offsets with zeros for the padding and for parameters
    without a gradient."""
    import torch
    from image_in_speech_watermarking_b200 import cnn_train as CT
    ps = [torch.nn.Parameter(torch.arange(n, dtype=torch.float32) + 10 * i) for i, n in enumerate((5, 64, 130, 1))]
    opt = CT.FlatAdam(ps)
    assert opt.offsets == [0, 64, 128, 320] and opt.flat.numel() == 384
    for p, o in zip(ps, opt.offsets):
        assert p.data_ptr() == opt.flat.data_ptr() + 4 * o and torch.equal(opt.flat[o:o + p.numel()], p.detach().reshape(-1))
    ps[0].grad, ps[2].grad = torch.ones(5), torch.full((130,), 2.0)
    g = opt.gather_grads()
    assert torch.equal(g[:5], torch.ones(5)) and float(g[5:128].abs().sum()) == 0.0
    assert torch.equal(g[128:258], torch.full((130,), 2.0)) and float(g[258:].abs().sum()) == 0.0
    opt.flat[0] = 99.0
    assert float(ps[0][0]) == 99.0                      # the fused optimiser kernel updates the parameters in place


def test_noiser_resolves_placeholders_and_draws_like_the_reference():
    """`hidden/noise_layers/noiser.py:8-31`: Identity first, the two `--noise` placeholders become their layers, anything else
    is a ValueError, and each call consumes exactly the draw `np.random.choice(layers, 1)` consumes (the reference's pick)."""
    from image_in_speech_watermarking_b200.hidden import noise_layers as NL
    n = NL.Noiser(['JpegPlaceholder', NL.Identity(), 'QuantizationPlaceholder'], None)
    assert [type(l).__name__ for l in n.noise_layers] == ["Identity", "JpegCompression", "Identity", "Quantization"]
    with pytest.raises(ValueError):
        NL.Noiser(['bogus'], None)

    class Tag:                                           # stands in for a layer: returns which one was picked
        def __init__(self, k):
            self.k = k

        def __call__(self, pair):
            return self.k
    tags = [Tag(k) for k in range(1, 5)]
    n = NL.Noiser(tags, None)
    for seed in range(20):
        np.random.seed(seed)
        want = [n.noise_layers.index(np.random.choice(n.noise_layers, 1)[0]) for _ in range(6)]
        after_ref = np.random.get_state()[1][:4].tolist(), np.random.get_state()[2]
        np.random.seed(seed)
        got = [n(None) for _ in range(6)]
        after = np.random.get_state()[1][:4].tolist(), np.random.get_state()[2]
        got = [0 if not isinstance(g, int) else g for g in got]          # Identity returns its input (None)
        assert got == want and after == after_ref


def test_audio_scale_grammar_matches_the_reference_rules():
    """`audio_scale` (`uformerWM/audio_test.py:33-55,329-341`): a one-character string is the identity, 'k' multiplies by
    float(k), 'a-b' is the min-max map of [data_min, data_max] onto [a, b]; the inverse of `:559-571` is (y - shift) / scale."""
    assert PT.scale_params('0') == (1.0, 0.0) and PT.scale_params('5') == (1.0, 0.0)          # len <= 1: untouched
    assert PT.scale_params('10') == (10.0, 0.0) and PT.scale_params('0.5') == (0.5, 0.0)
    sc, sh = PT.scale_params('0-1', data_min=-4.0, data_max=12.0)
    assert abs(sc * -4.0 + sh) < 1e-12 and abs(sc * 12.0 + sh - 1.0) < 1e-12
    sc, sh = PT.scale_params('2-6', data_min=1.0, data_max=3.0)
    x = np.linspace(1.0, 3.0, 7)
    y = x * sc + sh
    assert abs(y[0] - 2.0) < 1e-12 and abs(y[-1] - 6.0) < 1e-12
    assert np.allclose((y - sh) / sc, x, rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        PT.scale_params('0-1')                                                                # a range needs the dataset's min / max

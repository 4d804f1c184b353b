"""The UformerAudio training step (reference `uformerWM/audio_uformer_stft.py:418-549`) on libwmk kernels.

`lewin_block_train` is the LeWin block with its full backward pass (`wmk_lewin_block_train_f32`), the operator the step
differentiates 40 times per pass; around it `uformer_forward_train` assembles `UformerAudio.forward` in train mode (all
stages with stochastic depth, down / up-sampling, the image codec, the in-graph ISTFT -> STFT projection and its adjoint),
`training_losses` the four-term loss (`:463-482`) and `train_step` backward + gradient all-reduce + AdamW (`:538-539`,
`cnn_train.FlatAdam(decoupled=True)`).  torch.autograd only orders the calls: every forward and every gradient is a
hand-written fp32 kernel, pinned against one step of the unmodified reference and against float64 autograd of the oracle
(`tests/test_gpu_parity.py::test_uformer_training_step_*`)."""
import ctypes

import torch

from . import _lib

# parameter order of `wmk_lewin_block_train_f32` (suffixes of the block's state_dict prefix)
BLOCK_PARAMS = ("norm1.weight", "norm1.bias", "modulator.weight", "attn.relative_position_bias_table",
                "attn.qkv.to_q.weight", "attn.qkv.to_q.bias", "attn.qkv.to_kv.weight", "attn.qkv.to_kv.bias",
                "attn.proj.weight", "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.linear1.0.weight",
                "mlp.linear1.0.bias", "mlp.dwconv.0.weight", "mlp.dwconv.0.bias", "mlp.linear2.0.weight",
                "mlp.linear2.0.bias")


def lewin_block_train(x, params, heads, shift, dout=None, drop_scales=None):
    """x (n, H*H, C) CUDA fp32; params: {suffix: tensor} of one block (reference shapes; 'modulator.weight' optional).
    drop_scales (2, n): DropPath factors {0, 1 / keep} of the attention / MLP branch per sample (None: no DropPath).
    Returns out, or (out, dx, {suffix: gradient}) when `dout` is given."""
    lib = _lib.load()
    if not x.is_cuda:
        raise _lib.WmkError("lewin_block_train has no CPU implementation: inputs must be CUDA tensors")
    x = x.detach().contiguous().float()
    n, L, C = x.shape
    H = int(round(L ** 0.5))
    if H * H != L:
        raise ValueError("token count %d is not a square" % L)
    ps, gs, keep, grads = (ctypes.c_void_p * 18)(), (ctypes.c_void_p * 18)(), [], {}
    for i, k in enumerate(BLOCK_PARAMS):
        if k not in params:
            if k != "modulator.weight":
                raise KeyError(k)
            ps[i], gs[i] = None, None
            continue
        t = params[k].detach().contiguous().float().cuda()
        keep.append(t)
        ps[i] = t.data_ptr()
        if dout is not None:
            grads[k] = torch.empty_like(t)
            gs[i] = grads[k].data_ptr()
    out = torch.empty_like(x)
    ds = None
    if drop_scales is not None:
        ds = drop_scales.detach().contiguous().float().cuda()
        if ds.shape != (2, n):
            raise ValueError("drop_scales must be (2, %d)" % n)
    if dout is None:
        _lib.check(lib.wmk_lewin_block_train_f32(_lib.ptr(x), None, ps, None, _lib.ptr(out), None, n, H, C, heads, shift,
                                                 _lib.ptr(ds), _lib.stream_ptr()))
        return out
    dout = dout.detach().contiguous().float()
    dx = torch.empty_like(x)
    _lib.check(lib.wmk_lewin_block_train_f32(_lib.ptr(x), _lib.ptr(dout), ps, gs, _lib.ptr(out), _lib.ptr(dx), n, H, C, heads,
                                             shift, _lib.ptr(ds), _lib.stream_ptr()))
    return out, dx, grads


# ------------------------------------------------------------------------------------------------------------------
# The extractor (`EncoderTransformerWM` + the image codec's decoder, `uformerWM/model.py:1568-1583,1711-1718,2379-2382`)
# as a differentiable graph of libwmk kernels: torch.autograd only orders the backward calls, every forward and every
# gradient is a hand-written CUDA kernel.  A block is check-pointed: its backward call recomputes its forward.
# ------------------------------------------------------------------------------------------------------------------
DEPTHS, HEADS = (1, 2, 8, 8, 2), (1, 2, 4, 8, 16)


def _c(t):
    return t.detach().contiguous().float()


class _Block(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, heads, shift, names, scales, *tensors):
        out = lewin_block_train(x, dict(zip(names, tensors)), heads, shift, drop_scales=scales)
        ctx.save_for_backward(x, *tensors)
        ctx.meta = (heads, shift, names, scales)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, *tensors = ctx.saved_tensors
        heads, shift, names, scales = ctx.meta
        _, dx, grads = lewin_block_train(x, dict(zip(names, tensors)), heads, shift, dout=dout, drop_scales=scales)
        return (dx, None, None, None, None) + tuple(grads[k].reshape(t.shape) for k, t in zip(names, tensors))


class _Downsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        lib = _lib.load()
        x, w, b = _c(x), _c(w), _c(b)
        n, L, C = x.shape
        H = int(round(L ** 0.5))
        out = torch.empty((n, L // 4, 2 * C), device=x.device, dtype=torch.float32)
        _lib.check(lib.wmk_downsample_train_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), None, None, None, None,
                                                n, H, C, _lib.stream_ptr()))
        ctx.save_for_backward(x, w, b)
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = _lib.load()
        x, w, b = ctx.saved_tensors
        n, L, C = x.shape
        H = int(round(L ** 0.5))
        out = torch.empty((n, L // 4, 2 * C), device=x.device, dtype=torch.float32)
        dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
        _lib.check(lib.wmk_downsample_train_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), _lib.ptr(_c(dout)),
                                                _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), n, H, C, _lib.stream_ptr()))
        return dx, dw, db


class _Head(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conv4, w, b):
        lib = _lib.load()
        conv4, w, b = _c(conv4), _c(w), _c(b)
        n = conv4.shape[0]
        feat = torch.empty((n, 256), device=conv4.device, dtype=torch.float32)
        _lib.check(lib.wmk_extract_head_train_f32(_lib.ptr(conv4), _lib.ptr(w), _lib.ptr(b), _lib.ptr(feat), None, None, None, None,
                                                  n, _lib.stream_ptr()))
        ctx.save_for_backward(conv4, w, b)
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        lib = _lib.load()
        conv4, w, b = ctx.saved_tensors
        n = conv4.shape[0]
        feat = torch.empty((n, 256), device=conv4.device, dtype=torch.float32)
        dc, dw, db = torch.empty_like(conv4), torch.empty_like(w), torch.empty_like(b)
        _lib.check(lib.wmk_extract_head_train_f32(_lib.ptr(conv4), _lib.ptr(w), _lib.ptr(b), _lib.ptr(feat), _lib.ptr(_c(dfeat)),
                                                  _lib.ptr(dc), _lib.ptr(dw), _lib.ptr(db), n, _lib.stream_ptr()))
        return dc, dw, db


class _Leaky(torch.autograd.Function):
    """LeakyReLU(slope); slope 0 = ReLU."""

    @staticmethod
    def forward(ctx, x, slope):
        x = _c(x)
        y = torch.empty_like(x)
        _lib.check(_lib.load().wmk_leaky_relu_f32(_lib.ptr(x), None, _lib.ptr(y), x.numel(), slope, _lib.stream_ptr()))
        ctx.save_for_backward(x)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.empty_like(x)
        _lib.check(_lib.load().wmk_leaky_relu_f32(_lib.ptr(x), _lib.ptr(_c(dy)), _lib.ptr(dx), x.numel(), ctx.slope, _lib.stream_ptr()))
        return dx, None


class _Sigmoid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        y = torch.empty_like(x)
        _lib.check(_lib.load().wmk_sigmoid_f32(_lib.ptr(x), None, _lib.ptr(y), x.numel(), _lib.stream_ptr()))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dx = torch.empty_like(y)
        _lib.check(_lib.load().wmk_sigmoid_f32(_lib.ptr(y), _lib.ptr(_c(dy)), _lib.ptr(dx), y.numel(), _lib.stream_ptr()))
        return dx


class _Transpose(torch.autograd.Function):
    """(n, R, C) -> (n, C, R): NCHW planes <-> token layout."""

    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        n, R, C = x.shape
        y = torch.empty((n, C, R), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().wmk_transpose_batched_f32(_lib.ptr(x), _lib.ptr(y), n, R, C, _lib.stream_ptr()))
        return y

    @staticmethod
    def backward(ctx, dy):
        return _Transpose.apply(dy)


def extractor_forward_train(sd, y, drop_scales=None):
    """`UformerAudio.wm_decode(y)` (`model.py:2379-2382`) in training mode.  sd: {reference state_dict name: CUDA tensor}
    (those that require grad receive gradients); y (n, 2, 128, 128) CUDA; drop_scales: {block prefix: (2, n) DropPath
    factors} (None / missing prefix: no DropPath).  Returns (wm, logits), differentiable."""
    from . import cnn_train as CT
    p = "decoder_wm."
    n = y.shape[0]
    t = CT._Conv3x3.apply(y, sd[p + "input_proj.proj.0.weight"], sd[p + "input_proj.proj.0.bias"])         # model.py:824-829
    t = _Leaky.apply(t, 0.01)
    t = _Transpose.apply(t.reshape(n, 32, 128 * 128))                                                       # tokens (n, 16384, 32)
    for s in range(5):
        C = 32 << s
        lp = p + ("encoderlayer_%d." % s if s < 4 else "conv.")
        for i in range(DEPTHS[s]):
            bp = "%sblocks.%d." % (lp, i)
            names = tuple(k for k in BLOCK_PARAMS if bp + k in sd)
            t = _Block.apply(t, HEADS[s], 0 if i % 2 == 0 else 4, names, (drop_scales or {}).get(bp), *[sd[bp + k] for k in names])
        if s < 4:
            dp = "%sdowsample_%d.conv.0." % (p, s)
            t = _Downsample.apply(t, sd[dp + "weight"], sd[dp + "bias"])
    feat = _Head.apply(t, sd[p + "conv2.weight"], sd[p + "conv2.bias"])                                     # model.py:1580-1582
    q = "encoder_wm."
    h = CT._ConvT2x2.apply(feat.reshape(n, 4, 8, 8), sd[q + "t_conv1.weight"], sd[q + "t_conv1.bias"])      # model.py:1711-1718
    h = _Leaky.apply(h, 0.0)
    logits = CT._ConvT2x2.apply(h, sd[q + "t_conv2.weight"], sd[q + "t_conv2.bias"])
    return _Sigmoid.apply(logits), logits


# ------------------------------------------------------------------------------------------------------------------
# The whole UformerAudio forward in training mode (`uformerWM/model.py:2384-2511`) and the reference's training step
# (`uformerWM/audio_uformer_stft.py:418-549`: four-term loss, AdamW).  DropPath (`model.py:1004,1017`, stochastic depth
# p <= 0.1) is the identity here: pass drop_path_rate = 0 to the reference to compare.
# ------------------------------------------------------------------------------------------------------------------
DEC_DEPTHS, DEC_HEADS = (8, 8, 2, 1), (16, 8, 4, 2)


class _Upsample(torch.autograd.Function):
    """ConvTranspose2d(Cin, Cout, 2, stride 2) on tokens (`model.py:794-800`)."""

    @staticmethod
    def forward(ctx, x, w, b):
        x, w, b = _c(x), _c(w), _c(b)
        n, L, Cin = x.shape
        h, Cout = int(round(L ** 0.5)), w.shape[1]
        out = torch.empty((n, 4 * L, Cout), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().wmk_upsample_train_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), None, None, None, None,
                                                      n, h, Cin, Cout, _lib.stream_ptr()))
        ctx.save_for_backward(x, w, b)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w, b = ctx.saved_tensors
        n, L, Cin = x.shape
        h, Cout = int(round(L ** 0.5)), w.shape[1]
        out = torch.empty((n, 4 * L, Cout), device=x.device, dtype=torch.float32)
        dx, dw, db = torch.empty_like(x), torch.empty_like(w), torch.empty_like(b)
        _lib.check(_lib.load().wmk_upsample_train_f32(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), _lib.ptr(_c(dout)),
                                                      _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), n, h, Cin, Cout, _lib.stream_ptr()))
        return dx, dw, db


class _MaxPool16x8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, conv4):
        conv4 = _c(conv4)
        n = conv4.shape[0]
        out = torch.empty((n, 256), device=conv4.device, dtype=torch.float32)
        _lib.check(_lib.load().wmk_maxpool16x8_f32(_lib.ptr(conv4), None, _lib.ptr(out), n, _lib.stream_ptr()))
        ctx.save_for_backward(conv4)
        return out

    @staticmethod
    def backward(ctx, dy):
        (conv4,) = ctx.saved_tensors
        dc = torch.empty_like(conv4)
        _lib.check(_lib.load().wmk_maxpool16x8_f32(_lib.ptr(conv4), _lib.ptr(_c(dy)), _lib.ptr(dc), conv4.shape[0], _lib.stream_ptr()))
        return dc


class _Projection(torch.autograd.Function):
    """s = STFT(ISTFT(y)) on one-clip spectrograms (n, 2, 128, 128) (`model.py:2458-2463`); backward = the adjoint."""

    @staticmethod
    def forward(ctx, y):
        from . import audio_uformer_stft as FE
        y = _c(y)
        n = y.shape[0]
        wave = FE.istft_clips(y.reshape(n, 1, 2, 128, 128), 128, 8002)
        return FE.stft_clips(wave, 1).reshape(n, 2, 128, 128)

    @staticmethod
    def backward(ctx, ds):
        ds = _c(ds)
        dy = torch.empty_like(ds)
        _lib.check(_lib.load().wmk_stft_projection_adjoint_f32(_lib.ptr(ds), _lib.ptr(dy), ds.shape[0], _lib.stream_ptr()))
        return dy


def _blocks(sd, prefix, t, depth, heads, drop_scales=None):
    for i in range(depth):
        bp = "%sblocks.%d." % (prefix, i)
        names = tuple(k for k in BLOCK_PARAMS if bp + k in sd)
        t = _Block.apply(t, heads, 0 if i % 2 == 0 else 4, names, (drop_scales or {}).get(bp), *[sd[bp + k] for k in names])
    return t


def _encoder(sd, p, y, drop_scales=None):
    """input projection + the five encoder stages -> [conv0 .. conv4] (tokens)."""
    from . import cnn_train as CT
    n = y.shape[0]
    t = CT._Conv3x3.apply(y, sd[p + "input_proj.proj.0.weight"], sd[p + "input_proj.proj.0.bias"])
    t = _Transpose.apply(_Leaky.apply(t, 0.01).reshape(n, 32, 128 * 128))
    convs = []
    q = p if p else "encoder."
    for s in range(5):
        t = _blocks(sd, q + ("encoderlayer_%d." % s if s < 4 else "conv."), t, DEPTHS[s], HEADS[s], drop_scales)
        convs.append(t)
        if s < 4:
            dp = "%sdowsample_%d.conv.0." % (q, s)
            t = _Downsample.apply(t, sd[dp + "weight"], sd[dp + "bias"])
    return convs


def _codec_decode(sd, feat):
    from . import cnn_train as CT
    q = "encoder_wm."
    h = CT._ConvT2x2.apply(feat, sd[q + "t_conv1.weight"], sd[q + "t_conv1.bias"])
    return CT._ConvT2x2.apply(_Leaky.apply(h, 0.0), sd[q + "t_conv2.weight"], sd[q + "t_conv2.bias"])


def uformer_forward_train(sd, x, message, drop_scales=None):
    """`UformerAudio.forward(x, message)` -> (stft_new, noise, wm_pred, wm), differentiable w.r.t. every tensor of `sd`
    that requires grad.  x (n, 2, 128, 128), message (n, 1, 32, 32) CUDA; drop_scales {block prefix: (2, n)}: the DropPath
    factors of this step (`draw_drop_scales`; None = stochastic depth off)."""
    from . import cnn_train as CT
    n = x.shape[0]
    q = "encoder_wm."
    f = _Leaky.apply(CT._Conv3x3.apply(message, sd[q + "conv1.weight"], sd[q + "conv1.bias"]), 0.0)        # model.py:1720-1726
    f = CT._MaxPool.apply(f)
    f = _Leaky.apply(CT._Conv3x3.apply(f, sd[q + "conv2.weight"], sd[q + "conv2.bias"]), 0.0)
    feat_wm = CT._MaxPool.apply(f)                                                                        # (n, 4, 8, 8)
    feat_expand = feat_wm.reshape(n, 4, 64).repeat((1, 16, 8))                                             # (n, 64, 512)
    convs = _encoder(sd, "", x, drop_scales)
    conv4 = convs[4]
    c4ds = _MaxPool16x8.apply(conv4).reshape(n, 4, 8, 8)                                                   # model.py:2398-2400
    wm_pred = _Sigmoid.apply(_codec_decode(sd, feat_wm + c4ds))
    t = torch.cat([feat_expand, conv4], dim=2)                                                             # (n, 64, 1024)
    for s in range(4):                                                                                     # model.py:1221-1240
        up = "decoder.upsample_%d.deconv.0." % s
        t = _Upsample.apply(t, sd[up + "weight"], sd[up + "bias"])
        t = torch.cat([t, convs[3 - s]], dim=-1)
        t = _blocks(sd, "decoder.decoderlayer_%d." % s, t, DEC_DEPTHS[s], DEC_HEADS[s], drop_scales)
    img = _Transpose.apply(t).reshape(n, 64, 128, 128)
    noise = CT._Conv3x3.apply(img, sd["output_proj.proj.0.weight"], sd["output_proj.proj.0.bias"])         # model.py:857-865
    y = x + noise
    s_ = _Projection.apply(y)
    s_ = CT._Conv3x3.apply(s_, sd["stft_layer.0.weight"], sd["stft_layer.0.bias"])
    stft_new = CT._Conv3x3.apply(_Leaky.apply(s_, 0.0), sd["stft_layer.2.weight"], sd["stft_layer.2.bias"])
    wm, _ = extractor_forward_train(sd, y, drop_scales)                                                    # model.py:2508-2509 reads y
    return stft_new, noise, wm_pred, wm


def drop_path_rates(rate=0.1):
    """{block prefix: DropPath probability} as the reference constructors assign them (`model.py:1123-1125,1268-1270,
    1454-1456`): encoder / extractor blocks linspace(0, rate) over their first four stages and `rate` in the fifth, the
    decoder's the reversed encoder list."""
    enc = [float(v) for v in torch.linspace(0, rate, sum(DEPTHS[:4]))]
    out, k = {}, 0
    for s in range(5):
        for i in range(DEPTHS[s]):
            r = enc[k + i] if s < 4 else rate
            for p in ("encoder.", "decoder_wm."):
                out["%s%sblocks.%d." % (p, "encoderlayer_%d." % s if s < 4 else "conv.", i)] = r
        k += DEPTHS[s] if s < 4 else 0
    dec, k = enc[::-1], 0
    for s in range(4):
        for i in range(DEC_DEPTHS[s]):
            out["decoder.decoderlayer_%d.blocks.%d." % (s, i)] = dec[k + i]
        k += DEC_DEPTHS[s]
    return out


def draw_drop_scales(n, rate=0.1, generator=None):
    """One step's DropPath factors: per block (2, n) values in {0, 1 / keep} (Bernoulli(keep) per sample and branch)."""
    out = {}
    for p, r in drop_path_rates(rate).items():
        if r > 0.0:
            keep = 1.0 - r
            out[p] = torch.bernoulli(torch.full((2, n), keep), generator=generator) / keep
    return out


def training_losses(sd, x, message, drop_scales=None):
    """The four terms of `audio_uformer_stft.py:463-482`: loss1 = MSE(audio, target), loss2 = MSE(wm_gen, message),
    loss3 = MSE(wm_decode, message), loss4 = MSE(||noise|| / batch, 1).  Returns (loss, (l1, l2, l3, l4))."""
    from . import cnn_train as CT
    stft_new, noise, wm_pred, wm = uformer_forward_train(sd, x, message, drop_scales)
    l1 = CT.mse_loss(stft_new, x)
    l2 = CT.mse_loss(wm_pred, message)
    l3 = CT.mse_loss(wm, message)
    sq = CT.mse_loss(noise, torch.zeros_like(noise)) * noise.numel()          # sum of squares by the MSE kernel
    l4 = (torch.sqrt(sq) / noise.shape[0] - 1.0) ** 2
    return l1 + l2 + l3 + l4, (l1, l2, l3, l4)


def train_step(sd, optimizer, x, message, drop_path_rate=0.1):
    """One optimisation step (`audio_uformer_stft.py:418-549`, NativeScaler without autocast = plain backward + step);
    `optimizer` is a `cnn_train.FlatAdam(..., decoupled=True)` over the tensors of `sd`; gradients are summed over the
    ranks through `wmk_grad_allreduce_f32` (`sharding.allreduce_grads`)."""
    from . import sharding
    optimizer.zero_grad()
    loss, parts = training_losses(sd, x, message, draw_drop_scales(x.shape[0], drop_path_rate) if drop_path_rate > 0 else None)
    loss.backward()
    world = sharding.allreduce_grads(optimizer.gather_grads())
    optimizer.step(grad_scale=1.0 / world)
    return loss.detach(), tuple(p.detach() for p in parts)

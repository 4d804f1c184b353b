"""Training-mode pieces of UformerAudio (reference step `uformerWM/audio_uformer_stft.py:418-549`): the LeWin block with
its full backward pass on libwmk kernels (`wmk_lewin_block_train_f32`).  The complete training step (all stages,
down / up-sampling, the image codec, the in-graph ISTFT / STFT, the 4-term loss, AdamW) is not assembled yet
(DESIGN 7); this is the operator it differentiates 40 times per pass, pinned against autograd of the oracle."""
import ctypes

import torch

from . import _lib

# parameter order of `wmk_lewin_block_train_f32` (suffixes of the block's state_dict prefix)
BLOCK_PARAMS = ("norm1.weight", "norm1.bias", "modulator.weight", "attn.relative_position_bias_table",
                "attn.qkv.to_q.weight", "attn.qkv.to_q.bias", "attn.qkv.to_kv.weight", "attn.qkv.to_kv.bias",
                "attn.proj.weight", "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.linear1.0.weight",
                "mlp.linear1.0.bias", "mlp.dwconv.0.weight", "mlp.dwconv.0.bias", "mlp.linear2.0.weight",
                "mlp.linear2.0.bias")


def lewin_block_train(x, params, heads, shift, dout=None):
    """x (n, H*H, C) CUDA fp32; params: {suffix: tensor} of one block (reference shapes; 'modulator.weight' optional).
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
    if dout is None:
        _lib.check(lib.wmk_lewin_block_train_f32(_lib.ptr(x), None, ps, None, _lib.ptr(out), None, n, H, C, heads, shift,
                                                 _lib.stream_ptr()))
        return out
    dout = dout.detach().contiguous().float()
    dx = torch.empty_like(x)
    _lib.check(lib.wmk_lewin_block_train_f32(_lib.ptr(x), _lib.ptr(dout), ps, gs, _lib.ptr(out), _lib.ptr(dx), n, H, C, heads,
                                             shift, _lib.stream_ptr()))
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
    def forward(ctx, x, heads, shift, names, *tensors):
        out = lewin_block_train(x, dict(zip(names, tensors)), heads, shift)
        ctx.save_for_backward(x, *tensors)
        ctx.meta = (heads, shift, names)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, *tensors = ctx.saved_tensors
        heads, shift, names = ctx.meta
        _, dx, grads = lewin_block_train(x, dict(zip(names, tensors)), heads, shift, dout=dout)
        return (dx, None, None, None) + tuple(grads[k].reshape(t.shape) for k, t in zip(names, tensors))


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


def extractor_forward_train(sd, y):
    """`UformerAudio.wm_decode(y)` (`model.py:2379-2382`) in training mode.  sd: {reference state_dict name: CUDA tensor}
    (those that require grad receive gradients); y (n, 2, 128, 128) CUDA.  Returns (wm, logits), differentiable."""
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
            t = _Block.apply(t, HEADS[s], 0 if i % 2 == 0 else 4, names, *[sd[bp + k] for k in names])
        if s < 4:
            dp = "%sdowsample_%d.conv.0." % (p, s)
            t = _Downsample.apply(t, sd[dp + "weight"], sd[dp + "bias"])
    feat = _Head.apply(t, sd[p + "conv2.weight"], sd[p + "conv2.bias"])                                     # model.py:1580-1582
    q = "encoder_wm."
    h = CT._ConvT2x2.apply(feat.reshape(n, 4, 8, 8), sd[q + "t_conv1.weight"], sd[q + "t_conv1.bias"])      # model.py:1711-1718
    h = _Leaky.apply(h, 0.0)
    logits = CT._ConvT2x2.apply(h, sd[q + "t_conv2.weight"], sd[q + "t_conv2.bias"])
    return _Sigmoid.apply(logits), logits

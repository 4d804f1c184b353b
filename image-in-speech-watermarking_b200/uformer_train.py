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

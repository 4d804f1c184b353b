"""CUDA execution of small conv stacks described by torch.nn modules (which only hold the
parameters): Conv2d(3x3, pad 1) / ConvTranspose2d(2x2, s 2) + eval-mode BatchNorm2d + activation,
MaxPool2d(2).  Used by ModelA and the HiDDeN Decoder drop-ins; every tensor op is a libwmk kernel."""
import torch
import torch.nn as nn

from . import _lib

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_SIGMOID = 0, 1, 2, 3


def _cuda_f32(x):
    if not x.is_cuda:
        raise _lib.WmkError("the CNN path has no CPU implementation: inputs must be CUDA tensors")
    return x.detach().contiguous().float()


def bn_affine(bn):
    """eval-mode BatchNorm2d as y = scale * x + shift (running statistics)."""
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
    shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
    return scale, shift


def _act_of(mod):
    if isinstance(mod, nn.ReLU):
        return ACT_RELU, 0.0
    if isinstance(mod, nn.LeakyReLU):
        return ACT_LEAKY, float(mod.negative_slope)
    if isinstance(mod, nn.Sigmoid):
        return ACT_SIGMOID, 0.0
    return None


def conv3x3(x, conv, bn=None, act=ACT_NONE, slope=0.0, out=None, out_ch_offset=0):
    lib = _lib.load()
    x = _cuda_f32(x)
    B, Cin, H, W = x.shape
    Cout = conv.out_channels
    if conv.kernel_size != (3, 3) or conv.padding != (1, 1) or conv.stride != (1, 1) or conv.in_channels != Cin:
        raise NotImplementedError("conv3x3 kernel: 3x3, stride 1, padding 1 only")
    if out is None:
        out = torch.empty((B, Cout, H, W), device=x.device, dtype=torch.float32)
    scale, shift = bn_affine(bn) if bn is not None else (None, None)
    bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
    _lib.check(lib.wmk_conv3x3_f32(_lib.ptr(x), _lib.ptr(out), _lib.ptr(conv.weight.detach().float().contiguous()),
                                   _lib.ptr(bias), _lib.ptr(scale), _lib.ptr(shift), B, Cin, Cout, H, W, out_ch_offset,
                                   out.shape[1], act, slope, _lib.stream_ptr()))
    return out


def convT2x2(x, conv, bn=None, act=ACT_NONE, slope=0.0):
    lib = _lib.load()
    x = _cuda_f32(x)
    B, Cin, H, W = x.shape
    Cout = conv.out_channels
    if conv.kernel_size != (2, 2) or conv.stride != (2, 2) or conv.padding != (0, 0) or conv.in_channels != Cin:
        raise NotImplementedError("convT2x2 kernel: kernel 2, stride 2 only")
    out = torch.empty((B, Cout, 2 * H, 2 * W), device=x.device, dtype=torch.float32)
    scale, shift = bn_affine(bn) if bn is not None else (None, None)
    bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
    _lib.check(lib.wmk_convT2x2_f32(_lib.ptr(x), _lib.ptr(out), _lib.ptr(conv.weight.detach().float().contiguous()),
                                    _lib.ptr(bias), _lib.ptr(scale), _lib.ptr(shift), B, Cin, Cout, H, W, act, slope,
                                    _lib.stream_ptr()))
    return out


def maxpool2x2(x):
    lib = _lib.load()
    x = _cuda_f32(x)
    B, C, H, W = x.shape
    out = torch.empty((B, C, H // 2, W // 2), device=x.device, dtype=torch.float32)
    _lib.check(lib.wmk_maxpool2x2_f32(_lib.ptr(x), _lib.ptr(out), B * C, H, W, _lib.stream_ptr()))
    return out


def run_sequential(seq, x):
    """Execute an nn.Sequential of Conv2d/ConvTranspose2d [+ BatchNorm2d] [+ activation] / MaxPool2d /
    Dropout (eval: identity) / nested Sequential blocks with the CUDA kernels above."""
    mods = []

    def flatten(m):
        for c in m.children():
            if isinstance(c, nn.Sequential) or (len(list(c.children())) and not isinstance(
                    c, (nn.Conv2d, nn.ConvTranspose2d, nn.BatchNorm2d))):
                flatten(c)
            else:
                mods.append(c)
    flatten(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            bn, act, slope = None, ACT_NONE, 0.0
            j = i + 1
            if j < len(mods) and isinstance(mods[j], nn.BatchNorm2d):
                if mods[j].training:
                    raise NotImplementedError("BatchNorm in training mode (batch statistics / backward) is outside "
                                              "the inference hot path: call .eval()")
                bn = mods[j]
                j += 1
            if j < len(mods) and _act_of(mods[j]) is not None:
                act, slope = _act_of(mods[j])
                j += 1
            x = conv3x3(x, m, bn, act, slope) if isinstance(m, nn.Conv2d) else convT2x2(x, m, bn, act, slope)
            i = j
        elif isinstance(m, nn.MaxPool2d):
            x = maxpool2x2(x)
            i += 1
        elif isinstance(m, nn.Dropout):
            if m.training:
                raise NotImplementedError("Dropout in training mode is outside the inference hot path: call .eval()")
            i += 1
        else:
            raise NotImplementedError("no CUDA kernel for %s" % type(m).__name__)
    return x

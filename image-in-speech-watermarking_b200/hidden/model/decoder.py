"""The reference's modified HiDDeN `Decoder` (`hidden/model/decoder.py:6-40`): 1 input channel,
decoder_blocks x ConvBNRelu(64) -> ConvBNRelu(message_length) -> MaxPool2 -> ConvBNRelu(1) ->
MaxPool2, i.e. a (B,1,H/4,W/4) image."""
import torch
import torch.nn as nn

from ... import _lib, cnn
from ..options import HiDDenConfiguration
from .conv_bn_relu import ConvBNRelu


class Decoder(nn.Module):
    """precision='fp32' (default): every layer on the direct fp32 kernels (the 1e-3 parity mode).
    precision='bf16': activations NHWC bf16 and the 64 -> 64 / 64 -> message_length ConvBNRelu layers - 92 % of the
    decoder's 7.8 GFLOP per clip - as implicit GEMMs on the tcgen05 tensor cores (2e-2 tolerance class); needs
    decoder_channels == 64, message_length <= 32 and (B, 1, H, 128) inputs, falls back to fp32 kernels otherwise."""

    def __init__(self, config: HiDDenConfiguration, precision='fp32'):
        super().__init__()
        if precision not in ('fp32', 'bf16'):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        self._packed = None
        self.channels = config.decoder_channels
        layers = [ConvBNRelu(1, self.channels)]
        for _ in range(config.decoder_blocks - 1):
            layers.append(ConvBNRelu(self.channels, self.channels))
        layers.append(ConvBNRelu(self.channels, config.message_length))
        layers.append(nn.MaxPool2d(kernel_size=2, stride=2, padding=0))
        layers.append(ConvBNRelu(config.message_length, 1))
        layers.append(nn.MaxPool2d(kernel_size=2, stride=2, padding=0))
        self.layers = nn.Sequential(*layers)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, image_with_wm):
        x = image_with_wm
        mods = list(self.layers)
        if (self.precision == 'bf16' and x.is_cuda and x.dim() == 4 and x.shape[1] == 1 and x.shape[3] == 128 and
                x.shape[2] % 2 == 0 and self.channels == 64 and mods[-4].layers[0].out_channels <= 32 and not self.training):
            return self._forward_tc(x.detach().contiguous().float(), mods)
        return cnn.run_sequential(self.layers, image_with_wm)

    # ------------------------------------------------------------------ tensor-core path
    def _pack(self, mods, dev):
        ver = sum(int(p._version) for p in self.parameters()) + sum(int(b._version) for b in self.buffers())
        if self._packed is not None and self._packed["ver"] == ver and self._packed["dev"] == dev:
            return self._packed
        blocks = [m for m in mods if not isinstance(m, nn.MaxPool2d)]
        first, mids, last = blocks[0], blocks[1:-1], blocks[-1]
        sc0, sh0 = cnn.bn_affine(first.layers[1])
        pk = {"ver": ver, "dev": dev, "first": (first.layers[0].weight.detach().float().reshape(64, 9).contiguous(),
                                                  first.layers[0].bias.detach().float().contiguous(), sc0, sh0), "mids": []}
        for blk in mids:
            conv, bn = blk.layers[0], blk.layers[1]
            sc, sh = cnn.bn_affine(bn)
            co = conv.out_channels
            cp = 64 if co > 32 else 32
            w = conv.weight.detach().float().permute(0, 2, 3, 1).reshape(co, 576) * sc[:, None]      # k = (ky*3+kx)*64 + ci
            wp = torch.zeros((cp, 576), device=dev, dtype=torch.float32)
            wp[:co] = w
            bp = torch.zeros(cp, device=dev, dtype=torch.float32)
            bp[:co] = conv.bias.detach().float() * sc + sh
            pk["mids"].append((wp.to(torch.bfloat16).contiguous(), bp.contiguous(), cp))
        conv, bn = last.layers[0], last.layers[1]
        sc, sh = cnn.bn_affine(bn)
        ci = conv.in_channels
        wt = torch.zeros((9, 32), device=dev, dtype=torch.float32)
        wt[:, :ci] = conv.weight.detach().float()[0].permute(1, 2, 0).reshape(9, ci)
        pk["last"] = (wt.contiguous(), float(conv.bias.detach()[0]), float(sc[0]), float(sh[0]))
        self._packed = pk
        return pk

    def _forward_tc(self, x, mods):
        lib = _lib.load()
        B, _, H, W = x.shape
        dev = x.device
        pk = self._pack(mods, dev)
        st = _lib.stream_ptr()
        a = torch.empty((B, H, W, 64), device=dev, dtype=torch.bfloat16)
        b = torch.empty_like(a)
        w0, b0, sc0, sh0 = pk["first"]
        _lib.check(lib.wmk_conv3x3_c1_nhwc_bf16(_lib.ptr(x), _lib.ptr(a), _lib.ptr(w0), _lib.ptr(b0), _lib.ptr(sc0), _lib.ptr(sh0),
                                                B, H, W, st))
        cp = 64
        for wp, bp, cp in pk["mids"]:
            out = b if cp == 64 else torch.empty((B, H, W, cp), device=dev, dtype=torch.bfloat16)
            _lib.check(lib.wmk_conv3x3_nhwc_bf16_tc(_lib.ptr(a), _lib.ptr(out), _lib.ptr(wp), _lib.ptr(bp), B, H, W, cp, st))
            a, b = out, a
        pooled = torch.empty((B, H // 2, W // 2, cp), device=dev, dtype=torch.bfloat16)
        _lib.check(lib.wmk_maxpool2x2_nhwc_bf16(_lib.ptr(a), _lib.ptr(pooled), B, H, W, cp, st))
        wt, bias, sc, sh = pk["last"]
        y = torch.empty((B, 1, H // 2, W // 2), device=dev, dtype=torch.float32)
        _lib.check(lib.wmk_conv3x3_nhwc_to1_f32(_lib.ptr(pooled), _lib.ptr(y), _lib.ptr(wt), bias, sc, sh, B, H // 2, W // 2, cp, st))
        return cnn.maxpool2x2(y)

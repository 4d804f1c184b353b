"""`JpegCompression` (`hidden/noise_layers/jpeg_compression.py:65-160`): RGB->YUV, 8x8 DCT, keep the
first 25 / 9 / 9 zig-zag coefficients, inverse DCT, YUV->RGB - one fused CUDA kernel
(`wmk_noise_jpeg_f32`).  Like the reference it is hard-wired to 3-channel input (`:53-55`)."""
import torch
import torch.nn as nn

from ... import _lib
from .crop import _prep


class JpegCompression(nn.Module):
    def __init__(self, device=None, yuv_keep_weights=(25, 9, 9)):
        super().__init__()
        self.device = device
        self.yuv_keep_weighs = tuple(int(k) for k in yuv_keep_weights)       # (sic) the reference's attribute name

    def forward(self, noised_and_cover):
        x = _prep(noised_and_cover[0])
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("JpegCompression needs a (B,3,H,W) tensor (the reference indexes channels 0..2, "
                             "jpeg_compression.py:53-55); got %s" % (tuple(x.shape),))
        out = torch.empty_like(x)
        ky, ku, kv = self.yuv_keep_weighs
        _lib.check(_lib.load().wmk_noise_jpeg_f32(_lib.ptr(x), _lib.ptr(out), x.shape[0], x.shape[2], x.shape[3], ky, ku, kv,
                                                  _lib.stream_ptr()))
        noised_and_cover[0] = out
        return noised_and_cover

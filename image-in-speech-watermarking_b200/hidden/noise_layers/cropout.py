"""`Cropout` (`hidden/noise_layers/cropout.py:6-28`): noised inside a random rectangle, cover outside."""
import torch
import torch.nn as nn

from ... import _lib
from .crop import get_random_rectangle_inside, _prep


class Cropout(nn.Module):
    def __init__(self, height_ratio_range, width_ratio_range):
        super().__init__()
        self.height_ratio_range = height_ratio_range
        self.width_ratio_range = width_ratio_range

    def forward(self, noised_and_cover):
        x, c = _prep(noised_and_cover[0]), _prep(noised_and_cover[1])
        assert x.shape == c.shape
        h0, h1, w0, w1 = get_random_rectangle_inside(image=x, height_ratio_range=self.height_ratio_range,
                                                     width_ratio_range=self.width_ratio_range)
        B, C, H, W = x.shape
        out = torch.empty_like(x)
        _lib.check(_lib.load().wmk_noise_mix_f32(_lib.ptr(x), _lib.ptr(c), _lib.ptr(out), B * C, H, W, int(h0), int(h1),
                                                 int(w0), int(w1), None, _lib.stream_ptr()))
        noised_and_cover[0] = out
        return noised_and_cover

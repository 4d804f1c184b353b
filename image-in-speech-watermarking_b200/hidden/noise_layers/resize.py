"""`Resize` (`hidden/noise_layers/resize.py:6-26`): nearest-neighbour rescale by a random ratio."""
import math

import torch
import torch.nn as nn

from ... import _lib
from .crop import random_float, _prep


class Resize(nn.Module):
    def __init__(self, resize_ratio_range, interpolation_method='nearest'):
        super().__init__()
        if interpolation_method != 'nearest':
            raise NotImplementedError("Resize CUDA kernel: 'nearest' only (the reference's default)")
        self.resize_ratio_min = resize_ratio_range[0]
        self.resize_ratio_max = resize_ratio_range[1]
        self.interpolation_method = interpolation_method

    def forward(self, noised_and_cover):
        resize_ratio = random_float(self.resize_ratio_min, self.resize_ratio_max)
        x = _prep(noised_and_cover[0])
        B, C, H, W = x.shape
        Ho, Wo = int(math.floor(H * resize_ratio)), int(math.floor(W * resize_ratio))
        out = torch.empty((B, C, Ho, Wo), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().wmk_noise_resize_nearest_f32(_lib.ptr(x), _lib.ptr(out), B * C, H, W, Ho, Wo,
                                                            float(resize_ratio), _lib.stream_ptr()))
        noised_and_cover[0] = out
        return noised_and_cover

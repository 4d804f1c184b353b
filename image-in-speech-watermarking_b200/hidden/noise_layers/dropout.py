"""`Dropout` (`hidden/noise_layers/dropout.py:5-28`): random pixels of the noised image replaced by
the cover's; the (H, W) keep-mask is drawn with numpy's global RNG as in the reference."""
import numpy as np
import torch
import torch.nn as nn

from ... import _lib
from .crop import _prep


class Dropout(nn.Module):
    def __init__(self, keep_ratio_range):
        super().__init__()
        self.keep_min = keep_ratio_range[0]
        self.keep_max = keep_ratio_range[1]

    def forward(self, noised_and_cover):
        x, c = _prep(noised_and_cover[0]), _prep(noised_and_cover[1])
        mask_percent = np.random.uniform(self.keep_min, self.keep_max)
        mask = np.random.choice([0.0, 1.0], x.shape[2:], p=[1 - mask_percent, mask_percent])
        m = torch.tensor(mask, device=x.device, dtype=torch.float).contiguous()
        B, C, H, W = x.shape
        out = torch.empty_like(x)
        _lib.check(_lib.load().wmk_noise_mix_f32(_lib.ptr(x), _lib.ptr(c), _lib.ptr(out), B * C, H, W, 0, 0, 0, 0,
                                                 _lib.ptr(m), _lib.stream_ptr()))
        return [out, noised_and_cover[1]]

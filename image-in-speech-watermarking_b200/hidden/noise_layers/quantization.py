"""`Quantization` (`hidden/noise_layers/quantization.py:6-45`): min-max to [0,255], 10-term Fourier
soft rounding, min-max back to the input range."""
import torch
import torch.nn as nn

from ... import _lib
from .crop import _prep


class Quantization(nn.Module):
    def __init__(self, device=None):
        super().__init__()
        self.min_value, self.max_value, self.N = 0.0, 255.0, 10

    def forward(self, noised_and_cover):
        x = _prep(noised_and_cover[0])
        out = torch.empty_like(x)
        _lib.check(_lib.load().wmk_noise_quantize_f32(_lib.ptr(x), _lib.ptr(out), x.numel(), _lib.stream_ptr()))
        return [out, noised_and_cover[1]]

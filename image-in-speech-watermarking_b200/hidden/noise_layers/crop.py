"""`Crop` (`hidden/noise_layers/crop.py:15-75`).  The rectangle is drawn on the host with numpy's
global RNG exactly as the reference does (including its use of `width_ratio_range[0]` twice,
crop.py:32); the copy runs on the GPU."""
import numpy as np
import torch
import torch.nn as nn

from ... import _lib


def random_float(min, max):
    return np.random.rand() * (max - min) + min


def get_random_rectangle_inside(image, height_ratio_range, width_ratio_range):
    image_height, image_width = image.shape[2], image.shape[3]
    remaining_height = int(np.rint(random_float(height_ratio_range[0], height_ratio_range[1]) * image_height))
    remaining_width = int(np.rint(random_float(width_ratio_range[0], width_ratio_range[0]) * image_width))
    height_start = 0 if remaining_height == image_height else np.random.randint(0, image_height - remaining_height)
    width_start = 0 if remaining_width == image_width else np.random.randint(0, image_width - remaining_width)
    return height_start, height_start + remaining_height, width_start, width_start + remaining_width


def _prep(t):
    if not t.is_cuda:
        raise _lib.WmkError("noise layers have no CPU implementation: inputs must be CUDA tensors")
    return t.detach().contiguous().float()


class Crop(nn.Module):
    def __init__(self, height_ratio_range, width_ratio_range):
        super().__init__()
        self.height_ratio_range = height_ratio_range
        self.width_ratio_range = width_ratio_range

    def forward(self, noised_and_cover):
        x = _prep(noised_and_cover[0])
        h0, h1, w0, w1 = get_random_rectangle_inside(x, self.height_ratio_range, self.width_ratio_range)
        B, C, H, W = x.shape
        out = torch.empty((B, C, h1 - h0, w1 - w0), device=x.device, dtype=torch.float32)
        _lib.check(_lib.load().wmk_noise_crop_f32(_lib.ptr(x), _lib.ptr(out), B * C, H, W, int(h0), int(h1), int(w0),
                                                  int(w1), _lib.stream_ptr()))
        noised_and_cover[0] = out
        return noised_and_cover

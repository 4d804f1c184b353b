"""CPU restatement of the HiDDeN noise layers (`hidden/noise_layers/*.py`).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Random draws use numpy's global RNG in the same
order as the reference, so a common seed gives the same rectangle / mask / ratio."""
import numpy as np
import torch
import torch.nn.functional as F


def random_float(lo, hi):                                  # `crop.py:5-12`
    return np.random.rand() * (hi - lo) + lo


def random_rectangle(shape, height_ratio_range, width_ratio_range):   # `crop.py:15-46` (incl. the [0],[0] quirk)
    H, W = shape[2], shape[3]
    rh = int(np.rint(random_float(height_ratio_range[0], height_ratio_range[1]) * H))
    rw = int(np.rint(random_float(width_ratio_range[0], width_ratio_range[0]) * W))
    h0 = 0 if rh == H else np.random.randint(0, H - rh)
    w0 = 0 if rw == W else np.random.randint(0, W - rw)
    return h0, h0 + rh, w0, w0 + rw


def crop(noised, hr, wr):                                  # `crop.py:63-75`
    h0, h1, w0, w1 = random_rectangle(noised.shape, hr, wr)
    return noised[:, :, h0:h1, w0:w1].clone()


def cropout(noised, cover, hr, wr):                        # `cropout.py:16-28`
    h0, h1, w0, w1 = random_rectangle(noised.shape, hr, wr)
    m = torch.zeros_like(noised)
    m[:, :, h0:h1, w0:w1] = 1
    return noised * m + cover * (1 - m)


def dropout(noised, cover, keep_range):                    # `dropout.py:15-28`
    p = np.random.uniform(keep_range[0], keep_range[1])
    mask = np.random.choice([0.0, 1.0], noised.shape[2:], p=[1 - p, p])
    m = torch.tensor(mask, dtype=torch.float).expand_as(noised)
    return noised * m + cover * (1 - m)


def resize(noised, ratio_range):                           # `resize.py:17-26`
    r = random_float(ratio_range[0], ratio_range[1])
    return F.interpolate(noised, scale_factor=(r, r), mode="nearest")


def quantization(noised):                                  # `quantization.py:6-45`
    def transform(t, rng):
        lo, hi = t.min(), t.max()
        return (t - lo) / (hi - lo) * (rng[1] - rng[0]) + rng[0]
    w = torch.tensor([((-1) ** (n + 1)) / (np.pi * (n + 1)) for n in range(10)])
    s = torch.tensor([2 * np.pi * (n + 1) for n in range(10)])
    for _ in range(4):
        w.unsqueeze_(-1)
        s.unsqueeze_(-1)
    x = transform(noised, (0, 255)).clamp(0.0, 255.0)
    x = x + torch.sum(w * torch.sin(x * s), dim=0)
    return transform(x, (noised.min(), noised.max()))

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


def jpeg_compression(noised, yuv_keep=(25, 9, 9)):          # `jpeg_compression.py:6-160` (3-channel input only, `:53-55`)
    """float64 restatement: zero-pad to multiples of 8, RGB->YUV, per 8x8 block D = C P C^T with
    C[k][n] = cos(pi/8 (n+1/2) k), keep the first N zig-zag coefficients, P' = I^T D I with
    I[n][k] = ((n==0) * -1/2 + cos(pi/8 (k+1/2) n)) / 4, YUV->RGB, un-pad."""
    x = noised.double()
    B, C3, H, W = x.shape
    assert C3 == 3
    ph, pw = (8 - H % 8) % 8, (8 - W % 8) % 8
    x = F.pad(x, (0, pw, 0, ph))
    r, g, b = x[:, 0], x[:, 1], x[:, 2]
    yuv = torch.stack([0.299 * r + 0.587 * g + 0.114 * b, -0.14713 * r + -0.28886 * g + 0.436 * b,
                       0.615 * r + -0.51499 * g + -0.10001 * b], 1)
    n = torch.arange(8, dtype=torch.float64)
    Cm = torch.cos(np.pi / 8 * (n[None, :] + 0.5) * n[:, None])                       # [k][n]
    Im = ((n[:, None] == 0).double() * -0.5 + torch.cos(np.pi / 8 * (n[None, :] + 0.5) * n[:, None])) * 0.25   # [n][k]
    order = sorted(((i, j) for i in range(8) for j in range(8)), key=lambda p: (p[0] + p[1], -p[1] if (p[0] + p[1]) % 2 else p[1]))
    Hp, Wp = x.shape[2], x.shape[3]
    blocks = yuv.reshape(B, 3, Hp // 8, 8, Wp // 8, 8).permute(0, 1, 2, 4, 3, 5)      # [...][y][x]
    D = Cm @ blocks @ Cm.T
    mask = torch.zeros(3, 8, 8, dtype=torch.float64)
    for c, keep in enumerate(yuv_keep):
        for i, j in order[:keep]:
            mask[c, i, j] = 1
    D = D * mask[None, :, None, None]
    P = Im.T @ D @ Im
    yuv2 = P.permute(0, 1, 2, 4, 3, 5).reshape(B, 3, Hp, Wp)
    Y, U, V = yuv2[:, 0], yuv2[:, 1], yuv2[:, 2]
    rgb = torch.stack([Y + 1.13983 * V, Y + -0.39465 * U + -0.58060 * V, Y + 2.03211 * U], 1)
    return rgb[:, :, :H, :W].float()


def magphase_split(spec):
    """spec (n,2,F,T) re/im -> (mag, phase) each (n,1,F,T): np.hypot / np.arctan2 in float64 (new op,
    no reference call site: feeds STFT magnitudes to the 1-channel HiDDeN decoder)."""
    s = spec.double()
    return torch.hypot(s[:, 0:1], s[:, 1:2]).float(), torch.atan2(s[:, 1:2], s[:, 0:1]).float()


def magphase_merge(mag, phase):
    m, p = mag.double(), phase.double()
    return torch.cat([m * torch.cos(p), m * torch.sin(p)], 1).float()

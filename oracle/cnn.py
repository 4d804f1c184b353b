"""CPU restatement of the CNN rows of the hot path: `ModelA` (`uformerWM/model.py:3000-3066`) and the
reference's modified HiDDeN `Decoder` / `ConvBNRelu` (`hidden/model/decoder.py:6-40`,
`hidden/model/conv_bn_relu.py:3-18`) as plain torch modules evaluated on the CPU in eval mode.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned by tests/golden/cnn.npz, produced by the
unmodified reference classes (oracle/make_golden.py)."""
import torch
import torch.nn as nn


class ModelAOracle(nn.Module):
    """Same layer list as the reference (`model.py:3003-3041`), so state_dicts are interchangeable."""

    def __init__(self):
        super().__init__()
        self.embedder_encoder = nn.Sequential(
            nn.Conv2d(2, 16, 3, padding=1, stride=1), nn.BatchNorm2d(16), nn.LeakyReLU(0.2), nn.MaxPool2d(2, 2),
            nn.Conv2d(16, 32, 3, padding=1, stride=1), nn.BatchNorm2d(32), nn.LeakyReLU(0.2), nn.MaxPool2d(2, 2))
        self.embedder_decoder = nn.Sequential(
            nn.ConvTranspose2d(33, 16, 2, 2), nn.BatchNorm2d(16), nn.ReLU(), nn.Dropout(0.5),
            nn.ConvTranspose2d(16, 2, 2, 2), nn.BatchNorm2d(2), nn.Sigmoid())
        self.detector = nn.Sequential(
            nn.Conv2d(2, 16, 3, padding=1), nn.BatchNorm2d(16), nn.LeakyReLU(0.2), nn.MaxPool2d(2, 2),
            nn.Conv2d(16, 64, 3, padding=1), nn.BatchNorm2d(64), nn.LeakyReLU(0.2), nn.MaxPool2d(2, 2),
            nn.Conv2d(64, 1, 3, padding=1), nn.ReLU())

    def decode(self, x):                                  # `model.py:3044-3050`
        return self.detector(x)

    def encode(self, stft, watermark):                    # `model.py:3052-3059`
        x = self.embedder_encoder(stft)
        return self.embedder_decoder(torch.cat([x, watermark], 1))

    def forward(self, stft, watermark):                   # `model.py:3062-3066`
        e = self.encode(stft, watermark)
        return e, self.decode(e)


def modelA_train_step(model, x, wm, keep_mask, attack_noise=None):
    """One training forward/backward of `uformerWM/train_modelA.py:423-445` on the CPU with torch autograd:
    model in train() mode (BatchNorm batch statistics, running-stat update), Dropout(0.5) with the GIVEN
    keep mask (so that the CUDA path can replay it), optional additive `attack_noise` on the encoded
    spectrogram before the detector (BASELINE config 5), loss = MSE(target, encoded) + MSE(extracted, wm).
    Returns dict(encoded, extracted, loss1, loss2, grads{name: tensor})."""
    model.train()
    for p in model.parameters():
        p.grad = None
    h = model.embedder_encoder(x)
    h = torch.cat([h, wm], 1)
    dec = model.embedder_decoder
    h = dec[2](dec[1](dec[0](h)))                      # ConvT + BN + ReLU
    h = h * keep_mask * 2.0                            # nn.Dropout(0.5) with an explicit mask
    enc = dec[6](dec[5](dec[4](h)))                    # ConvT + BN + Sigmoid
    ext = model.detector(enc if attack_noise is None else enc + attack_noise)
    loss1 = torch.nn.functional.mse_loss(x, enc)
    loss2 = torch.nn.functional.mse_loss(ext, wm)
    (loss1 + loss2).backward()
    return {"encoded": enc.detach(), "extracted": ext.detach(), "loss1": float(loss1), "loss2": float(loss2),
            "grads": {n: p.grad.detach().clone() for n, p in model.named_parameters()}}


def _cbr(ci, co):
    return nn.Sequential()  # placeholder replaced below (keeps the 'layers.N.layers.M' key structure)


class _ConvBNRelu(nn.Module):
    def __init__(self, ci, co):
        super().__init__()
        self.layers = nn.Sequential(nn.Conv2d(ci, co, 3, 1, padding=1), nn.BatchNorm2d(co), nn.ReLU(inplace=True))

    def forward(self, x):
        return self.layers(x)


class HiddenDecoderOracle(nn.Module):
    """`hidden/model/decoder.py:12-40` with decoder_blocks / decoder_channels / message_length."""

    def __init__(self, decoder_blocks=7, decoder_channels=64, message_length=30):
        super().__init__()
        layers = [_ConvBNRelu(1, decoder_channels)]
        for _ in range(decoder_blocks - 1):
            layers.append(_ConvBNRelu(decoder_channels, decoder_channels))
        layers.append(_ConvBNRelu(decoder_channels, message_length))
        layers.append(nn.MaxPool2d(2, 2, 0))
        layers.append(_ConvBNRelu(message_length, 1))
        layers.append(nn.MaxPool2d(2, 2, 0))
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)


def randomize_(module, seed):
    """Deterministic non-trivial parameters AND BatchNorm running statistics (so the eval-mode
    affine is exercised)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in module.state_dict().items():
            if name.endswith("num_batches_tracked"):
                continue
            if name.endswith("running_var"):
                t.copy_(0.5 + torch.rand(t.shape, generator=g))
            elif name.endswith("running_mean"):
                t.copy_(0.2 * torch.randn(t.shape, generator=g))
            elif t.dim() == 1 and name.endswith("weight"):       # BN gamma
                t.copy_(1.0 + 0.2 * torch.randn(t.shape, generator=g))
            elif t.dim() == 1:
                t.copy_(0.1 * torch.randn(t.shape, generator=g))
            else:
                fan_in = t[0].numel() if t.dim() > 1 else t.numel()
                t.copy_(torch.randn(t.shape, generator=g) * (1.5 / fan_in ** 0.5))
    return module.eval()

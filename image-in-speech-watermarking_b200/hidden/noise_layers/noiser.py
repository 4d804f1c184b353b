"""`Noiser` (`hidden/noise_layers/noiser.py:8-31`): holds Identity plus the configured layers and sends each
batch through ONE of them, picked with one draw from numpy's global generator per call (the reference's
`np.random.choice(layers, 1)`, which consumes the same single bounded integer as drawing the index does).

The two string placeholders the `--noise` grammar emits (`hidden/noise_argparser.py`) are resolved here because
those layers need the device at construction time."""
import numpy as np
import torch.nn as nn

from .identity import Identity
from .jpeg_compression import JpegCompression
from .quantization import Quantization

# placeholder string -> constructor taking the device
_DEFERRED = {"JpegPlaceholder": JpegCompression, "QuantizationPlaceholder": Quantization}


def _resolve(entry, device):
    """A layer object passes through; a placeholder string becomes its layer; anything else is an error."""
    if not isinstance(entry, str):
        return entry
    make = _DEFERRED.get(entry)
    if make is None:
        expected = " or ".join('"%s"' % k for k in _DEFERRED)
        raise ValueError("Noiser: unknown layer placeholder %r (expected %s)" % (entry, expected))
    return make(device)


class Noiser(nn.Module):
    def __init__(self, noise_layers: list, device):
        super().__init__()
        # a plain list, as in the reference: the layers carry no parameters and `noise_layers` is part of the interface
        self.noise_layers = [Identity()] + [_resolve(entry, device) for entry in noise_layers]

    def forward(self, encoded_and_cover):
        pick = int(np.random.choice(len(self.noise_layers), 1)[0])
        return self.noise_layers[pick](encoded_and_cover)

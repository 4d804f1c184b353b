"""The reference's modified HiDDeN `Decoder` (`hidden/model/decoder.py:6-40`): 1 input channel,
decoder_blocks x ConvBNRelu(64) -> ConvBNRelu(message_length) -> MaxPool2 -> ConvBNRelu(1) ->
MaxPool2, i.e. a (B,1,H/4,W/4) image."""
import torch.nn as nn

from ... import cnn
from ..options import HiDDenConfiguration
from .conv_bn_relu import ConvBNRelu


class Decoder(nn.Module):
    def __init__(self, config: HiDDenConfiguration):
        super().__init__()
        self.channels = config.decoder_channels
        layers = [ConvBNRelu(1, self.channels)]
        for _ in range(config.decoder_blocks - 1):
            layers.append(ConvBNRelu(self.channels, self.channels))
        layers.append(ConvBNRelu(self.channels, config.message_length))
        layers.append(nn.MaxPool2d(kernel_size=2, stride=2, padding=0))
        layers.append(ConvBNRelu(config.message_length, 1))
        layers.append(nn.MaxPool2d(kernel_size=2, stride=2, padding=0))
        self.layers = nn.Sequential(*layers)
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, image_with_wm):
        return cnn.run_sequential(self.layers, image_with_wm)

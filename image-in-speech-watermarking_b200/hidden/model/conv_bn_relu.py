"""`ConvBNRelu` (`hidden/model/conv_bn_relu.py:3-18`) on libwmk's fused conv + BN(eval) + ReLU kernel."""
import torch.nn as nn

from ... import cnn


class ConvBNRelu(nn.Module):
    def __init__(self, channels_in, channels_out, stride=1):
        super().__init__()
        if stride != 1:
            raise NotImplementedError("ConvBNRelu CUDA kernel: stride 1 only (the reference never uses another)")
        self.layers = nn.Sequential(nn.Conv2d(channels_in, channels_out, 3, stride, padding=1),
                                    nn.BatchNorm2d(channels_out), nn.ReLU(inplace=True))

    def forward(self, x):
        return cnn.run_sequential(self.layers, x)

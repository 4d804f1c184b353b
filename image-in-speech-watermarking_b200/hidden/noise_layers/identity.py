"""`Identity` (`hidden/noise_layers/identity.py:4-12`)."""
import torch.nn as nn


class Identity(nn.Module):
    def forward(self, noised_and_cover):
        return noised_and_cover

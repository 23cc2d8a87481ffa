"""Position embeddings read by the blind cross-attention (reference: model/crossatt.py:21-48)."""
import math

import torch
from torch import nn


class ConvPos(nn.Module):
    """model/crossatt.py:21-33."""

    def __init__(self, dim, max_seq_len=2000, kernel_size=31):
        super().__init__()
        self.embed = nn.Embedding(max_seq_len, dim)
        self.dw_conv = nn.Conv1d(dim, dim, kernel_size, groups=dim, padding="same")

    def forward(self, x):
        return self.dw_conv(self.embed(x).transpose(1, 2)).transpose(1, 2)


class SinPos(nn.Module):
    """model/crossatt.py:36-48."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        e = 2 * torch.arange(self.dim // 2, device=x.device) / self.dim
        pos = x.unsqueeze(-1) * torch.pow(10000, -e).view(1, 1, -1)
        return torch.sin(torch.cat((pos, pos + math.pi / 2), dim=2))



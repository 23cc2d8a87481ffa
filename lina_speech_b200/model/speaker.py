"""Speaker-vector front end (the reference's ``SimpleSpeakerEncoder``, model/encoder.py:45-84): optional ``spk_encoder`` of
LinaModel, off the hot path.  A window of codec-token embeddings goes through a small bidirectional encoder; position 0 of
its output, projected back to the model width, replaces the first decoder input (model/modeling_lina.py:79-81)."""
import random

import torch
from torch import nn

from .base_blocks import MixingBlock, SelfAttention, SwiGLU


def _pick_window(n_frames: int, length: int, training: bool, skip: int) -> slice:
    """training: a random window that starts after ``skip`` frames; eval: the first ``length`` frames."""
    start = random.randint(skip, n_frames - length) if training else 0
    return slice(start, start + length)


class SimpleSpeakerEncoder(nn.Module):
    def __init__(self, dim: int, dim_inner: int, heads: int, n_layers=6, dropout=0.1, rotary=True,
                 window_length: int = 256, rank: int = 1):
        super().__init__()
        make_block = lambda: MixingBlock(lambda: SelfAttention(dim_inner, heads, rotary=rotary), lambda: SwiGLU(dim_inner),
                                         lambda: nn.LayerNorm(dim_inner), dropout)
        self.sa = nn.ModuleList(make_block() for _ in range(n_layers))       # parameter names as in the reference: sa.N.*
        self.window_length = window_length
        self.in_proj, self.out_proj = nn.Linear(dim, dim_inner), nn.Linear(dim_inner, dim)

    def forward(self, x, avoid_n_first_frames: int = 150, **kwargs):
        h = self.in_proj(x[:, _pick_window(x.shape[1], self.window_length, self.training, avoid_n_first_frames)])
        for block in self.sa:
            h = block(h)
        return self.out_proj(h[:, 0])

"""Codec-token embedding table with one sub-table per quantizer level.

Same parameter as the reference's ``MultiEmbedding`` (model/multiembed.py:7-23): a single ``weight`` of shape
[n_level, n_emb, d_emb] initialised N(0,1), so ``rvq_embed.weight`` loads from its checkpoints.  Level i of the index
tensor looks up sub-table i; the padding row receives no gradient (as with ``F.embedding(padding_idx=...)``).
"""
import torch
import torch.nn.functional as F
from torch import nn


class MultiEmbedding(nn.Module):
    def __init__(self, n_level: int, n_emb: int, d_emb: int, padding_idx=None):
        super().__init__()
        self.n_level, self.padding_idx = n_level, padding_idx
        table = torch.empty(n_level, n_emb, d_emb)
        nn.init.normal_(table)
        self.weight = nn.Parameter(table)

    def forward(self, idx: torch.Tensor) -> torch.Tensor:
        """idx [n_level, ...] (int64) -> [n_level, ..., d_emb]."""
        if idx.shape[0] != self.n_level:
            raise ValueError(f"expected {self.n_level} quantizer levels, got {idx.shape[0]}")
        levels = [F.embedding(idx[lv], self.weight[lv], padding_idx=self.padding_idx) for lv in range(self.n_level)]
        return torch.stack(levels, dim=0)

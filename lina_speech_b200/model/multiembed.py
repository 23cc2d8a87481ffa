"""Stacked per-quantizer embedding (model/multiembed.py:7-23): weight [n_level, n_emb, d_emb]."""
import torch
from torch import nn


class MultiEmbedding(nn.Module):
    def __init__(self, n_level, n_emb, d_emb, padding_idx=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n_level, n_emb, d_emb))
        self.n_level = n_level
        self.padding_idx = padding_idx
        nn.init.normal_(self.weight)

    def forward(self, idx):
        """idx [q, ...] -> [q, ..., d]: level i looks up weight[i] (padding row gets no gradient)."""
        return torch.stack([nn.functional.embedding(idx[i], self.weight[i], padding_idx=self.padding_idx)
                            for i in range(self.n_level)], dim=0)

"""Abstract forward / init_state / step contract of the backbone (model/attentive_rnn.py:6-17)."""
from abc import abstractmethod

import torch


class AttentiveRNN(torch.nn.Module):
    @abstractmethod
    def forward(self, x, ctx, x_mask, ctx_mask):
        ...

    @abstractmethod
    def init_state(self):
        ...

    @abstractmethod
    def step(self, x, ctx, crossatt_mask):
        ...

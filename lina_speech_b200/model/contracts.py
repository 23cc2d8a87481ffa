"""Interface every Lina backbone implements (the reference's ``AttentiveRNN`` ABC, model/attentive_rnn.py:6-17).

Three entry points, all driven by LinaModel:
  forward(x, ctx, ...)            teacher-forced pass over a whole sequence  -> (hidden, attention)
  init_state(max_seqlen, batch)   fresh per-layer recurrent state container  -> Cache
  step(x_t, ctx, t, state)        one autoregressive token                   -> (hidden, attention, state)
"""
import torch


class AttentiveRNN(torch.nn.Module):
    def forward(self, x, ctx, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} must implement forward()")

    def init_state(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} must implement init_state()")

    def step(self, x, ctx, *args, **kwargs):
        raise NotImplementedError(f"{type(self).__name__} must implement step()")

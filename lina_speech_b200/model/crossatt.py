"""Text<->audio alignment modules (model/crossatt.py), same parameter names.

BlindCrossAttention (model/crossatt.py:76-155): softmax(q k^T) reads positional embeddings, a GLA
MixingBlock (``pos_net``) carries the position estimate, softmax(x pos^T) reads the text values.
B200 note: in eval mode the text-side projections ln_k(k(ctx)), ln_v(v(ctx)) and the positional
embedding are memoised per ``ctx`` tensor -- the reference recomputes them for the whole text at
every decode step (crossatt.py:114-116,127), which costs more than the 13 GLA blocks of a step.
"""
import math

import torch
from torch import nn
from einops import rearrange

from .positional import ConvPos, SinPos


def exists(x):
    return x is not None


def tensor_version(t) -> int:
    """Version counter for memo keys; inference tensors do not track one (they are immutable enough)."""
    try:
        return t._version
    except RuntimeError:
        return -1


def scaled_dot_product_attention(query, key, value, mask=None):
    """model/crossatt.py:13-19 (eval branch: returns the attention weights too)."""
    w = query @ key.transpose(-2, -1) * (1 / math.sqrt(query.size(-1)))
    if exists(mask):
        w = w.masked_fill(~mask, -torch.finfo(w.dtype).max)
    w = torch.softmax(w, dim=-1)
    return w @ value, w


class BlindCrossAttention(nn.Module):
    def __init__(self, q_dim, k_dim, att_dim, heads, pos_net, dropout=0.1, pos_dim=64, rotary=False,
                 pos_type="sinusoidal"):
        super().__init__()
        self.q = nn.Linear(q_dim, att_dim)
        self.k = nn.Linear(k_dim, att_dim)
        self.v = nn.Linear(k_dim, att_dim)
        self.pos_net = pos_net
        if pos_type == "sinusoidal":
            self.pos_embed = SinPos(pos_dim)
        elif pos_type == "convolutional":
            self.pos_embed = ConvPos(pos_dim)
        else:
            raise ValueError(f"unknown pos_type {pos_type}")
        assert att_dim % heads == 0
        self.ln_q = nn.LayerNorm(att_dim)
        self.ln_k = nn.LayerNorm(att_dim)
        self.ln_v = nn.LayerNorm(att_dim)
        if rotary:
            raise NotImplementedError("BlindCrossAttention(rotary=True) is not used by the shipped model")
        self.rotary = None
        self.dropout_att = nn.Dropout(dropout)

    _memo = None          # class-level default: instances un-pickled from a reference checkpoint never ran __init__

    def clear_memo(self):
        self._memo = None

    def _memo_key(self, ctx):
        ps = [p for m in (self.k, self.v, self.ln_k, self.ln_v, self.pos_embed) for p in m.parameters()]
        return (ctx.data_ptr(), tensor_version(ctx), tuple(ctx.shape), tuple(ctx.stride()), ctx.dtype, ctx.device,
                tuple((p.data_ptr(), tensor_version(p), p.dtype) for p in ps))

    def _text_side(self, ctx, pos):
        """ln_k(k(ctx)), ln_v(v(ctx)), pos_emb -- memoised across decode steps in eval mode.

        The memo holds a reference to ``ctx`` itself, so its storage cannot be freed and handed to the next utterance's
        text tensor while the entry is alive (data_ptr alone is not an identity); callers that start a new utterance
        (LinaModel.forward / generate_batch) also clear it."""
        key = None
        if not self.training and pos is None and not torch.is_grad_enabled():
            key = self._memo_key(ctx)
            m = self._memo
            if m is not None and (m[0] is ctx or m[0].data_ptr() == ctx.data_ptr()) and m[1] == key:
                return m[2]
        v = self.ln_v(self.v(ctx)).unsqueeze(1)
        k = self.ln_k(self.k(ctx)).unsqueeze(1)
        if pos is None:
            pos = torch.arange(k.shape[2], device=k.device).unsqueeze(0)
        pos_emb = self.pos_embed(pos).unsqueeze(1)
        out = (k, v, pos_emb)
        if key is not None:
            self._memo = (ctx, key, out)
        return out

    def forward(self, q, k, mask=None, time_step=None, pos=None, **kwargs):
        q = self.ln_q(self.q(q)).unsqueeze(1)
        k, v, pos_emb = self._text_side(k, pos)
        if mask is not None:
            mask = mask.unsqueeze(1)
        if self.training:
            def sdpa(a, b, c):
                return nn.functional.scaled_dot_product_attention(
                    a, b, c.expand(a.shape[0], -1, -1, -1) if c.shape[0] != a.shape[0] else c,
                    attn_mask=mask, dropout_p=self.dropout_att.p), None
        else:
            def sdpa(a, b, c):
                return scaled_dot_product_attention(a, b, c, mask=mask)
        x, att1 = sdpa(q, k, pos_emb)
        x = self.pos_net(x.squeeze(1), **kwargs)
        x = x[0] if type(x) is tuple else x
        x, att2 = sdpa(x.unsqueeze(1), pos_emb.expand(x.shape[0], -1, -1, -1) if self.training else pos_emb, v)
        att = torch.cat((att1, att2), dim=1) if att1 is not None else None
        return x.squeeze(1), att


class CrossAttention(nn.Module):
    """Plain multi-head cross attention (model/crossatt.py:158-212), used when ``blind=False``."""

    def __init__(self, q_dim, k_dim, att_dim, heads, dropout=0.1, rotary=False):
        super().__init__()
        self.q = nn.Linear(q_dim, att_dim)
        self.k = nn.Linear(k_dim, att_dim)
        self.v = nn.Linear(k_dim, att_dim)
        assert att_dim % heads == 0
        self.heads = heads
        self.ln_q = nn.LayerNorm(att_dim)
        self.ln_k = nn.LayerNorm(att_dim)
        self.ln_v = nn.LayerNorm(att_dim)
        if rotary:
            raise NotImplementedError("CrossAttention(rotary=True) is not used by the shipped model")
        self.rotary = None
        self.dropout_att = dropout

    def forward(self, q, k, v=None, mask=None, time_step=None, **kwargs):
        if v is None:
            v = k
        q = self.ln_q(self.q(q))
        v = self.ln_v(self.v(v))
        k = self.ln_k(self.k(k))
        q, k, v = (rearrange(t, "b n (h d) -> b h n d", h=self.heads) for t in (q, k, v))
        if self.training:
            x = nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=self.dropout_att)
            att = None
        else:
            x, att = scaled_dot_product_attention(q, k, v, mask=mask)
        return rearrange(x, "b h n d -> b n (h d)"), att

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


import os

# LINA_FUSED_CROSSATT_STEP=1: the single-token cross attention runs as two launches of lina_cross_att_step instead of the torch
# sequence (LayerNorm, q k^T, scale, softmax, w v: ~10 launches).  Measured on B200 inside the CUDA-graphed decode step
# (profiles/ab_decode_r02.json): 0.93 ms per step against 0.89 ms op by op at batch 32 -- one CTA per sequence streams its
# 2 x 256 KB of keys / values at 17 us per launch, the batched library matmuls spread them over all SMs.  Default off.
FUSED_STEP = os.environ.get("LINA_FUSED_CROSSATT_STEP", "0") == "1"


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
        if key is not None:                       # memoised for the decode loop: the fused step kernel reads plain [n, d] rows
            k, v, pos_emb = k.contiguous(), v.contiguous(), pos_emb.contiguous()
        out = (k, v, pos_emb)
        if key is not None:
            self._memo = (ctx, key, out)
        return out

    def _can_step_fused(self, q, mask, pos) -> bool:
        return (FUSED_STEP and q.dim() == 3 and q.shape[1] == 1 and q.is_cuda and mask is None and pos is None
                and not self.training and not torch.is_grad_enabled() and not torch.is_autocast_enabled()
                and q.dtype in (torch.float32, torch.bfloat16, torch.float16) and self.q.weight.dtype == q.dtype
                and q.shape[-1] % (16 // q.element_size()) == 0)

    def _step_fused(self, q, ctx, **kwargs):
        """One token: q projection (GEMM), then ONE launch for LayerNorm + softmax(q k^T) + the read of the positional table,
        the pos_net block, and ONE launch for softmax(x pos^T) + the read of the text values (lina_cross_att_step); the two
        attention rows land in their slots of one [B, 2, 1, n] tensor."""
        from .. import _lib as L
        k, v, pos_emb = self._text_side(ctx, None)                      # [B,1,n,d], [B,1,n,d], [1,1,n,d]
        B, d = q.shape[0], k.shape[-1]
        n = k.shape[2]
        if not (k.is_contiguous() and v.is_contiguous() and pos_emb.is_contiguous() and k.dtype == q.dtype == v.dtype
                and pos_emb.dtype == q.dtype and k.shape[0] == B):
            return None
        qlin = self.q(q).view(B, d)
        att = torch.empty(B, 2, 1, n, dtype=q.dtype, device=q.device)
        x = torch.empty(B, d, dtype=q.dtype, device=q.device)
        scale = 1.0 / math.sqrt(d)
        lib = L.lib()
        rc = lib.lina_cross_att_step(L.ptr(qlin), qlin.stride(0), L.ptr(self.ln_q.weight), L.ptr(self.ln_q.bias), float(self.ln_q.eps),
                                     L.ptr(k), n * d, L.ptr(pos_emb), 0, L.ptr(att), 2 * n, L.ptr(x), d, B, n, d, scale, L.dt(q),
                                     L.stream(q))
        L.count_launches(1)
        L.check(rc, "lina_cross_att_step")
        y = self.pos_net(x.view(B, 1, d), **kwargs)
        y = (y[0] if type(y) is tuple else y).reshape(B, d)
        out = torch.empty(B, d, dtype=q.dtype, device=q.device)
        att2 = att[:, 1]
        rc = lib.lina_cross_att_step(L.ptr(y), y.stride(0), None, None, 0.0, L.ptr(pos_emb), 0, L.ptr(v), n * d, L.ptr(att2), 2 * n,
                                     L.ptr(out), d, B, n, d, scale, L.dt(q), L.stream(q))
        L.count_launches(1)
        L.check(rc, "lina_cross_att_step")
        return out.view(B, 1, d), att

    def forward(self, q, k, mask=None, time_step=None, pos=None, **kwargs):
        if self._can_step_fused(q, mask, pos):
            r = self._step_fused(q, k, **kwargs)
            if r is not None:
                return r
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

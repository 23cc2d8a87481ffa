"""CPU oracle for the Lina-Speech host path -- TEST INFRASTRUCTURE ONLY.

Functional (state-dict driven) restatement of the reference's model code for
the GLA backbone, in plain CPU torch with the GLA op of
:mod:`oracle.gla_oracle`.  Same import rule as gla_oracle.py: tests, smoke and
bench's cpu_baseline / ``--impl reference`` legs only.

A model is a ``dict[str, Tensor]`` with exactly the reference's state-dict key
names (``attentive_rnn.encoder.0.tmix.q_proj.weight`` ...), plus a small config
dict ``{"d_model", "n_layer", "heads", "n_quant", ...}``.  Pinned against the
reference's own classes by tests/golden/make_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import gla_oracle as G

SD = Dict[str, torch.Tensor]


def _lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd: SD, p: str, x, eps=1e-5):
    return F.layer_norm(x, x.shape[-1:], sd[p + ".weight"], sd[p + ".bias"], eps)


# ----------------------------------------------------------------------------
# a4: GatedLinearAttention.forward  (model/gla.py:131-227)
# ----------------------------------------------------------------------------
def gla_layer(sd: SD, p: str, x, heads: int, state: Optional[Tuple[torch.Tensor, ...]] = None,
              training: bool = False, use_short_conv: bool = True, normalizer: float = 16.0,
              eps: float = 1e-5):
    """``state`` = (conv_q, conv_k, conv_v, S) as built by
    GatedLinearAttention.init_state (model/gla.py:229-240); updated in place
    unless ``training`` (model/gla.py:205-213, FLA/fla/models/utils.py:39-74)."""
    B, T, _ = x.shape
    q, k, v = _lin(sd, p + ".q_proj", x), _lin(sd, p + ".k_proj", x), _lin(sd, p + ".v_proj", x)
    if use_short_conv:
        outs = []
        for i, (name, t) in enumerate((("q", q), ("k", k), ("v", v))):
            w = sd[f"{p}.{name}_conv1d.weight"][:, 0]
            cache = state[i] if state is not None else None
            if cache is not None and T == 1:      # convolution.py:161-162
                outs.append(G.short_conv_step(t, cache, w))
            else:                                  # convolution.py:163-178
                outs.append(G.short_conv_prefill(t, w, cache))
        q, k, v = outs
    hs = lambda t: t.view(B, T, heads, -1).transpose(1, 2)
    gk = _lin(sd, p + ".gk_proj.1", _lin(sd, p + ".gk_proj.0", x))
    gk = G.gate_logsigmoid(gk, normalizer)
    S0 = state[-1] if state is not None else None
    o, ST = G.recurrent_gla(hs(q), hs(k), hs(v), hs(gk), initial_state=S0, output_final_state=state is not None)
    if state is not None and not training:
        state[-1].copy_(ST)
    o = o.transpose(1, 2)                                             # b l h d
    g = _lin(sd, p + ".g_proj", x).view(B, T, heads, -1)
    o = G.rmsnorm_swish_gate(o, g, sd[p + ".g_norm_swish_gate.weight"], eps)
    return _lin(sd, p + ".o_proj", o.reshape(B, T, -1))


def swiglu(sd: SD, p: str, x):
    """model/base_blocks.py:42-50."""
    gate, u = _lin(sd, p + ".p_in", x).chunk(2, dim=-1)
    return _lin(sd, p + ".p_out", F.silu(gate) * u)


def gla_block(sd: SD, p: str, x, heads: int, state=None, training=False):
    """MixingBlock (model/base_blocks.py:56-69), dropout = 0."""
    x = gla_layer(sd, p + ".tmix", _ln(sd, p + ".norm1", x), heads, state, training) + x
    return swiglu(sd, p + ".cmix", _ln(sd, p + ".norm2", x)) + x


# ----------------------------------------------------------------------------
# BlindCrossAttention  (model/crossatt.py:76-155, eval branch :13-19,140-141)
# ----------------------------------------------------------------------------
def _sdpa(q, k, v, mask=None):
    w = q @ k.transpose(-2, -1) / math.sqrt(q.size(-1))
    if mask is not None:
        w = w.masked_fill(~mask, -torch.finfo(w.dtype).max)
    w = torch.softmax(w, dim=-1)
    return w @ v, w


def conv_pos(sd: SD, p: str, pos):
    """ConvPos (model/crossatt.py:21-33)."""
    y = F.embedding(pos, sd[p + ".embed.weight"]).transpose(1, 2)
    w = sd[p + ".dw_conv.weight"]
    y = F.conv1d(y, w, sd[p + ".dw_conv.bias"], padding=w.shape[-1] // 2, groups=w.shape[0])
    return y.transpose(1, 2)


def sin_pos(dim: int, pos):
    """SinPos (model/crossatt.py:36-48)."""
    e = 2 * torch.arange(dim // 2) / dim
    p = pos.unsqueeze(-1) * torch.pow(10000, -e).view(1, 1, -1)
    return torch.sin(torch.cat((p, p + math.pi / 2), dim=2))


def blind_cross_att(sd: SD, p: str, cfg, q, ctx, mask=None, pos=None, pos_state=None, training=False):
    qh = _ln(sd, p + ".ln_q", _lin(sd, p + ".q", q)).unsqueeze(1)
    vh = _ln(sd, p + ".ln_v", _lin(sd, p + ".v", ctx)).unsqueeze(1)
    kh = _ln(sd, p + ".ln_k", _lin(sd, p + ".k", ctx)).unsqueeze(1)
    j = kh.shape[2]
    if mask is not None:
        mask = mask.unsqueeze(1)
    if pos is None:
        pos = torch.arange(j).unsqueeze(0)
    if cfg.get("pos_type", "convolutional") == "convolutional":
        pe = conv_pos(sd, p + ".pos_embed", pos)
    else:
        pe = sin_pos(cfg["d_model"], pos)
    x, att1 = _sdpa(qh, kh, pe.unsqueeze(1), mask)
    x = gla_block(sd, p + ".pos_net", x.squeeze(1), cfg["heads"], pos_state, training)
    x, att2 = _sdpa(x.unsqueeze(1), pe.unsqueeze(1), vh, mask)
    return x.squeeze(1), torch.cat((att1, att2), dim=1)


# ----------------------------------------------------------------------------
# a8: AttentiveGLA.forward / init_state / step  (model/gla.py:287-313,358-365)
# ----------------------------------------------------------------------------
def init_state(cfg, batch_size: int, dtype=torch.float32) -> List[Tuple[torch.Tensor, ...]]:
    d, H = cfg["d_model"], cfg["heads"]
    kd, vd = int(d * cfg.get("expand_k", 1.0)), int(d * cfg.get("expand_v", 2.0))
    n = 2 * cfg["n_layer"] + 1
    z = lambda *s: torch.zeros(*s, dtype=dtype)
    return [(z(batch_size, kd, 4), z(batch_size, kd, 4), z(batch_size, vd, 4),
             z(batch_size, H, kd // H, vd // H)) for _ in range(n)]


def attentive_gla(sd: SD, p: str, cfg, x, ctx, mask=None, state=None, crossatt_pos=None,
                  training=False, step=False):
    """``step=False``: AttentiveGLA.forward -- the cross-attention pos_net never
    sees the cache (model/gla.py:294).  ``step=True``:
    AttentiveGLA.step -- every block incl. pos_net (layer_idx 2N) is stateful."""
    N, H = cfg["n_layer"], cfg["heads"]
    for i in range(N):
        x = gla_block(sd, f"{p}.encoder.{i}", x, H, state[i] if state is not None else None, training)
    pos_state = state[2 * N] if (step and state is not None) else None
    v, att = blind_cross_att(sd, p + ".cross_att", cfg, x, ctx, mask, crossatt_pos, pos_state, training)
    x = x + v
    for i in range(N):
        x = gla_block(sd, f"{p}.decoder.{i}", x, H, state[N + i] if state is not None else None, training)
    return x, att


# ----------------------------------------------------------------------------
# TextEncoder, rotary=False  (model/encoder.py:14-43, model/base_blocks.py:9-40)
# ----------------------------------------------------------------------------
def text_encoder(sd: SD, p: str, cfg, x, mask=None):
    heads = cfg.get("txt_heads", cfg["heads"])
    if mask is not None:
        mask = mask.unsqueeze(1) | torch.eye(mask.shape[-1], dtype=torch.bool).view(1, 1, *mask.shape[-2:])
    i = 0
    while f"{p}.sa.{i}.tmix.qkv.weight" in sd:
        b = f"{p}.sa.{i}"
        h = _ln(sd, b + ".norm1", x)
        B, n, d = h.shape
        qq, kk, vv = (t.view(B, n, heads, -1).transpose(1, 2) for t in _lin(sd, b + ".tmix.qkv", h).chunk(3, dim=-1))
        y = F.scaled_dot_product_attention(qq, kk, vv, attn_mask=mask)
        x = y.transpose(1, 2).reshape(B, n, d) + x
        x = swiglu(sd, b + ".cmix", _ln(sd, b + ".norm2", x)) + x
        i += 1
    return x


# ----------------------------------------------------------------------------
# a9: LinaModel.forward / generate_batch  (model/modeling_lina.py:61-108,112-192)
# ----------------------------------------------------------------------------
def rvq_embed(sd: SD, ids):
    """MultiEmbedding (model/multiembed.py:7-23) summed over quantizers; ids [q,b,n]."""
    w = sd["rvq_embed.weight"]
    return sum(F.embedding(ids[i], w[i], padding_idx=0) for i in range(w.shape[0]))


def logits_head(sd: SD, y):
    """EinMix 'b n d -> b n q l' with weight [q,l,d] (model/modeling_lina.py:51-57)."""
    return torch.einsum("bnd,qld->bnql", y, sd["logits_head.weight"])


def lina_forward(sd: SD, cfg, x, y, encoder_mask, crossatt_mask, logits_mask=None, state=None,
                 crossatt_pos=None, training=False):
    x_enc = text_encoder(sd, "txt_encoder", cfg, F.embedding(x, sd["txt_embed.weight"], padding_idx=0), encoder_mask)
    y_embd = rvq_embed(sd, y.permute(2, 0, 1))
    y_hat, att = attentive_gla(sd, "attentive_rnn", cfg, y_embd[:, :-1], x_enc, crossatt_mask[:, :-1],
                               state, crossatt_pos, training)
    logits = logits_head(sd, y_hat)
    if logits_mask is not None:
        ml, mt = logits[logits_mask[:, 1:]], y[:, 1:][logits_mask[:, 1:]]
    else:
        ml, mt = logits, y[:, 1:]
    loss = F.cross_entropy(ml.reshape(-1, ml.shape[-1]), mt.reshape(-1), ignore_index=1)
    return logits, loss, att


def lina_generate_greedy(sd: SD, cfg, x, batch_size: int, prompt=None, max_seqlen: int = 32, state=None):
    """generate_batch with k=1 (greedy) and force_max_seqlen=True; returns
    (qs [Q,B,steps], atts [B,2,steps,n], logits of every step [steps,B,Q,L])."""
    Q = sd["rvq_embed.weight"].shape[0]
    x = x.unsqueeze(0).expand(batch_size, -1)
    x_enc = text_encoder(sd, "txt_encoder", cfg, F.embedding(x, sd["txt_embed.weight"], padding_idx=0))
    y_embd = rvq_embed(sd, torch.ones(Q, batch_size, 1, dtype=torch.long))
    p_len = -1
    if prompt is not None:
        if prompt.shape[1] != batch_size:       # modeling_lina.py:135-136 (+3 only when broadcasting)
            prompt = prompt.expand(-1, batch_size, -1) + 3
        prompt = rvq_embed(sd, prompt)
        p_len = prompt.shape[1]
    if state is None:
        state = init_state(cfg, batch_size)
    qs, atts, all_logits = [], [], []
    for t in range(max_seqlen):
        y, att = attentive_gla(sd, "attentive_rnn", cfg, y_embd, x_enc, state=state, step=True)
        logits = logits_head(sd, y)                       # b 1 q l
        all_logits.append(logits[:, 0])
        tok = logits[:, 0].argmax(-1).t().unsqueeze(-1)   # q b 1
        qs.append(tok)
        atts.append(att)
        y_embd = prompt[:, [t]] if (prompt is not None and t < p_len) else rvq_embed(sd, tok)
    return torch.stack(qs, dim=2).squeeze(-1), torch.cat(atts, dim=2), torch.stack(all_logits)


def random_state_dict(cfg, n_codebook: int = 4096, n_special: int = 3, n_txt_vocab: int = 256, n_quant: int = 1,
                      seed: int = 0) -> SD:
    """A reference-keyed fp32 state dict with the shapes of ``LinaModel(AttentiveGLA(d, n_layer, heads, blind=True,
    use_short_conv=True, pos_type='convolutional'), ..., txt_encoder=TextEncoder(d, txt_heads, n_layers=txt_layers))``
    (model/modeling_lina.py:24-58, model/gla.py:35-129,252-285, model/crossatt.py:61-103, model/base_blocks.py:42-62) and
    torch's default initialisers' scales (uniform(+-1/sqrt(fan_in)) matrices, unit norms, N(0,1) embeddings).  Lets the CPU
    baseline build its model without touching the product package; tests check keys and shapes against the host classes."""
    g = torch.Generator().manual_seed(seed)
    d, H = cfg["d_model"], cfg["heads"]
    kd, vd = int(d * cfg.get("expand_k", 1.0)), int(d * cfg.get("expand_v", 2.0))
    hid = 4 * d // 3
    sd: SD = {}

    def mat(name, out_f, in_f, bias=False):
        sd[name + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) / math.sqrt(in_f)
        if bias:
            sd[name + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) / math.sqrt(in_f)

    def norm(name, n):
        sd[name + ".weight"], sd[name + ".bias"] = torch.ones(n), torch.zeros(n)

    def cmix(p):
        mat(p + "cmix.p_in", 2 * hid, d, True)
        mat(p + "cmix.p_out", d, hid, True)

    def block(p):
        mat(p + "tmix.q_proj", kd, d); mat(p + "tmix.k_proj", kd, d); mat(p + "tmix.v_proj", vd, d); mat(p + "tmix.g_proj", vd, d)
        mat(p + "tmix.gk_proj.0", 16, d); mat(p + "tmix.gk_proj.1", kd, 16, True)
        mat(p + "tmix.o_proj", d, vd)
        for n_, w in (("q", kd), ("k", kd), ("v", vd)):
            sd[p + f"tmix.{n_}_conv1d.weight"] = (torch.rand(w, 1, 4, generator=g) * 2 - 1) / 2.0
        sd[p + "tmix.g_norm_swish_gate.weight"] = torch.ones(vd // H)
        cmix(p)
        norm(p + "norm1", d); norm(p + "norm2", d)

    for i in range(cfg.get("txt_layers", 0)):
        p = f"txt_encoder.sa.{i}."
        mat(p + "tmix.qkv", 3 * d, d, True)
        cmix(p)
        norm(p + "norm1", d); norm(p + "norm2", d)
    for i in range(cfg["n_layer"]):
        block(f"attentive_rnn.encoder.{i}.")
    for i in range(cfg["n_layer"]):
        block(f"attentive_rnn.decoder.{i}.")
    ca = "attentive_rnn.cross_att."
    for n_ in "qkv":
        mat(ca + n_, d, d, True)
    block(ca + "pos_net.")
    sd[ca + "pos_embed.embed.weight"] = torch.randn(2000, d, generator=g)
    sd[ca + "pos_embed.dw_conv.weight"] = (torch.rand(d, 1, 31, generator=g) * 2 - 1) / math.sqrt(31)
    sd[ca + "pos_embed.dw_conv.bias"] = (torch.rand(d, generator=g) * 2 - 1) / math.sqrt(31)
    for n_ in "qkv":
        norm(ca + "ln_" + n_, d)
    sd["txt_embed.weight"] = torch.randn(n_txt_vocab, d, generator=g)
    sd["rvq_embed.weight"] = torch.randn(n_quant, n_codebook + n_special, d, generator=g)
    sd["logits_head.weight"] = torch.randn(n_quant, n_codebook + n_special, d, generator=g) / math.sqrt(d)
    return sd

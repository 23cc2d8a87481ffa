"""CPU oracle for the WavTokenizer decode path -- TEST INFRASTRUCTURE ONLY.

Functional restatement (state-dict driven, reference key names) of
``WavTokenizer.codes_to_features`` + ``decode`` (paths relative to
/root/reference/3rdparty/decoder).  Same import rule as gla_oracle.py.
Pinned against the reference's own modules by tests/golden/make_golden.py.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
CODEBOOK = "feature_extractor.encodec.quantizer.vq.layers.{}._codebook.embed"


def codes_to_features(sd: SD, codes):
    """pretrained.py:209-239 -- codes [K,L] or [K,B,L] -> features [B,C,L]."""
    if codes.dim() == 2:
        codes = codes.unsqueeze(1)
    K = codes.shape[0]
    books = [sd[CODEBOOK.format(i)] for i in range(K)]
    bins = books[0].shape[0]
    table = torch.cat(books, dim=0)
    idx = codes + (torch.arange(K) * bins).view(-1, 1, 1)
    return F.embedding(idx, table).sum(0).transpose(1, 2)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)   # models.py:15-16


def _swish(x):
    return x * torch.sigmoid(x)                                                # models.py:10-12


def _conv(sd, p, x, pad):
    return F.conv1d(x, sd[p + ".weight"], sd[p + ".bias"], padding=pad)


def resnet_block(sd, p, x):
    """models.py:58-78 (temb is None, dropout in eval)."""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)), 1)
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)), 1)
    return x + h


def attn_block(sd, p, x):
    """models.py:107-127."""
    h = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, f"{p}.{n}", h, 0) for n in "qkv")
    c = q.shape[1]
    w = torch.softmax(torch.bmm(q.permute(0, 2, 1), k) * (int(c) ** -0.5), dim=2)
    h = torch.bmm(v, w.permute(0, 2, 1))
    return x + _conv(sd, p + ".proj_out", h, 0)


def ada_ln(sd, p, x, bandwidth_id):
    """modules.py:81-86 -- x [B,T,C]."""
    scale = F.embedding(bandwidth_id, sd[p + ".scale.weight"])
    shift = F.embedding(bandwidth_id, sd[p + ".shift.weight"])
    return F.layer_norm(x, x.shape[-1:], eps=1e-6) * scale + shift


def convnext_block(sd, p, x, bandwidth_id):
    """modules.py:43-60 -- x [B,C,T]."""
    w = sd[p + ".dwconv.weight"]
    h = F.conv1d(x, w, sd[p + ".dwconv.bias"], padding=3, groups=w.shape[0]).transpose(1, 2)
    h = ada_ln(sd, p + ".norm", h, bandwidth_id)
    h = F.linear(h, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"])
    h = F.gelu(h)
    h = F.linear(h, sd[p + ".pwconv2.weight"], sd[p + ".pwconv2.bias"])
    h = sd[p + ".gamma"] * h
    return x + h.transpose(1, 2)


def backbone(sd: SD, x, bandwidth_id):
    """VocosBackbone.forward (models.py:223-235): [B,C_in,L] -> [B,L,dim]."""
    x = _conv(sd, "backbone.embed", x, 3)
    for i in (0, 1):
        x = resnet_block(sd, f"backbone.pos_net.{i}", x)
    x = attn_block(sd, "backbone.pos_net.2", x)
    for i in (3, 4):
        x = resnet_block(sd, f"backbone.pos_net.{i}", x)
    x = _gn(sd, "backbone.pos_net.5", x)
    x = ada_ln(sd, "backbone.norm", x.transpose(1, 2), bandwidth_id).transpose(1, 2)
    i = 0
    while f"backbone.convnext.{i}.gamma" in sd:
        x = convnext_block(sd, f"backbone.convnext.{i}", x, bandwidth_id)
        i += 1
    return F.layer_norm(x.transpose(1, 2), x.shape[1:2], sd["backbone.final_layer_norm.weight"],
                        sd["backbone.final_layer_norm.bias"], eps=1e-6)


def istft_same(spec, window, n_fft: int, hop: int):
    """ISTFT.forward, padding='same' (spectral_ops.py:33-75): irfft * window,
    overlap-add, trim (win-hop)/2 per side, divide by the window^2 envelope."""
    B, N, T = spec.shape
    win = window.shape[0]
    pad = (win - hop) // 2
    frames = torch.fft.irfft(spec, n_fft, dim=1, norm="backward") * window[None, :, None]
    out_size = (T - 1) * hop + win
    y = F.fold(frames, output_size=(1, out_size), kernel_size=(1, win), stride=(1, hop))[:, 0, 0, pad:-pad]
    env = F.fold(window.square().expand(1, T, -1).transpose(1, 2), output_size=(1, out_size),
                 kernel_size=(1, win), stride=(1, hop)).squeeze()[pad:-pad]
    return y / env


def istft_head(sd: SD, x, hop: int):
    """ISTFTHead.forward (heads.py:42-67): x [B,L,dim] -> wav [B, L*hop]."""
    h = F.linear(x, sd["head.out.weight"], sd["head.out.bias"]).transpose(1, 2)
    mag, p = h.chunk(2, dim=1)
    mag = torch.clip(torch.exp(mag), max=1e2)
    S = mag * (torch.cos(p) + 1j * torch.sin(p))
    n_fft = sd["head.out.weight"].shape[0] - 2
    return istft_same(S, sd["head.istft.window"], n_fft, hop)


def decode(sd: SD, features, bandwidth_id, hop: int = 320):
    """WavTokenizer.decode (pretrained.py:192-207)."""
    return istft_head(sd, backbone(sd, features, bandwidth_id), hop)

"""WavTokenizer decode on B200: ``codes_to_features`` + ``decode`` with the reference's interface.

Mirrors ``3rdparty/decoder`` of the reference: WavTokenizer (pretrained.py:32-239), VocosBackbone
(models.py:152-235) with ResnetBlock (:19-78) / AttnBlock (:80-127), ConvNeXtBlock and AdaLayerNorm
(modules.py:8-86), ISTFTHead (heads.py:24-67) and ISTFT (spectral_ops.py:7-75).  Module / parameter
names are the reference's, so its checkpoints load with ``load_state_dict`` (encoder-side keys, which
decode never touches, are dropped by :func:`WavTokenizer.load_reference_state_dict`).

Every arithmetic stage runs in liblina_b200.so -- no cuDNN / cuBLAS call on this path:

  * activations are channels-last ``[B, L, C]`` fp32 from end to end (the reference's [B, C, L] <-> [B, L, C] transposes
    around every ConvNeXt block and norm disappear);
  * every contraction -- the k = 7 embed conv, the eight k = 3 ResnetBlock convs, the attention block's 1x1 convs and its two
    batched products, the 24 point-wise linears, the head's linear -- is ``lina_gemm_bf16_terms`` (tcgen05 / TMEM / TMA,
    csrc/gemm_sm100.cu) on bf16 SPLITS of the fp32 tensors: with three parts and six part products per contraction
    (``gemm_precision = "bf16x3"``, the default) all 24 significand bits of both operands take part, i.e. the reference's
    fp32 arithmetic; ``"bf16x2"`` (two parts, three products) keeps 16 bits at half the tensor-core time;
  * bias, GELU, layer scale gamma and the residual add are GEMM epilogues; conv taps are stretches of the GEMM's K loop whose
    A tile is shifted by the tap (TMA zero fill = the conv's padding); the stages in between (GroupNorm, swish, depthwise conv,
    AdaLayerNorm, softmax, codebook gather) are one-pass row kernels (csrc/codec_cl.cu) that write the next GEMM's operand
    parts directly; the ISTFT head's tail is the warp-per-frame FFT + overlap-add of csrc/codec.cu.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Optional

import torch
from torch import nn

from .. import _lib as L
from . import gemm as G

# bench hook: a list collects (stage, algorithmic bytes, flops, start event, end event) around every launch group
PROFILE = None
PRECISIONS = {"bf16x3": 3, "bf16x2": 2}


class _timed:
    def __init__(self, name: str, nbytes: int, flops: int, ref: torch.Tensor):
        self.name, self.nbytes, self.flops, self.ref = name, nbytes, flops, ref

    def __enter__(self):
        if PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.ref.device))
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(torch.cuda.current_stream(self.ref.device))
            PROFILE.append((self.name, self.nbytes, self.flops, self.e0, self.e1))
        return False


def _ptr_array(ts):
    arr = (C.c_void_p * 3)()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


def _f32c(t: torch.Tensor) -> torch.Tensor:
    L.require_cuda(t)
    return t.detach().float().contiguous()


def rows(x, *, dw=None, gn=None, swish=False, ln=None, ln_eps=1e-6, out_f32=False, parts=0, name="rows"):
    """One pass per row of x [B, L, C] (lina_codec_cl_rows): dw = (weight [C,7], bias [C]); gn = (partials, weight, bias,
    groups, eps); ln = (scale [C], shift [C]).  Returns (fp32 [B,L,C] or None, tuple of bf16 parts)."""
    B, Ln, Cc = x.shape
    o32 = torch.empty_like(x) if out_f32 else None
    ps = tuple(torch.empty(B, Ln, Cc, dtype=torch.bfloat16, device=x.device) for _ in range(parts))
    dw_w, dw_b = dw if dw is not None else (None, None)
    gp, gw, gb, groups, geps = gn if gn is not None else (None, None, None, 0, 0.0)
    sc, sh = ln if ln is not None else (None, None)
    nbytes = x.numel() * 4 + (x.numel() * 4 if out_f32 else 0) + x.numel() * 2 * parts
    with _timed(name, nbytes, 0, x):
        rc = L.lib().lina_codec_cl_rows(L.ptr(x), L.ptr(dw_w), L.ptr(dw_b), L.ptr(gp), L.ptr(gw), L.ptr(gb), groups, geps,
                                        int(swish), L.ptr(sc), L.ptr(sh), ln_eps, L.ptr(o32), _ptr_array(ps), parts, B, Ln, Cc,
                                        L.stream(x))
    L.count_launches(1)
    L.check(rc, "lina_codec_cl_rows")
    return o32, ps


def gn_partials(x, groups: int):
    B, Ln, Cc = x.shape
    ws = torch.empty(int(L.lib().lina_codec_cl_gn_partials_bytes(B, Ln, groups)) // 4, dtype=torch.float32, device=x.device)
    with _timed("gn_stats", x.numel() * 4, 0, x):
        rc = L.lib().lina_codec_cl_gn_partials(L.ptr(x), L.ptr(ws), B, Ln, Cc, groups, L.stream(x))
    L.count_launches(1)
    L.check(rc, "lina_codec_cl_gn_partials")
    return ws


def gemm(a, w, *, name, NB, Ln, N, K, **kw):
    """lina_gemm_bf16_terms with the bench hook: flops = 2 * rows * N * K * taps * terms (tensor-core work actually issued)."""
    terms = kw.get("terms") or G.TERMS[min(len(a), len(w))]
    flops = 2 * NB * Ln * N * K * kw.get("taps", 1) * len(terms)
    with _timed(name, 0, flops, a[0]):
        return G.gemm_terms(a, w, NB=NB, Ln=Ln, N=N, K=K, **kw)


class ResnetBlock(nn.Module):
    """models.py:19-78 with in == out channels, temb unused, dropout in eval (parameters; the arithmetic is in
    VocosBackbone.forward)."""

    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout=0.1, temb_channels=0):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        if out_channels != in_channels or temb_channels > 0:
            raise NotImplementedError("VocosBackbone.pos_net only builds same-width ResnetBlocks without temb")
        self.norm1 = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.conv1 = nn.Conv1d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = nn.GroupNorm(32, out_channels, eps=1e-6, affine=True)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv1d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)


class AttnBlock(nn.Module):
    """models.py:80-127: single-head full softmax attention over the sequence (parameters)."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.q = nn.Conv1d(in_channels, in_channels, 1)
        self.k = nn.Conv1d(in_channels, in_channels, 1)
        self.v = nn.Conv1d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv1d(in_channels, in_channels, 1)


class AdaLayerNorm(nn.Module):
    """modules.py:63-86 (parameters only; the arithmetic is the LayerNorm stage of lina_codec_cl_rows)."""

    def __init__(self, num_embeddings: int, embedding_dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps, self.dim = eps, embedding_dim
        self.scale = nn.Embedding(num_embeddings, embedding_dim)
        self.shift = nn.Embedding(num_embeddings, embedding_dim)
        nn.init.ones_(self.scale.weight)
        nn.init.zeros_(self.shift.weight)

    def rows(self, cond_embedding_id):
        """(scale, shift) rows of the conditioning id: a device-side gather (no host read of the id tensor)."""
        if torch.is_tensor(cond_embedding_id):
            i = cond_embedding_id.reshape(-1)[:1].to(self.scale.weight.device)
            return (self.scale.weight.index_select(0, i)[0].float().contiguous(),
                    self.shift.weight.index_select(0, i)[0].float().contiguous())
        return (self.scale.weight[int(cond_embedding_id)].float().contiguous(),
                self.shift.weight[int(cond_embedding_id)].float().contiguous())


class ConvNeXtBlock(nn.Module):
    """modules.py:8-60 (parameters)."""

    def __init__(self, dim: int, intermediate_dim: int, layer_scale_init_value: Optional[float] = None,
                 adanorm_num_embeddings: Optional[int] = None):
        super().__init__()
        self.dwconv = nn.Conv1d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.adanorm = adanorm_num_embeddings is not None
        self.norm = AdaLayerNorm(adanorm_num_embeddings, dim, eps=1e-6) if self.adanorm else nn.LayerNorm(dim, eps=1e-6)
        self.pwconv1 = nn.Linear(dim, intermediate_dim)
        self.act = nn.GELU()
        self.pwconv2 = nn.Linear(intermediate_dim, dim)
        self.gamma = (nn.Parameter(layer_scale_init_value * torch.ones(dim))
                      if layer_scale_init_value is not None and layer_scale_init_value > 0 else None)


def _ver(t) -> int:
    try:
        return t._version
    except RuntimeError:
        return -1


class VocosBackbone(nn.Module):
    """models.py:152-235."""

    def __init__(self, input_channels: int, dim: int, intermediate_dim: int, num_layers: int,
                 layer_scale_init_value: Optional[float] = None, adanorm_num_embeddings: Optional[int] = None):
        super().__init__()
        self.input_channels = input_channels
        self.embed = nn.Conv1d(input_channels, dim, kernel_size=7, padding=3)
        self.adanorm = adanorm_num_embeddings is not None
        self.norm = AdaLayerNorm(adanorm_num_embeddings, dim, eps=1e-6) if self.adanorm else nn.LayerNorm(dim, eps=1e-6)
        layer_scale_init_value = layer_scale_init_value or 1 / num_layers
        self.convnext = nn.ModuleList([ConvNeXtBlock(dim, intermediate_dim, layer_scale_init_value,
                                                     adanorm_num_embeddings) for _ in range(num_layers)])
        self.final_layer_norm = nn.LayerNorm(dim, eps=1e-6)
        self.pos_net = nn.Sequential(ResnetBlock(in_channels=dim), ResnetBlock(in_channels=dim), AttnBlock(dim),
                                     ResnetBlock(in_channels=dim), ResnetBlock(in_channels=dim),
                                     nn.GroupNorm(32, dim, eps=1e-6, affine=True))

    _prepared = None      # class-level default: (key, parts, dict of split weights)

    # -- weights as GEMM operands: [N][tap][K] bf16 parts, built once per (parameters, precision) ----------------------
    def prepared(self, parts: int):
        ps = list(self.parameters())
        key = (parts, tuple((p.data_ptr(), _ver(p), p.dtype, p.device) for p in ps))
        if self._prepared is not None and self._prepared[0] == key:
            return self._prepared[1]

        def conv_w(m):          # [Co, Ci, k] -> [Co, k * Ci]: tap-major along K
            w = m.weight.detach().float()
            return G.split(w.permute(0, 2, 1).reshape(w.shape[0], -1), parts)

        def lin_w(m):
            return G.split(m.weight.detach().float(), parts)

        W = {"embed": (conv_w(self.embed), _f32c(self.embed.bias))}
        for i in (0, 1, 3, 4):
            b = self.pos_net[i]
            W[f"res{i}"] = dict(n1=(_f32c(b.norm1.weight), _f32c(b.norm1.bias)), c1=(conv_w(b.conv1), _f32c(b.conv1.bias)),
                                n2=(_f32c(b.norm2.weight), _f32c(b.norm2.bias)), c2=(conv_w(b.conv2), _f32c(b.conv2.bias)))
        at = self.pos_net[2]
        wqkv = torch.cat([m.weight.detach().float()[:, :, 0] for m in (at.q, at.k, at.v)], dim=0)
        W["attn"] = dict(n=(_f32c(at.norm.weight), _f32c(at.norm.bias)),
                         qkv=(G.split(wqkv, parts), torch.cat([_f32c(m.bias) for m in (at.q, at.k, at.v)]).contiguous()),
                         proj=(G.split(at.proj_out.weight.detach().float()[:, :, 0], parts), _f32c(at.proj_out.bias)))
        gn = self.pos_net[5]
        W["gn5"] = (_f32c(gn.weight), _f32c(gn.bias))
        for i, blk in enumerate(self.convnext):
            W[f"cnx{i}"] = dict(dw=(_f32c(blk.dwconv.weight).view(blk.dwconv.weight.shape[0], -1), _f32c(blk.dwconv.bias)),
                                p1=(lin_w(blk.pwconv1), _f32c(blk.pwconv1.bias)),
                                p2=(lin_w(blk.pwconv2), _f32c(blk.pwconv2.bias)),
                                gamma=_f32c(blk.gamma) if blk.gamma is not None else None)
        W["final"] = (_f32c(self.final_layer_norm.weight), _f32c(self.final_layer_norm.bias))
        self._prepared = (key, W)
        return W

    def forward(self, x: torch.Tensor, bandwidth_id: Optional[torch.Tensor] = None, parts: int = 3, f_parts=None):
        """x: features, channels-last [B, L, C_in] fp32 (``f_parts``: their bf16 parts when the gather already wrote them).
        Returns the bf16 parts of final_layer_norm(...) [B, L, dim] -- the head GEMM's operand."""
        B, Ln, Cin = x.shape
        D = self.embed.out_channels
        W = self.prepared(parts)
        if f_parts is None:
            _, f_parts = rows(x, parts=parts, name="split_features")
        kw = dict(NB=B, Ln=Ln)
        # embed: Conv1d(C_in, dim, 7, padding 3)                                       models.py:224
        w, b = W["embed"]
        x, _ = gemm(f_parts, w, name="embed_conv7", N=D, K=Cin, taps=7, pad=3, bias=b, **kw)
        for i in (0, 1, 2, 3, 4):
            if i == 2:                                                                 # AttnBlock, models.py:107-127
                A = W["attn"]
                _, h = rows(x, gn=(gn_partials(x, 32), *A["n"], 32, 1e-6), parts=parts, name="groupnorm")
                _, qkv = gemm(h, A["qkv"][0], name="attn_qkv", N=3 * D, K=D, bias=A["qkv"][1], out_f32=False, out_parts=parts, **kw)
                q = tuple(t[..., :D] for t in qkv)
                k = tuple(t[..., D:2 * D] for t in qkv)
                v = tuple(t[..., 2 * D:] for t in qkv)
                Lp = (Ln + 7) // 8 * 8
                S = torch.empty(B, Ln, Lp, dtype=torch.float32, device=x.device)
                gemm(q, k, name="attn_scores", N=Ln, K=D, b_batched=True, alpha=float(int(D) ** -0.5), lda=3 * D, ldb=3 * D,
                     a_batch_stride=Ln * 3 * D, b_batch_stride=Ln * 3 * D, out=S, **kw)
                P = tuple(torch.empty(B, Ln, Lp, dtype=torch.bfloat16, device=x.device) for _ in range(parts))
                with _timed("attn_softmax", S.numel() * 4 + P[0].numel() * 2 * parts, 0, S):
                    rc = L.lib().lina_codec_cl_softmax(L.ptr(S), _ptr_array(P), parts, B * Ln, Ln, Lp, Lp, L.stream(S))
                L.count_launches(1)
                L.check(rc, "lina_codec_cl_softmax")
                _, o = gemm(P, v, name="attn_pv", N=D, K=Ln, b_batched=True, b_mn=True, ldb=3 * D, b_batch_stride=Ln * 3 * D,
                            out_f32=False, out_parts=parts, **kw)
                gemm(o, A["proj"][0], name="attn_proj", N=D, K=D, bias=A["proj"][1], residual=x, out=x, **kw)
                continue
            R = W[f"res{i}"]                                                           # ResnetBlock, models.py:58-78
            _, h = rows(x, gn=(gn_partials(x, 32), *R["n1"], 32, 1e-6), swish=True, parts=parts, name="groupnorm_swish")
            t, _ = gemm(h, R["c1"][0], name="res_conv3", N=D, K=D, taps=3, pad=1, bias=R["c1"][1], **kw)
            _, h = rows(t, gn=(gn_partials(t, 32), *R["n2"], 32, 1e-6), swish=True, parts=parts, name="groupnorm_swish")
            gemm(h, R["c2"][0], name="res_conv3", N=D, K=D, taps=3, pad=1, bias=R["c2"][1], residual=x, out=x, **kw)
        # pos_net[5] GroupNorm, then backbone.norm (AdaLayerNorm / LayerNorm) in the same pass      models.py:226-230
        if self.adanorm:
            assert bandwidth_id is not None
            ln = self.norm.rows(bandwidth_id)
        else:
            ln = (_f32c(self.norm.weight), _f32c(self.norm.bias))
        x, _ = rows(x, gn=(gn_partials(x, 32), *W["gn5"], 32, 1e-6), ln=ln, out_f32=True, name="groupnorm_adaln")
        for i, blk in enumerate(self.convnext):                                        # ConvNeXtBlock, modules.py:43-60
            Cx = W[f"cnx{i}"]
            ln = blk.norm.rows(bandwidth_id) if blk.adanorm else (_f32c(blk.norm.weight), _f32c(blk.norm.bias))
            _, h = rows(x, dw=Cx["dw"], ln=ln, parts=parts, name="dwconv_adaln")
            I = blk.pwconv1.out_features
            _, g = gemm(h, Cx["p1"][0], name="pwconv1_gelu", N=I, K=D, bias=Cx["p1"][1], act="gelu", out_f32=False,
                        out_parts=parts, **kw)
            gemm(g, Cx["p2"][0], name="pwconv2_scale_residual", N=D, K=I, bias=Cx["p2"][1], gamma=Cx["gamma"], residual=x,
                 out=x, lda=g[0].stride(-2), **kw)
        _, y = rows(x, ln=W["final"], parts=parts, name="final_layernorm")
        return y


class ISTFT(nn.Module):
    """spectral_ops.py:7-75, padding='same' only (the shipped config)."""

    def __init__(self, n_fft: int, hop_length: int, win_length: int, padding: str = "same"):
        super().__init__()
        if padding != "same" or win_length != n_fft:
            raise NotImplementedError("only padding='same' with win_length == n_fft (WavTokenizer's config)")
        self.padding, self.n_fft, self.hop_length, self.win_length = padding, n_fft, hop_length, win_length
        self.register_buffer("window", torch.hann_window(win_length))


class ISTFTHead(nn.Module):
    """heads.py:24-67."""

    def __init__(self, dim: int, n_fft: int, hop_length: int, padding: str = "same"):
        super().__init__()
        self.out = nn.Linear(dim, n_fft + 2)
        self.istft = ISTFT(n_fft=n_fft, hop_length=hop_length, win_length=n_fft, padding=padding)

    _prepared = None

    def _weights(self, parts: int):
        w, b = self.out.weight, self.out.bias
        key = (parts, w.data_ptr(), _ver(w), b.data_ptr(), _ver(b), w.device)
        if self._prepared is None or self._prepared[0] != key:
            self._prepared = (key, G.split(w.detach().float(), parts), _f32c(b))
        return self._prepared[1], self._prepared[2]

    def forward(self, x, parts: int = 3) -> torch.Tensor:
        """x: [B, L, dim] fp32, or the tuple of its bf16 parts (what the backbone hands over)."""
        if torch.is_tensor(x):
            L.require_cuda(x)
            _, x = rows(x.float().contiguous(), parts=parts, name="split_head_input")
        B, Ln, D = x[0].shape
        n_fft, hop = self.istft.n_fft, self.istft.hop_length
        w, b = self._weights(len(x))
        ldh = (n_fft + 2 + 3) // 4 * 4
        h = torch.empty(B, Ln, ldh, dtype=torch.float32, device=x[0].device)
        gemm(x, w, name="head_linear", NB=B, Ln=Ln, N=n_fft + 2, K=D, bias=b, out=h)
        lib = L.lib()
        wav = torch.empty(B, Ln * hop, dtype=torch.float32, device=h.device)
        ws = torch.empty(int(lib.lina_codec_istft_workspace_bytes(B, Ln, n_fft)), dtype=torch.uint8, device=h.device)
        with _timed("istft_head", 4 * (B * Ln * (n_fft + 2) + wav.numel()), 0, h):
            rc = lib.lina_codec_istft_head_ld(L.ptr(h), ldh, L.ptr(_f32c(self.istft.window)), L.ptr(wav), L.ptr(ws), B, Ln,
                                              n_fft, hop, L.stream(h))
        L.count_launches(2)
        L.check(rc, "lina_codec_istft_head_ld")
        return wav


class _Codebook(nn.Module):
    def __init__(self, bins: int, dim: int):
        super().__init__()
        self.register_buffer("embed", torch.zeros(bins, dim))


class _Holder(nn.Module):
    pass


class CodebookFeatures(nn.Module):
    """Decode-side stand-in for EncodecFeatures (feature_extractors.py:54-141): only the quantizer codebooks
    (``encodec.quantizer.vq.layers[i]._codebook.embed`` [bins, 512]) are needed to decode."""

    def __init__(self, num_quantizers: int = 1, vq_bins: int = 4096, dimension: int = 512):
        super().__init__()
        self.encodec = _Holder()
        self.encodec.quantizer = _Holder()
        self.encodec.quantizer.bins = vq_bins
        self.encodec.quantizer.vq = _Holder()
        self.encodec.quantizer.vq.layers = nn.ModuleList()
        for _ in range(num_quantizers):
            layer = _Holder()
            layer._codebook = _Codebook(vq_bins, dimension)
            self.encodec.quantizer.vq.layers.append(layer)

    def codebooks(self):
        return torch.cat([l._codebook.embed for l in self.encodec.quantizer.vq.layers], dim=0)


class WavTokenizer(nn.Module):
    """pretrained.py:32-239 -- ``codes_to_features`` (:209-239) and ``decode`` (:192-207)."""

    def __init__(self, feature_extractor: CodebookFeatures, backbone: VocosBackbone, head: ISTFTHead):
        super().__init__()
        self.feature_extractor, self.backbone, self.head = feature_extractor, backbone, head
        # "bf16x3": every contraction as six bf16 part products (24 significand bits per operand = the reference's fp32);
        # "bf16x2": three part products (16 bits per operand), half the tensor-core time.
        self.gemm_precision = "bf16x3"

    @classmethod
    def from_hparams(cls, *, num_quantizers=1, vq_bins=4096, input_channels=512, dim=768, intermediate_dim=2304,
                     num_layers=12, adanorm_num_embeddings=4, n_fft=1280, hop_length=320, padding="same"):
        """The shipped config (wavtokenizer_mediumdata_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml)."""
        return cls(CodebookFeatures(num_quantizers, vq_bins, input_channels),
                   VocosBackbone(input_channels, dim, intermediate_dim, num_layers,
                                 adanorm_num_embeddings=adanorm_num_embeddings),
                   ISTFTHead(dim, n_fft, hop_length, padding))

    @classmethod
    def from_hparams0802(cls, config_path: str) -> "WavTokenizer":
        """pretrained.py:81-92: build from the training yaml (``model.init_args.{feature_extractor,backbone,head}.init_args``).
        Only the decode side is instantiated; encoder-only keys of the feature extractor are ignored."""
        import yaml
        with open(config_path, "r") as f:
            init = yaml.safe_load(f)["model"]["init_args"]
        fe = init["feature_extractor"].get("init_args", {})
        bb = init["backbone"].get("init_args", {})
        hd = init["head"].get("init_args", {})
        backbone = VocosBackbone(bb["input_channels"], bb["dim"], bb["intermediate_dim"], bb["num_layers"],
                                 bb.get("layer_scale_init_value"), bb.get("adanorm_num_embeddings"))
        head = ISTFTHead(hd["dim"], hd["n_fft"], hd["hop_length"], hd.get("padding", "same"))
        return cls(CodebookFeatures(fe.get("num_quantizers", 1), fe.get("vq_bins", 16384), bb["input_channels"]), backbone, head)

    @classmethod
    def from_pretrained0802(cls, config_path: str, model_path: str) -> "WavTokenizer":
        """pretrained.py:95-115 (what InferenceLina.ipynb calls): yaml + Lightning checkpoint; the decode-side tensors of
        ``state_dict`` (backbone.*, head.*, the quantizer codebooks) are loaded strictly, the model is returned in eval mode."""
        model = cls.from_hparams0802(config_path)
        raw = torch.load(model_path, map_location="cpu", weights_only=False)["state_dict"]
        model.load_reference_state_dict(raw)
        return model.eval()

    def load_reference_state_dict(self, state_dict):
        keep = {k: v for k, v in state_dict.items()
                if k.startswith(("backbone.", "head.")) or (k.startswith("feature_extractor.") and k.endswith("_codebook.embed"))}
        return self.load_state_dict(keep, strict=True)

    def _parts(self) -> int:
        try:
            return PRECISIONS[self.gemm_precision]
        except KeyError:
            raise ValueError(f"gemm_precision must be one of {sorted(PRECISIONS)}") from None

    @torch.inference_mode()
    def decode(self, features_input: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        """features [B, C, L] (what ``codes_to_features`` returns) -> waveform [B, L * hop]."""
        L.require_cuda(features_input)
        parts = self._parts()
        x = features_input
        if x.dim() != 3:
            raise ValueError("decode expects features [B, C, L]")
        xt = x.transpose(1, 2)                                  # channels-last view
        f_parts = getattr(features_input, "_lina_parts", None)
        if xt.dtype != torch.float32 or not xt.is_contiguous():
            xt, f_parts = xt.float().contiguous(), None
        if f_parts is not None and (len(f_parts) != parts or f_parts[0].shape != xt.shape):
            f_parts = None
        y = self.backbone(xt, parts=parts, f_parts=f_parts, **kwargs)
        return self.head(y, parts=parts)

    @torch.inference_mode()
    def codes_to_features(self, codes: torch.Tensor) -> torch.Tensor:
        """codes [K, L] or [K, B, L] -> features [B, C, L] (pretrained.py:209-239).  The memory behind the returned tensor
        is channels-last ([B, L, C], what the decoder's first contraction reads); the bf16 parts of the same values ride
        along on the tensor object, so ``decode(codes_to_features(c))`` never re-reads the features to split them."""
        L.require_cuda(codes)
        if codes.dim() == 2:
            codes = codes.unsqueeze(1)
        codes = codes.long().contiguous()
        Kq, B, Ln = codes.shape
        books = _f32c(self.feature_extractor.codebooks())
        bins, Cc = self.feature_extractor.encodec.quantizer.bins, books.shape[1]
        parts = self._parts()
        out = torch.empty(B, Ln, Cc, dtype=torch.float32, device=codes.device)
        ps = tuple(torch.empty(B, Ln, Cc, dtype=torch.bfloat16, device=codes.device) for _ in range(parts))
        with _timed("codes_to_features", 8 * codes.numel() + 4 * Kq * out.numel() + 4 * out.numel() + 2 * parts * out.numel(), 0, out):
            rc = L.lib().lina_codec_cl_gather(L.ptr(codes), L.ptr(books), L.ptr(out), _ptr_array(ps), parts, Kq, B, Ln, bins, Cc,
                                              L.stream(codes))
        L.count_launches(1)
        L.check(rc, "lina_codec_cl_gather")
        feats = out.transpose(1, 2)
        feats._lina_parts = ps
        return feats

"""WavTokenizer decode on B200: ``codes_to_features`` + ``decode`` with the reference's interface.

Mirrors ``3rdparty/decoder`` of the reference: WavTokenizer (pretrained.py:32-239), VocosBackbone
(models.py:152-235) with ResnetBlock (:19-78) / AttnBlock (:80-127), ConvNeXtBlock and AdaLayerNorm
(modules.py:8-86), ISTFTHead (heads.py:24-67) and ISTFT (spectral_ops.py:7-75).  Module / parameter
names are the reference's, so its checkpoints load with ``load_state_dict`` (encoder-side keys, which
decode never touches, are dropped by :func:`WavTokenizer.load_reference_state_dict`).

Dense convolutions and linears are library GEMMs (cuDNN / cuBLAS through torch); every stage between
them -- codebook gather+transpose, GroupNorm+swish, depthwise conv + transpose + AdaLayerNorm,
layer-scale + transpose + residual, final LayerNorm, and the whole ISTFT head tail (polar, inverse real
FFT, window, overlap-add, envelope normalisation) -- runs in liblina_b200.so.  fp32 like the reference.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib as L


# bench hook: a list collects (stage, algorithmic bytes, start event, end event) around every kernel of ours
PROFILE = None


class _timed:
    def __init__(self, name: str, nbytes: int, ref: torch.Tensor):
        self.name, self.nbytes, self.ref = name, nbytes, ref

    def __enter__(self):
        if PROFILE is not None:
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record(torch.cuda.current_stream(self.ref.device))
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e1.record(torch.cuda.current_stream(self.ref.device))
            PROFILE.append((self.name, self.nbytes, self.e0, self.e1))
        return False


def _f32c(t: torch.Tensor) -> torch.Tensor:
    L.require_cuda(t)
    return t.float().contiguous()


def groupnorm_swish(x, weight, bias, groups: int, eps: float, swish: bool):
    x = _f32c(x)
    B, C, Ln = x.shape
    y = torch.empty_like(x)
    with _timed("groupnorm_swish", 8 * x.numel(), x):
        rc = L.lib().lina_codec_groupnorm_swish(L.ptr(x), L.ptr(_f32c(weight)), L.ptr(_f32c(bias)), L.ptr(y), None,
                                                B, C, Ln, groups, eps, int(swish), L.stream(x))
    L.count_launches(1)
    L.check(rc, "lina_codec_groupnorm_swish")
    return y


def dwconv_adaln(x, dw_w, dw_b, scale, shift, eps: float):
    """x [B,C,L] -> [B,L,C]; dw_w None = no conv (plain transposing AdaLN / LayerNorm)."""
    x = _f32c(x)
    B, C, Ln = x.shape
    y = torch.empty(B, Ln, C, dtype=torch.float32, device=x.device)
    w = _f32c(dw_w).view(C, -1) if dw_w is not None else None
    if w is not None and w.shape[1] != 7:
        raise NotImplementedError("depthwise kernel size must be 7 (ConvNeXtBlock, modules.py:28)")
    lib = L.lib()
    if w is None:       # no conv: the single kernel (all channels of a 32-step tile per CTA) measures faster (0.052 vs 0.084 ms)
        with _timed("layernorm_t", 8 * x.numel(), x):
            rc = lib.lina_codec_dwconv_adaln(L.ptr(x), None, None, L.ptr(_f32c(scale)), L.ptr(_f32c(shift)), L.ptr(y), B, C, Ln,
                                             eps, L.stream(x))
        L.count_launches(1)
        L.check(rc, "lina_codec_dwconv_adaln")
        return y
    sc, sh = _f32c(scale), _f32c(shift)
    if sc.data_ptr() % 16 or sh.data_ptr() % 16:          # rows of an embedding table: the apply kernel reads them as float4
        sc, sh = sc.clone(), sh.clone()
    ws = torch.empty(int(lib.lina_codec_dwconv_adaln_workspace_bytes(B, C, Ln)), dtype=torch.uint8, device=x.device)
    with _timed("dwconv_adaln", 8 * x.numel(), x):
        rc = lib.lina_codec_dwconv_adaln_ws(L.ptr(x), L.ptr(w), L.ptr(_f32c(dw_b)) if dw_b is not None else None,
                                            L.ptr(sc), L.ptr(sh), L.ptr(y), L.ptr(ws), B, C, Ln, eps, L.stream(x))
    L.count_launches(2)
    L.check(rc, "lina_codec_dwconv_adaln_ws")
    return y


def scale_residual_t(h, gamma, res):
    """h [B,L,C], res [B,C,L] -> res + gamma * h^T."""
    h, res = _f32c(h), _f32c(res)
    B, Ln, C = h.shape
    out = torch.empty_like(res)
    with _timed("scale_residual_t", 12 * h.numel(), h):
        rc = L.lib().lina_codec_scale_residual_t(L.ptr(h), L.ptr(_f32c(gamma)) if gamma is not None else None,
                                                 L.ptr(res), L.ptr(out), B, C, Ln, L.stream(h))
    L.count_launches(1)
    L.check(rc, "lina_codec_scale_residual_t")
    return out


class ResnetBlock(nn.Module):
    """models.py:19-78 with in == out channels, temb unused, dropout in eval."""

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

    def forward(self, x, temb=None):
        h = self.conv1(groupnorm_swish(x, self.norm1.weight, self.norm1.bias, 32, 1e-6, True))
        h = self.conv2(groupnorm_swish(h, self.norm2.weight, self.norm2.bias, 32, 1e-6, True))
        return x + h


class AttnBlock(nn.Module):
    """models.py:80-127: single-head full softmax attention over the sequence."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.q = nn.Conv1d(in_channels, in_channels, 1)
        self.k = nn.Conv1d(in_channels, in_channels, 1)
        self.v = nn.Conv1d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv1d(in_channels, in_channels, 1)

    def forward(self, x):
        h = groupnorm_swish(x, self.norm.weight, self.norm.bias, 32, 1e-6, False)
        q, k, v = self.q(h), self.k(h), self.v(h)
        c = q.shape[1]
        w = torch.softmax(torch.bmm(q.permute(0, 2, 1), k) * (int(c) ** -0.5), dim=2)
        return x + self.proj_out(torch.bmm(v, w.permute(0, 2, 1)))


class AdaLayerNorm(nn.Module):
    """modules.py:63-86 (parameters only; the arithmetic is fused into dwconv_adaln)."""

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
            return self.scale.weight.index_select(0, i)[0], self.shift.weight.index_select(0, i)[0]
        return self.scale.weight[int(cond_embedding_id)], self.shift.weight[int(cond_embedding_id)]


class ConvNeXtBlock(nn.Module):
    """modules.py:8-60."""

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

    def forward(self, x, cond_embedding_id=None):
        if self.adanorm:
            assert cond_embedding_id is not None
            scale, shift = self.norm.rows(cond_embedding_id)
        else:
            scale, shift = self.norm.weight, self.norm.bias
        h = dwconv_adaln(x, self.dwconv.weight, self.dwconv.bias, scale, shift, 1e-6)     # [B,L,C]
        h = self.pwconv2(self.act(self.pwconv1(h)))
        return scale_residual_t(h, self.gamma, x)


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

    def forward(self, x: torch.Tensor, bandwidth_id: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self.embed(_f32c(x))
        for i in range(5):
            x = self.pos_net[i](x)
        gn = self.pos_net[5]
        x = groupnorm_swish(x, gn.weight, gn.bias, 32, 1e-6, False)
        if self.adanorm:
            assert bandwidth_id is not None
            scale, shift = self.norm.rows(bandwidth_id)
        else:
            scale, shift = self.norm.weight, self.norm.bias
        x = dwconv_adaln(x, None, None, scale, shift, 1e-6).transpose(1, 2).contiguous()       # [B,C,L]
        for blk in self.convnext:
            x = blk(x, cond_embedding_id=bandwidth_id)
        fl = self.final_layer_norm
        return dwconv_adaln(x, None, None, fl.weight, fl.bias, 1e-6)                           # [B,L,C]


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

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = _f32c(self.out(x))                                   # [B,L,n_fft+2]
        B, Ln, _ = h.shape
        n_fft, hop = self.istft.n_fft, self.istft.hop_length
        lib = L.lib()
        wav = torch.empty(B, Ln * hop, dtype=torch.float32, device=h.device)
        ws = torch.empty(int(lib.lina_codec_istft_workspace_bytes(B, Ln, n_fft)), dtype=torch.uint8, device=h.device)
        with _timed("istft_head", 4 * (h.numel() + wav.numel()), h):
            rc = lib.lina_codec_istft_head(L.ptr(h), L.ptr(_f32c(self.istft.window)), L.ptr(wav), L.ptr(ws), B, Ln,
                                           n_fft, hop, L.stream(h))
        L.count_launches(2)
        L.check(rc, "lina_codec_istft_head")
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
        # "fp32": library GEMMs / convolutions in full fp32 (bit-for-bit the reference's default math);
        # "tf32": let cuBLAS / cuDNN use TF32 tensor cores for them (10-bit mantissa operands, fp32 accumulate) --
        #         ~10x faster GEMMs, waveform error ~1e-3 relative.  The kernels of liblina_b200 are fp32 either way.
        self.gemm_precision = "fp32"

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

    @torch.inference_mode()
    def decode(self, features_input: torch.Tensor, **kwargs: Any) -> torch.Tensor:
        if self.gemm_precision not in ("fp32", "tf32"):
            raise ValueError("gemm_precision must be 'fp32' or 'tf32'")
        old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        tf32 = self.gemm_precision == "tf32"
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32, tf32
        try:
            return self.head(self.backbone(features_input, **kwargs))
        finally:
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old

    @torch.inference_mode()
    def codes_to_features(self, codes: torch.Tensor) -> torch.Tensor:
        L.require_cuda(codes)
        if codes.dim() == 2:
            codes = codes.unsqueeze(1)
        codes = codes.long().contiguous()
        Kq, B, Ln = codes.shape
        books = _f32c(self.feature_extractor.codebooks())
        bins, C = self.feature_extractor.encodec.quantizer.bins, books.shape[1]
        out = torch.empty(B, C, Ln, dtype=torch.float32, device=codes.device)
        with _timed("codes_to_features", 8 * codes.numel() + 4 * Kq * out.numel() + 4 * out.numel(), out):
            rc = L.lib().lina_codec_codes_to_features(L.ptr(codes), L.ptr(books), L.ptr(out), Kq, B, Ln, bins, C,
                                                      L.stream(codes))
        L.count_launches(1)
        L.check(rc, "lina_codec_codes_to_features")
        return out

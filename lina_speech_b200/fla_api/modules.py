"""Module-level drop-ins for the fla pieces ``model/gla.py:19`` imports:

  fla.modules.ShortConvolution        FLA/fla/modules/convolution.py:79-209
  fla.modules.FusedRMSNormSwishGate   FLA/fla/modules/fused_norm_gate.py:765-803
  fla.models.utils.Cache              FLA/fla/models/utils.py:11-107

Same constructor arguments, parameter names / shapes (so reference checkpoints load) and
call signatures; the arithmetic runs in liblina_b200.so.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from .. import _lib as L


# ------------------------------------------------------------------------------------------------
class _ShortConvFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, cache, silu):
        L.require_cuda(x, weight, cache)
        x = x.contiguous()
        w = weight.to(x.dtype).contiguous()                      # [D,1,W] == [D,W] in memory
        B, Ln, D = x.shape
        W = w.shape[-1]
        y = torch.empty_like(x)
        rc = L.lib().lina_short_conv_fwd(L.ptr(x), L.ptr(w), L.ptr(y), L.ptr(cache),
                                         L.dt(cache) if cache is not None else 0, B, Ln, D, W, int(silu),
                                         L.dt(x), L.stream(x))
        L.count_launches(1)
        L.check(rc, "lina_short_conv_fwd")
        ctx.save_for_backward(x, w)
        ctx.silu, ctx.wshape, ctx.wdtype = silu, weight.shape, weight.dtype
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        B, Ln, D = x.shape
        W = w.shape[-1]
        dy = dy.contiguous().to(x.dtype)
        dx = torch.empty_like(x)
        dw = torch.zeros(D, W, dtype=torch.float32, device=x.device)
        rc = L.lib().lina_short_conv_bwd(L.ptr(x), L.ptr(w), L.ptr(dy), L.ptr(dx), L.ptr(dw), B, Ln, D, W,
                                         int(ctx.silu), L.dt(x), L.stream(x))
        L.count_launches(1)
        L.check(rc, "lina_short_conv_bwd")
        return dx, dw.view(ctx.wshape).to(ctx.wdtype), None, None


class ShortConvolution(nn.Conv1d):
    """Depthwise causal conv (+SiLU) over ``[B, L, D]``; weight ``[D, 1, W]`` (state-dict key
    ``...conv1d.weight``), no bias by default -- FLA/fla/modules/convolution.py:84-139.  Lina never sets ``bias=True``
    (model/gla.py:106-108); that form runs the kernels without activation and adds bias + SiLU with torch ops."""

    def __init__(self, hidden_size: int, kernel_size: int, bias: bool = False,
                 activation: Optional[str] = "silu", use_fast_conv1d: Optional[bool] = True):
        super().__init__(in_channels=hidden_size, out_channels=hidden_size, kernel_size=kernel_size,
                         groups=hidden_size, bias=bias, padding=kernel_size - 1)
        self.hidden_size = hidden_size
        self.activation = None
        if activation is not None:
            assert activation in ["silu", "swish"], f"Activation `{activation}` not supported yet."
            self.activation = activation
        self.use_fast_conv1d = use_fast_conv1d

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None,
                cache: Optional[torch.Tensor] = None) -> torch.Tensor:
        """convolution.py:141-178: x ``[B, L, D]``; ``cache`` ``[B, D, W]`` is updated in place."""
        if mask is not None:
            x = x.mul_(mask.unsqueeze(-1))
        if cache is not None and x.shape[1] == 1:
            return self.step(x, cache)
        if self.bias is not None:
            return self._bias_act(_ShortConvFn.apply(x, self.weight, cache, False))
        return _ShortConvFn.apply(x, self.weight, cache, self.activation is not None)

    def _bias_act(self, y: torch.Tensor) -> torch.Tensor:
        y = y + self.bias.to(y.dtype)
        return torch.nn.functional.silu(y) if self.activation is not None else y

    def step(self, x: torch.Tensor, cache: torch.Tensor):
        """convolution.py:180-205 (one token, rolls ``cache``)."""
        assert x.shape[1] == 1, "Only support decoding with 1 token at a time for now"
        L.require_cuda(x, cache)
        xs = x.squeeze(1).contiguous()
        B, D = xs.shape
        W = self.kernel_size[0]
        w = self.weight.to(xs.dtype).contiguous()
        y = torch.empty_like(xs)
        if not cache.is_contiguous():
            raise ValueError("conv cache must be contiguous [B, D, W]")
        fused_act = self.activation is not None and self.bias is None
        rc = L.lib().lina_short_conv_update(L.ptr(xs), L.ptr(cache), L.dt(cache), L.ptr(w), L.ptr(y), B, D, W,
                                            int(fused_act), L.dt(xs), L.stream(xs))
        L.count_launches(1)
        L.check(rc, "lina_short_conv_update")
        if self.bias is not None:
            y = self._bias_act(y)
        return y.unsqueeze(1)

    @property
    def state_size(self) -> int:
        return self.hidden_size * self.kernel_size[0]


# ------------------------------------------------------------------------------------------------
class _NormGateFn(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, g, weight, eps):
        L.require_cuda(x, g, weight)
        shape = x.shape
        N = shape[-1]
        x2 = x.reshape(-1, N).contiguous()
        g2 = g.reshape(-1, N).to(x2.dtype).contiguous()
        w = weight.to(x2.dtype).contiguous() if weight is not None else None
        M = x2.shape[0]
        y = torch.empty_like(x2)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
        rc = L.lib().lina_rmsnorm_swishgate_fwd(L.ptr(x2), L.ptr(g2), L.ptr(w), L.ptr(y), L.ptr(rstd), M, N,
                                                float(eps), L.dt(x2), L.stream(x2))
        L.count_launches(1)
        L.check(rc, "lina_rmsnorm_swishgate_fwd")
        ctx.save_for_backward(x2, g2, w, rstd)
        ctx.shape, ctx.gdtype = shape, g.dtype
        ctx.wdtype = weight.dtype if weight is not None else None
        return y.view(shape)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        x2, g2, w, rstd = ctx.saved_tensors
        M, N = x2.shape
        dy2 = dy.reshape(M, N).to(x2.dtype).contiguous()
        dx, dg = torch.empty_like(x2), torch.empty_like(g2)
        dw = torch.zeros(N, dtype=torch.float32, device=x2.device) if w is not None else None
        rc = L.lib().lina_rmsnorm_swishgate_bwd(L.ptr(x2), L.ptr(g2), L.ptr(w), L.ptr(rstd), L.ptr(dy2), L.ptr(dx),
                                                L.ptr(dg), L.ptr(dw), M, N, L.dt(x2), L.stream(x2))
        L.count_launches(1)
        L.check(rc, "lina_rmsnorm_swishgate_bwd")
        return (dx.view(ctx.shape), dg.view(ctx.shape).to(ctx.gdtype),
                dw.to(ctx.wdtype) if dw is not None else None, None)


class FusedRMSNormSwishGate(nn.Module):
    """``y = RMSNorm(x) * weight * o * sigmoid(o)`` over the last dim --
    FLA/fla/modules/fused_norm_gate.py:765-803 (parameter ``weight[hidden_size]``, no bias)."""

    def __init__(self, hidden_size, elementwise_affine: bool = True, eps=1e-5):
        super().__init__()
        self.hidden_size = hidden_size
        self.elementwise_affine = elementwise_affine
        self.eps = eps
        if elementwise_affine:
            self.weight = nn.Parameter(torch.ones(hidden_size))
        else:
            self.register_parameter("weight", None)
        self.register_parameter("bias", None)

    def __repr__(self) -> str:
        s = f"{self.__class__.__name__}({self.hidden_size}"
        if not self.elementwise_affine:
            s += f", elementwise_affine={self.elementwise_affine}"
        return s + f", eps={self.eps})"

    def forward(self, x, o, residual=None, prenorm=False, residual_in_fp32=False):
        """Lina calls ``forward(x, o)`` only (model/gla.py:218).  The residual / prenorm options follow
        fused_norm_gate.py:100-111,460-480: the sum ``x + residual`` is formed in fp32 and normalised as such, stored as
        ``residual_out`` in the residual's dtype (fp32 with ``residual_in_fp32`` and no residual), ``y`` keeps x's dtype;
        they are torch ops around the same kernel."""
        if residual is None and not prenorm:
            return _NormGateFn.apply(x, o, self.weight, self.eps)
        if residual is not None:
            assert residual.shape == x.shape
            xs = x.float() + residual.float()
            residual_out = xs.to(residual.dtype)
        else:
            xs = x.float() if residual_in_fp32 else x
            residual_out = xs
        y = _NormGateFn.apply(xs, o, self.weight, self.eps).to(x.dtype)
        return (y, residual_out) if prenorm else y


# ------------------------------------------------------------------------------------------------
class Cache:
    """Per-layer recurrent state store -- FLA/fla/models/utils.py:11-107.

    ``states[layer_idx]`` is the tuple GatedLinearAttention.init_state builds
    (model/gla.py:229-240): (conv_q, conv_k, conv_v, S).  ``update`` copies in place
    (:61-66), so the tensors handed out by ``__getitem__`` stay valid across steps --
    which is what lets the B200 step kernel update them in place and skip the copy."""

    def __init__(self, seen_tokens: int = 0):
        self.states: List[Tuple[torch.Tensor, ...]] = []
        self._seen_tokens = seen_tokens

    def __getitem__(self, layer_idx: int):
        if layer_idx < len(self):
            return self.states[layer_idx]
        raise KeyError(f"Cache only has {len(self)} layers, attempted to access layer with index {layer_idx}")

    def __iter__(self):
        yield from self.states

    def __len__(self):
        return len(self.states)

    def update(self, state, layer_idx: int, offset: Optional[int] = 1,
               cache_kwargs: Optional[Dict[str, Any]] = None):
        if isinstance(state, torch.Tensor):
            state = (state,)
        if len(self.states) <= layer_idx:
            self.states.append(state)
        else:
            for i, s in enumerate(state):
                if s.data_ptr() != self.states[layer_idx][i].data_ptr():   # already updated in place
                    self.states[layer_idx][i].copy_(s)
            if layer_idx == len(self) - 1:
                self._seen_tokens += offset
        return state

    def get_seq_length(self, layer_idx: Optional[int] = 0) -> int:
        return 0 if len(self.states) <= layer_idx else self._seen_tokens

    def get_max_length(self) -> Optional[int]:
        return None

    def reorder_cache(self, beam_idx: torch.LongTensor):
        """models/utils.py:86-90 (beam search): batch rows re-selected; works on the per-layer tuples Lina stores (the
        reference's version assumes one tensor per layer)."""
        for layer_idx, st in enumerate(self.states):
            self.states[layer_idx] = tuple(t.index_select(0, beam_idx.to(t.device)) for t in st)

    def to_legacy_cache(self):
        return tuple(self.states)

    @classmethod
    def from_legacy_cache(cls, past_key_values=None, seen_tokens: int = 0) -> "Cache":
        cache = cls(seen_tokens)
        if past_key_values is not None:
            for layer_idx in range(len(past_key_values)):
                cache.update(past_key_values[layer_idx], layer_idx)
        return cache

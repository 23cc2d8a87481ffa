"""MixingBlock / SwiGLU / SelfAttention (model/base_blocks.py:9-69), same parameter names."""
from typing import Callable

import torch
import torch.nn as nn
import torch.nn.functional as F


import os

AUTOCAST_LN = os.environ.get("LINA_AUTOCAST_LN", "1") != "0"
# LINA_SKINNY_STEP=1: the single-token step (batch <= 32, bf16) runs its linears through lina_skinny_linear with the residual add +
# LayerNorm as prologue and SwiGLU's activation as epilogue (8 launches per block instead of 11).  Measured on B200 (decode
# loop, CUDA graph, profiles/README.md): 1.10 ms per step at batch 32 against 0.88 ms with library GEMMs + separate passes, equal
# at batch 8 -- the per-CTA prologue (every CTA normalises all rows) costs more than the three launches it saves.  Default off.
SKINNY_STEP = os.environ.get("LINA_SKINNY_STEP", "0") == "1"


class AsyncBound:
    """A scalar bound derived from parameters, refreshed WITHOUT stalling the stream.  The first value is read synchronously;
    afterwards, when the parameters' versions change (every optimizer step in training), the new value is computed on the
    device, copied to pinned host memory without blocking, and adopted once its event has completed -- until then the last
    known value answers.  Consumers apply a 2x safety margin for that one-step staleness."""

    def __init__(self):
        self.key, self.value, self.pending = None, None, None

    def get(self, key, compute):
        """``compute()`` -> 0-dim device tensor (or python float on CPU)."""
        if self.pending is not None and self.pending[2].query():
            self.value, self.pending = float(self.pending[1]), None
        if key != self.key:
            self.key = key
            with torch.no_grad():
                v = compute()
            if self.value is None or not torch.is_tensor(v) or not v.is_cuda:
                self.value, self.pending = float(v), None
            elif self.pending is None:
                host = torch.empty((), dtype=torch.float32, pin_memory=True)
                host.copy_(v.float(), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(v.device))
                self.pending = (key, host, ev)
        return self.value


def _ver(t) -> int:
    try:
        return t._version
    except RuntimeError:      # inference tensors do not track versions
        return -1


class SelfAttention(nn.Module):
    """model/base_blocks.py:9-40.  Bidirectional text-encoder attention (runs once per utterance,
    off the hot path).  ``rotary=True`` needs rotary-embedding-torch, exactly as in the reference."""

    def __init__(self, dim, heads, rotary=True, is_causal=False):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim)
        assert dim % heads == 0
        self.heads = heads
        self.rotary = None
        if rotary:
            from rotary_embedding_torch import RotaryEmbedding  # optional dependency of the reference too
            self.rotary = RotaryEmbedding((dim // heads) // 2)
        self.is_causal = is_causal

    def forward(self, x, mask=None, pos=None, cache=None, layer_idx=None, time_step=0):
        B, n, d = x.shape
        q, k, v = (t.view(B, n, self.heads, -1).transpose(1, 2) for t in self.qkv(x).chunk(3, dim=-1))
        if cache is not None:
            assert layer_idx is not None
            cache.update(k, v, layer_idx=layer_idx)
            k, v = cache[layer_idx]
        if self.rotary is not None:
            if pos is not None:
                from rotary_embedding_torch import apply_rotary_emb
                q, k = (apply_rotary_emb(self.rotary(pos).unsqueeze(1), t) for t in (q, k))
            else:
                q = self.rotary.rotate_queries_or_keys(q, offset=time_step)
                k = self.rotary.rotate_queries_or_keys(k)
        y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, is_causal=self.is_causal)
        return y.transpose(1, 2).reshape(B, n, d)


# hidden size of the inference-path SwiGLU GEMMs is zero-padded to a multiple of this: 1365 -> 1408 = 11 x 128, so N = 2816 = 11 x 256
# tiles for the library GEMM (with 1368 it picks a 240 x 192 kernel at 1150 TF/s; A/B in one process, profiles/step_breakdown.py:
# 36.1 / 36.3 ms per step with 8, 35.6 / 35.5 with 128 / 64).  LINA_SWIGLU_PAD=8 restores alignment-only padding.
SWIGLU_PAD = int(os.environ.get("LINA_SWIGLU_PAD", "128"))


class SwiGLU(nn.Module):
    """model/base_blocks.py:42-50: hidden = 4d//3, biases on both linears.

    Inference fast path (CUDA, no grad): the hidden size 4d//3 (1365 at d=1024) is odd, which sends both GEMMs
    down cuBLAS' unaligned legacy kernels; the weights are zero-padded ONCE to a multiple of 8 (mathematically a
    no-op: the padded gate columns are silu(0) * u = 0) and silu(gate) * u runs as one fused pass
    (lina_swiglu_act)."""

    def __init__(self, d_model):
        super().__init__()
        self.p_in = nn.Linear(d_model, (d_model * 4 // 3) * 2)
        self.p_out = nn.Linear(d_model * 4 // 3, d_model)

    _padded = None       # class-level default (un-pickled reference instances never ran this __init__)

    def _padded_weights(self):
        ps = (self.p_in.weight, self.p_in.bias, self.p_out.weight, self.p_out.bias)
        key = tuple((t.data_ptr(), _ver(t), t.dtype) for t in ps)
        if self._padded is None or self._padded[0] != key:
            hid = self.p_out.in_features
            hp = (hid + SWIGLU_PAD - 1) // SWIGLU_PAD * SWIGLU_PAD
            wi, bi, wo = ps[0].detach(), ps[1].detach(), ps[2].detach()
            wi_p = wi.new_zeros(2 * hp, wi.shape[1]); bi_p = bi.new_zeros(2 * hp)
            wi_p[:hid], wi_p[hp:hp + hid] = wi[:hid], wi[hid:]
            bi_p[:hid], bi_p[hp:hp + hid] = bi[:hid], bi[hid:]
            wo_p = wo.new_zeros(wo.shape[0], hp); wo_p[:, :hid] = wo
            self._padded = (key, hp, wi_p, bi_p, wo_p)
        return self._padded[1:]

    def forward(self, x):
        if (x.is_cuda and not torch.is_grad_enabled() and not torch.is_autocast_enabled()
                and x.dtype == self.p_in.weight.dtype and x.dtype in (torch.float32, torch.bfloat16, torch.float16)):
            from .. import _lib as L
            hp, wi_p, bi_p, wo_p = self._padded_weights()
            h = F.linear(x, wi_p, bi_p)
            h2 = h.reshape(-1, 2 * hp)
            a = torch.empty(h2.shape[0], hp, dtype=h.dtype, device=h.device)
            rc = L.lib().lina_swiglu_act(L.ptr(h2), L.ptr(a), h2.shape[0], hp, L.dt(h2), L.stream(h2))
            L.count_launches(1)
            L.check(rc, "lina_swiglu_act")
            return F.linear(a.view(*x.shape[:-1], hp), wo_p, self.p_out.bias)
        hid = self.p_out.in_features
        if x.is_cuda and hid % SWIGLU_PAD:
            # training / autograd path: the same zero padding, applied differentiably on the fly -- with hidden = 1365 the
            # unpadded GEMMs (N = 2730, K = 1365) run on cuBLAS' unaligned sm75/sm80 kernels, ~5x slower (56 ms of a
            # 300 ms training step at bs8 x seq4096)
            hp = (hid + SWIGLU_PAD - 1) // SWIGLU_PAD * SWIGLU_PAD
            padn = hp - hid
            wi, bi = self.p_in.weight, self.p_in.bias
            wi_p = torch.cat([F.pad(wi[:hid], (0, 0, 0, padn)), F.pad(wi[hid:], (0, 0, 0, padn))], dim=0)
            bi_p = torch.cat([F.pad(bi[:hid], (0, padn)), F.pad(bi[hid:], (0, padn))], dim=0)
            gate, u = F.linear(x, wi_p, bi_p).chunk(2, dim=-1)
            return F.linear(F.silu(gate) * u, F.pad(self.p_out.weight, (0, padn)), self.p_out.bias)
        gate, u = self.p_in(x).chunk(2, dim=-1)
        return self.p_out(F.silu(gate) * u)


class _AutocastLayerNorm(torch.autograd.Function):
    """nn.LayerNorm on the fp32 residual stream with the output written directly in the autocast dtype
    (lina_layernorm_f32in_fwd / _bwd).  Under autocast torch's LayerNorm returns fp32 and each consuming Linear casts it again;
    the values that reach the GEMMs are the same, the traffic is 6 instead of up to 34 bytes per element."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, x, weight, bias, eps, out_dtype):
        from .. import _lib as L
        N = x.shape[-1]
        x2 = x.reshape(-1, N).contiguous()
        M = x2.shape[0]
        w, b = weight.contiguous(), bias.contiguous()
        y = torch.empty(M, N, dtype=out_dtype, device=x.device)
        stats = torch.empty(2, M, dtype=torch.float32, device=x.device)
        rc = L.lib().lina_layernorm_f32in_fwd(L.ptr(x2), L.ptr(w), L.ptr(b), L.ptr(y), L.ptr(stats[0]), L.ptr(stats[1]), M, N,
                                              float(eps), L.dt(y), L.stream(x2))
        L.count_launches(1)
        L.check(rc, "lina_layernorm_f32in_fwd")
        ctx.save_for_backward(x2, w, stats)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        from .. import _lib as L
        x2, w, stats = ctx.saved_tensors
        M, N = x2.shape
        dy2 = dy.reshape(M, N).contiguous()
        dx = torch.empty_like(x2)
        dgb = torch.zeros(2, N, dtype=torch.float32, device=x2.device)
        rc = L.lib().lina_layernorm_f32in_bwd(L.ptr(x2), L.ptr(w), L.ptr(stats[0]), L.ptr(stats[1]), L.ptr(dy2), L.ptr(dx),
                                              L.ptr(dgb[0]), L.ptr(dgb[1]), M, N, L.dt(dy2), L.stream(x2))
        L.count_launches(1)
        L.check(rc, "lina_layernorm_f32in_bwd")
        return dx.view(ctx.shape), dgb[0], dgb[1], None, None


def autocast_layernorm(x, norm: nn.LayerNorm):
    """``norm(x)``; on the bf16 / fp16 autocast training path (fp32 stream, fp32 affine parameters) through the fused
    fp32-in kernel pair, otherwise plain ``norm(x)``."""
    if (AUTOCAST_LN and x.is_cuda and x.dtype == torch.float32 and torch.is_autocast_enabled()
            and isinstance(norm, nn.LayerNorm) and norm.elementwise_affine and norm.bias is not None
            and norm.weight.dtype == torch.float32 and x.shape[-1] % 4 == 0 and x.shape[-1] <= 1024
            and len(norm.normalized_shape) == 1):
        dt = torch.get_autocast_dtype("cuda")
        if dt in (torch.bfloat16, torch.float16):
            return _AutocastLayerNorm.apply(x, norm.weight, norm.bias, norm.eps, dt)
    return norm(x)


def unpack_ignore(x):
    """model/base_blocks.py:53-54: token mixers may return (output, extras); keep the output."""
    return x[0] if type(x) is tuple else x


class MixingBlock(nn.Module):
    """Pre-LN residual wrapper (model/base_blocks.py:56-69): tmix then cmix."""

    def __init__(self, tmix: Callable, cmix: Callable, norm: Callable, dropout: float = 0.0):
        super().__init__()
        self.tmix = tmix()
        self.cmix = cmix()
        self.norm1 = norm()
        self.norm2 = norm()
        self.drop = nn.Dropout(dropout)

    def forward(self, x, **kwargs):
        if hasattr(self.tmix, "gates_certified") and isinstance(self.norm1, nn.LayerNorm) and self.norm1.elementwise_affine \
                and self.norm1.bias is not None:
            kwargs = dict(kwargs, input_norm_bound=self._ln_output_norm_bound())
        t = self.tmix(autocast_layernorm(x, self.norm1), **kwargs)
        x = (t[0] if type(t) is tuple else t) + x
        x = self.cmix(autocast_layernorm(x, self.norm2)) + x
        return self.drop(x)

    _ln_bound = None

    def _ln_output_norm_bound(self) -> float:
        """||LayerNorm(.)||_2 <= sqrt(d) max|gamma| + ||beta||_2 (see AsyncBound for how it follows the parameters)."""
        n = self.norm1
        if self._ln_bound is None:
            self._ln_bound = AsyncBound()
        key = (n.weight.data_ptr(), _ver(n.weight), n.bias.data_ptr(), _ver(n.bias))
        d = n.weight.numel()
        return self._ln_bound.get(key, lambda: n.weight.detach().float().abs().max() * d ** 0.5 + n.bias.detach().float().norm())

    # -- inference fast path: residual adds fused into the following LayerNorm -------------------------------
    def can_fuse(self, x) -> bool:
        n = self.norm1
        return (not torch.is_grad_enabled() and not self.training and not torch.is_autocast_enabled()
                and x.is_cuda and isinstance(n, nn.LayerNorm)
                and n.elementwise_affine and n.bias is not None and x.dtype == n.weight.dtype
                and x.dtype in (torch.float32, torch.bfloat16, torch.float16)
                and x.shape[-1] % (16 // x.element_size()) == 0 and x.shape[-1] // (16 // x.element_size()) <= 256)

    # -- single-token step, batch <= 32, bf16: weight-streaming linears with the LayerNorms / SwiGLU fused in ---------------
    def _can_skinny(self, x, kwargs) -> bool:
        if not SKINNY_STEP or x.dim() != 3 or x.shape[1] != 1 or x.dtype != torch.bfloat16:
            return False
        t, c = self.tmix, self.cmix
        cache = kwargs.get("past_key_values")
        if not (isinstance(c, SwiGLU) and hasattr(t, "_step_core") and kwargs.get("use_cache") and cache is not None):
            return False
        from .. import _lib as L
        if x.shape[0] > L.lib().lina_skinny_linear_max_rows() or t.value_dim > 2048 or c.p_out.in_features > 2048:
            return False
        state = cache[t.layer_idx] if len(cache.states) > t.layer_idx else None
        return (state is not None and t._can_step(x, state, None, None) and self.norm2.weight.dtype == x.dtype
                and c.p_in.weight.dtype == x.dtype and t.o_proj.weight.dtype == x.dtype)

    def _forward_skinny(self, x, delta, **kwargs):
        """forward_fused for one token: 5 weight-streaming launches + the 3 step kernels per block (lina_skinny_linear)."""
        from .. import _lib as L
        t, c = self.tmix, self.cmix
        B, _, d = x.shape
        cache = kwargs["past_key_values"]
        state = cache[t.layer_idx]
        x2d = x.reshape(B, d)
        d2d = delta.reshape(B, d) if delta is not None else None
        new = lambda n: torch.empty(B, n, dtype=x.dtype, device=x.device)

        def run(xin, dl, norm, W, bias, N, K, pair=0):
            out = new(N)
            s_out = new(xin.shape[1]) if dl is not None else None
            rc = L.lib().lina_skinny_linear(L.ptr(xin), xin.stride(0), L.ptr(dl), dl.stride(0) if dl is not None else 0,
                                            L.ptr(norm.weight) if norm is not None else None,
                                            L.ptr(norm.bias) if norm is not None else None, float(norm.eps) if norm is not None else 0.0,
                                            L.ptr(s_out), L.ptr(W), W.stride(0), L.ptr(bias), L.ptr(out), out.stride(0), B, N, K, pair,
                                            L.stream(xin))
            L.count_launches(1)
            L.check(rc, "lina_skinny_linear")
            return out, (s_out if s_out is not None else xin)

        wcat = t._cat_weight()
        proj, x1 = run(x2d, d2d, self.norm1, wcat, None, wcat.shape[0], d)
        o = t._step_core(proj, state)
        cache.update(state, t.layer_idx, 1)                       # in place already: bumps seen_tokens only
        tt, _ = run(o, None, None, t.o_proj.weight, t.o_proj.bias, d, o.shape[1])
        hp, wi_p, bi_p, wo_p = c._padded_weights()
        a, x2 = run(x1, tt, self.norm2, wi_p, bi_p, hp, d, pair=hp)
        cc, _ = run(a, None, None, wo_p, c.p_out.bias, d, hp)
        return x2.view(B, 1, d), cc.view(B, 1, d)

    def forward_fused(self, x, delta=None, **kwargs):
        """Same arithmetic as ``forward`` for the block whose input is ``x + delta`` (``delta`` None = just ``x``),
        returned as the pair (residual, branch) with block output = residual + branch, so that the caller can hand
        the pending add to the next block's first LayerNorm (lina_add_layernorm: one pass instead of add + LN)."""
        if self._can_skinny(x, kwargs):
            return self._forward_skinny(x, delta, **kwargs)
        x1, h = add_layernorm(delta, x, self.norm1)
        if hasattr(self.tmix, "gates_certified"):          # ||LayerNorm(.)||_2 <= sqrt(d) max|gamma| + ||beta||_2
            kwargs = dict(kwargs, input_norm_bound=self._ln_output_norm_bound())
        t = self.tmix(h, **kwargs)
        t = t[0] if type(t) is tuple else t
        x2, h2 = add_layernorm(t, x1, self.norm2)
        return x2, self.cmix(h2)


def add_layernorm(a, x, norm: nn.LayerNorm):
    """(a + x, LayerNorm(a + x)); a None -> (x, LayerNorm(x)).  CUDA, no grad."""
    from .. import _lib as L
    xs = x.contiguous()
    N = xs.shape[-1]
    M = xs.numel() // N
    ln = torch.empty_like(xs)
    if norm.weight.dtype != xs.dtype:
        raise TypeError(f"add_layernorm: LayerNorm parameters are {norm.weight.dtype}, input is {xs.dtype}")
    if a is not None:
        a = a.to(xs.dtype).contiguous()          # one dtype argument describes every buffer the kernel reads
        s = torch.empty_like(xs)
    else:
        s = None
    rc = L.lib().lina_add_layernorm(L.ptr(a), L.ptr(xs), L.ptr(norm.weight), L.ptr(norm.bias), L.ptr(s), L.ptr(ln), M, N,
                                    float(norm.eps), L.dt(xs), L.stream(xs))
    L.count_launches(1)
    L.check(rc, "lina_add_layernorm")
    return (s if s is not None else xs), ln

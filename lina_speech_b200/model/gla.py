"""GLA token mixer and the AttentiveGLA backbone -- host side of the B200 hot path.

Mirrors the reference's ``model/gla.py`` (class names, constructor arguments, parameter names,
``forward / init_state / step / to_mode / get_init_state_tuning_params / get_state_from_params``)
so checkpoints and callers (LinaModel, train_initial_state, the notebook) carry over unchanged.
The arithmetic between the dense projections runs in liblina_b200.so:

  * multi-token calls: ShortConvolution -> fused_recurrent_gla / fused_chunk_gla / chunk_gla ->
    FusedRMSNormSwishGate, i.e. the reference's own sequence (model/gla.py:146-225) on our ops;
  * single-token calls with a cache in eval mode: ONE fused step (``lina_gla_step``) that rolls the
    three conv states, applies the gate non-linearity, updates the recurrent state in place and
    applies the norm-gate -- replacing ~12 launches and 4 state copies per layer per token.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from einops import einsum, rearrange, repeat

from .. import _lib as L
from ..fla_api import ops as fla_ops
from ..fla_api import (Cache, FusedRMSNormSwishGate, ShortConvolution, chunk_gla, fused_chunk_gla,
                       fused_recurrent_gla)
from .contracts import AttentiveRNN
from .base_blocks import MixingBlock, SwiGLU
from .crossatt import BlindCrossAttention, CrossAttention, tensor_version

# bring-up switches (A/B on the GPU box): LINA_FUSED_PREFILL=0 restores the op-by-op inference path,
# LINA_CAT5=1 folds the rank-16 gate projection into the concatenated GEMM as well (N = 6160)
FUSED_PREFILL = os.environ.get("LINA_FUSED_PREFILL", "1") != "0"
CAT5 = os.environ.get("LINA_CAT5", "0") == "1"
# LINA_PREGATED=0: the post-projection pass writes q, k, gk and the GLA kernel gates them itself (4x redundantly)
PREGATED = os.environ.get("LINA_PREGATED", "1") != "0"
# "cat4": one [q;k;v;g] GEMM; "split": four GEMMs (the pre-gated pass and the norm-gate read each output in place)
GEMM_GROUPING = os.environ.get("LINA_GEMM_GROUPING", "split")
# LINA_LOWRANK_KERNEL=0 sends gk_proj[1] of the whole-sequence path back through F.linear (A/B)
LOWRANK_KERNEL = os.environ.get("LINA_LOWRANK_KERNEL", "1") != "0"



class GateEnvelope:
    """Device-side record of "a chunk's summed log gate left the range of the single-pivot tensor-core kernel" (written by
    lina_gla_prefill_prep_gated) and the policy for reading it.  A lone GatedLinearAttention call reads the flag right away
    (one host read) and re-serves the call exactly; inside a backbone pass (``GateEnvelope.deferred``) the 13 mixers only
    accumulate into the flag and the backbone reads it ONCE at the end -- no per-layer host synchronisation."""
    _flags = {}
    depth = 0
    flagged_calls = 0        # mixer calls since the last read that relied on the device flag (no static certificate)

    @classmethod
    def flag(cls, device) -> torch.Tensor:
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        f = cls._flags.get(key)
        if f is None:
            with torch.inference_mode(False):
                f = cls._flags[key] = torch.zeros(1, dtype=torch.int32, device=device)
        return f

    @classmethod
    def tripped(cls, device) -> bool:
        """Read and clear (host synchronisation).  Never called during graph capture."""
        cls.flagged_calls = 0
        f = cls.flag(device)
        hit = bool(f.item())
        if hit:
            f.zero_()
        return hit

    class deferred:
        def __enter__(self):
            GateEnvelope.depth += 1
            return self

        def __exit__(self, *exc):
            GateEnvelope.depth -= 1
            return False


if "GRAD_CKPT" in os.environ:        # model/gla.py:26-33
    def maybe_grad_ckpt(f):
        def wrapped(*args, **kwargs):
            return torch.utils.checkpoint.checkpoint(f, *args, **kwargs, use_reentrant=False)
        return wrapped
else:
    def maybe_grad_ckpt(f):
        return f


class GatedLinearAttention(nn.Module):
    """model/gla.py:44-247."""

    def __init__(self, mode: str = "fused_chunk", hidden_size: int = 1024, expand_k: float = 1.0,
                 expand_v: float = 2.0, num_heads: int = 4, use_short_conv: bool = False, conv_size: int = 4,
                 conv_bias: bool = False, share_conv_kernel: bool = False, gate_fn: str = "swish",
                 layernorm_eps: float = 1e-5, gate_logit_normalizer: int = 16, gate_low_rank_dim: int = 16,
                 clamp_min: Optional[float] = None, fuse_norm: bool = True, layer_idx: int = None, **kwargs):
        super().__init__()
        assert mode in ["chunk", "fused_recurrent", "fused_chunk"], f"Not supported mode `{mode}`."
        if share_conv_kernel or conv_bias or gate_fn != "swish" or not fuse_norm:
            raise NotImplementedError("only the configuration AttentiveGLA builds (model/gla.py:270-274) is "
                                      "implemented: separate q/k/v short convs without bias, fused swish norm-gate")
        self.mode = mode
        self.hidden_size, self.expand_k, self.expand_v, self.num_heads = hidden_size, expand_k, expand_v, num_heads
        self.use_short_conv, self.conv_size, self.conv_bias = use_short_conv, conv_size, conv_bias
        self.share_conv_kernel = share_conv_kernel
        self.key_dim, self.value_dim = int(hidden_size * expand_k), int(hidden_size * expand_v)
        self.clamp_min, self.layer_idx = clamp_min, layer_idx
        assert self.key_dim % num_heads == 0 and self.value_dim % num_heads == 0
        self.head_qk_dim, self.head_v_dim = self.key_dim // num_heads, self.value_dim // num_heads

        self.q_proj = nn.Linear(hidden_size, self.key_dim, bias=False)
        self.k_proj = nn.Linear(hidden_size, self.key_dim, bias=False)
        self.v_proj = nn.Linear(hidden_size, self.value_dim, bias=False)
        self.g_proj = nn.Linear(hidden_size, self.value_dim, bias=False)
        self.gk_proj = nn.Sequential(nn.Linear(hidden_size, gate_low_rank_dim, bias=False),
                                     nn.Linear(gate_low_rank_dim, self.key_dim, bias=True))
        self.o_proj = nn.Linear(self.value_dim, hidden_size, bias=False)
        if use_short_conv:
            self.q_conv1d = ShortConvolution(self.key_dim, conv_size, activation="silu")
            self.k_conv1d = ShortConvolution(self.key_dim, conv_size, activation="silu")
            self.v_conv1d = ShortConvolution(self.value_dim, conv_size, activation="silu")
        self.g_norm_swish_gate = FusedRMSNormSwishGate(self.head_v_dim, eps=layernorm_eps)
        self.fuse_norm_and_gate = True
        self.gate_logit_normalizer = gate_logit_normalizer
        self.apply(self._initialize_weights)

    def _initialize_weights(self, module: nn.Module):       # model/gla.py:122-129
        if getattr(module, "_is_hf_initialized", False):
            return
        if isinstance(module, nn.Linear):
            nn.init.xavier_uniform_(module.weight, gain=2 ** -2.5)
            if module.bias is not None:
                nn.init.zeros_(module.bias)
        module._is_hf_initialized = True

    # lazily built weight caches; class-level defaults so that instances un-pickled from a reference checkpoint
    # (TrainLina.load_from_checkpoint restores __dict__ without running __init__) have them too
    _wcat = None
    _wcat4 = None

    _gate_cert = None

    def gate_preactivation_bound(self, input_norm_bound: float) -> float:
        """Upper bound of |gk_proj(x)| over every channel for any input with ||x||_2 <= input_norm_bound:
        max_c ( ||(W2 W1)_c||_2 * bound + |b_c| ), following the weights through AsyncBound (one small GEMM when they change).  With
        gate = logsigmoid(.) / normalizer a chunk of 64 tokens sums to at least 64 * logsigmoid(-bound) / normalizer, so a
        bound below ~20 (normalizer 16) CERTIFIES that the tensor-core kernels' single-pivot range (-80) cannot be left and
        no device-side check (and no host synchronisation) is needed for this layer."""
        from .base_blocks import AsyncBound
        w1, w2, b2 = self.gk_proj[0].weight, self.gk_proj[1].weight, self.gk_proj[1].bias
        key = tuple((t.data_ptr(), tensor_version(t), t.dtype) for t in (w1, w2) + ((b2,) if b2 is not None else ()))
        if self._gate_cert is None:
            self._gate_cert = (AsyncBound(), AsyncBound())
        rows = self._gate_cert[0].get(key, lambda: (w2.detach().float() @ w1.detach().float()).norm(dim=1).max())
        bias = self._gate_cert[1].get(key, lambda: b2.detach().float().abs().max() if b2 is not None else 0.0)
        return rows * input_norm_bound + bias

    def gates_certified(self, input_norm_bound) -> bool:
        if input_norm_bound is None:
            return False
        # 2x margin: in training the bounds may lag the parameters by one optimizer step (AsyncBound)
        x_min = -2.0 * self.gate_preactivation_bound(float(input_norm_bound))
        per_token = (x_min - math.log1p(math.exp(x_min))) if x_min > -30 else x_min      # logsigmoid
        return 64.0 * per_token / float(self.gate_logit_normalizer) >= fla_ops.GATE_SUM_LIMIT

    # -- single-token fast path --------------------------------------------------------------------
    def _cat_weight(self):
        """[q;k;v;g;gk0] projection weights stacked so a decode step needs one GEMM for them."""
        ws = (self.q_proj.weight, self.k_proj.weight, self.v_proj.weight, self.g_proj.weight, self.gk_proj[0].weight)
        key = tuple((w.data_ptr(), tensor_version(w), w.dtype) for w in ws)
        if self._wcat is None or self._wcat[0] != key:
            self._wcat = (key, torch.cat([w.detach() for w in ws], dim=0).contiguous())
        return self._wcat[1]

    def _step(self, x: torch.Tensor, state: Tuple[torch.Tensor, ...]) -> torch.Tensor:
        B = x.shape[0]
        proj = F.linear(x.view(B, -1), self._cat_weight())
        if proj.dtype != x.dtype:
            raise TypeError(f"GatedLinearAttention._step: projection is {proj.dtype}, input is {x.dtype}")
        return self.o_proj(self._step_core(proj, state)).view(B, 1, -1)

    def _step_core(self, proj: torch.Tensor, state: Tuple[torch.Tensor, ...]) -> torch.Tensor:
        """The step between the two GEMMs: proj [B, q;k;v;g;gk0] -> norm-gated o [B, value_dim]; states updated in place."""
        B = proj.shape[0]
        H, K, V, kd, vd = self.num_heads, self.head_qk_dim, self.head_v_dim, self.key_dim, self.value_dim
        xq, xk, xv, g, lo = torch.split(proj, [kd, kd, vd, vd, proj.shape[1] - 2 * kd - 2 * vd], dim=1)
        ldx = proj.shape[1]                            # the four slices are read in place with this row stride
        x = proj
        if self.use_short_conv:
            cq, ck, cv, S = state
            W = self.conv_size
            wq, wk, wv = (c.weight.to(x.dtype) for c in (self.q_conv1d, self.k_conv1d, self.v_conv1d))
        else:
            (S,) = state
            cq = ck = cv = wq = wk = wv = None
            W = 1
        if not S.is_contiguous() or (cq is not None and not (cq.is_contiguous() and ck.is_contiguous() and cv.is_contiguous())):
            raise ValueError("cache states must be contiguous")
        if cq is not None and not (cq.dtype == ck.dtype == cv.dtype == S.dtype):
            raise ValueError("conv states and recurrent state must share one dtype")
        lib = L.lib()
        out = torch.empty(B, vd, dtype=x.dtype, device=x.device)
        ws = torch.empty(int(lib.lina_gla_step_workspace_bytes(B, H, K, V)), dtype=torch.uint8, device=x.device)
        nw = self.g_norm_swish_gate.weight
        nw = nw.to(x.dtype) if nw is not None else None
        w2, b2 = self.gk_proj[1].weight, self.gk_proj[1].bias           # rank-R gate factors, applied inside the step kernel
        w2c = w2.to(x.dtype).contiguous()
        b2c = b2.to(x.dtype).contiguous() if b2 is not None else None
        rc = lib.lina_gla_step_lr(L.ptr(xq), L.ptr(xk), L.ptr(xv), L.ptr(lo), ldx, L.ptr(w2c),
                                  L.ptr(b2c), w2.shape[1], L.ptr(g), L.ptr(wq),
                                  L.ptr(wk), L.ptr(wv), L.ptr(cq), L.ptr(ck), L.ptr(cv), L.ptr(S), L.ptr(nw), L.ptr(out),
                                  L.ptr(ws), B, H, K, V, W, L.dt(x), L.dt(S), float(K) ** -0.5,
                                  float(self.gate_logit_normalizer), float(self.g_norm_swish_gate.eps), ldx, L.stream(x))
        L.count_launches(3)
        L.check(rc, "lina_gla_step_lr")
        return out

    # -- whole-sequence inference fast path ----------------------------------------------------------
    def _cat_weight4(self):
        """[q;k;v;g] stacked (N = 2*key_dim + 2*value_dim, a multiple of 256 at the shipped size): one GEMM reads
        the LayerNorm output once instead of four times."""
        ws = (self.q_proj.weight, self.k_proj.weight, self.v_proj.weight, self.g_proj.weight)
        key = tuple((w.data_ptr(), tensor_version(w), w.dtype) for w in ws)
        if getattr(self, "_wcat4", None) is None or self._wcat4[0] != key:
            self._wcat4 = (key, torch.cat([w.detach() for w in ws], dim=0).contiguous())
        return self._wcat4[1]

    def _can_prefill(self, x, reset_mask, attention_mask) -> bool:
        return (FUSED_PREFILL and self.use_short_conv and self.conv_size == 4 and not torch.is_grad_enabled()
                and not torch.is_autocast_enabled()      # raw-pointer path: every buffer must really be x.dtype
                and reset_mask is None and attention_mask is None
                and x.dtype in (torch.bfloat16, torch.float16, torch.float32)
                and self.key_dim % 8 == 0 and self.head_v_dim % 8 == 0
                and self.head_v_dim * x.element_size() // 16 <= 128
                and self.q_proj.weight.dtype == x.dtype)

    def _gate_expand(self, lo: torch.Tensor) -> torch.Tensor:
        """gk_proj[1] (model/gla.py:96-97) on a whole sequence: [B,T,R] (any row stride) -> [B,T,key_dim].  bf16 with R in
        {8,16,32} runs lina_lowrank_linear (the K = R library GEMM ran 10x off its write-bandwidth bound); else F.linear."""
        lin = self.gk_proj[1]
        R, N = lin.weight.shape[1], lin.weight.shape[0]
        if (LOWRANK_KERNEL and lo.dtype == torch.bfloat16 and lin.weight.dtype == torch.bfloat16 and R in (8, 16, 32) and N % 4 == 0
                and lo.dim() == 3 and lo.stride(2) == 1 and lo.stride(1) % 4 == 0 and lo.stride(0) == lo.shape[1] * lo.stride(1)
                and lo.data_ptr() % 8 == 0 and (lin.bias is None or lin.bias.dtype == torch.bfloat16)):
            Bn, Tn = lo.shape[0], lo.shape[1]
            out = torch.empty(Bn, Tn, N, dtype=lo.dtype, device=lo.device)
            w = lin.weight if lin.weight.is_contiguous() else lin.weight.contiguous()
            rc = L.lib().lina_lowrank_linear(L.ptr(lo), lo.stride(1), L.ptr(w), L.ptr(lin.bias), L.ptr(out), N, Bn * Tn, N, R,
                                             L.dt(lo), L.stream(lo))
            L.count_launches(1)
            L.check(rc, "lina_lowrank_linear")
            return out
        return F.linear(lo, lin.weight, lin.bias)

    def _prefill(self, x: torch.Tensor, last_state, use_cache: bool, past_key_values, input_norm_bound=None) -> torch.Tensor:
        """model/gla.py:146-225 for a whole sequence without autograd: one [q;k;v;g] GEMM, ONE pass for the three
        short convs + the gate non-linearity (lina_gla_prefill_prep), the GLA op on the [B,T,H,D] layout, the
        norm-gate reading g in place from the projection buffer, o_proj."""
        B, T, _ = x.shape
        H, K, V, kd, vd = self.num_heads, self.head_qk_dim, self.head_v_dim, self.key_dim, self.value_dim
        lib = L.lib()
        norm = float(self.gate_logit_normalizer)
        state_in = last_state[-1] if use_cache and last_state is not None else None
        pregated = (PREGATED and x.dtype == torch.bfloat16 and self.mode in ("fused_chunk", "chunk")
                    and self.clamp_min is None and norm > 0 and math.frexp(norm)[0] == 0.5
                    and bool(lib.lina_gla_chunk_fwd_uses_tensor_cores(B, H, T, K, V, L.BF16))
                    and (state_in is None or state_in.dtype in (torch.float32, torch.bfloat16, torch.float16)))
        if GEMM_GROUPING == "split" and not CAT5 and pregated:
            # four GEMMs, each output read in place with its own row stride (measured faster in-stream than one N=6144 GEMM)
            xq, xk, xv, g = self.q_proj(x), self.k_proj(x), self.v_proj(x), self.g_proj(x)
            lo = self.gk_proj[0](x)
        else:
            if CAT5:
                proj = F.linear(x, self._cat_weight())
                lo = proj[..., 2 * kd + 2 * vd:]
            else:
                proj = F.linear(x, self._cat_weight4())
                lo = self.gk_proj[0](x)
            xq, xk, xv, g = (proj[..., :kd], proj[..., kd:2 * kd], proj[..., 2 * kd:2 * kd + vd],
                             proj[..., 2 * kd + vd:2 * kd + 2 * vd])
        gk_raw = self._gate_expand(lo)
        for t_ in (xq, xk, xv, g, gk_raw):                   # raw-pointer path: one dtype argument describes them all
            if t_.dtype != x.dtype:
                raise TypeError(f"GatedLinearAttention._prefill: projection is {t_.dtype}, input is {x.dtype}")
        ldq, ldk, ldv, ldgate = xq.stride(1), xk.stride(1), xv.stride(1), g.stride(1)
        cq = ck = cv = None
        if use_cache and last_state is not None:
            cq, ck, cv = last_state[0], last_state[1], last_state[2]
            if not (cq.is_contiguous() and ck.is_contiguous() and cv.is_contiguous()):
                raise ValueError("conv caches must be contiguous [B, D, W]")
        wq, wk, wv = (c.weight.to(x.dtype).contiguous() for c in (self.q_conv1d, self.k_conv1d, self.v_conv1d))
        recurrent_state = last_state[-1] if use_cache else None
        q = torch.empty(B, T, kd, dtype=x.dtype, device=x.device)
        k = torch.empty_like(q)
        v = torch.empty(B, T, vd, dtype=x.dtype, device=x.device)
        if pregated:
            # q, k hold the gated MMA operands q~ = scale q e^G, k~ = k e^-G; gk is never materialised
            nt = (T + 63) // 64
            decay = torch.empty(B, H, nt, K, dtype=torch.float32, device=x.device)
            check = (fla_ops.GATE_CHECK and not torch.cuda.is_current_stream_capturing()
                     and not self.gates_certified(input_norm_bound))
            flag = GateEnvelope.flag(x.device) if check else None
            if check:
                GateEnvelope.flagged_calls += 1
            rc = lib.lina_gla_prefill_prep_gated(L.ptr(xq), ldq, L.ptr(xk), ldk, L.ptr(xv), ldv, L.ptr(wq), L.ptr(wk), L.ptr(wv),
                                                 L.ptr(gk_raw), gk_raw.stride(1), L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(decay),
                                                 L.ptr(cq), L.ptr(ck), L.ptr(cv), L.dt(cq) if cq is not None else 0,
                                                 B, T, H, K, V, self.conv_size, norm, float(K) ** -0.5, L.ptr(flag), L.stream(x))
            L.count_launches(2)
            L.check(rc, "lina_gla_prefill_prep_gated")
            if check and GateEnvelope.depth == 0 and GateEnvelope.tripped(x.device):
                pregated = False          # gates beyond e^-80 per chunk: serve this call with the exact kernels below
        if pregated:
            h0 = recurrent_state.contiguous() if recurrent_state is not None else None
            o = torch.empty(B, T, H, V, dtype=x.dtype, device=x.device)
            ht = torch.empty(B, H, K, V, dtype=torch.float32, device=x.device) if use_cache else None
            # hand-off space for tiles the kernel cuts in two along T (fills the last wave; 0 bytes when no cut applies)
            ws_bytes = lib.lina_gla_chunk_fwd_pregated_ws_bytes(B, H, T, K, V)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device) if ws_bytes else None
            prof = fla_ops.PROFILE
            if prof is not None:
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(torch.cuda.current_stream(x.device))
            rc = lib.lina_gla_chunk_fwd_pregated_bthd_ws(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(decay), L.ptr(h0),
                                                         L.dt(h0) if h0 is not None else 0, L.ptr(o), L.ptr(ht), L.ptr(ws),
                                                         ws_bytes, B, H, T, K, V, L.stream(x))
            L.count_launches(1)
            L.check(rc, "lina_gla_chunk_fwd_pregated_bthd_ws")
            if prof is not None:
                ev1.record(torch.cuda.current_stream(x.device))
                prof.append(("chunk_pregated", ev0, ev1))
            recurrent_state = ht
            if past_key_values is not None and not self.training:       # model/gla.py:205-213
                past_key_values.update((cq, ck, cv, recurrent_state), self.layer_idx, T)
        else:
            gk = torch.empty_like(q)
            if not (ldq == ldk == ldv):                   # this entry takes ONE row stride: pack the split GEMMs' outputs
                packed = torch.cat([xq, xk, xv], dim=-1)
                xq, xk, xv = packed[..., :kd], packed[..., kd:2 * kd], packed[..., 2 * kd:]
                ldq = packed.stride(1)
            ldx = ldq
            rc = lib.lina_gla_prefill_prep(L.ptr(xq), L.ptr(xk), L.ptr(xv), ldx, L.ptr(wq), L.ptr(wk), L.ptr(wv),
                                           L.ptr(gk_raw), gk_raw.stride(1), L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(gk),
                                           L.ptr(cq), L.ptr(ck), L.ptr(cv), L.dt(cq) if cq is not None else 0, B, T, kd, vd,
                                           self.conv_size, norm, float(self.clamp_min or 0.0),
                                           int(self.clamp_min is not None), L.dt(x), L.stream(x))
            L.count_launches(1)
            L.check(rc, "lina_gla_prefill_prep")
            q4, k4, gk4 = (t.view(B, T, H, K).transpose(1, 2) for t in (q, k, gk))
            v4 = v.view(B, T, H, V).transpose(1, 2)
            op = {"fused_recurrent": fused_recurrent_gla, "fused_chunk": fused_chunk_gla, "chunk": chunk_gla}[self.mode]
            o, recurrent_state = op(q4, k4, v4, gk4, initial_state=recurrent_state, output_final_state=use_cache)
            if past_key_values is not None and not self.training:       # model/gla.py:205-213
                past_key_values.update((cq, ck, cv, recurrent_state), self.layer_idx, T)
            o = o.transpose(1, 2)
        if not o.is_contiguous():
            o = o.contiguous()
        y = torch.empty_like(o)
        nw = self.g_norm_swish_gate.weight
        nw = nw.to(x.dtype) if nw is not None else None
        rc = lib.lina_rmsnorm_swishgate_fwd_ld(L.ptr(o), L.ptr(g), L.ptr(nw), L.ptr(y), None, B * T * H, V,
                                               float(self.g_norm_swish_gate.eps), H, ldgate, L.dt(x), L.stream(x))
        L.count_launches(1)
        L.check(rc, "lina_rmsnorm_swishgate_fwd_ld")
        return self.o_proj(y.view(B, T, vd))

    def _can_step(self, x, state, reset_mask, attention_mask) -> bool:
        return (x.shape[1] == 1 and state is not None and not self.training and not torch.is_grad_enabled()
                and not torch.is_autocast_enabled() and self.q_proj.weight.dtype == x.dtype
                and reset_mask is None and attention_mask is None and self.clamp_min is None
                and x.dtype in (torch.float32, torch.bfloat16) and state[-1].dtype in (torch.float32, torch.bfloat16)
                and self.head_v_dim % 8 == 0 and self.head_qk_dim <= 256)

    # -- general path ------------------------------------------------------------------------------
    def forward(self, hidden_states: torch.Tensor, reset_mask: Optional[torch.Tensor] = None,
                attention_mask: Optional[torch.Tensor] = None, reset_val: float = -20,
                past_key_values: Optional[Cache] = None, use_cache: Optional[bool] = False,
                output_attentions: Optional[bool] = False, **kwargs) -> torch.Tensor:
        """model/gla.py:131-227."""
        L.require_cuda(hidden_states)
        mode = self.mode
        last_state = past_key_values[self.layer_idx] if use_cache else None
        if self._can_step(hidden_states, last_state, reset_mask, attention_mask):
            o = self._step(hidden_states, last_state)
            past_key_values.update(last_state, self.layer_idx, 1)      # in place already: bumps seen_tokens only
            return o

        if self._can_prefill(hidden_states, reset_mask, attention_mask) and (not use_cache or last_state is not None):
            return self._prefill(hidden_states, last_state, use_cache, past_key_values, kwargs.get("input_norm_bound"))

        q, k, v = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
        if self.use_short_conv:
            cq, ck, cv = (last_state[0], last_state[1], last_state[2]) if use_cache else (None, None, None)
            q = self.q_conv1d(q, attention_mask, cq)
            k = self.k_conv1d(k, attention_mask, ck)
            v = self.v_conv1d(v, attention_mask, cv)
        if attention_mask is not None:
            v = v.mul_(attention_mask.unsqueeze(-1))
        H = self.num_heads
        q, k, v = (rearrange(t, "b l (h d) -> b h l d", h=H) for t in (q, k, v))
        gk = self.gk_proj(hidden_states)
        if not torch.is_grad_enabled() and gk.is_contiguous():
            gko = torch.empty_like(gk)                                   # one fused pass: logsigmoid / normalizer [clamp]
            rc = L.lib().lina_gate_logsigmoid(L.ptr(gk), L.ptr(gko), gk.numel(), float(self.gate_logit_normalizer),
                                              float(self.clamp_min or 0.0), int(self.clamp_min is not None),
                                              L.dt(gk), L.stream(gk))
            L.count_launches(1)
            L.check(rc, "lina_gate_logsigmoid")
            gk = rearrange(gko, "b n (h d) -> b h n d", h=H)
        else:
            gk = rearrange(gk, "b n (h d) -> b h n d", h=H)
            gk = F.logsigmoid(gk) / self.gate_logit_normalizer
            if self.clamp_min is not None:
                gk = torch.clamp_min(gk, self.clamp_min)
        if gk.dtype != q.dtype:          # autocast may leave the gate in fp32: keep one dtype so the tensor-core kernels apply
            gk = gk.to(q.dtype)
        if reset_mask is not None:
            gk = gk.masked_fill(reset_mask.unsqueeze(1).unsqueeze(3), reset_val)

        recurrent_state = last_state[-1] if use_cache else None
        # gates provably inside the tensor-core kernels' range (weights + the norm bound of the LayerNorm in front): the
        # operator then skips its per-call reduction and host read -- 13 synchronisations per training step otherwise
        cert = reset_mask is None and self.gates_certified(kwargs.get("input_norm_bound"))
        with fla_ops.gates_certified_scope(cert):
            if mode == "fused_recurrent":
                o, recurrent_state = fused_recurrent_gla(q, k, v, gk, initial_state=recurrent_state, output_final_state=use_cache)
            elif mode == "fused_chunk":
                o, recurrent_state = fused_chunk_gla(q, k, v, gk, initial_state=recurrent_state, output_final_state=use_cache)
            elif mode == "chunk":
                o, recurrent_state = chunk_gla(q, k, v, gk, initial_state=recurrent_state, output_final_state=use_cache)
            else:
                raise NotImplementedError(f"Not supported mode `{mode}`.")

        if past_key_values is not None and not self.training:          # model/gla.py:205-213
            if self.use_short_conv:
                new_state = (cq, ck, cv, recurrent_state)
            else:
                new_state = (recurrent_state,)
            past_key_values.update(new_state, self.layer_idx, q.shape[2])

        o = rearrange(o, "b h l d -> b l h d")
        g = rearrange(self.g_proj(hidden_states), "b l (h d) -> b l h d", h=H)
        o = self.g_norm_swish_gate(o, g)
        return self.o_proj(rearrange(o, "b l h d -> b l (h d)"))

    def init_state(self, batch_size: int) -> Tuple[torch.Tensor, ...]:
        """model/gla.py:229-240: zeros in the parameter dtype."""
        p = next(self.parameters())
        state = tuple()
        if self.use_short_conv:
            state += (p.new_zeros(batch_size, self.key_dim, self.conv_size),
                      p.new_zeros(batch_size, self.key_dim, self.conv_size),
                      p.new_zeros(batch_size, self.value_dim, self.conv_size))
        return state + (p.new_zeros(batch_size, self.num_heads, self.head_qk_dim, self.head_v_dim),)

    def state_size(self, **kwargs) -> int:
        n = self.key_dim * self.head_v_dim
        for m in self.children():
            if isinstance(m, ShortConvolution):
                n += m.state_size
        return n


class AttentiveGLA(AttentiveRNN):
    """model/gla.py:252-365: n_layer encoder blocks -> cross attention (+pos_net GLA block when
    ``blind``) -> n_layer decoder blocks; layer_idx 0..N-1, N..2N-1, 2N."""

    def __init__(self, d_model: int, n_layer: int, heads: int, dropout_att: float = 0.0, dropout: float = 0.0,
                 d_blind: int = None, blind: bool = False, cross_att_pp: bool = False, rotary: bool = False,
                 use_short_conv: bool = False, expand_k: float = 1.0, expand_v: float = 2.0,
                 pos_type="sinusoidal"):
        super().__init__()

        def block(d, h, i):
            return MixingBlock(lambda: GatedLinearAttention(hidden_size=d, num_heads=h, use_short_conv=use_short_conv,
                                                            expand_k=expand_k, expand_v=expand_v, layer_idx=i),
                               lambda: SwiGLU(d), lambda: nn.LayerNorm(d), dropout=dropout)

        self.encoder = nn.ModuleList([block(d_model, heads, i) for i in range(n_layer)])
        self.decoder = nn.ModuleList([block(d_model, heads, i) for i in range(n_layer, 2 * n_layer)])
        if d_blind is None:
            d_blind = d_model
        if blind:
            self.cross_att = BlindCrossAttention(d_model, d_model, d_model, 1, block(d_blind, heads, 2 * n_layer),
                                                 dropout_att, pos_dim=d_blind, rotary=rotary, pos_type=pos_type)
        elif cross_att_pp:
            raise NotImplementedError("cross_att_pp is an experimental variant outside the shipped model")
        else:
            self.cross_att = CrossAttention(d_model, d_model, d_model, heads, dropout_att)

    def forward(self, x, ctx, mask=None, pos=None, reset_mask=None, attention_only=None, forced_attention=None,
                init_state=None, crossatt_pos=None):
        """model/gla.py:287-300.  NB the cross attention's pos_net never sees ``init_state`` here."""
        if self.encoder[0].can_fuse(x):                       # inference: adds folded into the LayerNorms
            return self._forward_fused(x, ctx, mask, reset_mask, init_state, crossatt_pos)
        for e in self.encoder:
            if self.training:
                e = maybe_grad_ckpt(e)
            x = e(x, use_cache=init_state is not None, past_key_values=init_state)
        v, att = self.cross_att(x, ctx, mask=mask, reset_mask=reset_mask, pos=crossatt_pos)
        x = x + v
        for d in self.decoder:
            if self.training:
                d = maybe_grad_ckpt(d)
            x = d(x, use_cache=init_state is not None, past_key_values=init_state)
        return x, att

    def _all_certified(self) -> bool:
        """Every mixer's gates are bounded inside the tensor-core envelope by its weights alone (cached per weight version):
        no device flag will be written, so no cache snapshot and no host read are needed."""
        blocks = list(self.encoder) + list(self.decoder)
        if hasattr(self.cross_att, "pos_net"):
            blocks.append(self.cross_att.pos_net)
        return all(hasattr(b, "_ln_output_norm_bound") and b.tmix.gates_certified(b._ln_output_norm_bound()) for b in blocks)

    def _forward_fused(self, x, ctx, mask, reset_mask, init_state, crossatt_pos):
        """Inference pass with the residual adds folded into the LayerNorms.  The 13 mixers run on the pre-gated tensor-core
        kernels without looking at their gates one by one; the gate-envelope flag they accumulate is read ONCE here."""
        def run():
            kw = dict(use_cache=init_state is not None, past_key_values=init_state)
            xr, d = x, None
            for e in self.encoder:
                xr, d = e.forward_fused(xr, d, **kw)
            xe = xr + d
            v, att = self.cross_att(xe, ctx, mask=mask, reset_mask=reset_mask, pos=crossatt_pos)
            xr, d = xe, v
            for dd in self.decoder:
                xr, d = dd.forward_fused(xr, d, **kw)
            return xr + d, att

        capturing = torch.cuda.is_current_stream_capturing()
        snap = None
        if init_state is not None and not capturing and fla_ops.GATE_CHECK and x.shape[1] > 1 and not self._all_certified():
            snap = [tuple(t.clone() for t in st) for st in init_state.states]      # the pass updates the cache in place
        GateEnvelope.flagged_calls = 0
        with GateEnvelope.deferred():
            out = run()
        if (capturing or not fla_ops.GATE_CHECK or x.shape[1] == 1 or GateEnvelope.flagged_calls == 0
                or not GateEnvelope.tripped(x.device)):
            return out                                        # every mixer certified by its weights, or the flag is clear
        # some chunk's summed log gate fell below -80: redo the pass with the exact kernels (the reference is exact for any gate)
        if snap is not None:
            for st, sn in zip(init_state.states, snap):
                for a, b in zip(st, sn):
                    a.copy_(b)
        global PREGATED
        old, PREGATED = PREGATED, False
        try:
            return run()
        finally:
            PREGATED = old

    def init_state(self, max_seqlen=1000, batch_size=16):
        """model/gla.py:302-313."""
        cache = Cache()
        blocks = list(self.encoder) + list(self.decoder)
        for i, b in enumerate(blocks):
            cache.update(b.tmix.init_state(batch_size), i, offset=0)
        if hasattr(self.cross_att, "pos_net"):
            cache.update(self.cross_att.pos_net.tmix.init_state(batch_size), len(blocks), offset=0)
        return cache

    def get_state_from_params(self, params, batch_size, scale=0.02):
        """model/gla.py:315-325: S = (k (x) v summed over rank) * scale, written into slot [-1]."""
        cache = self.init_state(batch_size=batch_size)
        for i, x in enumerate(params):
            if len(x) == 2:
                state = einsum(*x, "b r h k vv, b r h kk v -> b h k v") * scale
            else:
                state = x[0]
            state = repeat(state, "1 ... -> bs ...", bs=batch_size).clone()
            cache.states[i] = cache.states[i][:-1] + (state,)
        return cache

    def to_mode(self, mode):
        """model/gla.py:327-333 (the reference sets ``pos_net.mode`` on the MixingBlock, which has no
        effect on its tmix; here the pos_net mixer really switches too)."""
        for b in list(self.encoder) + list(self.decoder):
            b.tmix.mode = mode
        if hasattr(self.cross_att, "pos_net"):
            self.cross_att.pos_net.mode = mode
            self.cross_att.pos_net.tmix.mode = mode

    def get_init_state_tuning_params(self, lora: Optional[int] = None, scale: float = 0.02, device=None):
        """model/gla.py:336-356: per enc/dec layer rank-r factors k [1,r,H,K,1], v [1,r,H,1,V]."""
        params = []
        for b in list(self.encoder) + list(self.decoder):
            t = b.tmix
            K, V, H = t.head_qk_dim, t.head_v_dim, t.num_heads
            if lora is not None:
                params.append((nn.Parameter(torch.randn(1, lora, H, K, 1, device=device)),
                               nn.Parameter(torch.randn(1, lora, H, 1, V, device=device) * scale)))
            else:
                params.append(nn.Parameter(torch.randn(1, H, K, V, device=device) * scale))
        return params

    def step(self, y_embd, x_enc, time_step, cache):
        """model/gla.py:358-365: one token through every block, all blocks stateful."""
        if self.encoder[0].can_fuse(y_embd):
            def run():
                kw = dict(past_key_values=cache, use_cache=True)
                xr, d = y_embd, None
                for e in self.encoder:
                    xr, d = e.forward_fused(xr, d, **kw)
                y = xr + d
                v, att = self.cross_att(y, x_enc, time_step=time_step, past_key_values=cache, use_cache=True)
                xr, d = y, v
                for dd in self.decoder:
                    xr, d = dd.forward_fused(xr, d, **kw)
                return xr + d, att, cache

            if (y_embd.shape[1] == 1 or torch.cuda.is_current_stream_capturing() or not fla_ops.GATE_CHECK
                    or self._all_certified()):
                return run()                                  # single-token steps run the exact recurrence (lina_gla_step)
            # multi-token prompt prefill: same deferred gate-envelope policy as _forward_fused
            snap = [tuple(t.clone() for t in st) for st in cache.states]
            seen = getattr(cache, "_seen_tokens", None)
            GateEnvelope.flagged_calls = 0
            with GateEnvelope.deferred():
                out = run()
            if GateEnvelope.flagged_calls == 0 or not GateEnvelope.tripped(y_embd.device):
                return out
            for st, sn in zip(cache.states, snap):
                for a, b in zip(st, sn):
                    a.copy_(b)
            if seen is not None:
                cache._seen_tokens = seen
            global PREGATED
            old, PREGATED = PREGATED, False
            try:
                return run()
            finally:
                PREGATED = old
        for e in self.encoder:
            y_embd = e(y_embd, past_key_values=cache, use_cache=True)
        v, att = self.cross_att(y_embd, x_enc, time_step=time_step, past_key_values=cache, use_cache=True)
        y_embd = y_embd + v
        for d in self.decoder:
            y_embd = d(y_embd, past_key_values=cache, use_cache=True)
        return y_embd, att, cache

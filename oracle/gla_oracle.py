"""CPU oracle for the GLA hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.  The product package
(``lina_speech_b200``) never does: it fails loudly when its CUDA library is
missing instead of falling back to anything in here.

Every function restates, in plain CPU torch, the arithmetic of one reference
function (paths relative to ``/root/reference``; ``FLA/`` =
``3rdparty/flash-linear-attention/``).  Parity of this restatement against the
reference itself is pinned by ``tests/golden/make_golden.py`` (run in the build
container where the reference is importable) and re-checked against the
committed fixtures by ``tests/test_oracle.py``.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# a0 / a1: the recurrence  (FLA/fla/ops/gla/naive.py:13-44)
# ----------------------------------------------------------------------------
def recurrent_gla(q, k, v, gk, scale: Optional[float] = None, initial_state=None,
                  output_final_state: bool = True, acc_dtype=torch.float32
                  ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """S_t = exp(gk_t)[:, None] * S_{t-1} + k_t^T v_t ;  o_t = (scale * q_t) S_t.

    Shapes: q, k, gk [B,H,T,K]; v [B,H,T,V]; initial_state [B,H,K,V].
    Follows FLA/fla/ops/gla/naive.py:22-44 (upcast, per-step update, output in
    the input dtype, final state in fp32); ``scale`` defaults to K**-0.5
    (naive.py:28, FLA/fla/ops/gla/recurrent_fuse.py:24-25).
    """
    odt = v.dtype
    q, k, v, gk = (x.to(acc_dtype) for x in (q, k, v, gk))
    B, H, T, K = q.shape
    V = v.shape[-1]
    if scale is None or scale == -1:
        scale = K ** -0.5
    S = torch.zeros(B, H, K, V, dtype=acc_dtype)
    if initial_state is not None:
        S = S + initial_state.to(acc_dtype)
    o = torch.empty(B, H, T, V, dtype=acc_dtype)
    for t in range(T):
        S = S * gk[:, :, t].exp().unsqueeze(-1) + k[:, :, t].unsqueeze(-1) * v[:, :, t].unsqueeze(-2)
        o[:, :, t] = torch.einsum("bhk,bhkv->bhv", q[:, :, t] * scale, S)
    return o.to(odt), (S.to(torch.float32) if output_final_state else None)


def recurrent_rwkv6(r, k, v, w, u, scale: Optional[float] = None, initial_state=None,
                    output_final_state: bool = True, acc_dtype=torch.float32):
    """RWKV6 form (FLA/fla/ops/rwkv6/recurrent_naive.py:8-42): output BEFORE the update, current token through the
    bonus: o_t = scale * r_t (S_{t-1} + diag(u) k_t^T v_t) ; S_t = diag(exp(w_t)) S_{t-1} + k_t^T v_t ; u [H,K]."""
    odt = v.dtype
    r, k, v, w, u = (x.to(acc_dtype) for x in (r, k, v, w, u))
    B, H, T, K = r.shape
    V = v.shape[-1]
    if scale is None or scale == -1:
        scale = K ** -0.5
    S = torch.zeros(B, H, K, V, dtype=acc_dtype)
    if initial_state is not None:
        S = S + initial_state.to(acc_dtype)
    o = torch.empty(B, H, T, V, dtype=acc_dtype)
    for t in range(T):
        kv = k[:, :, t].unsqueeze(-1) * v[:, :, t].unsqueeze(-2)
        o[:, :, t] = torch.einsum("bhk,bhkv->bhv", r[:, :, t] * scale, S + u[None, :, :, None] * kv)
        S = S * w[:, :, t].exp().unsqueeze(-1) + kv
    return o.to(odt), (S.to(torch.float32) if output_final_state else None)


def recurrent_gla_bwd(q, k, v, gk, h0, do, dht=None, scale: Optional[float] = None,
                      acc_dtype=torch.float64):
    """Explicit backward of :func:`recurrent_gla`.

    Restates the identities of FLA/fla/ops/common/fused_recurrent.py:172-257
    (forward sweep for dq, reverse sweep for dk/dv with dS carried, dh0 = dS
    after t = 0) and :335-342 (dgk = reversed cumsum of dq*q - dk*k).
    Returns (dq, dk, dv, dgk, dh0) in ``acc_dtype``.
    """
    q, k, v, gk, do = (x.to(acc_dtype) for x in (q, k, v, gk, do))
    B, H, T, K = q.shape
    V = v.shape[-1]
    if scale is None or scale == -1:
        scale = K ** -0.5
    S = torch.zeros(B, H, K, V, dtype=acc_dtype)
    if h0 is not None:
        S = S + h0.to(acc_dtype)
    dq = torch.empty_like(q)
    for t in range(T):
        S = S * gk[:, :, t].exp().unsqueeze(-1) + k[:, :, t].unsqueeze(-1) * v[:, :, t].unsqueeze(-2)
        dq[:, :, t] = scale * torch.einsum("bhkv,bhv->bhk", S, do[:, :, t])
    dS = torch.zeros(B, H, K, V, dtype=acc_dtype)
    if dht is not None:
        dS = dS + dht.to(acc_dtype)
    dk = torch.empty_like(k)
    dv = torch.empty_like(v)
    for t in range(T - 1, -1, -1):
        dS = dS + scale * q[:, :, t].unsqueeze(-1) * do[:, :, t].unsqueeze(-2)
        dk[:, :, t] = torch.einsum("bhkv,bhv->bhk", dS, v[:, :, t])
        dv[:, :, t] = torch.einsum("bhkv,bhk->bhv", dS, k[:, :, t])
        dS = dS * gk[:, :, t].exp().unsqueeze(-1)
    dgk = (dq * q - dk * k).flip(2).cumsum(2).flip(2)
    if dht is not None:
        # the reference formula (:335-342) drops this term; autograd through the
        # naive recurrence has it: d<dht, S_T>/dgk_t = sum_v dht * S_T for every t.
        dgk = dgk + (dht.to(acc_dtype) * S).sum(-1).unsqueeze(2)
    return dq, dk, dv, dgk, dS


def chunk_gla(q, k, v, gk, scale: Optional[float] = None, initial_state=None, chunk: int = 64,
              operand_dtype: Optional[torch.dtype] = None, acc_dtype=torch.float32):
    """Chunkwise-parallel form of the same function.

    Restates FLA/fla/ops/gla/chunk.py:110-136 (inter: (q*scale*e^G) @ S),
    :37-79 (intra scores A[t,s] = sum_k q_t k_s e^{G_t-G_s}, s <= t) and
    FLA/fla/ops/common/chunk_h.py:74-94 (S' = e^{G_C} S + (k e^{G_C-G})^T v),
    with a single pivot per chunk (chunk start).  ``operand_dtype`` rounds the
    gate-rescaled MMA operands the way a bf16 tensor-core kernel does
    (FLA/fla/ops/gla/chunk_util.py:57-60), to reason about tolerances.
    """
    odt = v.dtype
    q, k, v, gk = (x.to(acc_dtype) for x in (q, k, v, gk))
    B, H, T, K = q.shape
    V = v.shape[-1]
    if scale is None or scale == -1:
        scale = K ** -0.5
    rnd = (lambda x: x.to(operand_dtype).to(acc_dtype)) if operand_dtype is not None else (lambda x: x)
    S = torch.zeros(B, H, K, V, dtype=acc_dtype)
    if initial_state is not None:
        S = S + initial_state.to(acc_dtype)
    o = torch.empty(B, H, T, V, dtype=acc_dtype)
    for c0 in range(0, T, chunk):
        c1 = min(T, c0 + chunk)
        G = gk[:, :, c0:c1].cumsum(2)
        qg = rnd(q[:, :, c0:c1] * G.exp() * scale)
        kg = rnd(k[:, :, c0:c1] * (-G).exp())
        vv = rnd(v[:, :, c0:c1])
        A = torch.einsum("bhtk,bhsk->bhts", qg, kg).tril()
        o[:, :, c0:c1] = torch.einsum("bhtk,bhkv->bhtv", qg, rnd(S)) + torch.einsum("bhts,bhsv->bhtv", rnd(A), vv)
        GC = G[:, :, -1]
        S = (S + torch.einsum("bhsk,bhsv->bhkv", kg, vv)) * GC.exp().unsqueeze(-1)
    return o.to(odt), S.to(torch.float32)


def fused_chunk_gla_as_reference_rounds(q, k, v, gk, scale: Optional[float] = None, initial_state=None):
    """The reference's DEFAULT op ``fused_chunk_gla`` with every rounding it performs on low-precision inputs, in CPU torch
    (FLA/fla/ops/gla/chunk_fuse.py:302-399).  Used to bound OUR error by the REFERENCE's own: both are compared against
    the fp64 recurrence on identical inputs.  With dt = the input dtype (bf16 / fp16), BT = 16, BK = BV = 64:

      * g  = fp32 cumsum inside each 16-token chunk                          (chunk_util.py:5-26)
      * q_g = dt(q e^g scale), k_g = dt(k e^{g_last - g})                     (chunk_util.py:28-65)
      * inter: per K block, o_part = dt(q_g @ dt(h)), h = h e^{g_last} + k_g^T v in fp32, the NK partial outputs are
        stored in dt and summed by torch (fp32 accumulate, one rounding)       (chunk_fuse.py:77-99,323,367)
      * intra: A = dt(sum_k q scale k e^{g_t - g_s}) per K block (fp32 math, dt store), summed over blocks likewise,
        o2 = dt(A @ v), o = dt(o + o2)                                         (chunk_fuse.py:238-247,376-390)

    Pinned to the reference itself: tests/golden/gla_triton_bf16.npz holds the outputs of the reference's Triton kernels run
    on a B200 (profiles/triton_reference_bench.py); this function reproduces them (tests/test_oracle.py).

    T must be a multiple of 16 (the reference pads).  Returns (o in dt, final state fp32)."""
    dt = v.dtype
    B, H, T, K = q.shape
    V = v.shape[-1]
    BT, BK = 16, min(K, 64)
    assert T % BT == 0
    if scale is None or scale == -1:
        scale = K ** -0.5
    r = lambda x: x.to(dt).float()
    qf, kf, vf = q.float(), k.float(), v.float()
    n = T // BT
    g = gk.float().view(B, H, n, BT, K).cumsum(3)
    qc, kc, vc = qf.view(B, H, n, BT, K), kf.view(B, H, n, BT, K), vf.view(B, H, n, BT, V)
    g_last = g[:, :, :, -1:]
    qg = r(qc * g.exp() * scale)
    kg = r(kc * (g_last - g).exp())
    h = torch.zeros(B, H, K, V) if initial_state is None else initial_state.float().clone()
    o = torch.zeros(B, H, n, BT, V)
    for i in range(n):
        part = torch.zeros(B, H, BT, V)
        hr = r(h)
        for k0 in range(0, K, BK):             # NK partial outputs, each stored in dt; torch's sum(0) adds them in fp32
            part = part + r(torch.einsum("bhtk,bhkv->bhtv", qg[:, :, i, :, k0:k0 + BK], hr[:, :, k0:k0 + BK]))
        o[:, :, i] = r(part)                   # ... and rounds the sum once
        h = h * g_last[:, :, i, 0].exp().unsqueeze(-1) + torch.einsum("bhsk,bhsv->bhkv", kg[:, :, i], vc[:, :, i])
    A = torch.zeros(B, H, n, BT, BT)
    for k0 in range(0, K, BK):
        sl = slice(k0, k0 + BK)
        e = (g[:, :, :, :, None, sl] - g[:, :, :, None, :, sl]).exp()     # [.., t, s, k]
        a = ((qc[..., sl] * scale)[:, :, :, :, None] * kc[..., sl][:, :, :, None] * e).sum(-1).tril()
        A = A + r(a)
    A = r(A)
    o2 = r(torch.einsum("bhnts,bhnsv->bhntv", A, vc))
    return r(o + o2).view(B, H, T, V).to(dt), h


def pregated_chunk_fwd(qg, kg, v, decay, h0=None, row_decay: bool = False, chunk: int = 64, acc_dtype=torch.float32):
    """Contract of the pre-gated tensor-core kernel (lina_gla_chunk_fwd_pregated) restated in torch, for CPU tests of the
    host-side backward that is built from it: per 64-token chunk  o = qg S + tril(qg kg^T) v ;  S' = decay (.) (S + kg^T v),
    decay [B,H,NT,K] scaling the key dim (rows of S) or, with ``row_decay``, [B,H,NT,V] scaling the value dim.
    With qg = scale q e^G, kg = k e^-G, decay = e^{G_C} this is the chunk form of FLA/fla/ops/gla/chunk.py:110-136 +
    FLA/fla/ops/common/chunk_h.py:74-94.  T must be a multiple of ``chunk``.  Returns (o, final state), both acc_dtype."""
    qg, kg, v = (x.to(acc_dtype) for x in (qg, kg, v))
    B, H, T, K = qg.shape
    V = v.shape[-1]
    assert T % chunk == 0
    S = torch.zeros(B, H, K, V, dtype=acc_dtype) if h0 is None else h0.to(acc_dtype).clone()
    o = torch.empty(B, H, T, V, dtype=acc_dtype)
    for n in range(T // chunk):
        sl = slice(n * chunk, (n + 1) * chunk)
        P = torch.einsum("bhtk,bhsk->bhts", qg[:, :, sl], kg[:, :, sl]).tril()
        o[:, :, sl] = torch.einsum("bhtk,bhkv->bhtv", qg[:, :, sl], S) + torch.einsum("bhts,bhsv->bhtv", P, v[:, :, sl])
        S = S + torch.einsum("bhsk,bhsv->bhkv", kg[:, :, sl], v[:, :, sl])
        d = decay[:, :, n].to(acc_dtype)
        S = S * (d.unsqueeze(-2) if row_decay else d.unsqueeze(-1))
    return o, S


# ----------------------------------------------------------------------------
# a5: ShortConvolution  (FLA/fla/modules/convolution.py:141-205, torch branch)
# ----------------------------------------------------------------------------
def short_conv_prefill(x, weight, cache=None, activation: Optional[str] = "silu"):
    """x [B,L,D], weight [D,W] (the reference stores [D,1,W]); causal depthwise
    conv + SiLU.  ``cache`` [B,D,W] receives the last W inputs, left-padded with
    zeros (convolution.py:164-166).  Torch branch :175-178."""
    B, L, D = x.shape
    W = weight.shape[-1]
    xt = x.transpose(1, 2)
    if cache is not None:
        cache.copy_(F.pad(xt, (W - L, 0)) if L < W else xt[..., L - W:])
    y = F.conv1d(F.pad(xt.float(), (W - 1, 0)), weight.float().unsqueeze(1), groups=D)
    if activation is not None:
        y = F.silu(y)
    return y.transpose(1, 2).to(x.dtype)


def short_conv_step(x, cache, weight, activation: Optional[str] = "silu"):
    """x [B,1,D]; cache [B,D,W] rolled left by one, x inserted at the end, then
    dotted with the taps (convolution.py:197-204)."""
    cache.copy_(torch.roll(cache, shifts=-1, dims=-1))
    cache[:, :, -1] = x[:, 0].to(cache.dtype)
    y = (cache.float() * weight.float()).sum(-1)
    if activation is not None:
        y = F.silu(y)
    return y.to(x.dtype).unsqueeze(1)


# ----------------------------------------------------------------------------
# a6: FusedRMSNormSwishGate  (FLA/fla/modules/fused_norm_gate.py:41-55,120-139)
# ----------------------------------------------------------------------------
def rmsnorm_swish_gate(x, g, weight, eps: float = 1e-5):
    """y = x * rsqrt(mean(x^2) + eps) * w * g * sigmoid(g), fp32 math, output in
    x.dtype (kernel :120-139; reference restatement rms_norm_ref :41-55)."""
    xf, gf = x.float(), g.float()
    rstd = torch.rsqrt(xf.square().mean(-1, keepdim=True) + eps)
    y = xf * rstd
    if weight is not None:
        y = y * weight.float()
    return (y * gf * torch.sigmoid(gf)).to(x.dtype)


def rmsnorm_swish_gate_bwd(x, g, weight, dy, eps: float = 1e-5):
    """Gradients (dx, dg, dw) by autograd over the fp64 restatement."""
    xd = x.double().requires_grad_(True)
    gd = g.double().requires_grad_(True)
    wd = weight.double().requires_grad_(True)
    rstd = torch.rsqrt(xd.square().mean(-1, keepdim=True) + eps)
    y = xd * rstd * wd * gd * torch.sigmoid(gd)
    y.backward(dy.double())
    return xd.grad, gd.grad, wd.grad


# ----------------------------------------------------------------------------
# gate preparation  (model/gla.py:174-184)
# ----------------------------------------------------------------------------
def gate_logsigmoid(x, normalizer: float = 16.0, clamp_min: Optional[float] = None):
    gk = F.logsigmoid(x) / normalizer
    if clamp_min is not None:
        gk = torch.clamp_min(gk, clamp_min)
    return gk

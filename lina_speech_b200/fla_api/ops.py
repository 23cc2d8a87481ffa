"""Drop-in replacements for the reference's GLA operator API.

Same names, argument meaning, return values and autograd behaviour as

  fla.ops.gla.fused_recurrent_gla   FLA/fla/ops/gla/recurrent_fuse.py:13-27
  fla.ops.gla.fused_chunk_gla       FLA/fla/ops/gla/chunk_fuse.py:518-536
  fla.ops.gla.chunk_gla             FLA/fla/ops/gla/chunk.py:453-491

(FLA/ = 3rdparty/flash-linear-attention/ of the reference), backed by the CUDA
kernels of liblina_b200.so through the C ABI in include/lina_b200.h.

Contract: q, k, gk [B,H,T,K]; v [B,H,T,V]; head-first, made
contiguous here; any float dtype (kernels compute in fp32); ``o`` comes back in
v.dtype, ``final_state`` always fp32 [B,H,K,V]; ``initial_state`` any float dtype
or None; ``scale`` None / -1 means K**-0.5; gates are log-space (<= 0).
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from .. import _lib as L


def _is_bthd(x: torch.Tensor) -> bool:
    """[B,H,T,D] view of a contiguous [B,T,H,D] tensor (what rearrange('b l (h d) -> b h l d') returns)."""
    return (not x.is_contiguous()) and x.transpose(1, 2).is_contiguous()


def _prep(q, k, v, gk, h0, kind="recurrent"):
    L.require_cuda(q, k, v, gk, h0)
    if not (q.dtype == k.dtype == v.dtype == gk.dtype):
        q, k, v, gk = (x.float() for x in (q, k, v, gk))   # mixed dtypes: compute everything in fp32
    B_, H_, T_, K_ = q.shape
    bthd = (kind != "recurrent" and all(_is_bthd(x) for x in (q, k, v, gk)) and
            bool(L.lib().lina_gla_chunk_fwd_uses_tensor_cores(B_, H_, T_, K_, v.shape[-1], L._DT.get(q.dtype, -1))))
    if not bthd:
        q, k, v, gk = (x.contiguous() for x in (q, k, v, gk))
    if h0 is not None:
        h0 = h0.contiguous()
        if h0.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            h0 = h0.float()
    B, H, T, K = q.shape
    V = v.shape[-1]
    if k.shape != q.shape or gk.shape != q.shape or v.shape[:3] != q.shape[:3]:
        raise ValueError(f"GLA shapes disagree: q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)} gk{tuple(gk.shape)}")
    if h0 is not None and tuple(h0.shape) != (B, H, K, V):
        raise ValueError(f"initial_state must be [B,H,K,V]={B, H, K, V}, got {tuple(h0.shape)}")
    return q, k, v, gk, h0, bthd


# bench.py sets this to a list to collect (kind, start_event, end_event) of every forward launch
PROFILE = None


def _fwd(kind: str, q, k, v, gk, h0, scale: float, want_ht: bool, bthd: bool = False):
    lib = L.lib()
    if PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(torch.cuda.current_stream(q.device))
    B, H, T, K, V = q.shape[0], q.shape[1], q.shape[2], q.shape[3], v.shape[3]
    if bthd:     # inputs are [B,H,T,D] views of [B,T,H,D] memory; produce o the same way (no copies either side)
        o = torch.empty(B, T, H, V, dtype=v.dtype, device=v.device).transpose(1, 2)
    else:
        o = torch.empty_like(v)
    ht = torch.empty(B, H, K, V, dtype=torch.float32, device=q.device) if want_ht else None
    h0dt = L.dt(h0) if h0 is not None else 0
    if kind == "recurrent":
        rc = lib.lina_gla_recurrent_fwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(gk), L.ptr(h0), h0dt, L.ptr(o),
                                        L.ptr(ht), B, H, T, K, V, L.dt(q), scale, L.stream(q))
        L.count_launches(1)
    else:
        nbytes = lib.lina_gla_chunk_fwd_workspace_bytes(B, H, T, K, V, L.dt(q))
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=q.device)
        entry = lib.lina_gla_chunk_fwd_bthd if bthd else lib.lina_gla_chunk_fwd
        rc = entry(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(gk), L.ptr(h0), h0dt, L.ptr(o), L.ptr(ht),
                                    L.ptr(ws), B, H, T, K, V, L.dt(q), scale, L.stream(q))
        L.count_launches(1)
    L.check(rc, f"lina_gla_{kind}_fwd")
    if PROFILE is not None:
        ev1.record(torch.cuda.current_stream(q.device))
        PROFILE.append((kind, ev0, ev1))
    return o, ht


# The tensor-core chunk kernels use one pivot per 64-token chunk: k e^{-G} leaves the fp32 / bf16 exponent range when the
# summed log-gate of a channel inside one chunk goes below about -85 (DESIGN.md 4.1 "Numerics").  The reference is exact for any
# gate (chunk_fuse.py:196-238 computes the intra-chunk term pairwise in fp32), so the operator API checks the gates first and
# serves such inputs with the exact recurrence kernels instead.  One small reduction + one host read per call;
# LINA_GATE_CHECK=0 removes it (then the caller guarantees the envelope).
GATE_CHECK = os.environ.get("LINA_GATE_CHECK", "1") != "0"
GATE_SUM_LIMIT = -80.0
_CHUNK = 64


def _min_chunk_gate_sum(gk: torch.Tensor) -> torch.Tensor:
    """min over (b, h, chunk, channel) of the summed log-gates of one 64-token chunk; gk [B,H,T,K], any strides."""
    T = gk.shape[2]
    full = T // _CHUNK
    parts = []
    if full:
        parts.append(gk.narrow(2, 0, full * _CHUNK).unflatten(2, (full, _CHUNK)).sum(3, dtype=torch.float32).amin())
    if T % _CHUNK:
        parts.append(gk.narrow(2, full * _CHUNK, T % _CHUNK).sum(2, dtype=torch.float32).amin())
    return parts[0] if len(parts) == 1 else torch.minimum(parts[0], parts[1])


def _gates_in_envelope(gk: torch.Tensor) -> bool:
    return bool(_min_chunk_gate_sum(gk).item() >= GATE_SUM_LIMIT)        # NaN gates compare False -> exact path


_CERTIFIED = False        # set by gates_certified_scope: the caller proved the gates inside the envelope (no device read needed)


class gates_certified_scope:
    """``with gates_certified_scope(True): chunk_gla(...)`` -- the caller vouches that every 64-token chunk's summed log gate
    stays above GATE_SUM_LIMIT (GatedLinearAttention derives this from its weights, model/gla.py:gates_certified), so the
    operator skips its reduction + host read and the backward may use the tensor-core path."""

    def __init__(self, ok: bool):
        self.ok = bool(ok)

    def __enter__(self):
        global _CERTIFIED
        self.prev, _CERTIFIED = _CERTIFIED, self.ok
        return self

    def __exit__(self, *exc):
        global _CERTIFIED
        _CERTIFIED = self.prev
        return False


def _route(kind: str, q, v, gk, uses_tc):
    """'recurrent' | 'chunk' | 'fused_chunk' -> (kernel family that serves the forward, gates_ok): the chunk forms fall back
    to the exact recurrence when they would run on the tensor-core kernel with gates outside its numeric envelope.
    gates_ok is None when the gates were not looked at (the backward looks then, if it wants the tensor-core path)."""
    if kind == "recurrent" or not GATE_CHECK:
        return kind, (None if GATE_CHECK else True)
    if _CERTIFIED:
        return kind, True
    B, H, T, K = q.shape
    if not uses_tc(B, H, T, K, v.shape[-1], L._DT.get(q.dtype, -1)):
        return kind, None
    if q.is_cuda and torch.cuda.is_current_stream_capturing():
        return kind, None            # a host read is illegal during graph capture: the caller vouches for the gates
    ok = _gates_in_envelope(gk)
    return (kind if ok else "recurrent"), ok


TC_BWD = os.environ.get("LINA_TC_BWD", "1") != "0"
CONCURRENT_BWD = os.environ.get("LINA_CONCURRENT_BWD", "1") != "0"
_SIDE = {}


def _side_streams(dev, n):
    key = (dev.index if dev.index is not None else torch.cuda.current_device())
    pool = _SIDE.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def _tc_bwd_eligible(q, v) -> bool:
    """The backward-through-the-forward-kernel path: bf16, K a multiple of 128 (it becomes the value dim of the dq / dk
    runs), V split into pieces of 64 / 128 / 256 (the key dim of those runs), T long enough for the tensor-core kernel."""
    if not TC_BWD or q.dtype != torch.bfloat16:
        return False
    B, H, T, K = q.shape
    V = v.shape[-1]
    ns = (V + 255) // 256
    return (K in (128, 256) and V % 128 == 0 and V % ns == 0 and (V // ns) in (64, 128, 256) and T >= 64
            and B * H <= 65535)


def _run_pregated(qg, kg, v, decay, h0, o, ht, row_decay: bool, bthd: bool = False):
    """One launch of the pre-gated tensor-core kernel (the unit the backward is built from).  Operands are given in their
    MEMORY layout -- [B,H,T,D], or [B,T,H,D] with ``bthd`` -- and qg / kg / v may be last-dim slices of wider contiguous
    tensors: they are read in place (row stride = the parent's width)."""
    if bthd:
        B, T, H, K = qg.shape
    else:
        B, H, T, K = qg.shape
    V = v.shape[-1]

    def ld(t):
        if t.is_contiguous():
            return 0
        w = t.stride(2)                          # [B,H,T,W]: stride of T; [B,T,H,W]: stride of H -- the parent's row width
        exp = (H * T * w, T * w, w, 1) if not bthd else (T * H * w, H * w, w, 1)
        if tuple(t.stride()) != exp:
            raise ValueError("operand must be a last-dim slice of a contiguous tensor")
        return w

    assert decay.is_contiguous() and o.is_contiguous()
    rc = L.lib().lina_gla_chunk_fwd_pregated(L.ptr(qg), L.ptr(kg), L.ptr(v), L.ptr(decay), L.ptr(h0),
                                            L.dt(h0) if h0 is not None else 0, L.ptr(o), L.ptr(ht), B, H, T, K, V,
                                            int(bthd), int(row_decay), int(o.dtype == torch.float32), ld(qg), ld(kg), ld(v),
                                            L.stream(qg))
    L.count_launches(1)
    L.check(rc, "lina_gla_chunk_fwd_pregated")


def _bwd_tc(q, k, v, gk, h0, do, dht, scale: float, want_dh0: bool):
    """Product path of the tensor-core backward: same scheme as ``_bwd_tc_reference`` with fused element-wise kernels.
    9 launches of ours per call at V <= 512: prep, flip, 5 x tcgen05, post, finish (+ a tiny torch scan over the chunk
    totals and the transposes of the [K,V] states).  Works in the memory layout the operands arrive in: when q, k, v, gk
    are [B,H,T,D] views of [B,T,H,D] tensors (what the model passes) nothing is copied and the gradients come back as the
    same kind of view."""
    lib = L.lib()
    bthd = all(_is_bthd(x) for x in (q, k, v, gk))
    B, H, T, K = q.shape
    V = v.shape[-1]
    if bthd:                                       # memory-layout tensors [B,T,H,D]
        q, k, v, gk = (x.transpose(1, 2) for x in (q, k, v, gk))
        do = do.transpose(1, 2).contiguous()
    else:
        q, k, v, gk, do = (x.contiguous() for x in (q, k, v, gk, do))
    C = 64
    NT = (T + C - 1) // C
    Tp = NT * C
    dev, bf, f32 = q.device, q.dtype, torch.float32
    st = L.stream(q)
    tdim = 1 if bthd else 2

    def alloc(Tn, Dn, dtype):
        return torch.empty((B, Tn, H, Dn) if bthd else (B, H, Tn, Dn), dtype=dtype, device=dev)

    kt = alloc(T, K, bf)
    qh_r, kh_r = alloc(Tp, K, bf), alloc(Tp, K, bf)
    D, Dr = (torch.empty(B, H, NT, K, dtype=f32, device=dev) for _ in range(2))
    L.check(lib.lina_gla_bwd_prep(L.ptr(q), L.ptr(k), L.ptr(gk), L.ptr(kt), L.ptr(qh_r), L.ptr(kh_r), L.ptr(D), L.ptr(Dr),
                                  B, H, T, K, int(bthd), scale, st), "lina_gla_bwd_prep")
    do_r, v_r = alloc(Tp, V, bf), alloc(Tp, V, bf)
    L.check(lib.lina_time_reverse_pad2(L.ptr(do), L.ptr(v), L.ptr(do_r), L.ptr(v_r), B if bthd else B * H, T, Tp,
                                       H * V if bthd else V, st), "lina_time_reverse_pad2")
    L.count_launches(2)
    dht32 = dht.float().contiguous() if dht is not None else None

    # The five runs are independent and each fills only part of the machine at training batch sizes (dq / dk pieces: 2 x B*H
    # CTAs), so they are issued on side streams and joined before the post pass.  Everything they touch was allocated on the
    # calling stream before the fork and is released after the join, so no allocator cross-stream bookkeeping is needed.
    ns = (V + 255) // 256
    Vp = V // ns
    main = torch.cuda.current_stream(dev)
    sides = _side_streams(dev, 2 * ns) if CONCURRENT_BWD else []
    dv_r = alloc(Tp, V, bf)
    dh0 = torch.empty(B, H, K, V, dtype=f32, device=dev) if want_dh0 else None
    dq_parts = [alloc(T, K, f32) for _ in range(ns)]
    dk_parts = [alloc(Tp, K, f32) for _ in range(ns)]
    hts = [torch.empty(B, H, Vp, K, dtype=f32, device=dev) if dht is not None else None for _ in range(ns)]
    h0s = [h0[..., j * Vp:(j + 1) * Vp].float().transpose(-1, -2).contiguous() if h0 is not None else None for j in range(ns)]
    dhts = [dht32[..., j * Vp:(j + 1) * Vp].transpose(-1, -2).contiguous() if dht32 is not None else None for j in range(ns)]
    fork = main.record_event()

    def on(i):
        if not sides:
            return torch.cuda.stream(main)
        sides[i].wait_event(fork)
        return torch.cuda.stream(sides[i])

    # dv (+ dh0): reversed time, key-dim decay -- on the calling stream
    _run_pregated(kh_r, qh_r, do_r, Dr, dht32, dv_r, dh0, False, bthd)
    for j in range(ns):
        sl = slice(j * Vp, (j + 1) * Vp)
        with on(2 * j):       # dq~ piece (forward time, row decay)
            _run_pregated(do[..., sl], v[..., sl], kt, D, h0s[j], dq_parts[j], hts[j], True, bthd)
        with on(2 * j + 1):   # dk^ piece (reversed time)
            _run_pregated(v_r[..., sl], do_r[..., sl], qh_r, Dr, dhts[j], dk_parts[j], None, True, bthd)
    for sd in sides:
        main.wait_event(sd.record_event())
    dv = dv_r.flip(tdim).narrow(tdim, 0, T)
    ST = [ht.transpose(-1, -2) for ht in hts if ht is not None]
    while len(dq_parts) > 2:                       # V > 512: fold the extra pieces (the post kernel sums two)
        dq_parts[0].add_(dq_parts.pop())
        dk_parts[0].add_(dk_parts.pop())
    dq, dk, dgk = alloc(T, K, bf), alloc(T, K, bf), alloc(T, K, bf)
    dgk_local = alloc(T, K, f32)
    totals = torch.empty(B, H, NT, K, dtype=f32, device=dev)
    two = len(dq_parts) == 2
    L.check(lib.lina_gla_bwd_post(L.ptr(dq_parts[0]), L.ptr(dq_parts[1]) if two else None, L.ptr(dk_parts[0]),
                                  L.ptr(dk_parts[1]) if two else None, L.ptr(q), L.ptr(k), L.ptr(gk), L.ptr(dq), L.ptr(dk),
                                  L.ptr(dgk_local), L.ptr(totals), B, H, T, K, int(bthd), scale, st), "lina_gla_bwd_post")
    carry = totals.flip(2).cumsum(2).flip(2) - totals                     # sum over the LATER chunks, [B,H,NT,K]
    if dht32 is not None:
        carry = carry + (dht32 * torch.cat(ST, dim=-1)).sum(-1).unsqueeze(2)
    carry = carry.contiguous()
    L.check(lib.lina_gla_bwd_dgk_finish(L.ptr(dgk_local), L.ptr(carry), L.ptr(dgk), B, H, T, K, int(bthd), st),
            "lina_gla_bwd_dgk_finish")
    L.count_launches(2)
    if bthd:
        dq, dk, dgk, dv = (x.transpose(1, 2) for x in (dq, dk, dgk, dv))
    return dq, dk, dv, dgk, dh0


def _bwd_tc_reference(q, k, v, gk, h0, do, dht, scale: float, want_dh0: bool, run=None):
    """Chunked backward (dq, dk, dv, dgk, dh0) as FIVE runs of the pre-gated forward kernel (C = 64; G = in-chunk cumsum of
    gk, D = e^{G_C}; q^ = scale q e^{G-G_C}, k^ = k e^{G_C-G}, k~ = k e^{-G}; "rev" = time-reversed):

      dv, dh0 = kernel(qg = rev k^, kg = rev q^, v = rev do, decay = rev D, h0 = dht)                     (key-dim decay)
      dq~    += kernel(qg = do[:, Vj], kg = v[:, Vj], v = k~, decay = D as ROW decay, h0 = h0[:, :, :, Vj]^T)   per V piece j
      dk^    += kernel(qg = rev v[:, Vj], kg = rev do[:, Vj], v = rev q^, decay = rev D (ROW), h0 = dht[..., Vj]^T)
      dq = dq~ * scale e^G ;  dk = rev(dk^) * e^{G_C-G} ;  dgk = reversed cumsum_T(dq q - dk k) [+ sum_v dht S_T]

    (the identities of FLA/fla/ops/gla/chunk.py:140-341 / FLA/fla/ops/common/chunk_h.py:111-189 regrouped so that every
    contraction is the forward kernel's; verified against the recurrence's explicit backward in tests/test_host.py with the
    oracle's restatement of the kernel contract as ``run``).

    This torch-glue version is the readable statement of the scheme and what the CPU host-logic test exercises (with
    ``run`` = the oracle's restatement of the kernel contract); the product path is ``_bwd_tc`` below, which does the same
    with four fused element-wise kernels (csrc/gla_bwd_glue.cu) and strided operand reads instead of ~25 torch ops."""
    run = run or _run_pregated
    q, k, v, gk = (x.contiguous() for x in (q, k, v, gk))
    B, H, T, K = q.shape
    V = v.shape[-1]
    C = 64
    NT = (T + C - 1) // C
    Tp, pad = NT * C, NT * C - T
    lo = q.dtype

    def padT(x):
        return torch.nn.functional.pad(x, (0, 0, 0, pad)) if pad else x

    qf, kf, gf = (padT(x.float()) for x in (q, k, gk))
    vb, dob = padT(v).contiguous(), padT(do).contiguous()
    G = gf.view(B, H, NT, C, K).cumsum(3)
    GC = G[:, :, :, -1:, :]
    D = GC.squeeze(3).exp().contiguous()                                   # [B,H,NT,K]
    Dr = D.flip(2).contiguous()
    kt = (kf.view(B, H, NT, C, K) * (-G).exp()).to(lo).view(B, H, Tp, K)
    qh = (qf.view(B, H, NT, C, K) * ((G - GC).exp() * scale)).to(lo).view(B, H, Tp, K)
    e_gc_g = (GC - G).exp()
    kh = (kf.view(B, H, NT, C, K) * e_gc_g).to(lo).view(B, H, Tp, K)
    qh_r = qh.flip(2).contiguous()
    dht32 = dht.float().contiguous() if dht is not None else None

    dv_r = torch.empty(B, H, Tp, V, dtype=lo, device=q.device)
    dh0 = torch.empty(B, H, K, V, dtype=torch.float32, device=q.device) if want_dh0 else None
    run(kh.flip(2).contiguous(), qh_r, dob.flip(2).contiguous(), Dr, dht32, dv_r, dh0, False)
    dv = dv_r.flip(2)[:, :, :T]

    ns = (V + 255) // 256
    Vp = V // ns
    dqt = dkr = None
    ST = []
    for j in range(ns):
        sl = slice(j * Vp, (j + 1) * Vp)
        do_j, v_j = dob[..., sl].contiguous(), vb[..., sl].contiguous()
        h0_j = h0[..., sl].float().transpose(-1, -2).contiguous() if h0 is not None else None
        o = torch.empty(B, H, Tp, K, dtype=torch.float32, device=q.device)
        ht = torch.empty(B, H, Vp, K, dtype=torch.float32, device=q.device) if dht is not None else None
        run(do_j, v_j, kt, D, h0_j, o, ht, True)
        dqt = o if dqt is None else dqt.add_(o)
        if ht is not None:
            ST.append(ht.transpose(-1, -2))
        dht_j = dht32[..., sl].transpose(-1, -2).contiguous() if dht32 is not None else None
        o2 = torch.empty(B, H, Tp, K, dtype=torch.float32, device=q.device)
        run(v_j.flip(2).contiguous(), do_j.flip(2).contiguous(), qh_r, Dr, dht_j, o2, None, True)
        dkr = o2 if dkr is None else dkr.add_(o2)
    dq = (dqt.view(B, H, NT, C, K) * (G.exp() * scale)).view(B, H, Tp, K)[:, :, :T]
    dk = (dkr.flip(2).view(B, H, NT, C, K) * e_gc_g).view(B, H, Tp, K)[:, :, :T]
    dgk = (dq * q.float() - dk * k.float()).flip(2).cumsum(2).flip(2)
    if dht32 is not None:
        dgk = dgk + (dht32 * torch.cat(ST, dim=-1)).sum(-1).unsqueeze(2)
    return dq.to(lo), dk.to(lo), dv, dgk.to(lo), dh0


def _bwd(q, k, v, gk, h0, do, dht, scale: float, want_dh0: bool, allow_tc: bool = True):
    lib = L.lib()
    if allow_tc and _tc_bwd_eligible(q, v):
        if do.dtype != q.dtype:
            do = do.to(q.dtype)
        return _bwd_tc(q, k, v, gk, h0, do, dht, scale, want_dh0)
    q, k, v, gk = (x.contiguous() for x in (q, k, v, gk))     # the backward kernels read [B,H,T,D]
    B, H, T, K, V = q.shape[0], q.shape[1], q.shape[2], q.shape[3], v.shape[3]
    do = do.contiguous()
    if do.dtype != q.dtype:
        do = do.to(q.dtype)
    if dht is not None:
        dht = dht.contiguous().float()
    dq, dk, dgk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(gk), torch.empty_like(v)
    dh0 = torch.empty(B, H, K, V, dtype=torch.float32, device=q.device) if want_dh0 else None
    ws = torch.empty(int(lib.lina_gla_recurrent_bwd_workspace_bytes(B, H, T, K, V)), dtype=torch.uint8, device=q.device)
    rc = lib.lina_gla_recurrent_bwd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(gk), L.ptr(h0),
                                    L.dt(h0) if h0 is not None else 0, L.ptr(do), L.ptr(dht), L.ptr(dq), L.ptr(dk),
                                    L.ptr(dv), L.ptr(dgk), L.ptr(dh0), L.ptr(ws), B, H, T, K, V, L.dt(q), scale,
                                    L.stream(q))
    L.count_launches(3)
    L.check(rc, "lina_gla_recurrent_bwd")
    return dq, dk, dv, dgk, dh0


class _GLAFunction(torch.autograd.Function):
    """kind in {'recurrent','chunk','fused_chunk'}; mirrors FusedRecurrentFunction
    (FLA/fla/ops/common/fused_recurrent.py:261-343), ChunkGLAFunction (FLA/fla/ops/gla/chunk.py:343-450)
    and FusedChunkGLAFunction (FLA/fla/ops/gla/chunk_fuse.py:302-502)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, q, k, v, gk, scale, initial_state, output_final_state, kind):
        in_dtypes = (q.dtype, k.dtype, v.dtype, gk.dtype)
        L.require_cuda(q, k, v, gk, initial_state)
        served, ctx.gates_ok = (_route(kind, q, v, gk, L.lib().lina_gla_chunk_fwd_uses_tensor_cores)
                                if q.dtype == gk.dtype else (kind, None))
        q, k, v, gk, h0, bthd = _prep(q, k, v, gk, initial_state, served)
        o, ht = _fwd("recurrent" if served == "recurrent" else "chunk", q, k, v, gk, h0, scale, output_final_state, bthd)
        ctx.save_for_backward(q, k, v, gk, h0)
        ctx.scale, ctx.kind, ctx.in_dtypes = scale, kind, in_dtypes
        ctx.h0_dtype = initial_state.dtype if initial_state is not None else None
        ctx.set_materialize_grads(False)
        return o.to(in_dtypes[2]), ht

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, do, dht=None):
        q, k, v, gk, h0 = ctx.saved_tensors
        if do is None:
            do = torch.zeros_like(v)
        want_dh0 = h0 is not None and ctx.needs_input_grad[5] and ctx.kind != "fused_chunk"
        allow_tc = ctx.gates_ok
        if allow_tc is None:                                     # forward did not look at the gates (recurrent kind, ...)
            allow_tc = _gates_in_envelope(gk) if _tc_bwd_eligible(q, v) else False
        dq, dk, dv, dgk, dh0 = _bwd(q, k, v, gk, h0, do, dht, ctx.scale, want_dh0, allow_tc)
        dts = ctx.in_dtypes
        if dh0 is not None and ctx.h0_dtype is not None:
            dh0 = dh0.to(ctx.h0_dtype)
        return dq.to(dts[0]), dk.to(dts[1]), dv.to(dts[2]), dgk.to(dts[3]), None, dh0, None, None


def _scale(scale, K: int) -> float:
    return float(K) ** -0.5 if (scale is None or scale == -1) else float(scale)


def fused_recurrent_gla(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, gk: torch.Tensor = None,
                        gv: torch.Tensor = None, scale: Optional[float] = None,
                        initial_state: torch.Tensor = None, output_final_state: bool = False,
                        reverse: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """FLA/fla/ops/gla/recurrent_fuse.py:13-27.  ``gk=None`` is the ungated recurrence (zero log-gates);
    ``reverse=True`` runs the recurrence from t = T-1 down to 0 (common/fused_recurrent.py:44-51: pointers start at the
    last step and walk backwards) -- served by the same kernels on time-flipped operands, the flips being ordinary
    differentiable torch ops.  ``gv`` (value-side gates) is a generic fla option Lina never uses (model/gla.py:187-203);
    it raises instead of silently differing."""
    if gv is not None:
        raise NotImplementedError("lina_speech_b200.fused_recurrent_gla: value-side gates (gv) are not implemented; "
                                  "model/gla.py only uses the (q, k, v, gk) form")
    if gk is None:
        gk = torch.zeros_like(q)
    if reverse:
        q, k, v, gk = (x.flip(2) for x in (q, k, v, gk))
    o, ht = _GLAFunction.apply(q, k, v, gk, _scale(scale, q.shape[-1]), initial_state, output_final_state, "recurrent")
    return (o.flip(2) if reverse else o), ht


def fused_chunk_gla(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, g: torch.Tensor, scale: float = -1,
                    initial_state: torch.Tensor = None, output_final_state: bool = False
                    ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """FLA/fla/ops/gla/chunk_fuse.py:518-536 -- detaches ``initial_state`` (:529-530); any T (the
    reference pads to a multiple of 16, :505-511,:531; the kernels here mask instead)."""
    if initial_state is not None:
        initial_state = initial_state.detach()
    return _GLAFunction.apply(q, k, v, g, _scale(scale, q.shape[-1]), initial_state, output_final_state, "fused_chunk")


def chunk_gla(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, g: torch.Tensor, scale: Optional[float] = None,
              initial_state: torch.Tensor = None, output_final_state: bool = False,
              checkpoint_level: Optional[int] = 2) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """FLA/fla/ops/gla/chunk.py:453-491 (``checkpoint_level`` is accepted and validated; nothing is
    cached between forward and backward here, which is level 2's behaviour)."""
    assert checkpoint_level in [0, 1, 2]
    return _GLAFunction.apply(q, k, v, g, _scale(scale, q.shape[-1]), initial_state, output_final_state, "chunk")


def _rwkv6_through_gla(r, k, v, w, u, scale: float, initial_state, output_final_state: bool, kind: str):
    """RWKV6 with gradients, as the GLA operator on shifted queries plus the bonus term:
      o_t = scale r_t S_{t-1} + scale (r_t . u . k_t) v_t,   S_t = e^{w_t} S_{t-1} + k_t^T v_t
    and GLA with q'_s = r_{s+1} returns scale r_{s+1} S_s at step s, i.e. o's state part shifted by one step; step 0 reads
    the initial state.  Everything around the op is differentiable torch, so dr, dk, dv, dw, du and dh0 come from the GLA
    backward (tensor cores for bf16 at Lina's head sizes).  FLA/fla/ops/rwkv6/recurrent_naive.py:8-42 is the spec."""
    B, H, T, K = r.shape
    q_shift = torch.cat([r[:, :, 1:], r.new_zeros(B, H, 1, K)], dim=2)
    o_g, ht = _GLAFunction.apply(q_shift, k, v, w, scale, initial_state, output_final_state, kind)
    if initial_state is not None:
        first = scale * torch.einsum("bhk,bhkv->bhv", r[:, :, 0].float(), initial_state.float()).unsqueeze(2)
    else:
        first = torch.zeros(B, H, 1, v.shape[-1], dtype=torch.float32, device=v.device)
    bonus = scale * (r.float() * u.float()[None, :, None, :] * k.float()).sum(-1, keepdim=True) * v.float()
    o = torch.cat([first, o_g[:, :, :-1].float()], dim=2) + bonus
    return o.to(v.dtype), ht


def fused_recurrent_rwkv6(r: torch.Tensor, k: torch.Tensor, v: torch.Tensor, w: torch.Tensor, u: torch.Tensor,
                          scale: float = -1, initial_state: torch.Tensor = None, output_final_state: bool = False
                          ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """FLA/fla/ops/rwkv6/recurrent_fuse.py:335-368: ``w`` are log-space decays, ``u`` [H,K] the bonus.  Without autograd one
    dedicated kernel; when a gradient is required the op runs as GLA on shifted queries (:func:`_rwkv6_through_gla`)."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (r, k, v, w, u, initial_state)):
        return _rwkv6_through_gla(r, k, v, w, u, _scale(scale, r.shape[-1]), initial_state, output_final_state, "recurrent")
    L.require_cuda(r, k, v, w, u, initial_state)
    odt = v.dtype
    if not (r.dtype == k.dtype == v.dtype == w.dtype == u.dtype):
        r, k, v, w, u = (x.float() for x in (r, k, v, w, u))
    r, k, v, w, u = (x.contiguous() for x in (r, k, v, w, u))
    B, H, T, K = r.shape
    V = v.shape[-1]
    h0 = initial_state.contiguous() if initial_state is not None else None
    o = torch.empty_like(v)
    ht = torch.empty(B, H, K, V, dtype=torch.float32, device=r.device) if output_final_state else None
    rc = L.lib().lina_rwkv6_recurrent_fwd(L.ptr(r), L.ptr(k), L.ptr(v), L.ptr(w), L.ptr(u), L.ptr(h0),
                                          L.dt(h0) if h0 is not None else 0, L.ptr(o), L.ptr(ht), B, H, T, K, V,
                                          L.dt(r), _scale(scale, K), L.stream(r))
    L.count_launches(1)
    L.check(rc, "lina_rwkv6_recurrent_fwd")
    return o.to(odt), ht


def chunk_rwkv6(r, k, v, g, u, scale: float = -1, initial_state=None, output_final_state: bool = False,
                checkpoint_level: Optional[int] = 0):
    """FLA/fla/ops/rwkv6/chunk.py:803- : same function as the recurrent form.  Inference: the same dedicated kernel; with
    autograd: the chunkwise GLA operator on shifted queries."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (r, k, v, g, u, initial_state)):
        return _rwkv6_through_gla(r, k, v, g, u, _scale(scale, r.shape[-1]), initial_state, output_final_state, "chunk")
    return fused_recurrent_rwkv6(r, k, v, g, u, scale, initial_state, output_final_state)

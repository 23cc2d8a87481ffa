"""Host side of ``lina_gemm_bf16_terms`` (csrc/gemm_sm100.cu): fp32 tensors carried as bf16 parts, contractions as term lists.

``split(x, parts)`` writes x = p0 + p1 (+ p2) with p0 = bf16(x), p_i = bf16(remainder): 2 parts keep 16 significand bits,
3 parts all 24.  ``TERMS[parts]`` lists the part products a GEMM accumulates for that fidelity.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from .. import _lib as L

# (a part, b part) products, most significant first
TERMS = {1: ((0, 0),), 2: ((0, 0), (0, 1), (1, 0)), 3: ((0, 0), (0, 1), (1, 0), (0, 2), (2, 0), (1, 1))}
ACT = {None: 0, "none": 0, "gelu": 1, "swish": 2}
# K blocks (of 64) between two promotions of the tensor core's partial sums to fp32 registers; 0 = the library's default
DEFAULT_SPAN = int(os.environ.get("LINA_GEMM_SPAN", "0"))


def split(x: torch.Tensor, parts: int) -> Tuple[torch.Tensor, ...]:
    """fp32 -> ``parts`` bf16 tensors summing to x (to 8 * parts significand bits).  Torch ops: used for WEIGHTS (once per
    checkpoint); activations are split by the kernels that produce them."""
    out, r = [], x.float()
    for _ in range(parts):
        p = r.to(torch.bfloat16)
        out.append(p.contiguous())
        r = r - p.float()
    return tuple(out)


def gemm_terms(a: Sequence[torch.Tensor], b: Sequence[torch.Tensor], *, NB: int, Ln: int, N: int, K: int, taps: int = 1,
               pad: int = 0, terms=None, b_batched: bool = False, alpha: float = 1.0, bias=None, gamma=None, residual=None,
               act=None, out_f32: bool = True, out_parts: int = 0, lda: Optional[int] = None, ldb: Optional[int] = None,
               a_batch_stride: int = 0, b_batch_stride: int = 0, out: Optional[torch.Tensor] = None,
               ld_split: Optional[int] = None, b_mn: bool = False, span: int = 0):
    """One launch of the tensor-core contraction.  ``a``: bf16 parts viewed as [NB, Ln, lda]; ``b``: bf16 parts [N, ldb] or
    [NB, N, ldb] (``b_mn``: stored transposed, [K, ldb] / [NB, K, ldb]).  Returns (fp32 [NB, Ln, N] or None, tuple of bf16 parts [NB, Ln, ld_split])."""
    dev = a[0].device
    L.require_cuda(*a, *b)
    if terms is None:
        terms = TERMS[min(len(a), len(b))] if len(a) == len(b) else None
    if terms is None:
        raise ValueError("pass `terms` when A and B have different numbers of parts")
    g = L.GemmArgs()
    for i, t in enumerate(a):
        if t.dtype != torch.bfloat16:
            raise TypeError("gemm_terms: operand parts must be bf16")
        g.a[i] = t.data_ptr()
    for i, t in enumerate(b):
        if t.dtype != torch.bfloat16:
            raise TypeError("gemm_terms: operand parts must be bf16")
        g.b[i] = t.data_ptr()
    g.a_parts, g.b_parts = len(a), len(b)
    g.lda = lda if lda is not None else a[0].stride(-2)
    g.ldb = ldb if ldb is not None else b[0].stride(-2)
    g.a_batch_stride, g.b_batch_stride = a_batch_stride, b_batch_stride
    g.b_batched = int(b_batched)
    g.b_mn = int(b_mn)
    g.span = int(span) if span else DEFAULT_SPAN
    g.n_terms = len(terms)
    for i, (pa, pb) in enumerate(terms):
        g.term_a[i], g.term_b[i] = pa, pb
    g.NB, g.L, g.N, g.K, g.taps, g.pad = NB, Ln, N, K, taps, pad
    g.alpha = alpha
    keep = []
    for name, t in (("bias", bias), ("gamma", gamma), ("residual", residual)):
        if t is not None:
            if t.dtype != torch.float32 or not t.is_cuda:
                raise TypeError(f"gemm_terms: {name} must be a CUDA fp32 tensor")
            keep.append(t)
            setattr(g, name, t.data_ptr())
    if residual is not None:
        g.ld_res = residual.stride(-2)
    g.act = ACT[act]
    o32 = None
    if out is not None:
        o32 = out
    elif out_f32:
        o32 = torch.empty(NB, Ln, N, dtype=torch.float32, device=dev)
    if o32 is not None:
        g.out_f32, g.ld_out = o32.data_ptr(), o32.stride(-2)
    parts = ()
    if out_parts:
        ls = ld_split if ld_split is not None else (N + 7) // 8 * 8
        parts = tuple(torch.empty(NB, Ln, ls, dtype=torch.bfloat16, device=dev) for _ in range(out_parts))
        for i, t in enumerate(parts):
            g.out_split[i] = t.data_ptr()
        g.ld_split, g.out_parts = ls, out_parts
    rc = L.lib().lina_gemm_bf16_terms(C.byref(g), L.stream(a[0]))
    L.count_launches(1)
    L.check(rc, "lina_gemm_bf16_terms")
    return o32, parts

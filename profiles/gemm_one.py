"""One GEMM configuration in a loop (for ncu): python profiles/gemm_one.py <parts> <K> <N> <variant> [reps]"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lina_speech_b200.codec import gemm as G

parts, K, N, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dev, B, Ln = "cuda", 32, 750
torch.manual_seed(0)
x = G.split(torch.randn(B, Ln, K, device=dev), parts)
w = G.split(torch.randn(N, K, device=dev) / K ** 0.5, parts)
bias = torch.randn(N, device=dev)
resid = torch.randn(B, Ln, N, device=dev)
kw = {"f32": dict(), "res": dict(bias=bias, residual=resid), "gelu": dict(bias=bias, act="gelu"),
      "parts": dict(bias=bias, out_f32=False, out_parts=parts),
      "geluparts": dict(bias=bias, act="gelu", out_f32=False, out_parts=parts)}[variant]
for _ in range(reps):
    G.gemm_terms(x, w, NB=B, Ln=Ln, N=N, K=K, **kw)
torch.cuda.synchronize()

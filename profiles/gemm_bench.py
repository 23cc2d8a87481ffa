"""lina_gemm_bf16_terms in isolation at the ConvNeXt point-wise shapes (M = 32 x 750, K/N = 768 / 2304), epilogue variants."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lina_speech_b200.codec import gemm as G

dev = "cuda"
torch.manual_seed(0)
B, Ln = 32, 750
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, reps=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


res = {}
for parts in (1, 2, 3):
    for (K, N) in ((768, 2304), (2304, 768)):
        x = G.split(torch.randn(B, Ln, K, device=dev), parts)
        w = G.split(torch.randn(N, K, device=dev) / K ** 0.5, parts)
        bias = torch.randn(N, device=dev)
        resid = torch.randn(B, Ln, N, device=dev)
        flops = 2 * B * Ln * N * K * len(G.TERMS[parts])
        for name, kw in (("f32", dict()), ("f32+bias+res", dict(bias=bias, residual=resid)), ("gelu->f32", dict(bias=bias, act="gelu")),
                         ("parts_only", dict(bias=bias, out_f32=False, out_parts=parts)),
                         ("gelu->parts", dict(bias=bias, act="gelu", out_f32=False, out_parts=parts))):
            ms = bench(lambda: G.gemm_terms(x, w, NB=B, Ln=Ln, N=N, K=K, **kw))
            res[f"parts{parts} K{K} N{N} {name}"] = {"ms": round(ms, 4), "tflops_bf16": round(flops / ms / 1e9, 1)}
            print(f"parts{parts} K{K} N{N} {name:14s} {ms:.4f} ms  {flops / ms / 1e9:.0f} TF/s", flush=True)
# cuBLAS bf16 at the same shape, for scale
for (K, N) in ((768, 2304), (2304, 768)):
    a = torch.randn(B * Ln, K, device=dev, dtype=torch.bfloat16)
    wt = torch.randn(N, K, device=dev, dtype=torch.bfloat16)
    ms = bench(lambda: torch.nn.functional.linear(a, wt))
    print(f"cuBLAS bf16 K{K} N{N}: {ms:.4f} ms {2 * B * Ln * N * K / ms / 1e9:.0f} TF/s")
    res[f"cublas bf16 K{K} N{N}"] = {"ms": round(ms, 4), "tflops_bf16": round(2 * B * Ln * N * K / ms / 1e9, 1)}
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)

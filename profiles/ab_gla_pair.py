"""A/B of the pre-gated GLA kernel at the bench shape (B=32, T=2048, H=4, K=256, V=512, bf16) and two smaller ones:
round-1 one-CTA-per-tile kernel (variant key 9 = 1), CTA pairs sharing the score MMA (key 9 = 2: no T cut), pairs + T cut
(default, with workspace).  CUDA events per launch, L2 flushed between launches, median of n."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lina_speech_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
out = sys.argv[2] if len(sys.argv) > 2 else None
lib = L.lib()
dev, bf = "cuda", torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
for (B, T, H, K, V) in ((32, 2048, 4, 256, 512), (8, 4096, 4, 256, 512), (2, 1024, 4, 256, 512)):
    torch.manual_seed(0)
    nt = (T + 63) // 64
    qg = (torch.randn(B, T, H, K, device=dev) * 0.5).to(bf)
    kg = (torch.randn(B, T, H, K, device=dev) * 0.5).to(bf)
    v = torch.randn(B, T, H, V, device=dev).to(bf)
    decay = torch.rand(B, H, nt, K, device=dev) * 0.5 + 0.5
    o = torch.empty(B, T, H, V, dtype=bf, device=dev)
    ws_bytes = lib.lina_gla_chunk_fwd_pregated_ws_bytes(B, H, T, K, V)
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    nbytes = B * H * T * (2 * K + 2 * V) * 2 + decay.numel() * 4
    flops = B * H * T * (4 * K * V + 64 * (K + V))
    ref = None
    for name, key9, use_ws in (("one CTA per tile (round 1)", 1, False), ("CTA pairs", 2, False), ("CTA pairs + T cut", 0, True)):
        lib.lina_debug_set_variant(9, key9)
        def run():
            rc = lib.lina_gla_chunk_fwd_pregated_bthd_ws(L.ptr(qg), L.ptr(kg), L.ptr(v), L.ptr(decay), None, 0, L.ptr(o), None,
                                                         L.ptr(ws) if use_ws else None, ws_bytes if use_ws else 0, B, H, T, K, V, st)
            assert rc == 0, lib.lina_last_error_string()
        run(); run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        lib.lina_debug_set_variant(9, 0)
        ts.sort()
        ms = ts[len(ts) // 2]
        same = None
        if ref is None:
            ref = o.clone()
        else:
            same = bool(torch.equal(o, ref))
        key = f"B{B} T{T} {name}"
        res[key] = {"ms": round(ms, 4), "min_ms": round(ts[0], 4), "GBps_algorithmic": round(nbytes / ms / 1e6, 1),
                    "TFLOPs": round(flops / ms / 1e9, 1), "bit_identical_to_round1": same, "ws_bytes": ws_bytes if use_ws else 0}
        print(key, res[key], flush=True)
if out:
    json.dump(res, open(out, "w"), indent=1)

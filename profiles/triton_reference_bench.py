"""GPU-side reference timing (SURVEY 0 / BASELINE.md 3, VERDICT r1 "missing" #5): the reference's own Triton GLA ops, JIT-compiled
on the B200 under the installed Triton, timed beside this repo's kernels on identical inputs -- the shapes of the reference's
FLA/benchmarks/ops/benchmark_gla.py:13-82 (B16 H8 D128 bf16, T = 128 ... 16384, gates logsigmoid(N(0,1)).clamp_min(-5)) plus the
flagship shape (B32 H4 T2048 K256 V512, gates logsigmoid/16).

The vendored fla tree is NOT part of this repository: stage it in the git-ignored scratch directory before the gpurun call
(`cp -r /root/reference/3rdparty/flash-linear-attention/fla baseline/_ref/fla`).  The package __init__ files (which import HF
model classes) are bypassed by registering namespace stubs; only fla/ops/gla, fla/ops/common, fla/ops/utils.py, fla/utils.py run.

Also dumps the reference's bf16 outputs on seeded inputs (`--golden out.npz`): outputs of the reference itself run on the box,
committed as tests/golden/gla_triton_bf16.npz and compared with our kernels by tests/test_flagship_parity_gpu.py.
"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "baseline", "_ref")


def load_reference_ops():
    if not os.path.isdir(os.path.join(REF, "fla", "ops", "gla")):
        return None, "baseline/_ref/fla not staged"
    for name, sub in (("fla", ""), ("fla.ops", "ops"), ("fla.modules", "modules"), ("fla.models", "models")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, "fla", sub)] if sub else [os.path.join(REF, "fla")]
        sys.modules[name] = m
    try:
        import importlib
        gla = importlib.import_module("fla.ops.gla")
        return gla, None
    except Exception as e:      # noqa: BLE001
        return None, f"import failed: {e!r}"


def bench(fn, warmup=3, reps=10):
    """median ms of fn() with CUDA events; L2 flushed between repetitions."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def inputs(B, H, T, K, V, gates, seed=0, grad=False):
    g = torch.Generator().manual_seed(seed)
    q, k = (torch.randn(B, H, T, K, generator=g).bfloat16().cuda() for _ in range(2))
    v = torch.randn(B, H, T, V, generator=g).bfloat16().cuda()
    x = torch.randn(B, H, T, K, generator=g)
    gk = (F.logsigmoid(x).clamp_min(-5) if gates == "fla" else F.logsigmoid(x) / 16).bfloat16().cuda()
    if grad:
        for t in (q, k, v, gk):
            t.requires_grad_(True)
    return q, k, v, gk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "triton_reference_bench.json"))
    ap.add_argument("--golden", default=os.path.join(ROOT, "gpurun_out", "gla_triton_bf16.npz"))
    ap.add_argument("--max-t", type=int, default=16384)
    a = ap.parse_args()
    import triton
    from lina_speech_b200.fla_api import ops as ours
    ref, why = load_reference_ops()
    res = {"triton": triton.__version__, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0),
           "reference": "3rdparty/flash-linear-attention fla 0.1 (vendored), JIT-compiled here" if ref else None,
           "reference_unavailable": why, "rows": []}
    shapes = [(16, 8, T, 128, 128, "fla") for T in [128 * 2 ** i for i in range(8)] if T <= a.max_t]
    shapes += [(32, 4, 2048, 256, 512, "lina"), (8, 4, 4096, 256, 512, "lina"), (2, 4, 1024, 256, 512, "lina")]
    providers = {"ours.fused_chunk_gla": ours.fused_chunk_gla, "ours.chunk_gla": ours.chunk_gla,
                 "ours.fused_recurrent_gla": ours.fused_recurrent_gla}
    if ref is not None:
        providers.update({"ref.fused_chunk_gla": ref.fused_chunk_gla, "ref.chunk_gla": ref.chunk_gla,
                          "ref.fused_recurrent_gla": ref.fused_recurrent_gla})
    for (B, H, T, K, V, gates) in shapes:
        row = {"B": B, "H": H, "T": T, "K": K, "V": V, "gates": gates, "fwd_ms": {}, "fwd_bwd_ms": {}, "errors": {}}
        q, k, v, gk = inputs(B, H, T, K, V, gates)
        for name, fn in providers.items():
            if "recurrent" in name and B * H * T > 16 * 8 * 4096:
                continue                                   # serial in T: minutes at the long shapes, not informative
            try:
                with torch.no_grad():
                    row["fwd_ms"][name] = bench(lambda: fn(q, k, v, gk))
            except Exception as e:      # noqa: BLE001
                row["errors"][name] = repr(e)[:200]
        qg, kg_, vg, gg = inputs(B, H, T, K, V, gates, grad=True)
        do = torch.ones_like(vg)
        for name, fn in providers.items():
            if "recurrent" in name and B * H * T > 16 * 8 * 2048:
                continue
            if name in row["errors"]:
                continue
            try:
                def fb():
                    o = fn(qg, kg_, vg, gg)
                    o = o[0] if isinstance(o, tuple) else o
                    o.backward(do)
                row["fwd_bwd_ms"][name] = bench(fb, warmup=2, reps=5)
            except Exception as e:      # noqa: BLE001
                row["errors"][name + "_bwd"] = repr(e)[:200]
        print(json.dumps(row), flush=True)
        res["rows"].append(row)
        del q, k, v, gk, qg, kg_, vg, gg
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)

    # golden vectors from the reference's Triton path in bf16 (identical seeded inputs are rebuilt by the test)
    if ref is not None:
        gold = {}
        for tag, (B, H, T, K, V, gates) in {"a": (1, 1, 128, 256, 512, "lina"), "b": (1, 2, 192, 128, 128, "fla"),
                                            "c": (2, 1, 256, 256, 512, "lina")}.items():
            q, k, v, gk = inputs(B, H, T, K, V, gates, seed=100 + ord(tag))
            for opname in ("fused_chunk_gla", "chunk_gla"):
                try:
                    with torch.no_grad():
                        o, ht = getattr(ref, opname)(q, k, v, gk, output_final_state=True)
                    gold[f"{tag}_{opname}_o"] = o.float().cpu().numpy().astype(np.float32)
                    gold[f"{tag}_{opname}_ht"] = ht.float().cpu().numpy()
                except Exception as e:      # noqa: BLE001
                    print("golden", tag, opname, "failed:", repr(e)[:200], flush=True)
            gold[f"{tag}_shape"] = np.array([B, H, T, K, V, 0 if gates == "lina" else 1, 100 + ord(tag)])
        np.savez_compressed(a.golden, **gold)
        print("golden written:", sorted(gold), flush=True)


if __name__ == "__main__":
    main()

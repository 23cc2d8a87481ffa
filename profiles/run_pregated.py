"""The pre-gated GLA path alone at the bench shape (B=32, T=2048, H=4, K=256, V=512, bf16): lina_gla_prefill_prep_gated
(v conv kernel + q/k gate kernel) followed by lina_gla_chunk_fwd_pregated_bthd_ws -- the command ncu wraps for the per-kernel
captures under profiles/.  Prints CUDA-event time per launch of each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lina_speech_b200 import _lib as L

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
B, T, H, K, V = 32, 2048, 4, 256, 512
kd, vd = H * K, H * V
ldx = 2 * kd + 2 * vd
dev, bf = "cuda", torch.bfloat16
lib = L.lib()
torch.manual_seed(0)
proj = torch.randn(B, T, ldx, device=dev).to(bf)
gk_raw = torch.randn(B, T, kd, device=dev).to(bf)
wq, wk = (torch.randn(kd, 4, device=dev).mul(0.5).to(bf) for _ in range(2))
wv = torch.randn(vd, 4, device=dev).mul(0.5).to(bf)
q, k = (torch.empty(B, T, kd, dtype=bf, device=dev) for _ in range(2))
v = torch.empty(B, T, vd, dtype=bf, device=dev)
decay = torch.empty(B, H, (T + 63) // 64, K, dtype=torch.float32, device=dev)
o = torch.empty(B, T, H, V, dtype=bf, device=dev)
xq, xk, xv = proj[..., :kd], proj[..., kd:2 * kd], proj[..., 2 * kd:2 * kd + vd]
st = torch.cuda.current_stream().cuda_stream


def prep():
    rc = lib.lina_gla_prefill_prep_gated(L.ptr(xq), ldx, L.ptr(xk), ldx, L.ptr(xv), ldx, L.ptr(wq), L.ptr(wk), L.ptr(wv),
                                         L.ptr(gk_raw), kd, L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(decay), None, None, None, 0,
                                         B, T, H, K, V, 4, 16.0, K ** -0.5, None, st)
    assert rc == 0, lib.lina_last_error_string()


ws_bytes = lib.lina_gla_chunk_fwd_pregated_ws_bytes(B, H, T, K, V)
ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev)


def gla():
    rc = lib.lina_gla_chunk_fwd_pregated_bthd_ws(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(decay), None, 0, L.ptr(o), None, L.ptr(ws), ws_bytes,
                                                 B, H, T, K, V, st)
    assert rc == 0, lib.lina_last_error_string()


for fn, name, nbytes in ((prep, "prefill_prep_gated", (2 * (3 * kd + vd) + 2 * (2 * kd + vd)) * B * T),
                         (gla, "gla_chunk_fwd_pregated", B * H * T * (2 * K + 2 * V) * 2 + decay.numel() * 4)):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name}: {ms:.4f} ms/launch, {nbytes / ms / 1e6:.1f} GB/s algorithmic")

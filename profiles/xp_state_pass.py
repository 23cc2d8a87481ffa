"""(Historical: the kernel-side knob this drives was removed after the measurement -- numbers in DESIGN 4.1c; to repeat it, re-add
`xp` to SplitArgs and guard the three tcgen05.ld / st of the state pass with its bits.)
Timing experiment (results are WRONG on purpose): which part of the state pass of the pair kernel costs time?
variant key 10 bits: 1 = skip the fp32 write-back of ST, 2 = skip the bf16 SA write, 4 = skip the ST read."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lina_speech_b200 import _lib as L
lib = L.lib()
dev, bf = "cuda", torch.bfloat16
B, T, H, K, V = 32, 2048, 4, 256, 512
nt = T // 64
qg = (torch.randn(B, T, H, K, device=dev) * 0.5).to(bf); kg = (torch.randn(B, T, H, K, device=dev) * 0.5).to(bf)
v = torch.randn(B, T, H, V, device=dev).to(bf); decay = torch.rand(B, H, nt, K, device=dev) * 0.5 + 0.5
o = torch.empty(B, T, H, V, dtype=bf, device=dev)
wsb = lib.lina_gla_chunk_fwd_pregated_ws_bytes(B, H, T, K, V); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for xp in (0, 1, 2, 3, 4, 7, 0):
    lib.lina_debug_set_variant(10, xp)
    ts = []
    for i in range(12):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.lina_gla_chunk_fwd_pregated_bthd_ws(L.ptr(qg), L.ptr(kg), L.ptr(v), L.ptr(decay), None, 0, L.ptr(o), None, L.ptr(ws), wsb, B, H, T, K, V, st)
        e1.record(); torch.cuda.synchronize()
        assert rc == 0
        if i >= 2: ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"xp={xp}: median {ts[len(ts)//2]:.4f} ms  min {ts[0]:.4f}", flush=True)
lib.lina_debug_set_variant(10, 0)

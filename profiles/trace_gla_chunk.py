"""Print the per-role timeline (clock64, CTA (0,0)) of the tcgen05 GLA chunk kernel at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from lina_speech_b200 import _lib as L

B, H, T, K, V = 32, 4, 2048, 256, 512
torch.manual_seed(0)
q, k = (torch.randn(B, H, T, K, device="cuda", dtype=torch.bfloat16) for _ in range(2))
v = torch.randn(B, H, T, V, device="cuda", dtype=torch.bfloat16)
gk = (F.logsigmoid(torch.randn(B, H, T, K, device="cuda")) / 16).bfloat16()
o = torch.empty_like(v)
tr = torch.zeros(6, 64, 4, dtype=torch.int64, device="cuda")
for _ in range(2):
    L.check(L.debug_lib().lina_debug_gla_chunk_trace(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(gk), L.ptr(o), B, H, T, K, V,
                                                 K ** -0.5, L.ptr(tr), L.stream(q)), "trace")
torch.cuda.synchronize()
t = tr.cpu()
t0 = int(t[t > 0].min())
names = ["prep", "load", "mma", "mask", "epi", "corr"]
ev = {"prep": ["start", "end"], "load": ["start", "issued", "landed"], "mma": ["(0)", "(1)", "(2)", "done"],
      "mask": ["start", "end"], "epi": ["start", "end"], "corr": ["start", "end"]}
for n in list(range(0, 4)) + list(range(14, 24)) + [30, 31]:
    parts = []
    for r, nm in enumerate(names):
        vals = [int(t[r, n, e]) - t0 for e in range(len(ev[nm]))]
        parts.append(nm + ":" + "/".join(str(x) for x in vals))
    print(f"n={n:2d}  " + "  ".join(parts))
d = lambda r, e1, e0: (t[r, 8:30, e1] - t[r, 8:30, e0]).float().mean().item()
print("mean cycles (items 8..29): prep %.0f  load issue %.0f landed %.0f  mma (0)->(1) %.0f (1)->(2) %.0f (2)->done %.0f  "
      "mask %.0f  epi %.0f  corr %.0f" % (d(0, 1, 0), d(1, 1, 0), d(1, 2, 0), d(2, 1, 0), d(2, 2, 1), d(2, 3, 2),
                                          d(3, 1, 0), d(4, 1, 0), d(5, 1, 0)))
print("item period (mma done n+1 - n): %.0f cycles" % (t[2, 9:30, 3] - t[2, 8:29, 3]).float().mean().item())

"""Per-role timeline (clock64, CTA (0,0)) of the PRE-GATED tcgen05 GLA kernel at the bench shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lina_speech_b200 import _lib as L

B, H, T, K, V = 32, 4, 2048, 256, 512
torch.manual_seed(0)
bf = torch.bfloat16
qg, kg = (torch.randn(B, T, H, K, device="cuda").mul(0.5).to(bf) for _ in range(2))
v = torch.randn(B, T, H, V, device="cuda").to(bf)
decay = torch.rand(B, H, T // 64, K, device="cuda") * 0.2 + 0.7
o = torch.empty(B, T, H, V, device="cuda", dtype=bf)
tr = torch.zeros(6, 64, 4, dtype=torch.int64, device="cuda")
for _ in range(2):
    L.check(L.debug_lib().lina_debug_gla_pregated_trace(L.ptr(qg), L.ptr(kg), L.ptr(v), L.ptr(decay), L.ptr(o), B, H, T, K, V,
                                                  L.ptr(tr), L.stream(qg)), "trace")
torch.cuda.synchronize()
t = tr.cpu()
t0 = int(t[t > 0].min())
names = ["prep", "load", "mma", "mask", "epi", "state"]
ev = {"prep": [], "load": ["start", "issued", "v issued"], "mma": ["(0)", "(1)", "(2)", "done"],
      "mask": ["start", "end"], "epi": ["start", "end"], "state": ["start", "end"]}
for n in list(range(0, 3)) + list(range(14, 20)) + [30, 31]:
    parts = []
    for r, nm in enumerate(names):
        if not ev[nm]:
            continue
        vals = [int(t[r, n, e]) - t0 for e in range(len(ev[nm]))]
        parts.append(nm + ":" + "/".join(str(x) for x in vals))
    print(f"n={n:2d}  " + "  ".join(parts))
d = lambda r, e1, e0: (t[r, 8:30, e1] - t[r, 8:30, e0]).float().mean().item()
print("mean cycles (items 8..29): load wait->issued %.0f  mma (0)->(1) %.0f (1)->(2) %.0f (2)->done %.0f  mask %.0f  epi %.0f  "
      "state %.0f" % (d(1, 1, 0), d(2, 1, 0), d(2, 2, 1), d(2, 3, 2), d(3, 1, 0), d(4, 1, 0), d(5, 1, 0)))
print("item period (mma done n+1 - n): %.0f cycles" % (t[2, 9:30, 3] - t[2, 8:29, 3]).float().mean().item())
print("mma (0) start n+1 - done n: %.0f ; state start - mma done (same n): %.0f ; mma (1) start n+1 - state end n: %.0f" % (
    (t[2, 9:30, 0] - t[2, 8:29, 3]).float().mean().item(), (t[5, 8:30, 0] - t[2, 8:30, 3]).float().mean().item(),
    (t[2, 9:30, 1] - t[5, 8:29, 1]).float().mean().item()))

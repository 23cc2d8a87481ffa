"""Round-2 bring-up (not yet run on hardware): which TMEM (lane, column) holds element (m, n) of an M = 64 accumulator?
A[m, 0] = m + 1, A[m, 1] = 1/256, B[n, 0] = 1, B[n, 1] = n  =>  D[m, n] = m + 1 + n / 256 exactly; untouched cells = -12345."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lina_speech_b200 import _lib as L

for M in (128, 64):
    N, KD = 64, 16
    A = torch.zeros(M, KD, device="cuda"); B = torch.zeros(N, KD, device="cuda")
    A[:, 0] = torch.arange(1, M + 1, device="cuda"); A[:, 1] = 1.0 / 256
    B[:, 0] = 1.0; B[:, 1] = torch.arange(N, device="cuda")
    D = torch.empty(128, N, device="cuda")
    L.check(L.debug_lib().lina_debug_umma_probe_m(L.ptr(A), L.ptr(B), L.ptr(D), M, N, KD, None), "probe_m")
    torch.cuda.synchronize()
    d = D.cpu()
    written = d > -12000
    m = (d.floor() - 1).long()
    n = ((d - d.floor()) * 256).round().long()
    print(f"M={M}: lanes written {sorted(set(written.any(1).nonzero().flatten().tolist()))[:8]}... count {int(written.any(1).sum())}")
    for lane in range(0, 128, 8):
        row = [(int(m[lane, c]), int(n[lane, c])) if written[lane, c] else None for c in (0, 1, 31, 32, 63)]
        print(f"  lane {lane:3d}: (m, n) at columns 0,1,31,32,63 = {row}")

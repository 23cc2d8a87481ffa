"""Per-role timeline (clock64, rank 0 of the first CTA pair) of the CTA-pair pre-gated GLA kernel at the bench shape.
MMA-thread events per item n: A = O_EMPTY passed, B = first state block ready ((1) starts), C = all (1) issued, D = v landed ((3) issued
next), E = before the P wait, F = P there ((2) issued next), G = item issued; S0/S1 = state pass start (= (3) complete) / end;
E0/E1 = epilogue start (= (2) complete) / end; M0/M1 = mask start (= (0) complete) / end (own items only); X = (0) issue start."""
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
    tr.zero_()
    L.check(L.debug_lib().lina_debug_gla_pregated_trace(L.ptr(qg), L.ptr(kg), L.ptr(v), L.ptr(decay), L.ptr(o), B, H, T, K, V,
                                                  L.ptr(tr), L.stream(qg)), "trace")
torch.cuda.synchronize()
t = tr.cpu()
t0 = int(t[t > 0].min())
def g(r, n, e):
    x = int(t[r, n, e])
    return x - t0 if x > 0 else -1
for n in list(range(0, 4)) + list(range(12, 22)) + [30, 31]:
    print(f"n={n:2d} X={g(2,n,0):7d} A={g(0,n,0):7d} B={g(2,n,1):7d} C={g(0,n,1):7d} D={g(0,n,2):7d} E={g(0,n,3):7d} F={g(2,n,2):7d} G={g(2,n,3):7d} | "
          f"S0={g(5,n,0):7d} S1={g(5,n,1):7d}/{g(5,n,2):7d}/{g(5,n,3):7d} | E0={g(4,n,0):7d} E1={g(4,n,1):7d} | M0={g(3,n,0):7d} M1={g(3,n,1):7d} | "
          f"L0={g(1,n,0):7d} L1={g(1,n,1):7d} Lv={g(1,n,2):7d}")
import statistics as st
per = [g(2, n + 1, 3) - g(2, n, 3) for n in range(8, 29)]
print("item period:", st.mean(per))
def mean(f):
    return st.mean(f(n) for n in range(8, 29))
print("A->B (wait first state block) %.0f  B->C ((1) issue incl. block waits) %.0f  C->D (v wait) %.0f  D->E ((3) + scores issue) %.0f  E->F (P wait) %.0f  F->G %.0f  G->A(n+1) %.0f" % (
    mean(lambda n: g(2, n, 1) - g(0, n, 0)), mean(lambda n: g(0, n, 1) - g(2, n, 1)), mean(lambda n: g(0, n, 2) - g(0, n, 1)),
    mean(lambda n: g(0, n, 3) - g(0, n, 2)), mean(lambda n: g(2, n, 2) - g(0, n, 3)), mean(lambda n: g(2, n, 3) - g(2, n, 2)),
    mean(lambda n: g(0, n + 1, 0) - g(2, n, 3))))
print("(3) issued (D) -> state start (S0) %.0f  state S0->S1 %.0f  state end S1 -> next B %.0f  (2) issued (F) -> epilogue start E0 %.0f  epilogue %.0f" % (
    mean(lambda n: g(5, n, 0) - g(0, n, 2)), mean(lambda n: g(5, n, 1) - g(5, n, 0)), mean(lambda n: g(2, n + 1, 1) - g(5, n, 1)),
    mean(lambda n: g(4, n, 0) - g(2, n, 2)), mean(lambda n: g(4, n, 1) - g(4, n, 0))))

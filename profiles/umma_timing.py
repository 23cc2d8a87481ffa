"""cycles per tcgen05.mma (M=128, K=16, bf16) vs N and operand placement, one CTA alone on the chip."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lina_speech_b200 import _lib as L
out = torch.zeros(6, dtype=torch.int64, device="cuda")
print("N  A      B   issuer      issue/mma  total/mma")
for N in (16, 64, 128, 256):
    for a_tmem, a_mn, b_mn, name in ((0, 0, 0, "smemK smemK"), (1, 0, 0, "tmem  smemK"), (0, 1, 0, "smemMN smemK"), (0, 1, 1, "smemMN smemMN")):
        for elect in (0, 1):          # bit 1 of the `same_d` argument selects the elect.sync issuer
            L.check(L.debug_lib().lina_debug_umma_timing(L.ptr(out), N, a_tmem, a_mn, b_mn, 64, 1 | (elect << 1), None), "timing")
            torch.cuda.synchronize()
            o = out.cpu().tolist()
            print(f"{N:3d} {name:14s} {'elect.sync' if elect else 'tid == 0  '}  {o[4]/64:8.1f} {o[5]/64:9.1f}")

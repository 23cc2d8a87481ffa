"""Launch the GLA chunk-forward op alone at the bench shape (B=32,H=4,T=2048,K=256,V=512, bf16) -- the
command ncu wraps for the per-kernel captures under profiles/.  Prints CUDA-event time per launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from lina_speech_b200.fla_api import fused_chunk_gla

B, H, T, K, V = 32, 4, 2048, 256, 512
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
torch.manual_seed(0)
q, k = (torch.randn(B, H, T, K, device="cuda", dtype=torch.bfloat16) for _ in range(2))
v = torch.randn(B, H, T, V, device="cuda", dtype=torch.bfloat16)
gk = (F.logsigmoid(torch.randn(B, H, T, K, device="cuda")) / 16).bfloat16()
for _ in range(2):
    fused_chunk_gla(q, k, v, gk)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    fused_chunk_gla(q, k, v, gk)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
byt = B * H * T * (3 * K + 2 * V) * 2
print(f"gla_chunk_fwd: {ms:.3f} ms/launch, {byt / ms / 1e6:.1f} GB/s algorithmic, {B*H*T*(4*K*V+64*(K+V))/ms/1e9:.1f} TFLOP/s")

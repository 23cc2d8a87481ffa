"""A few autoregressive steps of the bench model (bs32, no CUDA graph) -- the command ncu wraps for the decode
launch list (profiles/launches_r01_decode.csv)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lm = bench.build_model(torch.device("cuda"), torch.bfloat16)
x, _, _, _ = bench.synth_inputs(1, 8, 128, 0)
qs, *_ = lm.generate_batch(x[0].cuda(), batch_size=B, max_seqlen=steps, k=100, force_max_seqlen=True, cuda_graph=False)
torch.cuda.synchronize()
print("decoded", tuple(qs.shape))

"""A/B of decode-step variants in ONE process on one box (CUDA-graphed loop, 400 steps, bs 32 and 128)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import lina_speech_b200.model.base_blocks as BB
import lina_speech_b200.model.crossatt as CA

dev = torch.device("cuda", 0)
lm = bench.build_model(dev, torch.bfloat16)
x, _, _, _ = bench.synth_inputs(1, 8, 128, 0)
xt = x[0].to(dev)
res = {}
for rep in range(2):
    for name, ca, sk in (("default", True, False), ("crossatt_op_by_op", False, False), ("skinny_linears", True, True)):
        CA.FUSED_STEP, BB.SKINNY_STEP = ca, sk
        for B in (32, 128):
            if sk and B > 32:
                continue
            tm = {}
            lm.generate_batch(xt, batch_size=B, max_seqlen=8, k=100, force_max_seqlen=True, cuda_graph=True)
            lm.generate_batch(xt, batch_size=B, max_seqlen=400, k=100, force_max_seqlen=True, cuda_graph=True,
                              stop_check_interval=1 << 30, _timing=tm)
            torch.cuda.synchronize()
            ms = tm["start"].elapsed_time(tm["end"]) / tm["steps"]
            res.setdefault(f"{name} B{B}", []).append(round(ms, 4))
            print(name, B, ms, flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)

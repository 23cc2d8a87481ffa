"""BASELINE configs[3]: one training step (forward + backward + AdamW) of the bench model at bs 8 x seq 4096, text 256,
bf16 autocast over fp32 parameters, synthetic (text, token) pairs.  CUDA-event ms per step and tokens/s, with the GLA
backward on the tensor-core path (five runs of the pre-gated tcgen05 kernel) and on the CUDA-core recurrence kernels.
usage: train_step.py [out.json] [B] [T]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lina_speech_b200.fla_api import ops

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "train_step.json")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
T = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
Tx = 256
dev = torch.device("cuda")
lm = bench.build_model(dev, torch.float32).train()
opt = torch.optim.AdamW(lm.parameters(), lr=2e-4, betas=(0.9, 0.95), weight_decay=0.1)
x, y, em, cm = bench.synth_inputs(B, T, Tx, seed=7)
xd, yd, emd, cmd = x.to(dev), y.to(dev), em.to(dev), cm.to(dev)
seen = {}
orig = ops._GLAFunction.forward


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = lm(xd, yd, emd, cmd)[1]
    loss.backward()
    opt.step()
    return loss


def run(n):
    for _ in range(2):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(loss)


res = {"config": {"workload": "LinaModel d1024 l12 train step (fwd+bwd+AdamW), bf16 autocast", "batch": B, "seq": T, "text_len": Tx}}
for name, tc, n in (("tensor_core_backward", True, 3), ("recurrence_backward", False, 1)):
    ops.TC_BWD = tc
    try:
        ms, loss = run(n)
        res[name] = {"ms_per_step": round(ms, 2), "tokens_per_s": round(B * T / ms * 1e3), "loss": loss,
                     "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}
        print(name, res[name], flush=True)
    except Exception as e:      # noqa: BLE001
        res[name] = {"error": repr(e)[:500]}
        print(name, "FAILED", repr(e)[:500], flush=True)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)

# kernel-level breakdown of one step (CUPTI through torch.profiler; never a timing source for the numbers above)
if os.environ.get("LINA_TRAIN_PROFILE", "1") != "0":
    ops.TC_BWD = True
    from torch.profiler import profile, ProfilerActivity
    step(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    tab = prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=90)
    with open(os.path.splitext(out_path)[0] + "_profile.txt", "w") as f:
        f.write(tab)
    print(tab[:6000])

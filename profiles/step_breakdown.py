"""Per-kernel GPU time of the bench forward step (bs32 x seq2048, bf16) as it runs in the pipelined step (torch.profiler /
CUPTI activity records: no serialisation, no cache flush between kernels -- unlike the ncu launch list), plus an A/B of the
rank-16 gate expansion (lina_lowrank_linear vs the library GEMM).  usage: step_breakdown.py [out.json]"""
import json, os, re, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import lina_speech_b200.model.gla as G

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "step_breakdown.json")
dev = torch.device("cuda")
lm = bench.build_model(dev, torch.bfloat16)
c = bench.CFG
x, y, em, cm = bench.synth_inputs(c["batch"], c["seq"], c["txt_len"], seed=1000)
xd, yd, emd, cmd = x.to(dev), y.to(dev), em.to(dev), cm.to(dev)


def step():
    with torch.inference_mode():
        return lm(xd, yd, emd, cmd)[1]


def timed(n=8):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


res = {}
for name, flag in (("lowrank kernel", True), ("library GEMM (K=16)", False), ("lowrank kernel (again)", True)):
    G.LOWRANK_KERNEL = flag
    res[name] = round(timed(), 3)
    print(f"{name:28s} {res[name]:8.3f} ms/step", flush=True)
G.LOWRANK_KERNEL = True
# SwiGLU hidden padding: 1365 -> 1368 (alignment only) vs 1408 = 11 x 128 (tile-friendly N = 2816 for the library GEMM)
import lina_speech_b200.model.base_blocks as BB
for pad in (8, 128, 64, 8):
    BB.SWIGLU_PAD = pad
    for m in lm.modules():
        if isinstance(m, BB.SwiGLU):
            m._padded = None
    res[f"swiglu pad {pad}"] = round(timed(), 3)
    print(f"swiglu hidden padded to a multiple of {pad:3d}: {res[f'swiglu pad {pad}']:8.3f} ms/step", flush=True)
from torch.profiler import profile, ProfilerActivity
NSTEP = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(NSTEP):
        step()
    torch.cuda.synchronize()
tot, cnt = collections.defaultdict(float), collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        nm = re.sub(r"\(.*", "", ev.name)[:80]
        tot[nm] += ev.device_time_total / 1e3 if hasattr(ev, "device_time_total") else ev.cuda_time_total / 1e3
        cnt[nm] += 1
table = sorted(tot.items(), key=lambda kv: -kv[1])
total = sum(tot.values()) / NSTEP
print(f"sum of kernel times per step: {total:.3f} ms")
res["kernels_ms_per_step"] = {}
for nm, ms in table[:30]:
    print(f"{ms / NSTEP:8.3f} ms  {cnt[nm] // NSTEP:4d} x  {nm}")
    res["kernels_ms_per_step"][nm] = [round(ms / NSTEP, 4), cnt[nm] // NSTEP]
res["sum_ms_per_step"] = round(total, 3)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)

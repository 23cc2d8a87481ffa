"""Per-kernel GPU time of one CUDA-graphed decode step (generate_batch, bs 32 and 128, bf16 cache) from CUPTI activity records
of the replays.  usage: decode_breakdown.py [out.json]"""
import json, os, re, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "decode_breakdown.json")
dev = torch.device("cuda")
lm = bench.build_model(dev, torch.bfloat16)
c = bench.CFG
x, y, em, cm = bench.synth_inputs(c["batch"], c["seq"], c["txt_len"], seed=1000)
xt = x[0].to(dev)
res = {}
from torch.profiler import profile, ProfilerActivity
for B in (32, 128):
    lm.generate_batch(xt, batch_size=B, max_seqlen=8, k=100, force_max_seqlen=True, cuda_graph=True)
    torch.cuda.synchronize()
    NSTEP = 64
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        lm.generate_batch(xt, batch_size=B, max_seqlen=NSTEP, k=100, force_max_seqlen=True, cuda_graph=True, stop_check_interval=1 << 30)
        torch.cuda.synchronize()
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            nm = re.sub(r"^void ", "", ev.name)
            nm = re.sub(r"\(anonymous namespace\)::|<unnamed>::|at::native::", "", nm)
            nm = re.sub(r"\(.*", "", nm)[:90]
            tot[nm] += ev.device_time_total / 1e3
            cnt[nm] += 1
    table = sorted(tot.items(), key=lambda kv: -kv[1])
    total = sum(tot.values()) / NSTEP
    print(f"bs{B}: sum of kernel times per step (incl. warm-up / capture kernels spread over {NSTEP} steps): {total * 1e3:.1f} us, launches per step ~{sum(cnt.values()) / NSTEP:.0f}")
    res[f"bs{B}"] = {"sum_us_per_step": round(total * 1e3, 1), "kernels": {}}
    for nm, ms in table[:28]:
        print(f"  {ms / NSTEP * 1e3:8.2f} us  {cnt[nm] / NSTEP:6.1f} x  {nm}")
        res[f"bs{B}"]["kernels"][nm] = [round(ms / NSTEP * 1e3, 2), round(cnt[nm] / NSTEP, 1)]
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)

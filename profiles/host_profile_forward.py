"""Host-side (Python / launch) cost of the EAGER teacher-forced pass at the bench shape: wall time to enqueue one pass without
waiting for the GPU, and a cProfile table of where it goes."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda")
lm = bench.build_model(dev, torch.bfloat16)
c = bench.CFG
x, y, em, cm = bench.synth_inputs(c["batch"], c["seq"], c["txt_len"], seed=1000)
xd, yd, emd, cmd = x.to(dev), y.to(dev), em.to(dev), cm.to(dev)


def step():
    with torch.inference_mode():
        return lm(xd, yd, emd, cmd)[1]


for _ in range(3):
    step()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
print("enqueue ms / total ms per pass:", [(round(a * 1e3, 1), round(b * 1e3, 1)) for a, b in ts])
pr = cProfile.Profile()
torch.cuda.synchronize()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])

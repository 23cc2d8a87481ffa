"""A/B of the whole teacher-forced forward step (bench model, bs32 x seq2048, bf16) under the inference-path switches:
the fused whole-sequence mixer path on/off, the 5-way concatenated GEMM, the short-conv tile height, the tcgen05
kernel OPT bits, and the fused cross entropy on/off.  CUDA-event ms per step, inputs resident.
usage: ab_forward.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from lina_speech_b200 import _lib as L
import lina_speech_b200.model.gla as G
import lina_speech_b200.model.modeling_lina as ML

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ab_forward.json")
dev = torch.device("cuda")
lm = bench.build_model(dev, torch.bfloat16)
c = bench.CFG
x, y, em, cm = bench.synth_inputs(c["batch"], c["seq"], c["txt_len"], seed=1000)
xd, yd, emd, cmd = x.to(dev), y.to(dev), em.to(dev), cm.to(dev)
lib = L.lib()
fused_ce = ML.LinaModel._fused_cross_entropy


def step():
    with torch.inference_mode():
        return lm(xd, yd, emd, cmd)[1]


def run(n=5):
    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(loss)


res = {}
configs = [
    ("round1_path (op by op, torch CE)", dict(fused=False, cat5=False, tl=0, opt=0, ce=False, pre=False)),
    ("op by op + fused CE", dict(fused=False, cat5=False, tl=0, opt=0, ce=True, pre=False)),
    ("fused_prefill + CE", dict(fused=True, cat5=False, tl=0, opt=0, ce=True, pre=False)),
    ("fused_prefill + CE + pregated", dict(fused=True, cat5=False, tl=0, opt=0, ce=True, pre=True)),
    ("fused_prefill + CE + pregated + split GEMMs", dict(fused=True, cat5=False, tl=0, opt=0, ce=True, pre=True, group="split")),
    ("... + split, one state warpgroup", dict(fused=True, cat5=False, tl=0, opt=0, ce=True, pre=True, group="split", state2=1)),
]
for name, cf in configs:
    G.FUSED_PREFILL, G.CAT5, G.PREGATED = cf["fused"], cf["cat5"], cf["pre"]
    G.GEMM_GROUPING = cf.get("group", "cat4")
    lib.lina_debug_set_variant(0, cf["tl"])
    lib.lina_debug_set_variant(2, cf["opt"])
    lib.lina_debug_set_variant(4, cf.get("state2", 0))
    ML.LinaModel._fused_cross_entropy = staticmethod(fused_ce) if cf["ce"] else staticmethod(lambda *a: None)
    try:
        ms, loss = run()
        res[name] = {"ms_per_step": round(ms, 3), "tokens_per_s": round(c["batch"] * c["seq"] / ms * 1e3), "loss": loss}
        print(f"{name:40s} {ms:8.3f} ms/step  {c['batch'] * c['seq'] / ms * 1e3 / 1e6:6.3f} M tok/s  loss {loss:.5f}", flush=True)
    except Exception as e:      # noqa: BLE001
        res[name] = {"error": repr(e)}
        print(name, "FAILED", repr(e), flush=True)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)

"""CUDA-event timing of the HBM-streaming kernels of one GLA block at the bench shape (B=32, T=2048, d_model 1024,
H=4, K=256, V=512, bf16) and of the tcgen05 GLA kernel's OPT variants, each against its algorithmic bytes and the
measured copy bandwidth (MEASURED_PEAKS.json).  Between timed launches a 512 MB buffer is rewritten so that no input
is L2-resident.  Writes one JSON document (argv[1], default gpurun_out/stream_bench.json)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from lina_speech_b200 import _lib as L
from lina_speech_b200.fla_api import fused_chunk_gla

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "stream_bench.json")
dev = "cuda"
lib = L.lib()
peak = 6558.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
B, T, D, H, K, V = 32, 2048, 1024, 4, 256, 512
kd, vd = H * K, H * V
M = B * T
bf = torch.bfloat16
torch.manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream


def timeit(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(2_000_000)          # ~1 ms of device-side spin: the host enqueues e0 / fn / e1 meanwhile
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


res = {"peak_gbs": peak, "shape": dict(B=B, T=T, D=D, H=H, K=K, V=V)}


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    res[name] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 3), "bytes": nbytes}
    print(f"{name:42s} {ms:8.4f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of peak", flush=True)


# ---- reference point: torch copy ----
a = torch.randn(M, 2048, device=dev).to(bf); b = torch.empty_like(a)
report("torch_copy_268MB", timeit(lambda: b.copy_(a)), 2 * a.numel() * 2)

# ---- post-projection pass: 3 convs + gate ----
ldx = 2 * kd + 2 * vd
proj = torch.randn(B, T, ldx, device=dev).to(bf)
gk_raw = torch.randn(B, T, kd, device=dev).to(bf)
wq, wk = (torch.randn(kd, 4, device=dev).to(bf) for _ in range(2))
wv = torch.randn(vd, 4, device=dev).to(bf)
q, k, gk = (torch.empty(B, T, kd, dtype=bf, device=dev) for _ in range(3))
v = torch.empty(B, T, vd, dtype=bf, device=dev)
xq, xk, xv, g = proj[..., :kd], proj[..., kd:2 * kd], proj[..., 2 * kd:2 * kd + vd], proj[..., 2 * kd + vd:]
prep_bytes = 2 * 2 * (3 * kd + vd) * M


def prep():
    rc = lib.lina_gla_prefill_prep(L.ptr(xq), L.ptr(xk), L.ptr(xv), ldx, L.ptr(wq), L.ptr(wk), L.ptr(wv), L.ptr(gk_raw), kd,
                                   L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(gk), None, None, None, 0, B, T, kd, vd, 4, 16.0, 0.0, 0,
                                   L.BF16, st())
    assert rc == 0, lib.lina_last_error_string()


for tl in (8, 16):
    lib.lina_debug_set_variant(0, tl)
    report(f"prefill_prep_TL{tl} (q,k,v conv + gate)", timeit(prep), prep_bytes)
lib.lina_debug_set_variant(0, 0)

# ---- the same pass with the chunk gating folded in (q~, k~, decay) ----
nt = (T + 63) // 64
decay = torch.empty(B, H, nt, K, dtype=torch.float32, device=dev)


def prep_gated():
    rc = lib.lina_gla_prefill_prep_gated(L.ptr(xq), ldx, L.ptr(xk), ldx, L.ptr(xv), ldx, L.ptr(wq), L.ptr(wk), L.ptr(wv), L.ptr(gk_raw), kd,
                                         L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(decay), None, None, None, 0, B, T, H, K, V, 4,
                                         16.0, K ** -0.5, st())
    assert rc == 0, lib.lina_last_error_string()
report("prefill_prep_gated (v conv + q~,k~,decay)", timeit(prep_gated), (2 * (3 * kd + vd) + 2 * (2 * kd + vd)) * M)

xq_c, xv_c = xq.contiguous(), xv.contiguous()
for variant, name in ((0, "tiles_f32x2"), (3, "tiles_scalar"), (1, "round1_sliding")):
    lib.lina_debug_set_variant(1, 1 if variant == 1 else 0)
    lib.lina_debug_set_variant(3, 1 if variant == 3 else 0)
    for x_, w_, y_, d_ in ((xq_c, wq, q, kd), (xv_c, wv, v, vd)):
        def conv(x_=x_, w_=w_, y_=y_, d_=d_):
            rc = lib.lina_short_conv_fwd(L.ptr(x_), L.ptr(w_), L.ptr(y_), None, 0, B, T, d_, 4, 1, L.BF16, st())
            assert rc == 0
        report(f"short_conv_fwd_{name}_D{d_}", timeit(conv), 2 * 2 * d_ * M)
lib.lina_debug_set_variant(1, 0)
lib.lina_debug_set_variant(3, 0)


def gate():
    rc = lib.lina_gate_logsigmoid(L.ptr(gk_raw), L.ptr(gk), gk_raw.numel(), 16.0, 0.0, 0, L.BF16, st())
    assert rc == 0
report("gate_logsigmoid (round 1)", timeit(gate), 2 * 2 * kd * M)

# ---- norm gate ----
o = torch.randn(M * H, V, device=dev).to(bf)
y = torch.empty_like(o)
nw = torch.ones(V, device=dev).to(bf)
g_c = g.contiguous()


def ng_strided():
    rc = lib.lina_rmsnorm_swishgate_fwd_ld(L.ptr(o), L.ptr(g), L.ptr(nw), L.ptr(y), None, M * H, V, 1e-5, H, ldx, L.BF16, st())
    assert rc == 0


def ng_dense():
    rc = lib.lina_rmsnorm_swishgate_fwd(L.ptr(o), L.ptr(g_c), L.ptr(nw), L.ptr(y), None, M * H, V, 1e-5, L.BF16, st())
    assert rc == 0
report("norm_gate_fwd (gate strided in proj)", timeit(ng_strided), 3 * 2 * vd * M)
report("norm_gate_fwd (dense)", timeit(ng_dense), 3 * 2 * vd * M)

# ---- add + layernorm ----
xa, xb = (torch.randn(M, D, device=dev).to(bf) for _ in range(2))
gam, bet = torch.ones(D, device=dev).to(bf), torch.zeros(D, device=dev).to(bf)
s_out, ln_out = torch.empty_like(xa), torch.empty_like(xa)


def addln():
    rc = lib.lina_add_layernorm(L.ptr(xa), L.ptr(xb), L.ptr(gam), L.ptr(bet), L.ptr(s_out), L.ptr(ln_out), M, D, 1e-5, L.BF16, st())
    assert rc == 0
report("add_layernorm", timeit(addln), 4 * 2 * D * M)

# ---- swiglu act ----
hp = 1368
hbuf = torch.randn(M, 2 * hp, device=dev).to(bf); abuf = torch.empty(M, hp, dtype=bf, device=dev)


def swi():
    rc = lib.lina_swiglu_act(L.ptr(hbuf), L.ptr(abuf), M, hp, L.BF16, st())
    assert rc == 0
report("swiglu_act", timeit(swi), 3 * 2 * hp * M)

# ---- cross entropy ----
Vn, ldl = 4099, 4104
logits = torch.randn(M, ldl, device=dev).to(bf)
tgt = torch.randint(3, Vn, (M,), device=dev)
rows = torch.empty(2, M, dtype=torch.float32, device=dev)


def ce():
    rc = lib.lina_cross_entropy_rows(L.ptr(logits), ldl, L.ptr(tgt), None, L.ptr(rows[0]), L.ptr(rows[1]), M, Vn, 1, L.BF16, st())
    assert rc == 0
report("cross_entropy_rows", timeit(ce), 2 * Vn * M)
ref = F.cross_entropy(logits[:2048, :Vn].float(), tgt[:2048])
ce()
got = rows[0][:2048].sum() / rows[1][:2048].sum()
res["cross_entropy_check"] = {"ours": float(got), "torch": float(ref)}
print("CE check", float(got), float(ref))


def ce_torch():
    return F.cross_entropy(logits[:, :Vn].reshape(-1, Vn).float(), tgt, ignore_index=1)
report("cross_entropy torch route (copy+float+lsm)", timeit(ce_torch, 3), 2 * Vn * M)
del logits, hbuf, abuf, xa, xb, s_out, ln_out, a, b

# ---- tcgen05 GLA kernel variants ----
q4, k4 = (torch.randn(B, H, T, K, device=dev).to(bf) for _ in range(2))
v4 = torch.randn(B, H, T, V, device=dev).to(bf)
gk4 = (F.logsigmoid(torch.randn(B, H, T, K, device=dev)) / 16).to(bf)
gla_bytes = B * H * T * (3 * K + 2 * V) * 2
o_ref = None
for opt in (0, 2):
    lib.lina_debug_set_variant(2, opt)
    o_, ht_ = fused_chunk_gla(q4, k4, v4, gk4, output_final_state=True)
    torch.cuda.synchronize()
    if o_ref is None:
        o_ref, ht_ref = o_.clone(), ht_.clone()
    d_o, d_h = (o_.float() - o_ref.float()).abs().max().item(), (ht_ - ht_ref).abs().max().item()
    report(f"gla_chunk_fwd_tcgen05_OPT{opt}", timeit(lambda: fused_chunk_gla(q4, k4, v4, gk4), 10), gla_bytes)
    res[f"gla_chunk_fwd_tcgen05_OPT{opt}"].update(max_diff_o_vs_opt0=d_o, max_diff_ht_vs_opt0=d_h)
    print(f"   OPT{opt}: max|o - o_opt0| = {d_o:.3e}, max|ht - ht_opt0| = {d_h:.3e}", flush=True)
lib.lina_debug_set_variant(2, 0)

# ---- pre-gated operands: prep_gated + the tcgen05 kernel without its gate pre-pass ----
del q4, k4, gk4
prep_gated()
v_bthd = v.view(B, T, H, V)
o_b = torch.empty(B, T, H, V, dtype=bf, device=dev)


def gla_pre():
    rc = lib.lina_gla_chunk_fwd_pregated_bthd(L.ptr(q), L.ptr(k), L.ptr(v), L.ptr(decay), None, 0, L.ptr(o_b), None, B, H, T, K, V, st())
    assert rc == 0, lib.lina_last_error_string()
report("gla_chunk_fwd_pregated_bthd (tcgen05)", timeit(gla_pre), B * H * T * (2 * K + 2 * V) * 2 + decay.numel() * 4)
res["gla_chunk_fwd_pregated_bthd (tcgen05)"]["finite"] = bool(torch.isfinite(o_b.float()).all())
o_one = o_b.clone()
for key, name in ((6, "3+1 stages instead of 2+2"), (7, "cluster multicast (2+2 stages)")):
    lib.lina_debug_set_variant(key, 1)
    nm = "gla_chunk_fwd_pregated_bthd " + name
    report(nm, timeit(gla_pre), B * H * T * (2 * K + 2 * V) * 2 + decay.numel() * 4)
    res[nm]["max_diff_vs_default"] = (o_b.float() - o_one.float()).abs().max().item()
    print("   max diff", res[nm]["max_diff_vs_default"])
    lib.lina_debug_set_variant(key, 0)
lib.lina_debug_set_variant(4, 1)
report("gla_chunk_fwd_pregated_bthd one state warpgroup", timeit(gla_pre), B * H * T * (2 * K + 2 * V) * 2 + decay.numel() * 4)
res["gla_chunk_fwd_pregated_bthd one state warpgroup"]["max_diff_vs_default"] = (o_b.float() - o_one.float()).abs().max().item()
print("   STATE2 max diff", res["gla_chunk_fwd_pregated_bthd one state warpgroup"]["max_diff_vs_default"])
lib.lina_debug_set_variant(4, 0)

os.makedirs(os.path.dirname(out_path), exist_ok=True)
with open(out_path, "w") as f:
    json.dump(res, f, indent=1)
print("wrote", out_path)

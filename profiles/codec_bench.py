"""Isolated CUDA-event timings of the WavTokenizer decode-tail kernels at the shipped size (B=32 sequences of L=750 frames,
dim 768, 32 groups, n_fft 1280 / hop 320; fp32) against their algorithmic bytes and the measured copy bandwidth, plus the
whole codes_to_features + decode call in both GEMM precisions.  512 MB L2 flush + device spin before every timed launch.
usage: codec_bench.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lina_speech_b200 import _lib as L
from lina_speech_b200.codec import WavTokenizer

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "codec_bench.json")
dev = "cuda"
lib = L.lib()
peak = 6558.0
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
B, C, Ln, G, NFFT, HOP = 32, 768, 750, 32, 1280, 320
torch.manual_seed(0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream
res = {"peak_gbs": peak, "shape": dict(B=B, C=C, L=Ln, groups=G, n_fft=NFFT, hop=HOP)}


def timeit(fn, n=10):
    fn(); fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    res[name] = {"ms": round(ms, 4), "GBps": round(gbs, 1), "frac": round(gbs / peak, 3), "bytes": nbytes}
    print(f"{name:46s} {ms:8.4f} ms  {gbs:8.1f} GB/s  {gbs / peak:5.2f} of peak", flush=True)


x = torch.randn(B, C, Ln, device=dev)
y = torch.empty_like(x)
ht = torch.randn(B, Ln, C, device=dev)
yt = torch.empty_like(ht)
gam, bet = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
dww, dwb = torch.randn(C, 7, device=dev) * 0.3, torch.randn(C, device=dev)
nb = x.numel() * 4


def check(rc):
    assert rc == 0, lib.lina_last_error_string()


report("groupnorm_swish", timeit(lambda: check(lib.lina_codec_groupnorm_swish(L.ptr(x), L.ptr(gam), L.ptr(bet), L.ptr(y), None, B, C, Ln, G, 1e-6, 1, st()))), 2 * nb)
wsb = torch.empty(int(lib.lina_codec_dwconv_adaln_workspace_bytes(B, C, Ln)), dtype=torch.uint8, device=dev)
report("dwconv_adaln_ws (2 kernels: conv+T+stats, apply)", timeit(lambda: check(lib.lina_codec_dwconv_adaln_ws(L.ptr(x), L.ptr(dww), L.ptr(dwb), L.ptr(gam), L.ptr(bet), L.ptr(yt), L.ptr(wsb), B, C, Ln, 1e-6, st()))), 2 * nb)
report("layernorm_t_ws (2 kernels)", timeit(lambda: check(lib.lina_codec_layernorm_t_ws(L.ptr(x), L.ptr(gam), L.ptr(bet), L.ptr(yt), L.ptr(wsb), B, C, Ln, 1e-6, st()))), 2 * nb)
report("dwconv_adaln (round-1 single kernel)", timeit(lambda: check(lib.lina_codec_dwconv_adaln(L.ptr(x), L.ptr(dww), L.ptr(dwb), L.ptr(gam), L.ptr(bet), L.ptr(yt), B, C, Ln, 1e-6, st()))), 2 * nb)
report("layernorm_t (round-1 single kernel)", timeit(lambda: check(lib.lina_codec_layernorm_t(L.ptr(x), L.ptr(gam), L.ptr(bet), L.ptr(yt), B, C, Ln, 1e-6, st()))), 2 * nb)
report("scale_residual_t", timeit(lambda: check(lib.lina_codec_scale_residual_t(L.ptr(ht), L.ptr(gam), L.ptr(x), L.ptr(y), B, C, Ln, st()))), 3 * nb)
hh = torch.randn(B, Ln, NFFT + 2, device=dev) * 0.5
win = torch.hann_window(NFFT, device=dev)
wav = torch.empty(B, Ln * HOP, device=dev)
ws = torch.empty(int(lib.lina_codec_istft_workspace_bytes(B, Ln, NFFT)), dtype=torch.uint8, device=dev)
report("istft_head (polar + irfft1280 + window + OLA)", timeit(lambda: check(lib.lina_codec_istft_head(L.ptr(hh), L.ptr(win), L.ptr(wav), L.ptr(ws), B, Ln, NFFT, HOP, st()))),
       hh.numel() * 4 + wav.numel() * 4)
if os.environ.get("LINA_BRINGUP"):                      # variant key 8: warp-per-frame fixed-radix FFT (not yet run on hardware)
    wav0 = wav.clone()
    lib.lina_debug_set_variant(8, 1)
    try:
        report("istft_head, warp-per-frame FFT (variant 8)", timeit(lambda: check(lib.lina_codec_istft_head(L.ptr(hh), L.ptr(win), L.ptr(wav), L.ptr(ws), B, Ln, NFFT, HOP, st()))),
               hh.numel() * 4 + wav.numel() * 4)
        res["istft variant 8 max diff"] = float((wav - wav0).abs().max())
        print("istft variant 8 max diff vs default:", res["istft variant 8 max diff"], flush=True)
    finally:
        lib.lina_debug_set_variant(8, 0)
codes = torch.randint(0, 4096, (1, B, Ln), device=dev)
books = torch.randn(4096, 512, device=dev)
feat = torch.empty(B, 512, Ln, device=dev)
report("codes_to_features", timeit(lambda: check(lib.lina_codec_codes_to_features(L.ptr(codes), L.ptr(books), L.ptr(feat), 1, B, Ln, 4096, 512, st()))),
       feat.numel() * 4 * 2)
report("torch copy (same bytes as groupnorm)", timeit(lambda: y.copy_(x)), 2 * nb)

wt = WavTokenizer.from_hparams().cuda().eval()
with torch.no_grad():
    wt.feature_extractor.encodec.quantizer.vq.layers[0]._codebook.embed.normal_()
bw = torch.tensor([0], device=dev)
for prec in ("fp32", "tf32"):
    wt.gemm_precision = prec
    def full():
        with torch.no_grad():
            return wt.decode(wt.codes_to_features(codes), bandwidth_id=bw)
    ms = timeit(full, 3)
    res[f"decode_{prec}"] = {"ms": round(ms, 3), "frames_per_s": round(B * Ln / ms * 1e3), "x_realtime": round(B * Ln / 75 / ms * 1e3)}
    print(f"decode {prec}: {ms:.3f} ms  {B * Ln / ms * 1e3:.0f} frames/s  {B * Ln / 75 / ms * 1e3:.0f} x real time", flush=True)
os.makedirs(os.path.dirname(out_path), exist_ok=True)
json.dump(res, open(out_path, "w"), indent=1)

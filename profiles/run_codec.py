"""WavTokenizer decode at the shipped size (dim 768, 12 ConvNeXt blocks, n_fft 1280 / hop 320; random weights),
B sequences of L frames: CUDA-event time of codes_to_features + decode, frames/s, x real time.  The command ncu
wraps for the codec launch list (profiles/launches_r01_codec.csv)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lina_speech_b200.codec import WavTokenizer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 750
n = int(sys.argv[3]) if len(sys.argv) > 3 else 5
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16x3"
torch.manual_seed(0)
wt = WavTokenizer.from_hparams().cuda().eval()
wt.gemm_precision = prec
with torch.no_grad():
    wt.feature_extractor.encodec.quantizer.vq.layers[0]._codebook.embed.normal_()
codes = torch.randint(0, 4096, (1, B, L), device="cuda")
bw = torch.tensor([0], device="cuda")
for _ in range(2):
    wav = wt.decode(wt.codes_to_features(codes), bandwidth_id=bw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    wav = wt.decode(wt.codes_to_features(codes), bandwidth_id=bw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"gemm_precision": prec, "B": B, "L": L, "ms": ms, "frames_per_s": B * L / ms * 1e3, "audio_s_per_s": B * L / 75 / ms * 1e3,
                  "wav_shape": list(wav.shape), "finite": bool(torch.isfinite(wav).all())}))

"""Where an iteration of train_initial_state (fp32 d1024 l12, batch 2) spends its GPU time: torch.profiler kernel table."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lina_speech_b200.tuning import train_initial_state

dev = torch.device("cuda", 0)
lm = bench.build_model(dev, torch.float32)
g = torch.Generator().manual_seed(4)
ds = []
for i in range(16):
    n = int(torch.randint(225, 751, (1,), generator=g))
    ds.append({"audio_token": torch.randint(0, 4096, (1, n), generator=g),
               "text": "".join(chr(97 + int(c)) for c in torch.randint(0, 26, (60 + 5 * i,), generator=g))})
tok = bench._ByteTokenizer()
train_initial_state(lm, ds, tok, 16)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    train_initial_state(lm, ds, tok, 16)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python profiles/train_step.py gpurun_out/r02u_train_step.json > gpurun_out/r02u_train.log 2>&1
grep -E "tensor_core_backward|recurrence_backward" gpurun_out/r02u_train.log
python - <<'PY'
import re
t=open('gpurun_out/r02u_train_step_profile.txt').read().splitlines()
# name col + Self CUDA + calls
hdr=t[1]
for l in t[3:48]:
    parts=re.split(r'\s{2,}', l.strip())
    if len(parts)>6: print(f"{parts[0][:95]:95s} {parts[-5]:>10s} {parts[-1]:>6s}")
PY

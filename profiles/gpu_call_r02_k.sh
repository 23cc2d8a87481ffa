#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_codec_gpu.py tests/test_gemm_gpu.py -m gpu -q > gpurun_out/r02k_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02k_tests.log
timeout 300 python - > gpurun_out/r02k_codec.json 2>gpurun_out/r02k_err.log <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
r = bench.codec_metrics(torch.device('cuda', 0))
print(json.dumps(r))
PY
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_codec.json'))
print('ms', round(d['ms'],3), 'bf16x2', d['two_part_mode'])
for k,v in d['stages'].items(): print(f"{k:28s} n={v['launch_groups']:3d} ms={v['ms']:.3f}  hbm={v.get('hbm_frac',0):.2f}  tf={v.get('tflops_bf16',0):.0f}")
PY

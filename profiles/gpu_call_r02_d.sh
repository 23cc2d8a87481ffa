#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -s > gpurun_out/r02d_gemm.log 2>&1; echo "gemm tests rc=$?"; grep -E "rel err|passed|failed|Error" gpurun_out/r02d_gemm.log | tail -20
timeout 900 python -m pytest tests/test_codec_gpu.py -m gpu -q -s > gpurun_out/r02d_codec.log 2>&1; echo "codec rc=$?"; grep -E "^E  |passed|failed|Error|bf16x2" gpurun_out/r02d_codec.log | tail -30
timeout 600 python -m pytest tests/test_flagship_parity_gpu.py -m gpu -q > gpurun_out/r02d_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r02d_parity.log
timeout 300 python - > gpurun_out/r02d_codec_bench.log 2>&1 <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
r = bench.codec_metrics(torch.device('cuda', 0))
print(json.dumps(r, indent=1))
PY
echo "codec bench rc=$?"; tail -120 gpurun_out/r02d_codec_bench.log

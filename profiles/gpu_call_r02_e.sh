#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_codec_gpu.py -m gpu -q -s > gpurun_out/r02e_tests.log 2>&1; echo "tests rc=$?"; grep -E "relative error|passed|failed|^E   +Assert|bf16x2" gpurun_out/r02e_tests.log | tail -12
for span in 2 4 8; do
LINA_GEMM_SPAN=$span timeout 300 python - > gpurun_out/r02e_codec_span$span.json 2>gpurun_out/r02e_err.log <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import bench
r = bench.codec_metrics(torch.device('cuda', 0))
print(json.dumps(r))
PY
python - <<PY
import json
d=json.load(open('gpurun_out/r02e_codec_span$span.json'))
print('span $span: ms', round(d['ms'],3), 'bf16x2', round(d['two_part_mode']['ms'],3), {k:(round(v['ms'],3), round(v.get('tflops_bf16',0)) or round(v.get('hbm_frac',0),2)) for k,v in d['stages'].items()})
PY
done

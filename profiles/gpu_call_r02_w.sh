#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flagship_parity_gpu.py -x -q -k "graphed" 2>&1 | tail -5
timeout 600 python bench.py --no-train --no-extras > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02w_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02w_bench.json'))
print('fwd ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['path'][:40], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['launch_ms'])
PY

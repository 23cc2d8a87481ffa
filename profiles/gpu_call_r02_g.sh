#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_flagship_parity_gpu.py -m gpu -q -x > gpurun_out/r02g_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r02g_tests.log
timeout 600 python bench.py --no-train > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench.json'))
print('fwd ms', d['ms_per_step'], 'value', d['value'])
for k in ('decode','decode_bs128','decode_prompt','init_state_tuning'):
    print(k, json.dumps(d.get(k))[:600])
print('codec ms', d['codec']['ms'])
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03a_tests.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r03a_tests.log
timeout 600 python bench.py > gpurun_out/r03a_bench.json 2> gpurun_out/r03a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r03a_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03a_bench.json'))
print('fwd ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print('roofline', d['roofline']['frac'], d['roofline']['launch_ms'])
for k in ('decode','decode_bs128','decode_prompt','train_step','init_state_tuning'):
    print(k, json.dumps(d.get(k))[:330])
print('codec ms', d['codec']['ms'])
PY

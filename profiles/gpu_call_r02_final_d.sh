#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03d_tests.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/r03d_tests.log
timeout 900 python bench.py > gpurun_out/r03d_bench.json 2> gpurun_out/r03d_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r03d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03d_bench.json'))
print('fwd ms', d['ms_per_step'], 'eager', d['eager_ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['launch_ms'])
print('decode', d['decode']['ms_per_step'], 'bs128', d['decode_bs128']['ms_per_step'], 'prompt', d['decode_prompt']['prefill']['total_ms'], 'codec', d['codec']['ms'], 'tuning', d['init_state_tuning']['it_per_s'], 'train', d['train_step']['ms_per_step'])
PY

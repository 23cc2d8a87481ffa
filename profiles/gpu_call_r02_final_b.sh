#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r03b_bench.json 2> gpurun_out/r03b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r03b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03b_bench.json'))
print('fwd ms', d['ms_per_step'], 'eager', d['eager_ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['launch_ms'])
for k in ('decode','decode_bs128','decode_prompt','train_step','init_state_tuning'):
    print(k, json.dumps(d.get(k))[:300])
print('codec ms', d['codec']['ms'], 'cpu', d.get('cpu_baseline'))
PY
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-400

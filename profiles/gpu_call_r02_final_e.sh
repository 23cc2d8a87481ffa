#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r03e_tests.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/r03e_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --no-train > gpurun_out/r03e_bench.json 2> gpurun_out/r03e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03e_bench.json'))
print('fwd ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'decode', d['decode']['ms_per_step'], d['decode_bs128']['ms_per_step'], 'roofline', d['roofline']['frac'])
PY

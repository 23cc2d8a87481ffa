#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytestG_all.log
timeout 600 python bench.py > gpurun_out/benchG.json 2> gpurun_out/benchG.err
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytestG_all.log; cut -c1-300 gpurun_out/benchG.json; python -c "
import json; d=json.load(open('gpurun_out/benchG.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['decode']['ms_per_step'], d['decode_bs128']['ms_per_step'], d['train_step'].get('ms_per_step'), d['clocks'])"; tail -2 gpurun_out/benchG.err

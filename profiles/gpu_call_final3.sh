#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1
timeout 400 python bench.py > gpurun_out/benchH.json 2> gpurun_out/benchH.err
tail -1 gpurun_out/smoke3.log; wc -l gpurun_out/benchH.json; python -c "
import json; d=json.load(open('gpurun_out/benchH.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['decode']['ms_per_step'], d['decode_bs128']['ms_per_step'], d['train_step'].get('ms_per_step'), d['clocks'], d['gpu_launches'])"; tail -2 gpurun_out/benchH.err

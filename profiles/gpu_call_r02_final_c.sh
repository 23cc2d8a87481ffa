#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r03c_bench_n2.json 2> gpurun_out/r03c_bench_n2.err
echo "rc=$?"; tail -2 gpurun_out/r03c_bench_n2.err; wc -l gpurun_out/r03c_bench_n2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r03c_bench_n2.json'))
print('N', d['n_gpus'], 'fwd ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print('decode', d['decode']['ms_per_step'], d['decode']['token_all_gather'], 'bs128', d['decode_bs128']['ms_per_step'])
print('extras', {k: (v.get('error') if isinstance(v, dict) and 'error' in v else 'ok') for k, v in d.items() if k in ('decode_prompt','codec','init_state_tuning','train_step')})
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | cut -c1-200

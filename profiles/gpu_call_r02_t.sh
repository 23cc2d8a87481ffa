#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-train --no-extras > gpurun_out/r02t_bench_n$N.json 2> gpurun_out/r02t_bench_n$N.err
echo "rc=$?"; tail -3 gpurun_out/r02t_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/r02t_bench_n$N.json'))
print('N', d['n_gpus'], 'fwd ms', d['ms_per_step'], 'value', d['value'])
print('decode', json.dumps(d['decode']))
print('decode128', json.dumps(d['decode_bs128']))
PY

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_skinny_gpu.py -m gpu -q -x > gpurun_out/r02h_skinny.log 2>&1; echo "skinny rc=$?"; tail -15 gpurun_out/r02h_skinny.log
timeout 900 python -m pytest tests/test_codec_gpu.py tests/test_model_gpu.py tests/test_flagship_parity_gpu.py tests/test_modules_gpu.py -m gpu -q > gpurun_out/r02h_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02h_tests.log
timeout 600 python bench.py --no-train > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02h_bench.json'))
print('fwd ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
for k in ('decode','decode_bs128','decode_prompt'):
    print(k, json.dumps(d.get(k))[:420])
c=d['codec']; print('codec ms', c['ms'], c['two_part_mode'])
for k,v in c['stages'].items(): print(f"{k:28s} n={v['launch_groups']:3d} ms={v['ms']:.3f}  hbm={v.get('hbm_frac',0):.2f}  tf={v.get('tflops_bf16',0):.0f}")
PY

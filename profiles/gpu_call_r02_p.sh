#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
# launch list of the bench step (2 timed + 1 warm-up resident steps, nothing else)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_bench_profile.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/r02p_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/launches_r02_bench_profile.csv
# full capture of the CTA-pair GLA kernel at the bench shape
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gla_chunk_fwd -s 2 -c 1 -o gpurun_out/ncu_gla_pair_r02 python profiles/run_pregated.py 3 > gpurun_out/r02p_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02p_ncu.log
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --no-train > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02p_bench.json'))
print('fwd ms', d['ms_per_step'], 'roofline', d['roofline']['frac'], d['roofline']['launch_ms'])
print(json.dumps(d.get('init_state_tuning'))[:400])
PY

#!/bin/bash
# round-1b GPU call 1: parity of the new kernels, stream A/B, forward A/B, bench, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in test_modules_gpu test_model_gpu test_gla_ops_gpu; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_$f.log
done
timeout 300 python profiles/stream_bench.py gpurun_out/stream_bench.json > gpurun_out/stream_bench.log 2>&1
timeout 400 python profiles/ab_forward.py gpurun_out/ab_forward.json > gpurun_out/ab_forward.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_call1.json 2> gpurun_out/bench_call1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_fwd.csv python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_fwd.log 2>&1
for f in test_fullsize_gpu test_codec_gpu test_umma_probe_gpu; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_$f.log
done
tail -3 gpurun_out/pytest_*.log; cat gpurun_out/stream_bench.log | tail -40; cat gpurun_out/ab_forward.log | tail -12; cat gpurun_out/bench_call1.json | cut -c1-1500

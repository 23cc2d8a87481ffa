#!/bin/bash
# round-1b GPU call 4: STATE2 parity + A/B, training-step profile, ncu capture of the pre-gated GLA kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_model_gpu.py -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest4.log
timeout 300 python profiles/stream_bench.py gpurun_out/stream_bench4.json > gpurun_out/stream_bench4.log 2>&1
timeout 400 python profiles/ab_forward.py gpurun_out/ab_forward4.json > gpurun_out/ab_forward4.log 2>&1
timeout 600 python profiles/train_step.py gpurun_out/train_step4.json > gpurun_out/train_step4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gla_chunk_fwd_sm100 -s 2 -c 1 -o gpurun_out/ncu_gla_pregated_r01 python profiles/run_pregated.py 1 > gpurun_out/ncu_gla_pregated.log 2>&1
ncu -i gpurun_out/ncu_gla_pregated_r01.ncu-rep --page raw --csv > gpurun_out/ncu_gla_pregated_r01_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_gla_pregated_r01.ncu-rep --page details > gpurun_out/ncu_gla_pregated_r01_details.txt 2>/dev/null
tail -12 gpurun_out/pytest4.log; tail -8 gpurun_out/stream_bench4.log; tail -9 gpurun_out/ab_forward4.log; head -70 gpurun_out/train_step4.log | cut -c1-200

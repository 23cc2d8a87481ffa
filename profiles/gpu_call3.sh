#!/bin/bash
# round-1b GPU call 3: tensor-core backward parity, full GPU suite, training step, split-GEMM A/B, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest3_all.log
timeout 600 python profiles/train_step.py gpurun_out/train_step.json > gpurun_out/train_step.log 2>&1
timeout 400 python profiles/ab_forward.py gpurun_out/ab_forward3.json > gpurun_out/ab_forward3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gla_chunk_fwd_sm100|qk_gate|conv4_silu" -c 3 -o gpurun_out/ncu_pregated_r01 python profiles/run_pregated.py 1 > gpurun_out/ncu_pregated.log 2>&1
ncu -i gpurun_out/ncu_pregated_r01.ncu-rep --page raw --csv > gpurun_out/ncu_pregated_r01_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
tail -30 gpurun_out/pytest3_all.log; cat gpurun_out/train_step.log | tail; tail -8 gpurun_out/ab_forward3.log; tail -5 gpurun_out/ncu_pregated.log

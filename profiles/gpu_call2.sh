#!/bin/bash
# round-1b GPU call 2: pregated path parity + A/B
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest2_test_modules_gpu.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_fullsize_gpu.py -q -m gpu 2>&1 | tail -30 > gpurun_out/pytest2_model.log
timeout 300 python profiles/stream_bench.py gpurun_out/stream_bench2.json > gpurun_out/stream_bench2.log 2>&1
timeout 400 python profiles/ab_forward.py gpurun_out/ab_forward2.json > gpurun_out/ab_forward2.log 2>&1
tail -25 gpurun_out/pytest2_*.log; tail -45 gpurun_out/stream_bench2.log; tail -12 gpurun_out/ab_forward2.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gla_ops_gpu.py tests/test_modules_gpu.py tests/test_fullsize_gpu.py -q -m gpu -x 2>&1 | tail -8 > gpurun_out/pytest10.log
timeout 300 python profiles/stream_bench.py gpurun_out/stream_bench10.json > gpurun_out/stream_bench10.log 2>&1
tail -5 gpurun_out/pytest10.log; grep -E "gla_chunk|prep_gated" gpurun_out/stream_bench10.log

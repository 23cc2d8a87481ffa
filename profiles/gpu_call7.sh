#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_gpu.py tests/test_model_gpu.py -q -m gpu 2>&1 | tail -10 > gpurun_out/pytest7.log
timeout 300 python profiles/stream_bench.py gpurun_out/stream_bench7.json > gpurun_out/stream_bench7.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
tail -5 gpurun_out/pytest7.log; grep -E "prep|pregated|conv" gpurun_out/stream_bench7.log; cut -c1-2500 gpurun_out/bench7.json; tail -3 gpurun_out/bench7.err

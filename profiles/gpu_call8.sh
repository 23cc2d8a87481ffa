#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_codec_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest8.log
timeout 300 python profiles/codec_bench.py gpurun_out/codec_bench8.json > gpurun_out/codec_bench8.log 2>&1
tail -8 gpurun_out/pytest8.log; tail -14 gpurun_out/codec_bench8.log

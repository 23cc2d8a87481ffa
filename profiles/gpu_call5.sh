#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest6_all.log
timeout 600 python profiles/train_step.py gpurun_out/train_step6.json > gpurun_out/train_step6.log 2>&1
tail -25 gpurun_out/pytest6_all.log; head -60 gpurun_out/train_step6.log | cut -c1-100,190-215

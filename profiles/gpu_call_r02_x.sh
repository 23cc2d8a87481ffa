#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python profiles/tuning_profile.py > gpurun_out/r02x_tuning_profile.log 2>&1
grep -v "^\[W\|warn" gpurun_out/r02x_tuning_profile.log | cut -c1-72,150-230 | head -34
grep "time total" gpurun_out/r02x_tuning_profile.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python profiles/step_breakdown.py gpurun_out/r02r_step_breakdown.json 2>&1 | grep -v Warning | tail -45

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python profiles/ab_gla_pair.py 20 gpurun_out/r02s_ab_gla_pair.json 2>&1 | grep -E "B32|B8"

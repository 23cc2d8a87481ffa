#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_modules_gpu.py -x -q -k "pair_kernel or pregated_prep" 2>&1 | tail -3
timeout 300 python profiles/ab_gla_pair.py 20 gpurun_out/r02s_ab_gla_pair.json 2>&1 | grep -E "B32|B8"

#!/bin/bash
# round 2, call l: CTA-pair GLA kernel bring-up: bit-identity test, A/B timing
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_modules_gpu.py -x -q -k "pair_kernel or pregated_prep" > gpurun_out/r02l_tests.log 2>&1
echo "tests rc=$?"
tail -15 gpurun_out/r02l_tests.log
timeout 300 python profiles/ab_gla_pair.py 20 gpurun_out/r02l_ab_gla_pair.json 2>&1 | tail -12

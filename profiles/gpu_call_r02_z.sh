#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02z_tests.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r02z_tests.log
timeout 600 python profiles/decode_breakdown.py gpurun_out/r02z_decode_breakdown.json 2>&1 | grep -E "bs32|bs128|topk|add_layernorm|state_kernel"

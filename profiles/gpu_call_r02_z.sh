#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_modules_gpu.py tests/test_fullsize_gpu.py tests/test_flagship_parity_gpu.py -q -x -k "step or decode or generat or greedy" > gpurun_out/r02z_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r02z_tests.log
timeout 600 python profiles/decode_breakdown.py gpurun_out/r02z_decode_breakdown.json 2>&1 | grep -E "bs32|bs128|step_prep"

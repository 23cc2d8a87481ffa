#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -s > gpurun_out/r02c_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -40 gpurun_out/r02c_gemm.log
timeout 900 python -m pytest tests/test_flagship_parity_gpu.py tests/test_modules_gpu.py -m gpu -q -s > gpurun_out/r02c_parity.log 2>&1; echo "parity rc=$?"; tail -15 gpurun_out/r02c_parity.log

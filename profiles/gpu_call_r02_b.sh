#!/bin/bash
# round 2, call 2: GPU suite with the new flagship-parity / full-size codec / envelope tests, the reference's Triton ops timed
# beside ours (needs baseline/_ref/fla staged), bench.py with the codec / prompted-decode / state-tuning legs.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02b_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -25 gpurun_out/r02b_gpu_tests.log
timeout 900 python profiles/triton_reference_bench.py > gpurun_out/r02b_triton.log 2>&1; echo "triton bench rc=$?"; tail -30 gpurun_out/r02b_triton.log
timeout 600 python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r02b_bench.err; cat gpurun_out/r02b_bench.json

#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_modules_gpu.py -x -q -k "pair_kernel" 2>&1 | tail -3
timeout 120 python profiles/trace_gla_pair.py > gpurun_out/r02m_trace_pair.log 2>&1
cat gpurun_out/r02m_trace_pair.log | cut -c1-260
timeout 300 python profiles/ab_gla_pair.py 20 gpurun_out/r02m_ab_gla_pair.json 2>&1 | grep B32

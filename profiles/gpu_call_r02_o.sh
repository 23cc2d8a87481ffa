#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 300 python profiles/tuning_profile.py > gpurun_out/r02o_tuning_profile.log 2>&1
grep -v "^\[W\|warn" gpurun_out/r02o_tuning_profile.log | cut -c1-200 | head -45
# ncu: the pair kernel at the bench shape (run_pregated.py launches prep + the GLA kernel)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gla_chunk_fwd -s 2 -c 1 -o gpurun_out/ncu_gla_pair_r02 python profiles/run_pregated.py 3 > gpurun_out/r02o_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r02o_ncu.log
ls -la gpurun_out/*.ncu-rep

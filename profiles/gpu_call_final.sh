#!/bin/bash
# end-of-session capture: full GPU suite, bench line, ncu launch list of one forward step, ncu --set full of the pre-gated kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytestF_all.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/benchF.json 2> gpurun_out/benchF.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_fwd_final.csv python bench.py --profile --steps 1 --warmup 1 --no-train > gpurun_out/ncu_fwdF.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gla_chunk_fwd_sm100 -s 2 -c 1 -o gpurun_out/ncu_gla_pregated_final python profiles/run_pregated.py 1 > gpurun_out/ncu_gla_pregatedF.log 2>&1
ncu -i gpurun_out/ncu_gla_pregated_final.ncu-rep --page raw --csv > gpurun_out/ncu_gla_pregated_final_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_gla_pregated_final.ncu-rep --page details > gpurun_out/ncu_gla_pregated_final_details.txt 2>/dev/null
timeout 300 python profiles/train_step.py gpurun_out/train_stepF.json > gpurun_out/train_stepF.log 2>&1
tail -4 gpurun_out/pytestF_all.log; cut -c1-2800 gpurun_out/benchF.json; tail -2 gpurun_out/benchF.err; grep -E "tensor_core|recurrence" gpurun_out/train_stepF.log

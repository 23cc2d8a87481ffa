#!/bin/bash
# First GPU call of round 2: (1) the regular GPU suite, (2) the items written at the end of round 1 (parity green in the last
# 14 s call, never timed): ISTFT warp-per-frame FFT (variant key 8; LINA_BRINGUP adds the multi-pass size), M = 64 layout probe.
# Each step under its own timeout so that a hang costs minutes, not the call.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gpu_tests.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/r02_gpu_tests.log
LINA_BRINGUP=1 timeout 300 python -m pytest tests/test_codec_gpu.py tests/test_gla_ops_gpu.py tests/test_model_gpu.py -m gpu -q -k "warp_per_frame or outside_the_tensor_core_envelope" > gpurun_out/r02_bringup_tests.log 2>&1; echo "bring-up tests rc=$?"; tail -3 gpurun_out/r02_bringup_tests.log
timeout 120 python profiles/probe_m64.py > gpurun_out/r02_probe_m64.log 2>&1; echo "probe_m64 rc=$?"; tail -40 gpurun_out/r02_probe_m64.log
LINA_BRINGUP=1 timeout 300 python profiles/codec_bench.py gpurun_out/r02_codec_bench.json > gpurun_out/r02_codec_bench.log 2>&1; echo "codec_bench rc=$?"; grep -i istft gpurun_out/r02_codec_bench.log
timeout 400 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; cat gpurun_out/r02_bench.json | head -c 1500

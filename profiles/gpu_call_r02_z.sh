#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 python profiles/decode_breakdown.py gpurun_out/r02z_decode_breakdown.json 2>&1 | grep -v Warn | tail -64

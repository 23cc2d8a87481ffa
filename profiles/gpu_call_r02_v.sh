#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python profiles/xp_state_pass.py 2>&1 | tail -10

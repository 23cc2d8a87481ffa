#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python profiles/host_profile_forward2.py 2>&1 | grep -v Warn | tail -8

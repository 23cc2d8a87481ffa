#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 120 python profiles/trace_gla_pair.py 2>&1 | cut -c1-330 | tail -24

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/e2e_breakdown.py 4 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g37_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g37_tests.log

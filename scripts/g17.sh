#!/bin/bash
# batch 2: full GPU test suite with the lazy two-array label state + 8-bit import / export, bench A/B, e2e breakdown
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g17_tests.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/g17_tests.log
VARIANTS="prev default" STEPS=5 bash scripts/g4.sh
timeout 300 python scripts/e2e_breakdown.py 4 2>&1 | tail -3

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/fp32_switch.py 50 2>&1 | grep "^fp64 x" | awk '{for(i=3;i<=NF-4;i+=3) printf "%s ", $i; print ""}'
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g24_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/g24_tests.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
VARIANTS="default:--no-spare-sms default default:--no-spare-sms default" STEPS=5 bash scripts/g4.sh
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g35_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g35_tests.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "asymmetric or borderline or large_k or l2 or L2 or grid or estimate or invariants" > gpurun_out/g55_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g55_tests.log
{
timeout 300 python scripts/kmix.py "64:64" 80
CHAINS=256 timeout 300 python scripts/kbench.py 64 48
} > gpurun_out/g55.log 2>&1
cat gpurun_out/g55.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CHAINS=256 timeout 300 python scripts/kbench.py 64 48 32 > gpurun_out/g60.log 2>&1; cat gpurun_out/g60.log
timeout 300 python scripts/kmix.py "64:64" 80 >> gpurun_out/g60.log 2>&1; tail -4 gpurun_out/g60.log

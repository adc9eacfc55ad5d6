#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
CHAINS=256 timeout 300 python scripts/kbench.py 64
CHAINS=96 timeout 300 python scripts/kbench.py 64 48:64
CHAINS=32 timeout 300 python scripts/kbench.py 64
CHAINS=96 SCHED=abrupt timeout 300 python scripts/kbench.py 64
CHAINS=256 SCHED=abrupt timeout 300 python scripts/kbench.py 32
} > gpurun_out/g43.log 2>&1
cat gpurun_out/g43.log

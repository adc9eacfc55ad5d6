#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python scripts/c4_overhead.py > gpurun_out/g49.log 2>&1; cat gpurun_out/g49.log

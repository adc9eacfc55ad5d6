#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 300 python scripts/kmix.py "64:64" 80
timeout 300 python scripts/kmix.py "48:32,48:48,48:64,64:24,64:32,64:48,64:64,32:48,24:64,32:64" 8
timeout 300 python scripts/kmix.py "32:32" 256
} > gpurun_out/g44.log 2>&1
cat gpurun_out/g44.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep2_kernel -s 3 -c 1 -o gpurun_out/r02_sweep2_l2_k64 -f python scripts/kbench.py 64 > gpurun_out/g9_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/g9_ncu.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --sweeps-per-step 1 --no-cpu-baseline --no-fp32-extra $BARGS"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep2_kernel -s 300 -c 1 -o gpurun_out/${NAME:-prof} -f $B > gpurun_out/g5_ncu.log 2>&1
echo "ncu rc=$?"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
CHAINS=256 timeout 300 python scripts/kbench.py 64 > gpurun_out/g59_plain.log 2>&1; cat gpurun_out/g59_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep2_kernel -s 3 -c 1 -o gpurun_out/r02c_sweep2_l2_k64 -f python scripts/kbench.py 64 > gpurun_out/g59_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/g59_ncu.log
ncu -i gpurun_out/r02c_sweep2_l2_k64.ncu-rep --page raw --csv > gpurun_out/r02c_l2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02c_sweep2_l2_k64.ncu-rep --page source --csv --print-source sass > gpurun_out/r02c_l2_sass.csv 2>/dev/null
ls -la gpurun_out/r02c_*

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/full_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/full_tests.log
timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 --sweeps-per-step 2 > gpurun_out/g13_c5.json 2> gpurun_out/g13_c5.err; echo "c5 rc=$?"; cut -c1-2500 gpurun_out/g13_c5.json; tail -3 gpurun_out/g13_c5.err

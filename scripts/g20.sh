#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python scripts/repro_k0.py 2>&1 | tail -5
timeout 900 python -m pytest tests/test_parallel_gpu.py tests/test_parity_operating_point.py -m gpu -x -q -k "large_k or cluster or asymmetric or kats" > gpurun_out/g20_tests.log 2>&1; echo "tests rc=$?"; tail -n 8 gpurun_out/g20_tests.log

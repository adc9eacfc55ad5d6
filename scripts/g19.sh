#!/bin/bash
# batch 4: the cluster (DSMEM) form of sweep2 for K = 33 .. 64: tests, then K classes on the C3 graph against counts in L2
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_operating_point.py -m gpu -x -q -k "kats and cluster" > gpurun_out/g19_kat.log 2>&1; echo "kat rc=$?"; tail -n 5 gpurun_out/g19_kat.log
timeout 900 python -m pytest tests/test_parallel_gpu.py -m gpu -x -q -k "large_k or cluster or asymmetric" > gpurun_out/g19_tests.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/g19_tests.log
timeout 600 python scripts/kbench.py 48 64 40 2>&1 | tail -4
KERNEL=5 timeout 600 python scripts/kbench.py 48 64 2>&1 | tail -3

#!/bin/bash
# batch 3: histogram bin remap + folded "+1" (default) and the early log q / eta of the source block (lqr variant)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
VARIANTS="magic default lqr" STEPS=5 bash scripts/g4.sh
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g18_tests.log 2>&1; echo "tests(default) rc=$?"; tail -n 3 gpurun_out/g18_tests.log
BISBM_LIB=build/variants/libbisbm_lqr.so timeout 900 python -m pytest tests/test_parity_operating_point.py tests/test_parallel_gpu.py -m gpu -x -q > gpurun_out/g18_tests_lqr.log 2>&1; echo "tests(lqr) rc=$?"; tail -n 3 gpurun_out/g18_tests_lqr.log

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
VARIANTS="default bddearly default bddearly" STEPS=5 bash scripts/g4.sh
BISBM_LIB=build/variants/libbisbm_bddearly.so timeout 900 python -m pytest tests/test_parity_operating_point.py tests/test_parallel_gpu.py -m gpu -x -q -k "kats or invariants or large_graph" 2>&1 | tail -2

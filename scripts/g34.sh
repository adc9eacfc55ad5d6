#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
VARIANTS="default tqahead default tqahead" STEPS=5 bash scripts/g4.sh
BISBM_LIB=build/variants/libbisbm_tqahead.so timeout 900 python -m pytest tests/test_parity_operating_point.py tests/test_parallel_gpu.py -m gpu -x -q -k "kats or invariants or hub or k32 or isolated" 2>&1 | tail -3

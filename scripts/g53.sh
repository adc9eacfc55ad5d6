#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "edge_list_text or cli or grid" > gpurun_out/g53_tests.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/g53_tests.log

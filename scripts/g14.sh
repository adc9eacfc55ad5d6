#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/parity_operating_point.jsonl
timeout 2400 python -m pytest tests -x -q -m gpu -s > gpurun_out/full_tests_s.log 2>&1; echo "tests rc=$?"; grep -E "passed|failed|^FAILED|Error" gpurun_out/full_tests_s.log | tail -5
grep -E "vs oracle|burn-in:|known answers|marginals:|entropy oracle|grid minimum|fp32 vs fp64" gpurun_out/full_tests_s.log | cut -c1-400

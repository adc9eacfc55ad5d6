#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
{
timeout 300 python scripts/kmix.py "32:32" 256
LOGQ_EVERY=1000000 timeout 300 python scripts/kmix.py "32:32" 256
} > gpurun_out/g47.log 2>&1
cat gpurun_out/g47.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-fp32-extra > gpurun_out/g47_c3.json 2> gpurun_out/g47_c3.err; echo "c3 rc=$?"; python -c "
import json; r=json.load(open('gpurun_out/g47_c3.json')); print('%.4e'%r['value'], r['roofline']['frac'], '%.4e'%r['e2e']['value'])"
timeout 600 python -m pytest tests -x -q -m gpu -k "parity or kat or transition or statistical or invariants or expansion" > gpurun_out/g47_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g47_tests.log

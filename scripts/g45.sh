#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/g45_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g45_tests.log
{
timeout 300 python scripts/kmix.py "64:64" 80
timeout 300 python scripts/kmix.py "64:64" 96
timeout 300 python scripts/kmix.py "32:32" 256
} > gpurun_out/g45.log 2>&1
cat gpurun_out/g45.log
timeout 900 python bench.py --workload c4 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/g45_c4.json 2> gpurun_out/g45_c4.err; echo "c4 rc=$?"; tail -n 2 gpurun_out/g45_c4.err
python -c "
import json; r=json.load(open('gpurun_out/g45_c4.json')); print('%.4e'%r['value'], r['ms_per_step'])
for b in r['buckets_rank0_last_step']: print(b)"

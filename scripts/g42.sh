#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --workload c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g42_c4.json 2> gpurun_out/g42_c4.err; echo "c4 rc=$?"; tail -n 2 gpurun_out/g42_c4.err
python -c "
import json; r=json.load(open('gpurun_out/g42_c4.json')); print('%.4e'%r['value'], r['ms_per_step'])
for b in r['buckets_rank0_last_step']: print(b)"

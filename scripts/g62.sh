#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q -m gpu -k "grid" > gpurun_out/g62_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/g62_tests.log
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g62_c4.json 2> gpurun_out/g62_c4.err; echo "c4 rc=$?"; tail -n 2 gpurun_out/g62_c4.err
python -c "
import json; r=json.load(open('gpurun_out/g62_c4.json')); print('%.4e'%r['value'], r['ms_per_step'], r['imbalance'])
for b in r['buckets_rank0_last_step']: print(b)"

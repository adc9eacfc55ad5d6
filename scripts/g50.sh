#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parallel_gpu.py -x -q -m gpu -k "grid" > gpurun_out/g50_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g50_tests.log
timeout 600 python scripts/c4_overhead.py > gpurun_out/g50.log 2>&1; cat gpurun_out/g50.log
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g50_c4.json 2> gpurun_out/g50_c4.err; echo "c4 rc=$?"; tail -n 2 gpurun_out/g50_c4.err
python -c "
import json; r=json.load(open('gpurun_out/g50_c4.json')); print('%.4e'%r['value'], r['ms_per_step'], r['imbalance'])"

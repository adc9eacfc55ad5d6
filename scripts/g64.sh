#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/g64_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/g64_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 120 python bench.py --workload c4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/g64_c4.json 2> gpurun_out/g64_c4.err; echo "c4 rc=$?"; python -c "
import json; r=json.load(open('gpurun_out/g64_c4.json')); print('%.4e'%r['value'], r['ms_per_step'], sum(b['score_teardown_ms'] for b in r['buckets_rank0_last_step']))"

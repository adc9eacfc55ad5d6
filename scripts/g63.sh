#!/bin/bash
# final check of HEAD: GPU tests, smoke, default bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g63_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/g63_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/g63_bench.json 2> gpurun_out/g63_bench.err; echo "bench rc=$?"; python -c "
import json; r=json.load(open('gpurun_out/g63_bench.json')); print('value %.4e frac %.3f e2e %.4e (%.3f) launches %d' % (r['value'], r['roofline']['frac'], r['e2e']['value'], r['e2e']['value']/r['value'], r['gpu_launches']))"

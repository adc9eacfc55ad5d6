#!/bin/bash
# final check of HEAD: GPU tests, smoke, default bench at the driver's settings, marginalize workload, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g54_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/g54_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/g54_bench.json 2> gpurun_out/g54_bench.err; echo "bench rc=$?"; python -c "
import json; r=json.load(open('gpurun_out/g54_bench.json')); print('value %.4e frac %.3f e2e %.4e (%.3f) launches %d' % (r['value'], r['roofline']['frac'], r['e2e']['value'], r['e2e']['value']/r['value'], r['gpu_launches']))"
timeout 600 python bench.py --workload marginalize --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g54_marg.json 2> gpurun_out/g54_marg.err; echo "marg rc=$?"; cut -c1-200 gpurun_out/g54_marg.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/g54_launches.csv python bench.py --steps 1 --warmup 1 --sweeps-per-step 1 --no-cpu-baseline --no-fp32-extra > gpurun_out/g54_ncu.log 2>&1; echo "ncu rc=$?"

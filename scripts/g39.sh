#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g39_c3.json 2> gpurun_out/g39_c3.err; echo "c3 rc=$?"; python -c "
import json; r=json.load(open('gpurun_out/g39_c3.json')); print('%.4e'%r['value'], r['config']['sweep_plan'])"
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/g39_c4.json 2> gpurun_out/g39_c4.err; echo "c4 rc=$?"; tail -n 2 gpurun_out/g39_c4.err; cut -c1-500 gpurun_out/g39_c4.json
timeout 600 python bench.py --workload marginalize --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g39_marg.json 2> gpurun_out/g39_marg.err; echo "marg rc=$?"; cut -c1-400 gpurun_out/g39_marg.json

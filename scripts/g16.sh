#!/bin/bash
# batch 1 of the re-entry session: A/B of the beta hoist / 2^52 splice, e2e breakdown, fp32 instantiation check
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
VARIANTS="prev default magic" STEPS=3 bash scripts/g4.sh
timeout 300 python scripts/e2e_breakdown.py 4 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --precision fp32 > gpurun_out/g16_fp32.json 2> gpurun_out/g16_fp32.err; python -c "
import json; r=json.load(open('gpurun_out/g16_fp32.json')); print('fp32 bench', r['value'], r['dtype'], r['config']['sweep_plan'])"

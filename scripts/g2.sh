#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
rm -f gpurun_out/parity_operating_point.jsonl
timeout 900 python -m pytest -x -q -m gpu -s tests/test_parity_operating_point.py -k "parity_with_oracle" > gpurun_out/g2_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/g2_parity.log
grep -E "^\{|passed|failed|rc=" gpurun_out/g2_parity.log | cut -c1-700
B="python bench.py --steps 1 --warmup 1 --sweeps-per-step 1 --no-cpu-baseline --no-fp32-extra"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep2_kernel -s 300 -c 1 -o gpurun_out/r02_sweep2_f64 -f $B > gpurun_out/g2_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/g2_ncu2.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out | tail -8

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu -s tests/test_parallel_gpu.py -k "grid_search or large_k or asymmetric or u8_labels" > gpurun_out/g6_tests.log 2>&1
echo "tests rc=$?"; grep -E "grid minimum|passed|failed|Error" gpurun_out/g6_tests.log | tail -5
timeout 900 python -m pytest -x -q -m gpu -s tests/test_parity_operating_point.py -k "burn_in or stationary" > gpurun_out/g6_parity.log 2>&1
echo "parity rc=$?"; grep -E "burn-in:|passed|failed|Error" gpurun_out/g6_parity.log | tail -5
timeout 600 python bench.py --workload c4 --steps 2 --warmup 1 > gpurun_out/g6_c4.json 2> gpurun_out/g6_c4.err; echo "c4 rc=$?"; cat gpurun_out/g6_c4.json | cut -c1-1500; tail -3 gpurun_out/g6_c4.err
timeout 600 python bench.py --workload marginalize --steps 2 --warmup 1 --sweeps-per-step 4 > gpurun_out/g6_marg.json 2> gpurun_out/g6_marg.err; echo "marg rc=$?"; cat gpurun_out/g6_marg.json | cut -c1-1500; tail -3 gpurun_out/g6_marg.err

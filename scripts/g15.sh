#!/bin/bash
# round-end evidence: default bench, reference arm, ncu launch list + full capture of the sweep kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02b_bench_default.json
timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r02b_bench_reference.json
B="python bench.py --steps 1 --warmup 1 --sweeps-per-step 1 --no-cpu-baseline --no-fp32-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02b_launches_bench.csv $B > gpurun_out/g15_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sweep2_kernel -s 300 -c 1 -o gpurun_out/r02b_sweep2_f64_final -f $B > gpurun_out/g15_ncu2.log 2>&1; echo "ncu full rc=$?"

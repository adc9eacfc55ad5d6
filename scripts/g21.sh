#!/bin/bash
# C5 at the NAMED size on one GPU (50M nodes / 1B edges, Ka = Kb = 128, 32 chains): device ingest, sweeps, counts == rebuild.
# Falls back to 20M / 400M when the box has less than 150 GB of free host memory.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
free -g | head -2; nproc; nvidia-smi --query-gpu=memory.total,memory.used --format=csv,noheader
FREE=$(free -g | awk '/^Mem:/{print $7}')
if [ "$FREE" -ge 150 ]; then N=50000000; E=1000000000; else N=20000000; E=400000000; fi
echo "host free ${FREE} GB -> nodes $N edges $E"
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; free -g | awk '/^Mem:/{print "host used GB", $3}'; sleep 20; done ) > gpurun_out/g21_mem.log 2>&1 &
MON=$!
timeout 1500 python bench.py --workload c5 --c5-nodes $N --c5-edges $E --c5-chains 32 --steps 2 --warmup 1 --sweeps-per-step 1 --no-cpu-baseline > gpurun_out/r02_c5_full.json 2> gpurun_out/r02_c5_full.err; echo "c5 rc=$?"
kill $MON
tail -n 3 gpurun_out/r02_c5_full.err; cut -c1-900 gpurun_out/r02_c5_full.json; sort -u gpurun_out/g21_mem.log | tail -6

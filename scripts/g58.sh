#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 8 --workload c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g58_c4_n8.json 2> gpurun_out/g58_c4_n8.err; echo "c4 n8 rc=$?"; tail -n 3 gpurun_out/g58_c4_n8.err
python -c "
import json; r=json.load(open('gpurun_out/g58_c4_n8.json')); print('%.4e'%r['value'], r['ms_per_step'], r['imbalance'], r['best'])
for b in r['buckets_rank0_last_step']: print(b)"

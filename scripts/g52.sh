#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --workload c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/g52_c5_n8.json 2> gpurun_out/g52_c5_n8.err; echo "c5 n8 rc=$?"; tail -n 3 gpurun_out/g52_c5_n8.err
python -c "
import json
for l in open('gpurun_out/g52_c5_n8.json'):
    if l.startswith('{'):
        r=json.loads(l); print('%.4e'%r['value'], r['ms_per_step'], r['allreduce'], r['invariants_ok'], r['config']['graph_generate_s'], r['config']['graph_build_s'])"

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g56_n2.json 2> gpurun_out/g56_n2.err; echo "n2 rc=$?"; tail -n 2 gpurun_out/g56_n2.err
wc -l gpurun_out/g56_n2.json; python -c "
import json; r=json.load(open('gpurun_out/g56_n2.json')); print('value %.4e e2e %.4e n_gpus %d %s' % (r['value'], r['e2e']['value'], r['n_gpus'], r['config'].get('collective')))"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/g56_ref_n2.json 2> gpurun_out/g56_ref_n2.err; echo "ref n2 rc=$?"; wc -l gpurun_out/g56_ref_n2.json; cut -c1-200 gpurun_out/g56_ref_n2.json

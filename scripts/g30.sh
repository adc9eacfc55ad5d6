#!/bin/bash
# eight GPUs: bench.py under torchrun (N = 8), the driver's settings
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc=$?"
python -c "
import json
r=json.loads([l for l in open('gpurun_out/r02_bench_n8.json') if l.startswith('{')][-1]); print('N=8 value %.4e e2e %.4e (%.3f of value) frac %.3f clocks %s' % (r['value'], r['e2e']['value'], r['e2e']['value']/r['value'], r['roofline']['frac'], r['clocks']))"

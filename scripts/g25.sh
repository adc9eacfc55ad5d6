#!/bin/bash
# two GPUs: the tests that need a second device, then bench.py under torchrun (N = 2)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_cli.py tests/test_parallel_gpu.py -m gpu -x -q -k "gpus_mode or two_handles" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json
r=json.loads([l for l in open('gpurun_out/r02_bench_n2.json') if l.startswith('{')][-1]); print('N=2 value %.4e e2e %.4e (%.3f) frac %.3f collective %s' % (r['value'], r['e2e']['value'], r['e2e']['value']/r['value'], r['roofline']['frac'], r['config'].get('collective')))"
tail -n 3 gpurun_out/r02_bench_n2.err

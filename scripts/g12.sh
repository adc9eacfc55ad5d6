#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m pytest -x -q -m gpu tests/test_cli.py -k gpus_mode > gpurun_out/g12_cli.log 2>&1; echo "cli rc=$?"; tail -3 gpurun_out/g12_cli.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 2 --warmup 2 --no-fp32-extra > gpurun_out/g12_bench$N.json 2> gpurun_out/g12_bench$N.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/g12_bench$N.json; tail -3 gpurun_out/g12_bench$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c4 --steps 1 --warmup 1 > gpurun_out/g12_c4_$N.json 2> gpurun_out/g12_c4_$N.err; echo "c4 rc=$?"; cut -c1-1800 gpurun_out/g12_c4_$N.json; tail -3 gpurun_out/g12_c4_$N.err

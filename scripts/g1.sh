#!/bin/bash
# first GPU contact of the round-2 kernel: focused tests, then a short bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g1_smi.txt 2>&1
timeout 600 python -m pytest -x -q -m gpu -s \
  "tests/test_parity_operating_point.py::test_parallel_kernel_transition_kats" \
  "tests/test_parallel_gpu.py::test_invariants_small_graph" \
  "tests/test_parallel_gpu.py::test_heterogeneous_k_and_single_block_types" \
  "tests/test_parallel_gpu.py::test_k32_specialised_kernel" \
  "tests/test_parallel_gpu.py::test_large_graph_invariants_and_logq_expansion" \
  "tests/test_parallel_gpu.py::test_asymmetric_and_borderline_k" \
  "tests/test_parity_operating_point.py::test_parallel_kernel_transition_kats_large_blocks" \
  > gpurun_out/g1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g1_tests.log
tail -30 gpurun_out/g1_tests.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err
echo "bench rc=$?"
cat gpurun_out/g1_bench.json | head -c 3000
tail -5 gpurun_out/g1_bench.err

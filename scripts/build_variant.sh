#!/bin/bash
# build an experimental variant of libbisbm.so with extra -D flags:  scripts/build_variant.sh NAME -DFOO=1 ...
# -> build/variants/libbisbm_NAME.so ; run with BISBM_LIB=build/variants/libbisbm_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -o build/variants/libbisbm_$name.so bipartitesbm-mcmc_b200/csrc/capi.cu
echo build/variants/libbisbm_$name.so

#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in default xlogq xsmall; do
  if [ $v = default ]; then unset BISBM_LIB; else export BISBM_LIB=build/variants/libbisbm_$v.so; fi
  echo "== $v"; timeout 600 python scripts/fp32_switch.py 50 2>&1 | grep "^fp64 x" | awk '{for(i=NF-24;i<=NF-4;i+=2) printf "%s ", $i; print ""}'
done

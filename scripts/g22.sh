#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in default xlogq xsmall; do
  if [ $v = default ]; then unset BISBM_LIB; else export BISBM_LIB=build/variants/libbisbm_$v.so; fi
  echo "== $v"; timeout 600 python scripts/fp32_drift.py 40 2>&1 | grep "^fp64"
done

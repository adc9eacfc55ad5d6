#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g29_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g29_tests.log
bash scripts/g27.sh 2>&1 | grep -E "real|entropy|NMI" | head -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"

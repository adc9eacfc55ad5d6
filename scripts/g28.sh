#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
timeout 900 python -m pytest tests/test_merge_gpu.py tests/test_cli.py tests/test_replay_gpu.py -m gpu -x -q 2>&1 | tail -3
bash scripts/g27.sh 2>&1 | grep -E "real|entropy|NMI" | head -5

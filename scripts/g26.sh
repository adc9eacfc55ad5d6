#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
env | grep -i nccl
timeout 600 python -m pytest tests/test_cli.py -m gpu -x -q -k "gpus_mode" 2>&1 | tail -40

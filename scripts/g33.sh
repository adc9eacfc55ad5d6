#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
VARIANTS="default predc default predc" STEPS=5 bash scripts/g4.sh

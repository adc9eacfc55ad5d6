#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_operating_point.py -x -q -m gpu -k "quench" -s > gpurun_out/g57_tests.log 2>&1; echo "tests rc=$?"; grep -v "^{" gpurun_out/g57_tests.log | tail -n 8
python - <<'PY'
import json
for l in open('gpurun_out/g57_tests.log'):
    if l.startswith('{'):
        r = json.loads(l)
        print(r['test'], r['protocol'], 'oracle %.0f +- %.0f acc %.4f nmi %.4f | gpu %.0f +- %.0f acc %.4f nmi %.4f | p %s' % (
            r['oracle']['entropy_mean'], r['oracle']['entropy_sd'], r['oracle']['accept_mean'], r['oracle']['nmi_mean'],
            r['gpu']['entropy_mean'], r['gpu']['entropy_sd'], r['gpu']['accept_mean'], r['gpu']['nmi_mean'], {k: round(v, 3) for k, v in r['ks_p'].items()}))
PY

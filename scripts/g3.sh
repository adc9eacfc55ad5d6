#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest -x -q -m gpu tests/test_parity_operating_point.py -k "kats" > gpurun_out/g3_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/g3_tests.log
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
for v in default $VARIANTS; do
  if [ $v = default ]; then unset BISBM_LIB; else export BISBM_LIB=build/variants/libbisbm_$v.so; fi
  timeout 300 $B > gpurun_out/g3_$v.json 2> gpurun_out/g3_$v.err
  python - <<PY
import json
try:
    r=json.load(open('gpurun_out/g3_$v.json'))
    print('$v', 'value %.4e frac %.3f launch %.1f us e2e %.3e fp32 %s' % (r['value'], r['roofline']['frac'], r['roofline']['avg_launch_ms']*1e3, r['e2e']['value'], r.get('extra',{}).get('fp32_move_arithmetic',{}).get('value')))
except Exception as e:
    print('$v failed', e); print(open('gpurun_out/g3_$v.err').read()[-800:])
PY
done

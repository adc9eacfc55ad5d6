#!/bin/bash
# A/B of library variants and bench options on ONE box: VARIANTS="name[:bench args] ..." (name "default" = in-tree lib)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
STEPS=${STEPS:-3}
for spec in $VARIANTS; do
  v=${spec%%:*}; extra=""
  if [[ "$spec" == *:* ]]; then extra=$(echo "${spec#*:}" | tr ',' ' '); fi
  if [ $v = default ]; then unset BISBM_LIB; else export BISBM_LIB=build/variants/libbisbm_$v.so; fi
  tag=$(echo $spec | tr ':,=' '___' | tr -d '-')
  timeout 300 python bench.py --steps $STEPS --warmup 2 --no-cpu-baseline $extra > gpurun_out/g4_$tag.json 2> gpurun_out/g4_$tag.err
  python - <<PY
import json
try:
    r=json.load(open('gpurun_out/g4_$tag.json'))
    print('%-28s value %.4e frac %.3f launch %.1f us slice %d e2e %.3e acc %.4f fp32 %s' % ('$spec', r['value'], r['roofline']['frac'], r['roofline']['avg_launch_ms']*1e3, r['config']['sweep_plan']['slice_vertices_per_launch'], r['e2e']['value'], r['acceptance'], r.get('extra',{}).get('fp32_move_arithmetic',{}).get('value')))
except Exception as e:
    print('$spec failed', e); print(open('gpurun_out/g4_$tag.err').read()[-800:])
PY
done

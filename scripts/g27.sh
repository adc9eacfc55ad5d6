#!/bin/bash
# bin/mcmc -g on bisbm-1000 (singleton blocks: K = 500 + 500 at the start, ~480 rungs of the sigma = 1.01 ladder)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python - <<'PY'
import numpy as np
g = np.load("tests/golden/c2_const_k46.npz")
with open("/tmp/b1000.edgelist", "w") as f:
    for a, b in g["edges"]:
        f.write("%d\t%d\n" % (a, b))
PY
( time timeout 1200 bin/mcmc -e /tmp/b1000.edgelist -y 500 500 -n 500 500 -z 4 6 -g -c abrupt_cool -a 2000 -t 20000 -x 100000 -d 3 --gen_seed 12345 > /tmp/out.txt 2> /tmp/err.txt ) 2>&1 | tail -4; echo "rc done"
head -c 1500 /tmp/err.txt; grep -E "Ka, Kb|entropy|Elapsed|Maximum resident" /tmp/err.txt; python -c "
import numpy as np
lab=np.array(open('/tmp/out.txt').read().split(),dtype=int); g=np.load('tests/golden/c2_const_k46.npz')
print('labels', lab.size, 'blocks a', len(set(lab[:500])), 'b', len(set(lab[500:])), 'agreement with planted (best-match not computed): nmi-ish', )
import sys; sys.path.insert(0,'tests'); from helpers import nmi; print('NMI vs planted', nmi(lab, g['labels0']))"

"""fp32 vs fp64 move arithmetic of the parallel sweep on the statistical-parity workload of
tests/test_parallel_gpu.py (bisbm-1000, (Ka,Kb)=(4,6), randomised starts, abrupt_cool 1e5 steps then
greedy, 200 sweeps): final entropy / acceptance per precision over several seed sets, and the
oracle's numbers beside them.  Usage: python scripts/precision_study.py [n_seed_sets] [chains]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
host = importlib.import_module("bipartitesbm-mcmc_b200.host")
from conftest import load_golden  # noqa: E402

sets = int(sys.argv[1]) if len(sys.argv) > 1 else 4
R = int(sys.argv[2]) if len(sys.argv) > 2 else 96
g = load_golden("c2_const_k46")
na, nb, edges, mb = g["na"], g["nb"], g["edges"], g["labels0"]
n = na + nb
graph = host.Graph(edges, na, nb)
for prec in ("fp64", "fp32"):
    for inflight in (0, 1):
        ents, accs = [], []
        for k in range(sets):
            pool = host.ChainPool(graph, np.tile(mb, (R, 1)), 4, 6, 1.0)
            pool.set_precision(prec)
            seeds = np.arange(R, dtype=np.uint64) + 1000 * (k + 1)
            pool.randomize(seeds)
            acc, _ = pool.anneal("abrupt_cool", 1e5, 0.0, 200 * n, 10 ** 9, seeds, max_inflight=inflight)
            ents.append(pool.entropy())
            accs.append(acc)
        e = np.concatenate(ents)
        a = np.concatenate(accs)
        print("%s inflight=%d: entropy %.1f +- %.1f (sem %.1f)  acceptance %.4f +- %.4f  [%d chains]" % (
            prec, inflight, e.mean(), e.std(), e.std() / np.sqrt(len(e)), a.mean(), a.std(), len(e)))
if len(sys.argv) > 3:
    from oracle import port
    eo, ao = [], []
    for s in range(int(sys.argv[3])):
        o = port.PortChain(n, na, nb, edges, mb, 4, 6, 1.0, 1000 + s, 2000 + s)
        o.init(True)
        ao.append(o.anneal("abrupt_cool", 1e5, 0, 200 * n, 10 ** 9))
        eo.append(o.entropy())
    eo, ao = np.array(eo), np.array(ao)
    print("oracle: entropy %.1f +- %.1f (sem %.1f)  acceptance %.4f +- %.4f  [%d chains]" % (
        eo.mean(), eo.std(), eo.std() / np.sqrt(len(eo)), ao.mean(), ao.std(), len(eo)))

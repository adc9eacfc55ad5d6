"""moves/s of successive 4-sweep calls from a randomised start on the C3 graph, fp32 vs fp64 move arithmetic:
does the throughput depend on how far the chains have equilibrated?  usage: python scripts/fp32_drift.py [calls]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import planted
host = importlib.import_module("bipartitesbm-mcmc_b200").host
na = nb = 500000; n = na + nb; ka = kb = 32; C = 256
edges = planted(na, nb, ka, kb, 10_000_000, 0)
graph = host.Graph(edges, na, nb)
lab = np.concatenate([np.arange(na) * ka // na, ka + np.arange(nb) * kb // nb]).astype(np.uint32)
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for prec in ("fp32", "fp64"):
    pool = host.ChainPool(graph, np.broadcast_to(lab, (C, n)), ka, kb, 1.0)
    pool.set_precision(prec)
    seeds = np.arange(C, dtype=np.uint64) + 1
    pool.randomize(seeds)
    out = []
    for k in range(calls):
        acc, _ = pool.anneal("constant", 1.0, 0.0, 4 * n, 10 ** 18, seeds)
        ms, la, mv = pool.last_timing()
        out.append((mv / ms * 1e3, float(acc.mean())))
    print(prec, " ".join("%.2e/%.2f" % o for o in out[::4]), flush=True)
    del pool
# block sizes of a few chains after the last call: does a block shrink out of the log q expansion's range (e_r < 1024 d)?
pool = host.ChainPool(graph, np.broadcast_to(lab, (C, n)), ka, kb, 1.0)
seeds = np.arange(C, dtype=np.uint64) + 1
pool.randomize(seeds)
for k in range(calls):
    pool.anneal("constant", 1.0, 0.0, 4 * n, 10 ** 18, seeds)
    if k % 8 == 7 or k == calls - 1:
        nr = np.stack([pool.n_r(c) for c in (0, 100, 255)]); er = np.stack([pool.m_r(c) for c in (0, 100, 255)])
        print("after %d sweeps: n_r min %s max %s   e_r min %s" % (4 * (k + 1), nr.min(1), nr.max(1), er.min(1)), flush=True)

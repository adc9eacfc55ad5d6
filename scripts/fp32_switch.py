"""the bench's 'extra' measurement in isolation: evolve 256 chains in double, switch the same pool to fp32 and back"""
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
pool = host.ChainPool(graph, np.broadcast_to(lab, (C, n)), ka, kb, 1.0)
seeds = np.arange(C, dtype=np.uint64) + 1
pool.randomize(seeds)
def run(k, tag):
    out = []
    for _ in range(k):
        pool.anneal("constant", 1.0, 0.0, 4 * n, 10 ** 18, seeds)
        ms, la, mv = pool.last_timing()
        out.append(mv / ms * 1e3)
    print(tag, " ".join("%.2e" % o for o in out), pool.sweep_info(), flush=True)
sw = int(sys.argv[1]) if len(sys.argv) > 1 else 45
run(sw, "fp64 x%d" % sw)
pool.set_precision("fp32"); run(4, "fp32")
pool.set_precision("fp64"); run(3, "fp64")
pool.set_precision("fp32"); run(3, "fp32")

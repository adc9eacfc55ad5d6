import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import planted
host = importlib.import_module("bipartitesbm-mcmc_b200").host
na = nb = 3000; ka = kb = 48
edges = planted(na, nb, 8, 8, 60000, 11)
for kernel in (5, 0, -1, 0):
    graph = host.Graph(edges, na, nb)
    C = 33
    lab0 = np.concatenate([np.arange(na) % ka, ka + np.arange(nb) % kb]).astype(np.uint32)
    pool = host.ChainPool(graph, np.tile(lab0, (C, 1)), ka, kb, 1.0)
    pool.set_option("kernel", kernel)
    seeds = np.arange(C, dtype=np.uint64) + 31
    pool.randomize(seeds)
    try:
        pool.anneal("constant", 1.0, 0.0, 4 * (na + nb), 10 ** 9, seeds)
        print(kernel, "ok", pool.sweep_info())
    except Exception as e:
        print(kernel, "FAILED", e)

"""moves/s of one pool whose chains have DIFFERENT (Ka, Kb) (a grid-search bucket) on the C3 graph.
usage: python scripts/kmix.py "48:32,48:48,64:64" [restarts]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import planted
host = importlib.import_module("bipartitesbm-mcmc_b200").host
na = nb = 500000
n = na + nb
edges = planted(na, nb, 32, 32, 10_000_000, 0)
graph = host.Graph(edges, na, nb)
pts = [tuple(int(x) for x in p.split(":")) for p in sys.argv[1].split(",")]
restarts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kas = np.array([a for a, b in pts for _ in range(restarts)], dtype=np.uint32)
kbs = np.array([b for a, b in pts for _ in range(restarts)], dtype=np.uint32)
C = len(kas)
lab = np.stack([np.concatenate([np.arange(na) * a // na, a + np.arange(nb) * b // nb]) for a, b in zip(kas, kbs)]).astype(np.uint32)
pool = host.ChainPool(graph, lab, kas, kbs, 1.0)
seeds = np.arange(C, dtype=np.uint64) + 1
if os.environ.get("LOGQ_EVERY"):
    pool.set_option("logq_every", int(os.environ["LOGQ_EVERY"]))
pool.randomize(seeds)
for what, args in (("first 2 sweeps, T=1", ("constant", 1.0, 0.0, 2 * n)), ("next 2 sweeps, T=1", ("constant", 1.0, 0.0, 2 * n)),
                   ("2 greedy sweeps", ("abrupt_cool", 0.0, 0.0, 2 * n)), ("2 more greedy sweeps", ("abrupt_cool", 0.0, 0.0, 2 * n))):
    acc, _ = pool.anneal(args[0], args[1], args[2], args[3], 10 ** 18, seeds)
    ms, la, mv = pool.last_timing()
    print("%s chains=%d plan=%s  %-22s %.3e moves/s  acc %.3f  occupied %s" % (sys.argv[1][:40], C, pool.sweep_info(), what, mv / ms * 1e3, acc.mean(),
          pool.occupied_blocks().min(0) if hasattr(pool, "occupied_blocks") else ""), flush=True)

"""moves/s of one pool of chains per K on the C3 graph (device events): which kernel / plan each K class gets.
usage: python scripts/kbench.py [K ...]   (K = Ka = Kb; "a:b" for asymmetric)"""
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
C = int(os.environ.get("CHAINS", "256"))
for arg in (sys.argv[1:] or ["8", "16", "32", "48", "64"]):
    ka, kb = (int(x) for x in arg.split(":")) if ":" in arg else (int(arg), int(arg))
    lab = np.concatenate([np.arange(na) * ka // na, ka + np.arange(nb) * kb // nb]).astype(np.uint32)
    pool = host.ChainPool(graph, np.broadcast_to(lab, (C, n)), ka, kb, 1.0)
    if os.environ.get("GENERIC"):
        pool.set_option("generic", 1)
    if os.environ.get("KERNEL"):
        pool.set_option("kernel", int(os.environ["KERNEL"]))
    seeds = np.arange(C, dtype=np.uint64) + 1
    pool.randomize(seeds)
    if os.environ.get("INFLIGHT_DIV"):
        pool.set_option("inflight_div", int(os.environ["INFLIGHT_DIV"]))
    pool.anneal("constant", 1.0, 0.0, 1 * n, 10 ** 18, seeds)
    if os.environ.get("SCHED") == "abrupt":      # one hot + one greedy sweep, as a grid-search step has them
        acc, _ = pool.anneal("abrupt_cool", 1.0 * n, 0.0, 2 * n, 10 ** 18, seeds)
    else:
        acc, _ = pool.anneal("constant", 1.0, 0.0, 2 * n, 10 ** 18, seeds)
    ms, la, mv = pool.last_timing()
    print("K=%s chains=%d kernel/wpc/cpg/slice=%s  %.3e moves/s  acc %.3f  launches %d" % (arg, C, pool.sweep_info(), mv / ms * 1e3, acc.mean(), la), flush=True)
    del pool

"""BASELINE configs[1]: bisbm-1000 marginalisation, Ka=Kb=10, 1000 burn-in sweeps, 20000 sampling sweeps, one sample
every 10 sweeps, over a pool of chains (default 256).  Prints wall time and moves/s.  Needs a GPU."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bipartitesbm-mcmc_b200")
host = pkg.host
z = np.load(os.path.join(ROOT, "tests", "golden", "c2_abrupt.npz"))
edges, lab0 = z["edges"], z["labels0"]
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
graph = host.Graph(edges, 500, 500)
pool = host.ChainPool(graph, np.tile(lab0, (C, 1)), 10, 10, 1.0)
seeds = np.arange(C, dtype=np.uint64) + 1
pool.marginals_clear()
t0 = time.perf_counter()
pool.marginalize(sweeps // 20, sweeps, 10, seeds)
dt = time.perf_counter() - t0
ms, launches, moves = pool.last_timing()
hist = pool.marginals()
print("chains %d sweeps %d+%d: wall %.2f s, device %.2f s, %d launches, %.3e moves/s, samples/node %d" % (
    C, sweeps // 20, sweeps, dt, ms / 1e3, launches, moves / dt, hist[0].sum()))

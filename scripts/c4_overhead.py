"""where a bisbm_grid_search step's wall time goes: per step, wall clock vs the per-bucket report (set-up, anneal, teardown)."""
import importlib, os, subprocess, sys, time
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
vals = (2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64)
points = [(a, b) for a in vals for b in vals]
smi = None
for i in range(8):
    if i == 4:
        smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-lms", "200"],
                               stdout=subprocess.DEVNULL)
    t0 = time.perf_counter()
    ent, acc, best, lab, st = host.grid_search(graph, points, 8, 1.0, "abrupt_cool", 2.0 * n, 0.0, 4 * n, 10 ** 18, seed=i + 1)
    wall = (time.perf_counter() - t0) * 1e3
    rep = st["report"]
    print("step %d smi=%s wall %.1f ms  device %.1f  setup %.1f  anneal(host) %.1f  score+teardown %.1f  unaccounted %.1f" % (
        i, smi is not None, wall, st["device_ms"], sum(r["setup_ms"] for r in rep), sum(r["anneal_ms"] for r in rep),
        sum(r["score_teardown_ms"] for r in rep),
        wall - sum(r["setup_ms"] + r["anneal_ms"] + r["score_teardown_ms"] for r in rep)), flush=True)
if smi:
    smi.terminate()

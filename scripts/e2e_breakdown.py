"""Where an end-to-end step of bench.py goes: host labels in (pinned u8) -> set_labels -> anneal (4 sweeps) -> labels out.
usage: python scripts/e2e_breakdown.py [steps]   (C3 graph, Ka = Kb = 32, 256 chains)"""
import importlib, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import planted
host = importlib.import_module("bipartitesbm-mcmc_b200").host
na = nb = 500000
n = na + nb
ka = kb = 32
C = 256
edges = planted(na, nb, ka, kb, 10_000_000, 0)
graph = host.Graph(edges, na, nb)
lab = np.concatenate([np.arange(na) * ka // na, ka + np.arange(nb) * kb // nb]).astype(np.uint32)
pool = host.ChainPool(graph, np.broadcast_to(lab, (C, n)), ka, kb, 1.0)
seeds = np.arange(C, dtype=np.uint64) + 1
pool.randomize(seeds)
pool.anneal("constant", 1.0, 0.0, 4 * n, 10 ** 18, seeds)
a = torch.empty((C, n), dtype=torch.uint8).pin_memory()
b = torch.empty((C, n), dtype=torch.uint8).pin_memory()
pool.labels(out=a.numpy())
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = {"set_labels": 0.0, "anneal": 0.0, "anneal_dev_ms": 0.0, "labels": 0.0}
# raw copy speeds for reference
d = torch.empty((C, n), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); h2d = time.perf_counter() - t0
t0 = time.perf_counter(); b.copy_(d, non_blocking=True); torch.cuda.synchronize(); d2h = time.perf_counter() - t0
print("raw torch copies of %d MB: h2d %.2f ms (%.1f GB/s)  d2h %.2f ms (%.1f GB/s)" % (C * n >> 20, h2d * 1e3, C * n / h2d / 1e9, d2h * 1e3, C * n / d2h / 1e9))
for step in range(steps + 1):
    v = step % n
    a.numpy()[0, v] = (a.numpy()[0, v] + 1) % ka
    t0 = time.perf_counter(); pool.set_labels(a.numpy()); t1 = time.perf_counter()
    pool.anneal("constant", 1.0, 0.0, 4 * n, 10 ** 18, seeds); t2 = time.perf_counter()
    pool.labels(out=b.numpy()); t3 = time.perf_counter()
    a, b = b, a
    if step:   # first pass = warm-up
        T["set_labels"] += t1 - t0; T["anneal"] += t2 - t1; T["labels"] += t3 - t2
        T["anneal_dev_ms"] += pool.last_timing()[0]
tot = T["set_labels"] + T["anneal"] + T["labels"]
print("per step (ms): set_labels %.2f  anneal %.2f (device events %.2f)  labels %.2f  total %.2f  -> e2e/resident %.3f" % (
    T["set_labels"] / steps * 1e3, T["anneal"] / steps * 1e3, T["anneal_dev_ms"] / steps, T["labels"] / steps * 1e3, tot / steps * 1e3,
    T["anneal_dev_ms"] / steps / (tot / steps * 1e3)))

"""How much concurrency (moves of one chain evaluated against counts that do not include each other) can the
parallel sweep take before the sampled distribution moves?  Runs the same pool of chains with
max_inflight = 1 (strictly sequential chains), n/64 (default), n/16, n/4 and compares the samples of the
description length with two-sample KS tests against the sequential run.  Needs a GPU:
    python scripts/staleness_study.py"""
import importlib
import os
import sys

import numpy as np
from scipy.stats import ks_2samp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from helpers import planted, planted_labels  # noqa: E402

pkg = importlib.import_module("bipartitesbm-mcmc_b200")
host = pkg.host
na = nb = 40000
ka = kb = 8
edges = planted(na, nb, ka, kb, 600000, 5)
graph = host.Graph(edges, na, nb)
C = 256
n = na + nb
base = np.tile(planted_labels(na, nb, ka, kb), (C, 1))
out = {}
for name, infl, sweeps in [("seq", 1, 40), ("n/64", na // 64, 40), ("n/16", na // 16, 40), ("n/4", na // 4, 40), ("n/1", na, 40)]:
    pool = host.ChainPool(graph, base, ka, kb, 1.0)
    seeds = np.arange(C, dtype=np.uint64) + 1000
    pool.randomize(seeds)
    acc, _ = pool.anneal("constant", 1.0, 0.0, sweeps * n, 10 ** 18, seeds + (0 if name == "seq" else 50000), max_inflight=infl)
    ent = pool.entropy()
    ms, _, mv = pool.last_timing()
    out[name] = (ent, acc)
    print("%-5s inflight %6d  entropy %.1f +- %.1f  accept %.4f  %.2e moves/s" % (name, infl, ent.mean(), ent.std(), acc.mean(), mv / ms * 1e3))
for name in out:
    if name == "seq":
        continue
    print("KS vs sequential: %-5s entropy p=%.3f  accept p=%.3f" % (name, ks_2samp(out["seq"][0], out[name][0]).pvalue,
                                                                     ks_2samp(out["seq"][1], out[name][1]).pvalue))

"""Generate tests/golden/merge.npz from the UNMODIFIED reference (oracle/_ref/libref_log.so): the agglomerative merge /
split initialiser, blockmodel_t::agg_merge (src/blockmodel.cc:109-256) and the merge path of main (src/mcmc_main.cc:406-451).
Run in the authoring container only:   python tests/golden/make_merge_fixture.py

Every case stores its inputs (graph = one of the committed goldens, initial labels, seeds, arguments) and the reference's
outputs after each call: labels, (ka, kb), m_rs, e_r, n_r, RNG word counts of both engines, entropy()."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def graph(name):
    g = np.load(os.path.join(OUT, name + ".npz"))
    return g["edges"], int(g["na"]), int(g["nb"])


def state(ch, tag, out):
    out[tag + "_labels"] = ch.labels()
    out[tag + "_k"] = np.array([ch.ka, ch.kb])
    out[tag + "_m"] = ch.m()
    out[tag + "_m_r"] = ch.m_r()
    out[tag + "_n_r"] = ch.n_r()
    out[tag + "_words"] = np.array(ch.rng_words(), dtype=np.uint64)
    out[tag + "_entropy"] = ch.entropy()


out = {}

# ---- case sw: southernWomen from singletons (the -g start), two merge calls, then the --nature form
edges, na, nb = graph("c1_seed1")
n = na + nb
lab = np.arange(n, dtype=np.uint32)
ch = ref.RefChain(n, na, nb, edges, lab, na, nb, 1.0, 7, log_rng=True)
ch.init(False)
out["sw_labels0"] = lab
ch.agg_merge(6, 4, 10); state(ch, "sw_s1", out)
ch.agg_merge(5, 0, 10); state(ch, "sw_s2", out)
ch.agg_merge(0, 4, 10); state(ch, "sw_s3", out)
ch.agg_merge(3, None, 10); state(ch, "sw_s4", out)
ch.anneal("abrupt_cool", 0.0, 0.0, n, 10 ** 9); state(ch, "sw_s5", out)
ch.agg_merge(2, 2, 10); state(ch, "sw_s6", out)

# ---- case b1000: bisbm-1000 with every planted block cut in two (8 + 12 blocks), merged back to (4, 6)
edges, na, nb = graph("c2_const_k46")
n = na + nb
g = np.load(os.path.join(OUT, "c2_const_k46.npz"))
l0 = g["labels0"].astype(np.int64)            # (4, 6) labels
par = np.arange(n) % 2
lab = np.where(np.arange(n) < na, 2 * l0 + par, 8 + 2 * (l0 - 4) + par).astype(np.uint32)
out["b1000_labels0"] = lab
ch = ref.RefChain(n, na, nb, edges, lab, 8, 12, 1.0, 3, log_rng=True)
ch.init(False)
ch.agg_merge(4, 6, 10); state(ch, "b1000_s1", out)
# the whole merge path of main for the same start (ladder with sigma = 1.01, greedy sweeps, final abrupt_cool anneal)
ch = ref.RefChain(n, na, nb, edges, lab, 8, 12, 1.0, 11, log_rng=True)
ch.init(False)
out["b1000_path_args"] = np.array([4, 6, 2000.0, 20 * n, 10 ** 6], dtype=np.float64)   # KA, KB, p0, sampling_steps, steps_await
out["b1000_path_entropy"] = ch.merge_path(4, 6, 2000.0, 20 * n, 10 ** 6)
state(ch, "b1000_path", out)

# ---- split (negative diff): the reference's compute_dS(mb, split_move) reads its vector out of bounds, so only what is
#      well defined is recorded: the engine words the shuffles consume and the new block counts
edges, na, nb = graph("c1_seed1")
n = na + nb
g = np.load(os.path.join(OUT, "c1_seed1.npz"))
lab = g["labels0"].astype(np.uint32)
ka0 = int(lab[:na].max()) + 1
kb0 = int(lab.max()) + 1 - ka0
out["split_labels0"] = lab
out["split_k0"] = np.array([ka0, kb0])
ch = ref.RefChain(n, na, nb, edges, lab, ka0, kb0, 1.0, 5, log_rng=True)
ch.init(False)
ch.agg_merge(-1, 0, 10)
out["split_s1_k"] = np.array([ch.ka, ch.kb])
out["split_s1_words"] = np.array(ch.rng_words(), dtype=np.uint64)
out["split_s1_n_r"] = ch.n_r()

# ---- the -g path (singletons -> (2, 3)) and the -g -u path of main on southernWomen; the CLI tests print these label lines
edges, na, nb = graph("c1_seed1")
n = na + nb
lab = np.arange(n, dtype=np.uint32)
ch = ref.RefChain(n, na, nb, edges, lab, na, nb, 1.0, 9, log_rng=True)
ch.init(False)
out["swg_args"] = np.array([2, 3, 100.0, 10 * n, 1000], dtype=np.float64)
out["swg_entropy"] = ch.merge_path(2, 3, 100.0, 10 * n, 1000)
state(ch, "swg", out)
ch = ref.RefChain(n, na, nb, edges, lab, na, nb, 1.0, 9, log_rng=True)
ch.init(False)
out["swu_entropy"] = ch.nature_path(100.0, 10 * n, 1000)
state(ch, "swu", out)

np.savez_compressed(os.path.join(OUT, "merge.npz"), **out)
for k in sorted(out):
    v = np.asarray(out[k])
    print(k, v.shape, v.ravel()[:8])

"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref*.so, built by
`make -C oracle ref` from /root/reference/src with the Boost stand-ins and the deterministic
random_device header).  Run in the authoring container only:

    python tests/golden/make_golden.py

Each file holds the inputs (edge list in file order, initial labels, parameters) AND the
reference's outputs, so the tests need neither /root/reference nor the reference build.
The shipped datasets are read from /root/reference/dataset and stored as integer arrays.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
DS = "/root/reference/dataset/"
SW = DS + "southernWomen.edgelist"
BI = DS + "bisbm-n_1000-ka_4-kb_6-r-1.0-Ka_30-Ir_1.75.gt.edgelist"
MB = DS + "optional_membership_file.txt"


def planted(na, nb, ka, kb, n_edges, seed, ratio=10.0):
    """Synthetic planted bipartite SBM: block of type-a node i = i*ka//na,
    block pair weight `ratio` on the paired diagonal, 1 elsewhere; multi-edges kept."""
    rng = np.random.default_rng(seed)
    w = np.ones((ka, kb))
    for r in range(ka):
        w[r, r * kb // ka] = ratio
    p = (w / w.sum()).ravel()
    pair = rng.choice(ka * kb, size=n_edges, p=p)
    r, s = pair // kb, pair % kb
    ba = np.arange(na) * ka // na
    bb = np.arange(nb) * kb // nb
    a_start = np.searchsorted(ba, np.arange(ka))
    a_cnt = np.bincount(ba, minlength=ka)
    b_start = np.searchsorted(bb, np.arange(kb))
    b_cnt = np.bincount(bb, minlength=kb)
    ea = a_start[r] + (rng.random(n_edges) * a_cnt[r]).astype(np.int64)
    eb = na + b_start[s] + (rng.random(n_edges) * b_cnt[s]).astype(np.int64)
    return np.stack([ea, eb], 1).astype(np.uint32)


def trajectory(name, edges, na, nb, labels, ka, kb, eps, seed, randomize, schedule, p0, p1, duration, steps_await,
               gen_seed=12345, kats=()):
    n = na + nb
    c = ref.RefChain(n, na, nb, edges, labels, ka, kb, eps, seed, gen_seed, log_rng=True)
    c.init(randomize)
    out = dict(edges=edges, na=na, nb=nb, labels0=np.asarray(labels, dtype=np.uint32), ka=ka, kb=kb, eps=eps, seed=seed,
               gen_seed=gen_seed, randomize=int(randomize), schedule=ref.SCHEDULES[schedule], p0=np.float32(p0),
               p1=np.float32(p1), duration=duration, steps_await=steps_await)
    out["init_labels"] = c.labels()
    out["init_entropy"] = c.entropy()
    out["init_m"] = c.m()
    out["init_m_r"] = c.m_r()
    out["init_n_r"] = c.n_r()
    out["init_eta"] = c.eta()
    kv, ks, kd, ka_ = [], [], [], []
    for (v, s) in kats:  # RNG-free transition_ratio known answers on the initial state
        dS, ar = c.transition(v, s)
        kv.append(v); ks.append(s); kd.append(dS); ka_.append(ar)
    out["kat_v"] = np.array(kv, dtype=np.uint32)
    out["kat_s"] = np.array(ks, dtype=np.uint32)
    out["kat_dS"] = np.array(kd, dtype=np.float64)
    out["kat_accu"] = np.array(ka_, dtype=np.float64)
    out["accept"] = c.anneal(schedule, p0, p1, duration, steps_await)
    out["labels"] = c.labels()
    out["entropy"] = c.entropy()
    out["entropy_accum"] = c.entropy_accum()
    out["m"] = c.m()
    out["m_r"] = c.m_r()
    out["n_r"] = c.n_r()
    out["eta"] = c.eta()
    out["vlist"] = c.vlist()
    out["rng_words"] = np.array(c.rng_words(), dtype=np.uint64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "accept", out["accept"], "entropy", out["entropy"], "dS", out["entropy_accum"], "words", out["rng_words"])
    c.close()


def all_kats(labels, n, ka, kb, step):
    lab = np.asarray(labels)
    out = []
    for v in range(0, n, step):
        for s in range(ka + kb):
            out.append((v, s))
    return out


def main():
    if not ref.available(True):
        raise SystemExit("build the reference first: make -C oracle ref")
    sw = ref.load_edge_list(SW)
    bi = ref.load_edge_list(BI)
    mb = np.loadtxt(MB, dtype=np.uint32)
    lab_sw = ref.labels_from_block_sizes([4, 4, 4, 3, 3, 3, 3, 3, 3, 2])
    lab_bi = ref.labels_from_block_sizes([50] * 20)
    # C1: the Travis / README maximisation command (reference .travis.yml:27), seeds 1 and 42
    for seed in (1, 42):
        trajectory("c1_seed%d" % seed, sw, 18, 14, lab_sw, 5, 5, 1e-3, seed, True, "exponential", 10, 0.1, 1000, 100,
                   kats=all_kats(lab_sw, 32, 5, 5, 1))
    # C2 as the code runs the README marginalisation command (README.md:87)
    trajectory("c2_abrupt", bi, 500, 500, lab_bi, 10, 10, 1.0, 1, False, "abrupt_cool", 100, 0, 20000, 100,
               kats=all_kats(lab_bi, 1000, 10, 10, 37))
    # C2 graph, planted partition file, T = 1 sampling with randomised start, (4, 6)
    trajectory("c2_const_k46", bi, 500, 500, mb, 4, 6, 1.0, 7, True, "constant", 1, 0, 5000, 10 ** 9,
               kats=all_kats(mb, 1000, 4, 6, 91))
    # the remaining schedules on southernWomen
    trajectory("c1_linear", sw, 18, 14, lab_sw, 5, 5, 0.5, 3, True, "linear", 3.0, 0.002, 1280, 10 ** 9)
    trajectory("c1_log", sw, 18, 14, lab_sw, 5, 5, 2.0, 5, False, "logarithmic", 1.5, 2, 960, 10 ** 9)
    trajectory("c1_abrupt_stop", sw, 18, 14, lab_sw, 5, 5, 1e-3, 9, True, "abrupt_cool", 200, 0, 3200, 64)
    # single-block types: proposal short-circuit (src/blockmodel.cc:614-615)
    trajectory("c1_ka1", sw, 18, 14, np.array([0] * 18 + [1] * 7 + [2] * 7, dtype=np.uint32), 1, 2, 1.0, 11, True,
               "constant", 1, 0, 640, 10 ** 9)
    # large blocks: e_r >= 10001 so log q takes the asymptotic branch (src/support/int_part.cc:88-98)
    big = planted(600, 600, 2, 2, 60000, 0)
    lab_big = np.concatenate([np.arange(600) * 2 // 600, 2 + np.arange(600) * 2 // 600]).astype(np.uint32)
    trajectory("big_blocks", big, 600, 600, lab_big, 2, 2, 1.0, 2, True, "constant", 1, 0, 2400, 10 ** 9,
               kats=all_kats(lab_big, 1200, 2, 2, 53))
    # isolated nodes (degree 0) and a node count that makes std::shuffle take the odd/even paths
    iso = planted(40, 31, 3, 2, 90, 4)
    lab_iso = np.concatenate([np.arange(45) * 3 // 45, 3 + np.arange(36) * 2 // 36]).astype(np.uint32)
    iso_edges = iso.copy()
    iso_edges[:, 1] += 5  # type-a ids 0..44 (40..44 isolated), type-b ids 45..80 (76..80 isolated)
    trajectory("isolated", iso_edges, 45, 36, lab_iso, 3, 2, 0.1, 13, True, "exponential", 5, 0.999, 810, 10 ** 9,
               kats=all_kats(lab_iso, 81, 3, 2, 8))

    # math known answers
    ref.init_tables(20000)
    ns = [1, 2, 3, 5, 10, 17, 100, 999, 5000, 9999, 10000]
    qn, qk, qv = [], [], []
    for n in ns:
        for k in sorted(set([1, 2, 3, n // 2 if n > 1 else 1, max(1, n - 1), n, n + 5])):
            qn.append(n); qk.append(k); qv.append(ref.log_q(n, k))
    an, ak, av = [], [], []
    for n in [10001, 12345, 65536, 100000, 312500, 10 ** 6, 10 ** 7, 2 * 10 ** 8]:
        for k in [1, 2, 5, 9, 10, 11, 17, 18, 31, 32, 100, 1000, 15625, n // 3, n]:
            if k <= n:
                an.append(n); ak.append(k); av.append(ref.log_q_approx(n, k))
    xs = np.concatenate([np.linspace(0, 4, 81), np.array([1e-300, 1e-12, 0.4999, 0.5, 1.5, 1.5001, 2.0, 2.0001, 10, 1e6])])
    sp = np.array([ref.spence(float(x)) for x in xs])
    lg_i = np.array([1, 2, 3, 10, 171, 172, 1000, 30001, 2 * 10 ** 6], dtype=np.uint64)
    lg = np.array([ref.lgamma_fast(int(i)) for i in lg_i])
    sc = []
    for sid, (p0, p1) in enumerate([(10, 0.1), (3.0, 0.002), (1.5, 2), (0.75, 0), (100, 0)]):
        for t in [0, 1, 2, 31, 99, 100, 101, 1000, 123457, 2 ** 24 + 1, 3 * 10 ** 9]:
            sc.append((sid, p0, p1, t, ref.schedule(sid, p0, p1, t)))
    sc = np.array(sc, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "math.npz"), q_n=np.array(qn), q_k=np.array(qk), q_v=np.array(qv),
                        a_n=np.array(an, dtype=np.uint64), a_k=np.array(ak, dtype=np.uint64), a_v=np.array(av),
                        sp_x=xs, sp_v=sp, lg_i=lg_i, lg_v=lg, sched=sc)
    print("math.npz written")


if __name__ == "__main__":
    main()

"""numpy restatements used by several tests (not product code)."""
import numpy as np


def counts_from_labels(edges, na, nb, labels, ka, kb):
    """m (K x K symmetric), e_r, n_r, eta rebuilt from scratch like compute_m / compute_m_r /
    compute_n_r / compute_eta_rk (reference src/blockmodel.cc:691-746)."""
    K = ka + kb
    n = na + nb
    labels = np.asarray(labels, dtype=np.int64)
    ea, eb = edges[:, 0].astype(np.int64), edges[:, 1].astype(np.int64)
    m = np.zeros((K, K), dtype=np.int64)
    np.add.at(m, (labels[ea], labels[eb]), 1)
    np.add.at(m, (labels[eb], labels[ea]), 1)
    deg = np.bincount(np.concatenate([ea, eb]), minlength=n)
    eta = np.zeros((K, deg.max() + 1), dtype=np.int64)
    np.add.at(eta, (labels, deg), 1)
    return m, m.sum(1), np.bincount(labels, minlength=K), eta


def planted(na, nb, ka, kb, n_edges, seed, ratio=10.0):
    """Synthetic planted bipartite SBM of SURVEY.md 8(d) (same generator as tests/golden/make_golden.py
    and bench.py)."""
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


def planted_labels(na, nb, ka, kb):
    return np.concatenate([np.arange(na) * ka // na, ka + np.arange(nb) * kb // nb]).astype(np.uint32)


def nmi(a, b):
    a = np.asarray(a); b = np.asarray(b)
    n = len(a)
    ua, ia = np.unique(a, return_inverse=True)
    ub, ib = np.unique(b, return_inverse=True)
    c = np.zeros((len(ua), len(ub)))
    np.add.at(c, (ia, ib), 1)
    pa, pb = c.sum(1) / n, c.sum(0) / n
    p = c / n
    nz = p > 0
    mi = (p[nz] * np.log(p[nz] / (pa[:, None] * pb[None, :])[nz])).sum()
    ha = -(pa[pa > 0] * np.log(pa[pa > 0])).sum()
    hb = -(pb[pb > 0] * np.log(pb[pb > 0])).sum()
    return 2 * mi / (ha + hb) if ha + hb > 0 else 1.0

"""Parity tests proper: the CUDA replay path, called through the C ABI (libbisbm.so), against
the golden fixtures from the unmodified reference and against the oracle on fresh seeded
cases.  Integer state (labels, m_rs, e_r, n_r, eta, vlist, RNG word counts) bit-exact; the
accumulated dS bit-exact where log q comes from the exact table, 1e-12 relative where the
asymptotic branch runs on device libm; entropy() within 1e-9 relative (north-star)."""
import numpy as np
import pytest

from conftest import TRAJECTORIES, load_golden
from oracle import port

pytestmark = pytest.mark.gpu


def make_pool(host, g, extra_chains=0):
    graph = host.Graph(g["edges"], g["na"], g["nb"])
    lab = np.asarray(g["labels0"], dtype=np.uint32)
    if extra_chains:
        lab = np.tile(lab, (1 + extra_chains, 1))
    return graph, host.ChainPool(graph, lab, g["ka"], g["kb"], g["eps"])


@pytest.mark.parametrize("name", TRAJECTORIES)
def test_replay_matches_reference(host, name):
    g = load_golden(name)
    graph, pool = make_pool(host, g, extra_chains=2 if name.startswith("c1") else 0)
    chain = 1 if pool.n_chains > 1 else 0
    pool.replay_init(chain, g["seed"], g["gen_seed"], bool(g["randomize"]))
    assert (pool.labels(chain) == g["init_labels"]).all()
    assert (pool.m(chain) == g["init_m"]).all() and (pool.m_r(chain) == g["init_m_r"]).all()
    assert (pool.n_r(chain) == g["init_n_r"]).all() and (pool.eta(chain) == g["init_eta"]).all()
    assert abs(pool.entropy(chain) - g["init_entropy"]) <= 1e-9 * abs(g["init_entropy"])
    exact = name != "big_blocks"
    for v, s, dS, ar in list(zip(g["kat_v"], g["kat_s"], g["kat_dS"], g["kat_accu"]))[:60]:
        d, a = pool.replay_transition(chain, int(v), int(s))
        if np.isinf(dS):
            assert np.isinf(d)
            continue
        if exact:
            assert d == dS and a == ar
        else:
            assert abs(d - dS) <= 1e-12 * max(1.0, abs(dS)) and a == ar
    # the KAT calls must not have disturbed the chain: re-init and run the trajectory
    graph2, pool = make_pool(host, g, extra_chains=2 if name.startswith("c1") else 0)
    pool.replay_init(chain, g["seed"], g["gen_seed"], bool(g["randomize"]))
    acc, sweeps = pool.replay_anneal(chain, int(g["schedule"]), float(g["p0"]), float(g["p1"]), int(g["duration"]),
                                     int(g["steps_await"]))
    assert (pool.labels(chain) == g["labels"]).all()
    assert (pool.m(chain) == g["m"]).all() and (pool.m_r(chain) == g["m_r"]).all()
    assert (pool.n_r(chain) == g["n_r"]).all() and (pool.eta(chain) == g["eta"]).all()
    assert (pool.replay_vlist(chain) == g["vlist"]).all()
    assert pool.replay_rng_words(chain) == tuple(int(x) for x in g["rng_words"])
    assert acc == g["accept"]
    if exact:
        assert pool.entropy_accum(chain) == g["entropy_accum"]
    else:
        assert abs(pool.entropy_accum(chain) - g["entropy_accum"]) <= 1e-12 * abs(g["entropy_accum"]) + 1e-10
    assert abs(pool.entropy(chain) - g["entropy"]) <= 1e-9 * abs(g["entropy"])
    if pool.n_chains > 1:  # the other chains were not touched
        assert (pool.labels(0) == g["labels0"]).all()


def test_reference_shaped_interface(pkg):
    """The same run written against the Python mirror of the reference's classes."""
    g = load_golden("c1_seed1")
    N = g["na"] + g["nb"]
    adj = pkg.edge_to_adj(g["edges"], N, g["na"], g["nb"])
    types = [0] * g["na"] + [1] * g["nb"]
    bm = pkg.blockmodel_t(g["labels0"], types, g["ka"] + g["kb"], g["ka"], g["kb"], g["eps"], adj, gen_seed=g["gen_seed"])
    engine = pkg.mt19937(g["seed"])
    bm.shuffle_bisbm(engine, g["na"], g["nb"])
    mh = pkg.metropolis_hasting()
    rate = mh.anneal(bm, pkg.exponential_schedule, [10, 0.1], 1000, 100, engine)
    assert rate == g["accept"]
    assert (bm.get_memberships() == g["labels"]).all()
    assert (bm.get_m() == g["m"]).all() and (bm.get_m_r() == g["m_r"]).all() and (bm.get_n_r() == g["n_r"]).all()
    assert bm.get_entropy() == g["entropy_accum"]
    assert abs(bm.entropy() - g["entropy"]) <= 1e-9 * g["entropy"]


def test_replay_matches_oracle_on_fresh_cases(host):
    rng = np.random.default_rng(1)
    for case in range(6):
        na, nb = int(rng.integers(3, 60)), int(rng.integers(3, 60))
        ka, kb = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        ka, kb = min(ka, na), min(kb, nb)
        ne = int(rng.integers(5, 500))
        edges = np.stack([rng.integers(0, na, ne), na + rng.integers(0, nb, ne)], 1).astype(np.uint32)
        labels = np.concatenate([np.arange(na) % ka, ka + np.arange(nb) % kb]).astype(np.uint32)
        sched = int(rng.integers(0, 5))
        p0, p1 = [(5, 0.99), (2.0, 0.001), (1.0, 2), (1.3, 0), (150, 0)][sched]
        n = na + nb
        eps = float(rng.choice([1e-3, 0.5, 1.0, 10.0]))
        o = port.PortChain(n, na, nb, edges, labels, ka, kb, eps, 100 + case, 4242)
        o.init(True)
        acc_o = o.anneal(sched, p0, p1, 20 * n, 40)
        graph = host.Graph(edges, na, nb)
        pool = host.ChainPool(graph, labels, ka, kb, eps)
        pool.replay_init(0, 100 + case, 4242, True)
        acc, sw = pool.replay_anneal(0, sched, p0, p1, 20 * n, 40)
        assert acc == acc_o and sw == o.sweeps_done()
        assert (pool.labels(0) == o.labels()).all() and (pool.m(0) == o.m()).all()
        assert (pool.m_r(0) == o.m_r()).all() and (pool.n_r(0) == o.n_r()).all() and (pool.eta(0) == o.eta()).all()
        assert pool.entropy_accum(0) == o.entropy_accum()
        assert abs(pool.entropy(0) - o.entropy()) <= 1e-9 * abs(o.entropy())
        # a second anneal call continues both RNG streams and the vlist permutation
        acc_o2 = o.anneal(3, 1.0, 0, 3 * n, 10 ** 9)
        acc2, _ = pool.replay_anneal(0, 3, 1.0, 0, 3 * n, 10 ** 9)
        assert acc2 == acc_o2 and (pool.labels(0) == o.labels()).all()


def test_argument_errors(host):
    edges = np.array([[0, 2], [1, 3], [0, 3]], dtype=np.uint32)
    g = host.Graph(edges, 2, 2)
    with pytest.raises(host.BisbmError):
        host.ChainPool(g, np.array([0, 1, 1, 2], dtype=np.uint32), 2, 2, 1.0)  # type-b node in a type-a block
    with pytest.raises(host.BisbmError):
        host.ChainPool(g, np.array([0, 1, 2, 3], dtype=np.uint32), 2, 2, 0.0)  # epsilon must be > 0
    with pytest.raises(host.BisbmError):
        host.Graph(np.array([[0, 1]], dtype=np.uint32), 2, 2)  # edge inside one type
    pool = host.ChainPool(g, np.array([0, 1, 2, 3], dtype=np.uint32), 2, 2, 1.0)
    with pytest.raises(host.BisbmError):
        pool.replay_anneal(0, 3, 1.0, 0, 10, 10)  # replay_init not called


def test_replay_at_scale_matches_oracle(host):
    """A graph far beyond the shipped datasets (50k nodes / 500k edges, K = 8 + 8, block degree totals
    above the 10001 threshold so log q takes the asymptotic branch): two replay sweeps on the GPU vs
    the oracle.  Labels and all counts bit-exact; accumulated dS to 1e-11 relative (device libm in
    log_q_approx)."""
    from helpers import planted, planted_labels
    na = nb = 25000
    ka = kb = 8
    edges = planted(na, nb, ka, kb, 500000, 21)
    labels = planted_labels(na, nb, ka, kb)
    n = na + nb
    o = port.PortChain(n, na, nb, edges, labels, ka, kb, 1.0, 77, 78)
    o.init(True)
    acc_o = o.anneal("constant", 1.0, 0, 2 * n, 10 ** 18)
    graph = host.Graph(edges, na, nb)
    pool = host.ChainPool(graph, labels, ka, kb, 1.0)
    pool.replay_init(0, 77, 78, True)
    acc, sw = pool.replay_anneal(0, "constant", 1.0, 0, 2 * n, 10 ** 18)
    assert sw == 2
    assert (pool.labels(0) == o.labels()).all()
    assert (pool.m(0) == o.m()).all() and (pool.m_r(0) == o.m_r()).all() and (pool.n_r(0) == o.n_r()).all()
    assert (pool.eta(0) == o.eta()).all() and (pool.replay_vlist(0) == o.vlist()).all()
    assert acc == acc_o
    assert pool.replay_rng_words(0) == o.rng_words()
    assert abs(pool.entropy_accum(0) - o.entropy_accum()) <= 1e-11 * abs(o.entropy_accum())
    assert abs(pool.entropy(0) - o.entropy()) <= 1e-9 * abs(o.entropy())

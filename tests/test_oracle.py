"""The oracle (oracle/bisbm_oracle.c, the plain-C restatement) against the golden fixtures
generated from the unmodified reference build, and -- where oracle/_ref is present -- against
that build directly.  Bit-exact everywhere (integer state AND doubles)."""
import numpy as np
import pytest

from conftest import TRAJECTORIES, load_golden
from oracle import port, ref


def run_port(g):
    n = g["na"] + g["nb"]
    c = port.PortChain(n, g["na"], g["nb"], g["edges"], g["labels0"], g["ka"], g["kb"], g["eps"], g["seed"], g["gen_seed"])
    c.init(bool(g["randomize"]))
    return c


@pytest.mark.parametrize("name", TRAJECTORIES)
def test_port_matches_reference_trajectory(name):
    g = load_golden(name)
    c = run_port(g)
    assert (c.labels() == g["init_labels"]).all()
    assert c.entropy() == g["init_entropy"]
    assert (c.m() == g["init_m"]).all() and (c.m_r() == g["init_m_r"]).all()
    assert (c.n_r() == g["init_n_r"]).all() and (c.eta() == g["init_eta"]).all()
    for v, s, dS, ar in zip(g["kat_v"], g["kat_s"], g["kat_dS"], g["kat_accu"]):
        d, a = c.transition(int(v), int(s))
        assert d == dS or (np.isinf(d) and np.isinf(dS))
        if not np.isinf(dS):
            assert a == ar
    acc = c.anneal(int(g["schedule"]), float(g["p0"]), float(g["p1"]), int(g["duration"]), int(g["steps_await"]))
    assert acc == g["accept"]
    assert (c.labels() == g["labels"]).all()
    assert (c.m() == g["m"]).all() and (c.m_r() == g["m_r"]).all() and (c.n_r() == g["n_r"]).all()
    assert (c.eta() == g["eta"]).all() and (c.vlist() == g["vlist"]).all()
    assert c.entropy_accum() == g["entropy_accum"]
    assert c.entropy() == g["entropy"]
    assert tuple(c.rng_words()) == tuple(int(x) for x in g["rng_words"])


def test_survey_known_answers():
    """The RNG-free values quoted in SURVEY.md 8(c)."""
    g = load_golden("c1_seed1")
    c = port.PortChain(32, 18, 14, g["edges"], g["labels0"], 5, 5, 1e-3, 1)
    c.init(False)
    assert c.entropy() == 229.73989707660991
    assert list(c.m_r()) == [30, 15, 18, 20, 6, 12, 20, 36, 15, 6]
    assert c.transition(0, 1) == (1.2974035277270204, 1.7724979575577744)
    assert c.transition(31, 5) == (5.077254981437644, 1643.869891129752)
    dS, _ = c.transition(0, 7)
    assert np.isinf(dS)


def test_math_known_answers():
    m = load_golden("math")
    tab = port.log_q_table(10000, 10000)
    for n, k, v in zip(m["q_n"], m["q_k"], m["q_v"]):
        kk = min(int(k), int(n))
        assert tab[int(n), kk] == v
    for n, k, v in zip(m["a_n"], m["a_k"], m["a_v"]):
        assert port.log_q_approx(int(n), int(k)) == v
    for x, v in zip(m["sp_x"], m["sp_v"]):
        assert port.spence(float(x)) == v
    for sid, p0, p1, t, v in m["sched"]:
        got = port.schedule(int(sid), float(p0), float(p1), int(t))
        assert got == v or (np.isnan(got) and np.isnan(v)) or (np.isinf(got) and np.isinf(v))


def test_rng_primitives_against_numpy_mt19937():
    """Raw MT19937 words agree with numpy's implementation of the same generator."""
    for seed in (0, 1, 12345, 4294967295):
        r = port.Rng(seed)
        bg = np.random.MT19937()
        bg._legacy_seeding(seed)
        want = bg.random_raw(2000)
        got = np.array([r.word() for _ in range(2000)], dtype=np.uint64)
        assert (got == want).all()


def test_shuffle_is_permutation_and_deterministic():
    for n in (0, 1, 2, 3, 32, 33, 1000, 65535, 65536, 70001):
        a = port.Rng(5).shuffle(np.arange(n, dtype=np.uint32))
        b = port.Rng(5).shuffle(np.arange(n, dtype=np.uint32))
        assert (a == b).all()
        assert (np.sort(a) == np.arange(n)).all()


@pytest.mark.skipif(not ref.available(), reason="reference build (oracle/_ref) not present")
def test_port_matches_reference_build_on_random_cases():
    """Fresh seeded cases beyond the committed fixtures, port vs the reference .so."""
    rng = np.random.default_rng(0)
    for case in range(4):
        na, nb = int(rng.integers(5, 40)), int(rng.integers(5, 40))
        ka, kb = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        ne = int(rng.integers(10, 300))
        edges = np.stack([rng.integers(0, na, ne), na + rng.integers(0, nb, ne)], 1).astype(np.uint32)
        labels = np.concatenate([np.arange(na) % ka, ka + np.arange(nb) % kb]).astype(np.uint32)
        sched = int(rng.integers(0, 5))
        p0, p1 = [(5, 0.99), (2.0, 0.001), (1.0, 2), (1.3, 0), (150, 0)][sched]
        n = na + nb
        a = ref.RefChain(n, na, nb, edges, labels, ka, kb, 0.5, case + 1, 777)
        b = port.PortChain(n, na, nb, edges, labels, ka, kb, 0.5, case + 1, 777)
        a.init(True); b.init(True)
        assert a.anneal(sched, p0, p1, 12 * n, 50) == b.anneal(sched, p0, p1, 12 * n, 50)
        assert (a.labels() == b.labels()).all() and (a.m() == b.m()).all() and (a.eta() == b.eta()).all()
        assert a.entropy_accum() == b.entropy_accum() and a.entropy() == b.entropy()


def test_merge_fixture_is_self_consistent():
    """tests/golden/merge.npz (reference outputs of agg_merge and of the merge paths of main): every recorded state is a
    valid bipartite partition with the recorded block counts, and its m_rs / e_r / n_r equal a rebuild from its labels."""
    from helpers import counts_from_labels
    g = load_golden("merge")
    graphs = {"sw": "c1_seed1", "swg": "c1_seed1", "swu": "c1_seed1", "b1000": "c2_const_k46"}
    tags = sorted(k[:-len("_labels")] for k in g if k.endswith("_labels") and (k[:-len("_labels")] + "_k") in g)
    assert len(tags) >= 10
    for tag in tags:
        gr = load_golden(graphs[tag.split("_")[0]])
        na, nb = gr["na"], gr["nb"]
        ka, kb = (int(x) for x in g[tag + "_k"])
        lab = g[tag + "_labels"]
        assert set(lab[:na]) == set(range(ka)) and set(lab[na:]) == set(range(ka, ka + kb)), tag
        m, e, nr, eta = counts_from_labels(gr["edges"], na, nb, lab, ka, kb)
        assert (g[tag + "_m"] == m).all() and (g[tag + "_m_r"] == e).all() and (g[tag + "_n_r"] == nr).all(), tag


@pytest.mark.skipif(not ref.available(), reason="reference build (oracle/_ref) not present")
def test_merge_fixture_matches_reference_build():
    """The fixture is what the reference build produces today (first two agg_merge calls of the southernWomen case)."""
    g = load_golden("merge")
    gr = load_golden("c1_seed1")
    na, nb = gr["na"], gr["nb"]
    n = na + nb
    ch = ref.RefChain(n, na, nb, gr["edges"], g["sw_labels0"], na, nb, 1.0, 7, log_rng=True)
    ch.init(False)
    ch.agg_merge(6, 4, 10)
    assert (ch.labels() == g["sw_s1_labels"]).all() and tuple(ch.rng_words()) == tuple(int(x) for x in g["sw_s1_words"])
    ch.agg_merge(5, 0, 10)
    assert (ch.labels() == g["sw_s2_labels"]).all() and ch.entropy() == g["sw_s2_entropy"]


def test_parity_fixture_quench_chains_come_from_the_oracle():
    """tests/golden/parity_mid.npz is what tests/golden/make_parity_fixture.py writes: chain 0 of the quench protocol
    (randomised start, abrupt_cool, 10 hot + 10 greedy sweeps on the 40k + 40k graph), reference visiting order and the
    type-alternating test aid, recomputed here with the oracle must give the stored numbers bit for bit."""
    import hashlib
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    fx = load_golden("parity_mid")
    if "qu_entropy" not in fx:
        pytest.skip("fixture without the quench protocol")
    import make_parity_fixture as mk
    edges, _ = mk._graph()
    assert hashlib.sha1(np.ascontiguousarray(edges).tobytes()).hexdigest() == str(fx["edges_sha1"])
    assert int(fx["sweeps_qu"]) == mk.SWEEPS_QU and int(fx["hot_qu"]) == mk.HOT_QU
    for proto in ("qu", "qu_alt"):
        ent, acc, nm = mk.run_chain((proto, 0))
        assert ent == fx["%s_entropy" % proto][0] and acc == fx["%s_accept" % proto][0] and nm == fx["%s_nmi" % proto][0]

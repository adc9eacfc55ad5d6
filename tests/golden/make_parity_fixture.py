"""Oracle side of the statistical-parity tests at the BENCHMARKED operating point (sliced multi-CTA plan).

Runs R independent chains of the oracle (oracle/liboracle.so, the plain-C restatement pinned bit for bit to
the reference build, tests/test_oracle.py) on a mid-size planted graph where the CPU is still feasible and the
GPU already takes the sliced plan (several CTAs per chain group), and stores the per-chain observables as
tests/golden/parity_mid.npz.  The GPU tests (tests/test_parity_operating_point.py) regenerate the graph from
the same seed, check its checksum, and compare their own chains against these samples with two-sample KS tests.
Also stores transition_ratio known answers on that graph (large blocks: log q takes the asymptotic branch,
src/support/int_part.cc:73-98, which is where the parallel kernels use their per-block expansions).

Run in the authoring container only (about 10 CPU-minutes on 8 cores):
    python tests/golden/make_parity_fixture.py
"""
import hashlib
import os
import sys
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import nmi, planted, planted_labels  # noqa: E402
from oracle import port  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# the graph of scripts/staleness_study.py
NA = NB = 40000
KA = KB = 8
NE = 600000
GSEED = 5
R = 128
SWEEPS_EQ = 30      # protocol "eq": planted start, T = 1: samples of the stationary distribution
SWEEPS_TR = 40      # protocol "tr": randomised start, T = 1: the burn-in transient of the staleness study

HOT_QU, SWEEPS_QU = 10, 20   # protocol "qu": randomised start, abrupt_cool -- 10 sweeps at T = 1, then 10 greedy sweeps (T = 0)

_G = {}


def _graph():
    if not _G:
        _G["edges"] = planted(NA, NB, KA, KB, NE, GSEED)
        _G["lab"] = planted_labels(NA, NB, KA, KB)
    return _G["edges"], _G["lab"]


def run_chain(args):
    proto, s = args
    edges, lab = _graph()
    n = NA + NB
    o = port.PortChain(n, NA, NB, edges, lab, KA, KB, 1.0, 7000 + s, 9000 + s)
    o.init(proto.startswith("tr") or proto.startswith("qu"))
    if proto.startswith("qu"):      # the quench of the reference's default schedule (src/metropolis_hasting.cc:33-37)
        acc = o.anneal("abrupt_cool", float(HOT_QU * n), 0.0, SWEEPS_QU * n, 10 ** 18, alternate=proto.endswith("_alt"))
        out = (o.entropy(), acc, nmi(o.labels(), lab))
        o.close()
        return out
    sweeps = SWEEPS_TR if proto.startswith("tr") else SWEEPS_EQ
    # "tr_alt": the same burn-in with the type-alternating visiting order of the GPU's parallel mode (oracle test aid)
    acc = o.anneal("constant", 1.0, 0.0, sweeps * n, 10 ** 18, alternate=proto.endswith("_alt"))
    out = (o.entropy(), acc, nmi(o.labels(), lab))
    o.close()
    return out


def run_small(s):
    """bisbm-1000 at (4,6): randomised start, abrupt_cool T0 = 1e5 steps then greedy, 200 sweeps
    (the protocol of tests/test_parallel_gpu.py::test_statistical_parity_with_oracle_nmi_and_entropy)."""
    z = np.load(os.path.join(OUT, "c2_const_k46.npz"))
    na, nb, edges, mb = int(z["na"]), int(z["nb"]), z["edges"], z["labels0"]
    n = na + nb
    o = port.PortChain(n, na, nb, edges, mb, 4, 6, 1.0, 1000 + s, 2000 + s)
    o.init(True)
    acc = o.anneal("abrupt_cool", 1e5, 0, 200 * n, 10 ** 9)
    out = (o.entropy(), acc, nmi(o.labels(), mb))
    o.close()
    return out


def main():
    if "--only-alt" in sys.argv:      # add the tr_alt_* arrays to an existing fixture
        z = dict(np.load(os.path.join(OUT, "parity_mid.npz")))
        with Pool(min(8, os.cpu_count() or 1)) as pool:
            res = np.array(pool.map(run_chain, [("tr_alt", s) for s in range(R)], chunksize=1))
        z["tr_alt_entropy"], z["tr_alt_accept"], z["tr_alt_nmi"] = res[:, 0], res[:, 1], res[:, 2]
        print("tr_alt entropy %.1f +- %.1f  accept %.4f +- %.4f  nmi %.4f" % (res[:, 0].mean(), res[:, 0].std(), res[:, 1].mean(), res[:, 1].std(), res[:, 2].mean()))
        np.savez_compressed(os.path.join(OUT, "parity_mid.npz"), **z)
        return
    if "--only-quench" in sys.argv:   # add the qu_* / qu_alt_* arrays to an existing fixture
        z = dict(np.load(os.path.join(OUT, "parity_mid.npz")))
        with Pool(min(8, os.cpu_count() or 1)) as pool:
            for proto in ("qu", "qu_alt"):
                res = np.array(pool.map(run_chain, [(proto, s) for s in range(R)], chunksize=1))
                z["%s_entropy" % proto], z["%s_accept" % proto], z["%s_nmi" % proto] = res[:, 0], res[:, 1], res[:, 2]
                print(proto, "entropy %.1f +- %.1f  accept %.4f +- %.4f  nmi %.4f" % (res[:, 0].mean(), res[:, 0].std(), res[:, 1].mean(), res[:, 1].std(), res[:, 2].mean()), flush=True)
        z["sweeps_qu"], z["hot_qu"] = SWEEPS_QU, HOT_QU
        np.savez_compressed(os.path.join(OUT, "parity_mid.npz"), **z)
        return
    edges, lab = _graph()
    n = NA + NB
    out = dict(na=NA, nb=NB, ka=KA, kb=KB, n_edges=NE, graph_seed=GSEED, R=R, sweeps_eq=SWEEPS_EQ, sweeps_tr=SWEEPS_TR, sweeps_qu=SWEEPS_QU, hot_qu=HOT_QU,
               edges_sha1=hashlib.sha1(np.ascontiguousarray(edges).tobytes()).hexdigest())
    # transition_ratio known answers on the planted state (RNG-free)
    o = port.PortChain(n, NA, NB, edges, lab, KA, KB, 1.0, 1, 2)
    o.init(False)
    kv, ks, kd, ka_ = [], [], [], []
    for v in list(range(0, NA, 1999)) + list(range(NA, n, 1999)):
        for s in range(KA + KB):
            dS, ar = o.transition(v, s)
            if np.isfinite(dS) and s != lab[v]:
                kv.append(v); ks.append(s); kd.append(dS); ka_.append(ar)
    out.update(kat_v=np.array(kv, dtype=np.uint32), kat_s=np.array(ks, dtype=np.uint32),
               kat_dS=np.array(kd), kat_accu=np.array(ka_), init_entropy=o.entropy())
    o.close()
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        for proto in ("eq", "tr", "tr_alt", "qu", "qu_alt"):
            res = np.array(pool.map(run_chain, [(proto, s) for s in range(R)], chunksize=1))
            out["%s_entropy" % proto], out["%s_accept" % proto], out["%s_nmi" % proto] = res[:, 0], res[:, 1], res[:, 2]
            print(proto, "entropy %.1f +- %.1f  accept %.4f +- %.4f  nmi %.4f" % (
                res[:, 0].mean(), res[:, 0].std(), res[:, 1].mean(), res[:, 1].std(), res[:, 2].mean()), flush=True)
        res = np.array(pool.map(run_small, range(R), chunksize=4))
        out["q46_entropy"], out["q46_accept"], out["q46_nmi"] = res[:, 0], res[:, 1], res[:, 2]
        print("q46 entropy %.1f +- %.1f  accept %.4f  nmi %.4f" % (res[:, 0].mean(), res[:, 0].std(), res[:, 1].mean(),
                                                                res[:, 2].mean()), flush=True)
    np.savez_compressed(os.path.join(OUT, "parity_mid.npz"), **out)
    print("parity_mid.npz written")


if __name__ == "__main__":
    main()

"""Host compilation of the DEVICE headers (csrc/replay.cuh, sweep.cuh, devmath.cuh) checked
against the golden fixtures.  This is a debugging aid for a container without a GPU: it
proves the text of the replay routine and of the parallel kernel's per-move arithmetic is
right before GPU time is spent.  It is not a product path: libbisbm.so has no host
execution route (see test_capi_cpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, TRAJECTORIES, load_golden
from oracle import port

EMUL_DIR = os.path.join(ROOT, "tests", "emul")
LIB = os.path.join(EMUL_DIR, "libemul.so")
pu = C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(EMUL_DIR, "emul.cc")
    deps = [src] + [os.path.join(ROOT, "bipartitesbm-mcmc_b200", "csrc", f) for f in
                    ("devmath.cuh", "state.cuh", "replay.cuh", "sweep.cuh", "sweep2.cuh", "gltable.h")]
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                               "-o", LIB, src, "-lm"])
    L = C.CDLL(LIB)
    L.emul_create.restype = C.c_void_p
    L.emul_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, pu, pu, pu, C.c_uint32, C.c_uint32, C.c_double,
                              C.c_uint32, C.c_uint32, C.c_int]
    L.emul_destroy.argtypes = [C.c_void_p]
    L.emul_anneal.restype = C.c_double
    L.emul_anneal.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_uint64, C.c_uint64]
    L.emul_entropy_accum.restype = C.c_double
    L.emul_entropy_accum.argtypes = [C.c_void_p]
    L.emul_get_labels.argtypes = [C.c_void_p, pu]
    L.emul_get_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emul_words.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.emul_transition.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.emul_par_dS.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.emul_par2_dS.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_double),
                               C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.emul_logq_expansion.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]
    for fn in (L.emul_dlog, L.emul_dexp):
        fn.restype = C.c_double
        fn.argtypes = [C.c_double]
    L.emul_lgamma_diff.restype = C.c_double
    L.emul_lgamma_diff.argtypes = [C.c_double, C.c_double]
    L.emul_block_degree_delta.restype = C.c_double
    L.emul_block_degree_delta.argtypes = [C.c_int, C.c_int, C.c_int]
    L.emul_log_q_approx.restype = C.c_double
    L.emul_log_q_approx.argtypes = [C.c_uint64, C.c_uint64]
    L.emul_log_q_approx_tab.restype = C.c_double
    L.emul_log_q_approx_tab.argtypes = [C.c_uint64, C.c_uint64]
    L.emul_gl_table_stats.argtypes = [C.c_void_p] * 4
    L.emul_feistel.restype = C.c_uint32
    L.emul_feistel.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    return L


def _create(L, g, randomize=None):
    ea = np.ascontiguousarray(g["edges"][:, 0], dtype=np.uint32)
    eb = np.ascontiguousarray(g["edges"][:, 1], dtype=np.uint32)
    lab = np.ascontiguousarray(g["labels0"], dtype=np.uint32)
    rnd = int(g["randomize"]) if randomize is None else int(randomize)
    return L.emul_create(g["na"], g["nb"], len(ea), ea.ctypes.data_as(pu), eb.ctypes.data_as(pu), lab.ctypes.data_as(pu),
                         g["ka"], g["kb"], g["eps"], g["seed"], g["gen_seed"], rnd)


@pytest.mark.parametrize("name", TRAJECTORIES)
def test_replay_text_is_bit_exact_on_host(emul, name):
    g = load_golden(name)
    n = g["na"] + g["nb"]
    h = _create(emul, g)
    K = g["ka"] + g["kb"]
    for v, s, dS, ar in zip(g["kat_v"], g["kat_s"], g["kat_dS"], g["kat_accu"]):
        d, a = C.c_double(), C.c_double()
        emul.emul_transition(h, int(v), int(s), C.byref(d), C.byref(a))
        assert d.value == dS or (np.isinf(d.value) and np.isinf(dS))
        if not np.isinf(dS):
            assert a.value == ar
    sched = int(g["schedule"])
    temps = None
    if sched in (0, 2):
        sweeps = int(g["duration"]) // n
        t = np.array([port.schedule(sched, float(g["p0"]), float(g["p1"]), i) for i in range(sweeps * n)])
        temps = t.ctypes.data_as(C.c_void_p)
    acc = emul.emul_anneal(h, sched, float(g["p0"]), float(g["p1"]), temps, int(g["duration"]), int(g["steps_await"]))
    assert acc == g["accept"]
    out = np.empty(n, dtype=np.uint32)
    emul.emul_get_labels(h, out.ctypes.data_as(pu))
    assert (out == g["labels"]).all()
    assert emul.emul_entropy_accum(h) == g["entropy_accum"]
    m = np.zeros((g["ka"], g["kb"]), dtype=np.int32)
    e = np.zeros(K, dtype=np.int32)
    nr = np.zeros(K, dtype=np.int32)
    emul.emul_get_counts(h, m.ctypes.data, e.ctypes.data, nr.ctypes.data)
    assert (m == g["m"][:g["ka"], g["ka"]:]).all() and (e == g["m_r"]).all() and (nr == g["n_r"]).all()
    a, b = C.c_uint64(), C.c_uint64()
    emul.emul_words(h, C.byref(a), C.byref(b))
    assert (a.value, b.value) == tuple(int(x) for x in g["rng_words"])
    emul.emul_destroy(h)


@pytest.mark.parametrize("name,taylor,tol", [("c1_seed1", 0, 1e-9), ("c2_abrupt", 0, 1e-9), ("c2_const_k46", 0, 1e-9),
                                             ("big_blocks", 0, 1e-9), ("big_blocks", 1, 2e-6), ("isolated", 0, 1e-9)])
def test_parallel_move_arithmetic_matches_reference(emul, name, taylor, tol):
    """dS / accu_r as the parallel kernel computes them (log-products, Stirling differences,
    log q expansion) vs the reference's transition_ratio known answers.  Tolerance: 1e-9
    relative (north-star) on the exact paths; 2e-6 ABSOLUTE for the second-order log q
    expansion evaluated 2.5% away from its expansion point."""
    g = load_golden(name)
    h = _create(emul, g)
    checked = 0
    for v, s, dS, ar in zip(g["kat_v"], g["kat_s"], g["kat_dS"], g["kat_accu"]):
        r = g["init_labels"][v]
        if np.isinf(dS) or r == s:
            continue
        d, a = C.c_double(), C.c_double()
        emul.emul_par_dS(h, int(v), int(s), taylor, C.byref(d), C.byref(a))
        if taylor:
            assert abs(d.value - dS) < tol
        else:
            assert abs(d.value - dS) <= tol * max(1.0, abs(dS))
        assert abs(a.value - ar) <= 1e-12 * abs(ar)
        checked += 1
    assert checked > 0
    emul.emul_destroy(h)


@pytest.mark.parametrize("name,taylor", [("c1_seed1", 0), ("c2_abrupt", 0), ("c2_const_k46", 0), ("big_blocks", 0),
                                         ("big_blocks", 1), ("isolated", 0)])
@pytest.mark.parametrize("fp32", [0, 1])
def test_staged_kernel_move_arithmetic_matches_reference(emul, name, taylor, fp32):
    """The staged sweep kernel's per-move arithmetic (sweep2.cuh, R = double / float) vs the reference's transition_ratio
    known answers.  double: 1e-9 relative on dS (north-star), 1e-12 relative on accu_r (2e-6 absolute on dS where the
    second-order log q expansion is evaluated 2.5% away from its expansion point).  float: what the accept test consumes is
    a = log(accu_r) - dS / T, so the tolerance is ABSOLUTE, 2e-5 * max(1, |dS| / 8) -- an acceptance probability off by a
    factor exp(2e-5) at worst -- with the host's log2f / exp2f / division standing in for lg2/ex2/rcp.approx."""
    g = load_golden(name)
    h = _create(emul, g)
    checked = 0
    worst = 0.0
    for v, s, dS, ar in zip(g["kat_v"], g["kat_s"], g["kat_dS"], g["kat_accu"]):
        r = g["init_labels"][v]
        if np.isinf(dS) or r == s:
            continue
        d, a, fb = C.c_double(), C.c_double(), C.c_int()
        emul.emul_par2_dS(h, int(v), int(s), taylor, fp32, C.byref(d), C.byref(a), C.byref(fb))
        if fp32:
            tol = 2e-5 * max(1.0, abs(dS) / 8.0) + (2e-6 if taylor else 0.0)
            assert abs(d.value - dS) <= tol, (v, s, d.value, dS)
            assert abs(a.value - np.log(ar)) <= 2e-5 * max(1.0, abs(np.log(ar))), (v, s, a.value, np.log(ar))
        else:
            tol = 2e-6 if taylor else 1e-9 * max(1.0, abs(dS))
            assert abs(d.value - dS) <= tol, (v, s, d.value, dS)
            assert abs(a.value - np.log(ar)) <= 1e-12 * max(1.0, abs(np.log(ar))), (v, s, a.value, np.log(ar))
        worst = max(worst, abs(d.value - dS))
        checked += 1
    print(name, "fp32" if fp32 else "fp64", "dS worst abs error", worst, "over", checked)
    assert checked > 0
    emul.emul_destroy(h)


def test_branch_free_log_and_exp(emul):
    """dlog / dexp of sweep2.cuh (the double kernel's logarithm and exponential) against libm: < 1.5 ulp."""
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(0.5, 2.0, 20000), 1.0 + rng.uniform(-1e-6, 1e-6, 2000),
                         np.array([1.0, 2.0, 0.5, np.sqrt(2.0), np.nextafter(np.sqrt(2.0), 2), 1e-300, 1e300, 3.0, 7.6e4])])
    for x in xs:
        got, want = emul.emul_dlog(float(x)), np.log(x)
        assert abs(got - want) <= 3.4e-16 * max(abs(want), 1e-300) + 1e-323, (x, got, want)
    ys = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-1, 1, 20000), np.array([0.0, -1e-17, 1e-17, -700.0, 700.0, -745.0])])
    for y in ys:
        got, want = emul.emul_dexp(float(y)), np.exp(max(min(y, 700.0), -700.0))
        assert abs(got - want) <= 3.4e-16 * want, (y, got, want)


def test_logq_expansion_accuracy(emul):
    """The per-block expansion of log q (sweep.cuh logq_expand / sweep2.cuh logq_fast) against differences of the
    reference's asymptotic formula: at the expansion point (what bisbm_parallel_transition sees right after a refresh)
    1e-9 on the term; drifted by 1% of the block (typical within one half sweep) 1e-6; at the edge of the validity range
    (6% drift) 1e-4 (small blocks; large ones are 10-100x better, see the printed worst cases)."""
    worst = {}
    for e0, n0 in [(75000, 5000), (312500, 15625), (20000, 1100), (5 * 10 ** 6, 2 * 10 ** 5), (40000, 20000)]:
        for frac, tol in [(0.0, 1e-9), (0.01, 1e-6), (1.0 / 16.0, 1e-4)]:
            for sx in (-1, 1):
                for sy in (-1, 0, 1):
                    x, y = int(sx * frac * e0), int(sy * frac * n0 * 0.999)
                    for de, dn in [(-1, -1), (20, 1), (-20, -1), (60, 1), (-200, -1), (255, 1)]:
                        a, ex, ok = C.c_double(), C.c_double(), C.c_int()
                        emul.emul_logq_expansion(e0, n0, x, y, de, dn, C.byref(a), C.byref(ex), C.byref(ok))
                        if not ok.value:
                            continue
                        err = abs(a.value - ex.value)
                        scale = max(1.0, abs(de) / 20.0)
                        worst[frac] = max(worst.get(frac, 0.0), err / scale)
                        assert err <= tol * scale, (e0, n0, x, y, de, dn, a.value, ex.value)
    print("log q expansion: worst |error| (scaled to degree 20) by drift:", worst)


def test_lgamma_diff(emul):
    from math import lgamma
    for x in [1, 2, 5, 31, 32, 33, 100, 12345, 10 ** 6, 3 * 10 ** 8]:
        for d in [0, 1, 2, 7, 20, 57, 500]:
            want = lgamma(x + d) - lgamma(x)
            got = emul.emul_lgamma_diff(float(x), float(d))
            # the direct difference itself loses ~1e-16 * lgamma(x) to cancellation
            assert abs(got - want) <= 1e-11 * max(1.0, abs(want)) + 4e-16 * abs(lgamma(x + d))


def test_block_degree_delta(emul):
    """Euler-Maclaurin midpoint form of the e_r terms vs lgamma, across the switch at c = 32 d."""
    from math import lgamma
    for e_r in [60, 100, 640, 641, 2000, 10 ** 5, 312500, 2 * 10 ** 9 - 100]:
        for e_s in [0, 5, 639, 640, 5000, 312500, 2 * 10 ** 9 - 100]:
            for d in [0, 1, 2, 7, 20, 57]:
                if d > e_r or e_s + d > 2 ** 31 - 2:
                    continue
                want = (lgamma(e_s + d + 1) - lgamma(e_s + 1)) - (lgamma(e_r + 1) - lgamma(e_r - d + 1))
                got = emul.emul_block_degree_delta(e_r, e_s, d)
                slack = 4e-16 * (abs(lgamma(e_s + d + 1)) + abs(lgamma(e_r + 1)))  # cancellation in `want` itself
                assert abs(got - want) <= 2e-10 * max(1.0, abs(want)) + slack, (e_r, e_s, d, got, want)


def test_log_q_approx_device_text(emul):
    m = load_golden("math")
    for n, k, v in zip(m["a_n"], m["a_k"], m["a_v"]):
        assert emul.emul_log_q_approx(int(n), int(k)) == v


def test_feistel_is_a_permutation(emul):
    for n in (1, 2, 3, 18, 500, 4097, 100000):
        for key in (0, 1, 0xDEADBEEFCAFE):
            p = np.array([emul.emul_feistel(i, n, key) for i in range(n)])
            assert (np.sort(p) == np.arange(n)).all()
    a = np.array([emul.emul_feistel(i, 1000, 1) for i in range(1000)])
    b = np.array([emul.emul_feistel(i, 1000, 2) for i in range(1000)])
    assert (a != b).mean() > 0.9


def test_tabulated_asymptotic_log_q(emul):
    """Blocks without a valid expansion (small or drifted) evaluate the asymptotic log q from the tabulated g(u), lf(u) of
    u = k / sqrt(n) (devmath.cuh log_q_approx_tab: 6-point Lagrange in log u inside one run of equal iteration count of the
    reference's get_v) instead of iterating on the device.  Against the formula itself (the reference's int_part.cc:73-98
    restated, which the oracle pins): 1e-13 relative on the value and 1e-10 absolute on the DIFFERENCES a move needs,
    log q(e +- d, n +- 1) - log q(e, n), over block shapes from 10^4 to 4 * 10^9 edges ends."""
    import time
    rng = np.random.default_rng(5)
    t0 = time.perf_counter()
    emul.emul_log_q_approx_tab(20000, 100)          # builds the table (once per process in the library)
    assert time.perf_counter() - t0 < 5.0
    st = (C.c_uint32 * 4)()
    emul.emul_gl_table_stats(C.byref(st, 0), C.byref(st, 4), C.byref(st, 8), C.byref(st, 12))
    print("g / lf table: %d nodes, %d runs of equal iteration count (%d shorter than the stencil), %d flagged intervals" % tuple(st))
    assert st[1] < 100 and st[3] < 20
    worst_v, worst_d = 0.0, 0.0
    for _ in range(6000):
        e = int(10 ** rng.uniform(4.01, 9.6))
        # a block's degree sum is at most 255 n in the kernels that use this path (u8 histogram bins: degree <= 255)
        n = int(min(e, max(e // 255 + 1, 10 ** rng.uniform(0.3, np.log10(e)))))
        a, b = emul.emul_log_q_approx_tab(e, n), emul.emul_log_q_approx(e, n)
        assert abs(a - b) <= 1e-13 * max(1.0, abs(b)), (e, n, a, b)
        worst_v = max(worst_v, abs(a - b) / max(1.0, abs(b)))
        d = int(rng.integers(1, 256))
        if e - d < 10001 or n < 2:
            continue
        for de, dn in ((-d, -1), (d, 1)):
            da = emul.emul_log_q_approx_tab(e + de, n + dn) - a
            db = emul.emul_log_q_approx(e + de, n + dn) - b
            tol = 1e-10 if e < 10 ** 7 else 1e-9 * (e / 1e7) ** 0.5      # (sqrt(e) g(u): a few ulp of g times sqrt(e))
            assert abs(da - db) <= tol, (e, n, de, dn, da, db)
            worst_d = max(worst_d, abs(da - db))
    print("tabulated log q: worst relative error of the value %.2e, worst absolute error of a move's difference %.2e" % (worst_v, worst_d))

"""ctypes front-end of oracle/_ref/libref*.so -- the UNMODIFIED reference, compiled in place.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module.  The product (bipartitesbm-mcmc_b200/, bin/mcmc) never does.

The library is produced by `make -C oracle ref` from /root/reference/src (see
oracle/Makefile and oracle/ref_driver.cc); it is git-ignored and travels to the GPU box as
a prebuilt file.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

EXPONENTIAL, LINEAR, LOGARITHMIC, CONSTANT, ABRUPT_COOL = range(5)
SCHEDULES = {"exponential": 0, "linear": 1, "logarithmic": 2, "constant": 3, "abrupt_cool": 4}


def lib_path(log_rng=False):
    return os.path.join(_HERE, "_ref", "libref_log.so" if log_rng else "libref.so")


def available(log_rng=False):
    return os.path.exists(lib_path(log_rng))


_libs = {}


def _load(log_rng):
    if log_rng in _libs:
        return _libs[log_rng]
    L = C.CDLL(lib_path(log_rng))
    u64, u32, dbl, vp = C.c_uint64, C.c_uint32, C.c_double, C.c_void_p
    pu32 = C.POINTER(C.c_uint32)
    pi32 = C.POINTER(C.c_int32)
    L.ref_create.restype = vp
    L.ref_create.argtypes = [u64, u64, u64, u64, pu32, pu32, pu32, u64, u64, dbl, u32, u64]
    L.ref_destroy.argtypes = [vp]
    L.ref_init.argtypes = [vp, C.c_int]
    L.ref_anneal.restype = dbl
    L.ref_anneal.argtypes = [vp, C.c_int, C.c_float, C.c_float, u64, u64]
    L.ref_last_anneal_seconds.restype = dbl
    L.ref_last_anneal_seconds.argtypes = [vp]
    L.ref_schedule.restype = dbl
    L.ref_schedule.argtypes = [C.c_int, C.c_float, C.c_float, u64]
    L.ref_step.restype = C.c_int
    L.ref_step.argtypes = [vp, u64, dbl]
    L.ref_transition.argtypes = [vp, u64, u64, C.POINTER(dbl), C.POINTER(dbl)]
    for name in ("ref_entropy", "ref_entropy_accum", "ref_entropy_min"):
        getattr(L, name).restype = dbl
        getattr(L, name).argtypes = [vp]
    L.ref_get_labels.argtypes = [vp, pu32]
    L.ref_get_vlist.argtypes = [vp, pu32]
    L.ref_get_m.argtypes = [vp, pi32]
    L.ref_get_m_r.argtypes = [vp, pi32]
    L.ref_get_n_r.argtypes = [vp, pi32]
    L.ref_eta_width.restype = u64
    L.ref_eta_width.argtypes = [vp]
    L.ref_get_eta.argtypes = [vp, pu32]
    L.ref_get_k.argtypes = [vp, u64, pi32]
    L.ref_rng_words.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.ref_log_q.restype = dbl
    L.ref_log_q.argtypes = [C.c_int, C.c_int]
    L.ref_log_q_approx.restype = dbl
    L.ref_log_q_approx.argtypes = [u64, u64]
    L.ref_lgamma_fast.restype = dbl
    L.ref_lgamma_fast.argtypes = [u64]
    L.ref_safelog_fast.restype = dbl
    L.ref_safelog_fast.argtypes = [u64]
    L.ref_spence.restype = dbl
    L.ref_spence.argtypes = [dbl]
    L.ref_init_tables.argtypes = [u64]
    L.ref_merge_path.restype = dbl
    L.ref_merge_path.argtypes = [vp, u64, u64, C.c_float, u64, u64]
    L.ref_agg_merge.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    L.ref_agg_merge_total.argtypes = [vp, C.c_int, C.c_int]
    L.ref_nature_path.restype = dbl
    L.ref_nature_path.argtypes = [vp, C.c_float, u64, u64]
    L.ref_get_ka.restype = u64
    L.ref_get_ka.argtypes = [vp]
    L.ref_get_kb.restype = u64
    L.ref_get_kb.argtypes = [vp]
    _libs[log_rng] = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class RefChain:
    """One reference chain: blockmodel_t + metropolis_hasting + mt19937 engine."""

    def __init__(self, n, na, nb, edges, labels, ka, kb, eps, engine_seed, gen_seed=12345, log_rng=False):
        self.L = _load(log_rng)
        edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
        ea = np.ascontiguousarray(edges[:, 0])
        eb = np.ascontiguousarray(edges[:, 1])
        labels = np.ascontiguousarray(labels, dtype=np.uint32)
        assert labels.size == n == na + nb
        self.n, self.na, self.nb, self.ka, self.kb = n, na, nb, ka, kb
        self.K = ka + kb
        self.h = self.L.ref_create(n, na, nb, len(ea), _p(ea, C.c_uint32), _p(eb, C.c_uint32),
                                   _p(labels, C.c_uint32), ka, kb, float(eps), gen_seed, engine_seed)

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, randomize):
        self.L.ref_init(self.h, 1 if randomize else 0)

    def anneal(self, schedule, p0, p1, duration, steps_await):
        if isinstance(schedule, str):
            schedule = SCHEDULES[schedule]
        return self.L.ref_anneal(self.h, schedule, p0, p1, duration, steps_await)

    def merge_path(self, KA, KB, p0, sampling_steps, steps_await):
        """reference src/mcmc_main.cc:406-451 (labels with more blocks than -z): ladder of agg_merge + greedy sweeps, final
        abrupt_cool anneal; returns entropy().  Afterwards labels() has (KA, KB) blocks."""
        s = self.L.ref_merge_path(self.h, KA, KB, p0, sampling_steps, steps_await)
        self.ka, self.kb = int(self.L.ref_get_ka(self.h)), int(self.L.ref_get_kb(self.h))
        self.K = self.ka + self.kb
        return s

    def nature_path(self, p0, sampling_steps, steps_await):
        """reference src/mcmc_main.cc:360-379 + 397-402 (-g -u); the chain must start from singleton blocks"""
        s = self.L.ref_nature_path(self.h, p0, sampling_steps, steps_await)
        self._refresh_k()
        return s

    def _refresh_k(self):
        self.ka, self.kb = int(self.L.ref_get_ka(self.h)), int(self.L.ref_get_kb(self.h))
        self.K = self.ka + self.kb

    def agg_merge(self, diff_a, diff_b=None, nm=10):
        """blockmodel_t::agg_merge, both overloads (src/blockmodel.cc:109-256); diff_b None = the --nature form"""
        if diff_b is None:
            self.L.ref_agg_merge_total(self.h, diff_a, nm)
        else:
            self.L.ref_agg_merge(self.h, diff_a, diff_b, nm)
        self._refresh_k()

    def anneal_seconds(self):
        return self.L.ref_last_anneal_seconds(self.h)

    def step(self, v, T):
        return bool(self.L.ref_step(self.h, v, T))

    def transition(self, v, s):
        dS, ar = C.c_double(), C.c_double()
        self.L.ref_transition(self.h, v, s, C.byref(dS), C.byref(ar))
        return dS.value, ar.value

    def entropy(self):
        return self.L.ref_entropy(self.h)

    def entropy_accum(self):
        return self.L.ref_entropy_accum(self.h)

    def labels(self):
        out = np.empty(self.n, dtype=np.uint32)
        self.L.ref_get_labels(self.h, _p(out, C.c_uint32))
        return out

    def vlist(self):
        out = np.empty(self.n, dtype=np.uint32)
        self.L.ref_get_vlist(self.h, _p(out, C.c_uint32))
        return out

    def m(self):
        out = np.empty((self.K, self.K), dtype=np.int32)
        self.L.ref_get_m(self.h, _p(out, C.c_int32))
        return out

    def m_r(self):
        out = np.empty(self.K, dtype=np.int32)
        self.L.ref_get_m_r(self.h, _p(out, C.c_int32))
        return out

    def n_r(self):
        out = np.empty(self.K, dtype=np.int32)
        self.L.ref_get_n_r(self.h, _p(out, C.c_int32))
        return out

    def eta(self):
        w = self.L.ref_eta_width(self.h)
        out = np.empty((self.K, w), dtype=np.uint32)
        self.L.ref_get_eta(self.h, _p(out, C.c_uint32))
        return out

    def k(self, v):
        out = np.empty(self.K, dtype=np.int32)
        self.L.ref_get_k(self.h, v, _p(out, C.c_int32))
        return out

    def rng_words(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.ref_rng_words(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


def schedule(schedule_id, p0, p1, t, log_rng=False):
    return _load(log_rng).ref_schedule(schedule_id, p0, p1, t)


def log_q(n, k):
    return _load(False).ref_log_q(n, k)


def log_q_approx(n, k):
    return _load(False).ref_log_q_approx(n, k)


def lgamma_fast(x):
    return _load(False).ref_lgamma_fast(x)


def spence(x):
    return _load(False).ref_spence(x)


def init_tables(num_edges):
    _load(False).ref_init_tables(num_edges)


def load_edge_list(path):
    """Whitespace-separated unsigned pairs, one edge per line, file order kept
    (reference src/graph_utilities.cc:20-34)."""
    return np.loadtxt(path, dtype=np.uint32).reshape(-1, 2)


def labels_from_block_sizes(sizes):
    """Block r repeated sizes[r] times (reference src/mcmc_main.cc:310-317)."""
    return np.repeat(np.arange(len(sizes), dtype=np.uint32), sizes)

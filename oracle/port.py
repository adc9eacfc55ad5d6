"""ctypes front-end of oracle/liboracle.so -- the plain-C restatement (oracle/bisbm_oracle.c).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module.  The product (bipartitesbm-mcmc_b200/, bin/mcmc) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "liboracle.so")

EXPONENTIAL, LINEAR, LOGARITHMIC, CONSTANT, ABRUPT_COOL = range(5)
SCHEDULES = {"exponential": 0, "linear": 1, "logarithmic": 2, "constant": 3, "abrupt_cool": 4}

_lib = None


def build():
    """(Re)build liboracle.so with gcc if it is missing or older than its source."""
    src = os.path.join(_HERE, "bisbm_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "port"], stdout=subprocess.DEVNULL)


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB)
    u64, u32, dbl, vp = C.c_uint64, C.c_uint32, C.c_double, C.c_void_p
    pu32, pi32 = C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
    L.ora_create.restype = vp
    L.ora_create.argtypes = [u32, u32, u32, u64, pu32, pu32, pu32, u32, u32, dbl, u32, u32]
    L.ora_destroy.argtypes = [vp]
    L.ora_init.argtypes = [vp, C.c_int]
    L.ora_anneal.restype = dbl
    L.ora_anneal.argtypes = [vp, C.c_int, C.c_float, C.c_float, u64, u64]
    L.ora_anneal_alternating.restype = dbl
    L.ora_anneal_alternating.argtypes = [vp, C.c_int, C.c_float, C.c_float, u64, u64]
    L.ora_step.restype = C.c_int
    L.ora_step.argtypes = [vp, u32, dbl]
    L.ora_transition.argtypes = [vp, u32, u32, C.POINTER(dbl), C.POINTER(dbl)]
    L.ora_log_q.restype = dbl
    L.ora_log_q.argtypes = [vp, C.c_int, C.c_int]
    for name in ("ora_entropy", "ora_entropy_accum"):
        getattr(L, name).restype = dbl
        getattr(L, name).argtypes = [vp]
    L.ora_sweeps_done.restype = u64
    L.ora_sweeps_done.argtypes = [vp]
    L.ora_get_labels.argtypes = [vp, pu32]
    L.ora_get_vlist.argtypes = [vp, pu32]
    L.ora_get_m.argtypes = [vp, pi32]
    L.ora_get_m_r.argtypes = [vp, pi32]
    L.ora_get_n_r.argtypes = [vp, pi32]
    L.ora_eta_width.restype = u32
    L.ora_eta_width.argtypes = [vp]
    L.ora_get_eta.argtypes = [vp, pu32]
    L.ora_get_k.argtypes = [vp, u32, pi32]
    L.ora_rng_words.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.ora_schedule.restype = dbl
    L.ora_schedule.argtypes = [C.c_int, C.c_float, C.c_float, u64]
    L.ora_spence.restype = dbl
    L.ora_spence.argtypes = [dbl]
    L.ora_log_q_approx.restype = dbl
    L.ora_log_q_approx.argtypes = [u64, u64]
    L.ora_build_log_q_table.restype = C.POINTER(dbl)
    L.ora_build_log_q_table.argtypes = [u32, u32]
    L.ora_free.argtypes = [vp]
    L.ora_mt_seed.argtypes = [vp, u32]
    L.ora_mt_next.restype = u32
    L.ora_mt_next.argtypes = [vp]
    L.ora_canon.restype = dbl
    L.ora_canon.argtypes = [vp]
    L.ora_nd.restype = u32
    L.ora_nd.argtypes = [vp, u32]
    L.ora_shuffle.argtypes = [pu32, u64, vp]
    L.ora_categorical.restype = u32
    L.ora_categorical.argtypes = [pi32, u32, vp]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class MT(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("idx", C.c_int), ("words", C.c_uint64)]


class Rng:
    """std::mt19937 + the libstdc++ distribution transforms of oracle/bisbm_oracle.c."""

    def __init__(self, seed):
        self.L = _load()
        self.s = MT()
        self.L.ora_mt_seed(C.byref(self.s), seed)

    def word(self):
        return self.L.ora_mt_next(C.byref(self.s))

    def canon(self):
        return self.L.ora_canon(C.byref(self.s))

    def nd(self, r):
        return self.L.ora_nd(C.byref(self.s), r)

    def shuffle(self, x):
        x = np.ascontiguousarray(x, dtype=np.uint32)
        self.L.ora_shuffle(_p(x, C.c_uint32), x.size, C.byref(self.s))
        return x

    def categorical(self, w):
        w = np.ascontiguousarray(w, dtype=np.int32)
        return self.L.ora_categorical(_p(w, C.c_int32), w.size, C.byref(self.s))

    @property
    def words(self):
        return self.s.words


class PortChain:
    """One chain of the C restatement; same surface as oracle.ref.RefChain."""

    def __init__(self, n, na, nb, edges, labels, ka, kb, eps, engine_seed, gen_seed=12345):
        self.L = _load()
        edges = np.ascontiguousarray(edges, dtype=np.uint32).reshape(-1, 2)
        ea = np.ascontiguousarray(edges[:, 0])
        eb = np.ascontiguousarray(edges[:, 1])
        labels = np.ascontiguousarray(labels, dtype=np.uint32)
        assert labels.size == n == na + nb
        self.n, self.na, self.nb, self.ka, self.kb = n, na, nb, ka, kb
        self.K = ka + kb
        self.h = self.L.ora_create(n, na, nb, len(ea), _p(ea, C.c_uint32), _p(eb, C.c_uint32),
                                   _p(labels, C.c_uint32), ka, kb, float(eps), engine_seed, gen_seed)

    def close(self):
        if self.h:
            self.L.ora_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init(self, randomize):
        self.L.ora_init(self.h, 1 if randomize else 0)

    def anneal(self, schedule, p0, p1, duration, steps_await, alternate=False):
        """alternate=True: the type-alternating visiting order of the GPU's parallel mode instead of the reference's
        (ora_anneal_alternating; a test aid to isolate that documented deviation)."""
        if isinstance(schedule, str):
            schedule = SCHEDULES[schedule]
        fn = self.L.ora_anneal_alternating if alternate else self.L.ora_anneal
        return fn(self.h, schedule, p0, p1, duration, steps_await)

    def step(self, v, T):
        return bool(self.L.ora_step(self.h, v, T))

    def transition(self, v, s):
        dS, ar = C.c_double(), C.c_double()
        self.L.ora_transition(self.h, v, s, C.byref(dS), C.byref(ar))
        return dS.value, ar.value

    def log_q(self, n, k):
        return self.L.ora_log_q(self.h, n, k)

    def entropy(self):
        return self.L.ora_entropy(self.h)

    def entropy_accum(self):
        return self.L.ora_entropy_accum(self.h)

    def sweeps_done(self):
        return self.L.ora_sweeps_done(self.h)

    def labels(self):
        out = np.empty(self.n, dtype=np.uint32)
        self.L.ora_get_labels(self.h, _p(out, C.c_uint32))
        return out

    def vlist(self):
        out = np.empty(self.n, dtype=np.uint32)
        self.L.ora_get_vlist(self.h, _p(out, C.c_uint32))
        return out

    def m(self):
        out = np.empty((self.K, self.K), dtype=np.int32)
        self.L.ora_get_m(self.h, _p(out, C.c_int32))
        return out

    def m_r(self):
        out = np.empty(self.K, dtype=np.int32)
        self.L.ora_get_m_r(self.h, _p(out, C.c_int32))
        return out

    def n_r(self):
        out = np.empty(self.K, dtype=np.int32)
        self.L.ora_get_n_r(self.h, _p(out, C.c_int32))
        return out

    def eta(self):
        w = self.L.ora_eta_width(self.h)
        out = np.empty((self.K, w), dtype=np.uint32)
        self.L.ora_get_eta(self.h, _p(out, C.c_uint32))
        return out

    def k(self, v):
        out = np.empty(self.K, dtype=np.int32)
        self.L.ora_get_k(self.h, v, _p(out, C.c_int32))
        return out

    def rng_words(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.ora_rng_words(self.h, C.byref(a), C.byref(b))
        return a.value, b.value


def schedule(schedule_id, p0, p1, t):
    return _load().ora_schedule(schedule_id, p0, p1, t)


def spence(x):
    return _load().ora_spence(x)


def log_q_approx(n, k):
    return _load().ora_log_q_approx(n, k)


def log_q_table(n_max, k_max):
    L = _load()
    p = L.ora_build_log_q_table(n_max, k_max)
    a = np.ctypeslib.as_array(p, shape=(n_max + 1, k_max + 1)).copy()
    L.ora_free(p)
    return a

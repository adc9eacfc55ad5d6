// replay.cuh -- sequential-replay mode: ONE chain, moves strictly in the reference's order,
// consuming the reference's two mt19937 streams through libstdc++-exact transforms, with
// every floating-point operation rounded separately in the reference's order.  Reproduces
// metropolis_hasting::anneal / step / transition_ratio (reference
// src/metropolis_hasting.cc:42-192) and blockmodel_t::single_vertex_change /
// apply_mcmc_moves / shuffle_bisbm (reference src/blockmodel.cc:461-503, 613-637, 672-679)
// bit for bit on labels, m_rs, e_r, n_r, eta and the accumulated dS.
//
// This mode exists for parity, not throughput: the chain's logic runs on lane 0 of one
// warp.  lgamma and log q(n<10001,k) come from host-built glibc tables so dS is bit-exact;
// log(accu_r) / exp(a) use CUDA libm (<= 1 ulp from glibc; can flip an accept only when
// the uniform draw lands within 1 ulp of exp(a)).
#pragma once
#include "state.cuh"

namespace bisbm {

struct ReplayState {
    uint32_t engine[MT_STATE_WORDS];  // the seeded `engine` of src/mcmc_main.cc:242
    uint32_t gen[MT_STATE_WORDS];     // blockmodel_t::gen, src/blockmodel.hh:17-18
    double entropy_accum;             // blockmodel_t::entropy_ (sum of accepted dS)
    double entropy_min;               // metropolis_hasting::entropy_min_
    double accu_r;                    // metropolis_hasting::accu_r_
    uint64_t accepted, u, sweeps_done;
    uint32_t stopped;
    int32_t last_accept;
    double last_dS;
};

enum : uint32_t { RP_KT_MAX = 64u, RP_KT_OVER = 0xffffffffu };

struct ReplayCtx {
    GraphView g;
    ChainRef c;
    Tables tb;
    ReplayState* rs;
    uint32_t* vlist;  // [n], persists across sweeps and anneal calls (src/metropolis_hasting.cc:78)
    int32_t* kh;      // [max(KA,KB)] neighbour-block histogram scratch: zero outside the bins listed in kt
    uint32_t* kt;     // [1 + RP_KT_MAX]: kt[0] = number of non-zero bins of kh (RP_KT_OVER: too many to list), then the bins, ascending
    double eps;
};

BISBM_HD uint32_t rp_label(const ReplayCtx& x, uint32_t v) {  // global block id of v
    int32_t l = x.c.labels[(size_t)v * x.c.C];
    return v < x.g.na ? (uint32_t)l : x.c.ka + (uint32_t)l;
}

// neighbour-block histogram of v over the opposite type's blocks (reference k_[v], kept as
// an N x K matrix there; rebuilt from the adjacency here)
// The non-zero bins are listed in ascending order (kt), so the loops over "every block t with k_t != 0" of transition_ratio and
// apply_mcmc_moves cost O(degree) instead of O(K) and still add their terms in the reference's order.
BISBM_HD void rp_hist(const ReplayCtx& x, uint32_t v) {
    uint32_t n = x.kt[0];
    if (n == RP_KT_OVER) {
        uint32_t kmax = x.c.KA > x.c.KB ? x.c.KA : x.c.KB;      // (the strides: K may have changed since the list overflowed)
        for (uint32_t t = 0; t < kmax; ++t) x.kh[t] = 0;
    } else {
        for (uint32_t j = 0; j < n; ++j) x.kh[x.kt[1 + j]] = 0;
    }
    n = 0;
    bool over = false;
    for (uint32_t e = x.g.row_ptr[v]; e < x.g.row_ptr[v + 1]; ++e) {
        uint32_t nb = x.g.col[e];
        uint32_t t = (uint32_t)x.c.labels[(size_t)nb * x.c.C];
        if (x.kh[t]++ == 0 && !over) {
            if (n == RP_KT_MAX) { over = true; continue; }
            uint32_t j = n++;                                   // insert t, keeping the list ascending
            while (j > 0 && x.kt[j] > t) { x.kt[1 + j] = x.kt[j]; --j; }
            x.kt[1 + j] = t;
        }
    }
    x.kt[0] = over ? RP_KT_OVER : n;
}
// the loops below visit bin number j of v's histogram: the j-th listed bin, or simply bin j when the list overflowed
BISBM_HD uint32_t rp_bins(const ReplayCtx& x, uint32_t kopp) { return x.kt[0] == RP_KT_OVER ? kopp : x.kt[0]; }
BISBM_HD uint32_t rp_bin(const ReplayCtx& x, uint32_t j) { return x.kt[0] == RP_KT_OVER ? j : x.kt[1 + j]; }

// transition_ratio (src/metropolis_hasting.cc:103-192); kh must hold v's histogram when
// r != s are of the same type.
BISBM_HD double rp_transition(const ReplayCtx& x, uint32_t v, uint32_t r, uint32_t s) {
    ReplayState* rs = x.rs;
    if (r == s) { rs->accu_r = 1.0; return 0.0; }
    const ChainRef& c = x.c;
    uint32_t ka = c.ka, K = c.ka + c.kb;
    if ((r < ka && s >= ka) || (r >= ka && s < ka)) return BISBM_INF;  // accu_r stays stale
    double eps = x.eps, Kd = (double)K;
    int deg = (int)(x.g.row_ptr[v + 1] - x.g.row_ptr[v]);
    uint32_t didx = x.g.degidx[v];
    uint32_t sr = slot_of(c, r), ss = slot_of(c, s);
    int n_rr = nr_ref(c, sr), n_rs = nr_ref(c, ss);
    int eta_r = eta_ref(c, sr, didx), eta_s = eta_ref(c, ss, didx);
    int e0r = e_ref(c, sr), e1r = e0r - deg, e0s = e_ref(c, ss), e1s = e0s + deg;
    double a0 = 0.0, a1 = 0.0, S0 = 0.0, S1 = 0.0;
    bool va = r < ka;
    uint32_t kopp = va ? c.kb : c.ka;
    for (uint32_t j = 0, nbins = rp_bins(x, kopp); j < nbins; ++j) {  // ascending global id of the opposite type
        uint32_t t = rp_bin(x, j);
        int kk = x.kh[t];
        if (kk == 0) continue;
        uint32_t gi = va ? ka + t : t;
        int m_r = m_at(c, r, gi), m_s = m_at(c, s, gi);
        int e_i = e_ref(c, slot_of(c, gi));
        double den = dadd((double)e_i, dmul(eps, Kd));
        a0 = dadd(a0, ddiv(ddiv(dmul((double)kk, dadd((double)m_s, eps)), den), (double)deg));
        a1 = dadd(a1, ddiv(ddiv(dmul((double)kk, dadd((double)(m_r - kk), eps)), den), (double)deg));
        S0 = dsub(S0, lgamma_int(x.tb, (int64_t)m_r + 1));
        S0 = dsub(S0, lgamma_int(x.tb, (int64_t)m_s + 1));
        S1 = dsub(S1, lgamma_int(x.tb, (int64_t)m_r - kk + 1));
        S1 = dsub(S1, lgamma_int(x.tb, (int64_t)m_s + kk + 1));
    }
    S0 = dsub(S0, -lgamma_int(x.tb, (int64_t)e0r + 1));
    S0 = dsub(S0, -lgamma_int(x.tb, (int64_t)e0s + 1));
    S1 = dsub(S1, -lgamma_int(x.tb, (int64_t)e1r + 1));
    S1 = dsub(S1, -lgamma_int(x.tb, (int64_t)e1s + 1));
    S0 = dadd(S0, -lgamma_int(x.tb, (int64_t)eta_r + 1));
    S0 = dadd(S0, -lgamma_int(x.tb, (int64_t)eta_s + 1));
    S1 = dadd(S1, -lgamma_int(x.tb, (int64_t)eta_r - 1 + 1));
    S1 = dadd(S1, -lgamma_int(x.tb, (int64_t)eta_s + 1 + 1));
    S0 = dadd(S0, log_q(x.tb, e0r, n_rr));
    S0 = dadd(S0, log_q(x.tb, e0s, n_rs));
    S1 = dadd(S1, log_q(x.tb, e1r, n_rr - 1));
    S1 = dadd(S1, log_q(x.tb, e1s, n_rs + 1));
    rs->accu_r = (deg == 0) ? 1.0 : ddiv(a1, a0);
    return dsub(S1, S0);
}

// std::discrete_distribution<size_t>(m_[t].begin(), m_[t].end())(gen)
// (libstdc++ bits/random.tcc:2657-2714) over the FULL K-long row of the symmetric matrix
BISBM_HD uint32_t rp_categorical(const ReplayCtx& x, uint32_t t) {
    const ChainRef& c = x.c;
    uint32_t K = c.ka + c.kb;
    if (K < 2) return 0;
    // the weights are integers: their sequential double sum is exact and equals the row sum m_r_[t] the state carries
    const double sum = (double)e_ref(c, slot_of(c, t));
    double u = mt_canon(x.rs->gen);
    // row t is zero outside the other type's blocks [lo, hi): those entries add +0.0 to the running sum.  Leading zeros
    // have cp = 0 (taken only by u == 0); behind hi the partial sum stays where it is until the forced cp = 1 of the last
    // entry.  Walking [lo, hi) alone therefore returns what lower_bound over the full K-long vector returns.
    const uint32_t lo = t < c.ka ? c.ka : 0u, hi = t < c.ka ? K : c.ka;
    if (!(sum > 0.0)) return 0;              // (0 / 0 weights: every cp is NaN and lower_bound stays at the first entry)
    if (lo > 0 && !(0.0 < u)) return 0;
    double acc = 0.0;
    for (uint32_t i = lo; i < hi; ++i) {
        double p = ddiv((double)m_at(c, t, i), sum);
        acc = (i == 0) ? p : dadd(acc, p);
        double cp = (i == K - 1) ? 1.0 : acc;
        if (!(cp < u)) return i;  // lower_bound: first cp >= u
    }
    return K - 1;
}

// single_vertex_change (src/blockmodel.cc:613-637)
BISBM_HD uint32_t rp_propose(const ReplayCtx& x, uint32_t v, uint32_t r) {
    const ChainRef& c = x.c;
    uint32_t K = c.ka + c.kb;
    bool tb = v >= x.g.na;
    if ((!tb && c.ka == 1) || (tb && c.kb == 1)) return r;
    uint32_t row = x.g.row_ptr[v], d = x.g.row_ptr[v + 1] - row;
    if (d == 0) return (uint32_t)(uint64_t)dmul(mt_canon(x.rs->engine), (double)K);
    uint64_t which = (uint64_t)dmul(mt_canon(x.rs->engine), (double)d);
    uint32_t j = x.g.col[row + which];
    uint32_t t = rp_label(x, j);
    double eK = dmul(x.eps, (double)K);
    double R = ddiv(eK, dadd((double)e_ref(c, slot_of(c, t)), eK));
    if (mt_canon(x.rs->engine) < R) return (uint32_t)(uint64_t)dmul(mt_canon(x.rs->engine), (double)K);
    return rp_categorical(x, t);
}

// apply_mcmc_moves (src/blockmodel.cc:461-503)
BISBM_HD bool rp_apply(const ReplayCtx& x, uint32_t v, uint32_t r, uint32_t s, double dS) {
    const ChainRef& c = x.c;
    uint32_t sr = slot_of(c, r), ss = slot_of(c, s);
    if (nr_ref(c, sr) - 1 == 0) return false;  // would empty block r
    if (r != s) {
        nr_ref(c, sr)--; nr_ref(c, ss)++;
        uint32_t didx = x.g.degidx[v];
        eta_ref(c, sr, didx)--;
        eta_ref(c, ss, didx)++;
        bool va = r < c.ka;
        uint32_t kopp = va ? c.kb : c.ka;
        for (uint32_t j = 0, nbins = rp_bins(x, kopp); j < nbins; ++j) {
            uint32_t t = rp_bin(x, j);
            int kk = x.kh[t];
            if (kk == 0) continue;
            uint32_t gi = va ? c.ka + t : t;
            *m_ptr(c, r, gi) -= kk;
            *m_ptr(c, s, gi) += kk;
        }
        int deg = (int)(x.g.row_ptr[v + 1] - x.g.row_ptr[v]);
        e_ref(c, sr) -= deg;
        e_ref(c, ss) += deg;
        c.labels[(size_t)v * c.C] = (int32_t)(va ? s : s - c.ka);
    }
    x.rs->entropy_accum = dadd(x.rs->entropy_accum, dS);
    return true;
}

// step (src/metropolis_hasting.cc:42-62)
BISBM_HD bool rp_step(const ReplayCtx& x, uint32_t v, double T) {
    uint32_t r = rp_label(x, v);
    uint32_t s = rp_propose(x, v, r);
    uint32_t ka = x.c.ka;
    bool same_type = (r < ka) == (s < ka);
    if (r != s && same_type) rp_hist(x, v);
    double dS = rp_transition(x, v, r, s);
    x.rs->last_dS = dS;
    if (T == 0.0) {
        if (dS < 0) return rp_apply(x, v, r, s, dS);
        return false;
    }
    double a = dadd(dmul(ddiv(-1.0, T), dS), log(x.rs->accu_r));
    if (a > 0.0) return rp_apply(x, v, r, s, dS);
    if (mt_canon(x.rs->engine) < exp(a)) return rp_apply(x, v, r, s, dS);
    return false;
}

// cooling schedules whose arithmetic is exact on the device (src/metropolis_hasting.cc:15-37);
// exponential / logarithmic temperatures come from the host (glibc pow / log) via `temps`.
BISBM_HD double rp_schedule(int schedule, float p0, float p1, uint64_t t) {
    switch (schedule) {
        case 1: {  // linear, float arithmetic
#ifdef __CUDA_ARCH__
            float r = __fsub_rn(p0, __fmul_rn(p1, (float)t));
#else
            volatile float pr = p1 * (float)t; volatile float r = p0 - pr;
#endif
            return (double)r;
        }
        case 3: return (double)p0;
        default: return ((float)t < p0) ? 1.0 : 0.0;  // abrupt_cool
    }
}

// anneal (src/metropolis_hasting.cc:64-101), sweeps [first_sweep, first_sweep + n_sweeps)
BISBM_HD void rp_anneal_sweeps(const ReplayCtx& x, int schedule, float p0, float p1, const double* temps,
                               uint64_t first_sweep, uint64_t n_sweeps, uint64_t steps_await) {
    ReplayState* rs = x.rs;
    uint64_t N = x.g.n;
    for (uint64_t sw = 0; sw < n_sweeps; ++sw) {
        uint64_t sweep = first_sweep + sw;
        mt_shuffle(x.vlist, N, 1, rs->engine);
        uint64_t base = N * sweep;
        for (uint64_t vi = 0; vi < N; ++vi) {
            double T = temps ? temps[sw * N + vi] : rp_schedule(schedule, p0, p1, base + vi);
            if (rp_step(x, x.vlist[vi], T)) {
                ++rs->accepted;
                if (rs->entropy_accum < rs->entropy_min) { rs->entropy_min = rs->entropy_accum; rs->u = 0; }
            }
            if (T < 1.0) ++rs->u;
        }
        rs->sweeps_done = sweep + 1;
        if (rs->u >= steps_await) { rs->stopped = 1; return; }
    }
}

// shuffle_bisbm's two std::shuffle calls (src/blockmodel.cc:672-674) on the chain-minor labels
BISBM_HD void rp_shuffle_labels(const ReplayCtx& x) {
    uint32_t* lab = (uint32_t*)x.c.labels;
    mt_shuffle(lab, x.g.na, x.c.C, x.rs->engine);
    mt_shuffle(lab + (size_t)x.g.na * x.c.C, x.g.nb, x.c.C, x.rs->engine);
}

#ifdef __CUDACC__
__global__ void replay_init_kernel(ReplayCtx x, uint32_t engine_seed, uint32_t gen_seed, int randomize) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ReplayState* rs = x.rs;
    mt_seed(rs->engine, engine_seed);
    mt_seed(rs->gen, gen_seed);
    rs->entropy_accum = 0.0;
    rs->entropy_min = BISBM_INF;
    rs->accu_r = 0.0;
    rs->accepted = 0; rs->u = 0; rs->sweeps_done = 0; rs->stopped = 0;
    for (uint32_t v = 0; v < x.g.n; ++v) x.vlist[v] = v;
    if (randomize) rp_shuffle_labels(x);
}

__global__ void replay_anneal_kernel(ReplayCtx x, int schedule, float p0, float p1, const double* temps,
                                     uint64_t first_sweep, uint64_t n_sweeps, uint64_t steps_await) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    rp_anneal_sweeps(x, schedule, p0, p1, temps, first_sweep, n_sweeps, steps_await);
}

__global__ void replay_step_kernel(ReplayCtx x, uint32_t v, double T) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    x.rs->last_accept = rp_step(x, v, T) ? 1 : 0;
}

__global__ void replay_transition_kernel(ReplayCtx x, uint32_t v, uint32_t s) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t r = rp_label(x, v);
    rp_hist(x, v);
    x.rs->last_dS = rp_transition(x, v, r, s);
}
#endif

}  // namespace bisbm

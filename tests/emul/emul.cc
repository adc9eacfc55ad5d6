// tests/emul/emul.cc -- HOST compilation of the device headers, for debugging only.
//
// The container that authors this code has no GPU.  The replay logic (replay.cuh) and the
// per-move arithmetic of the parallel kernel (sweep.cuh: lgamma_diff, logq_delta) are plain
// BISBM_HD functions, so g++ can compile the very same text and a CPU test can compare it
// with the oracle before any GPU time is spent.  TEST INFRASTRUCTURE: built only by
// tests/test_emul.py into tests/emul/libemul.so; libbisbm.so never contains or calls it.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../bipartitesbm-mcmc_b200/csrc/replay.cuh"
#include "../../bipartitesbm-mcmc_b200/csrc/sweep.cuh"
#include "../../bipartitesbm-mcmc_b200/csrc/sweep2.cuh"
#include "../../bipartitesbm-mcmc_b200/csrc/gltable.h"

using namespace bisbm;

struct Emul {
    uint32_t n, na, nb, ka, kb, C, chain, W, maxdeg;
    uint64_t E;
    std::vector<uint32_t> row_ptr, col, degidx, degvals, vlist;
    std::vector<int32_t> labels, m, e, nr, eta, kh;
    std::vector<uint32_t> kt;
    std::vector<double> lg, qtab;
    uint32_t qn, qk;
    ReplayState rs;
    double eps;
};

static void build_qtab(std::vector<double>& q, uint32_t qn, uint32_t qk) {
    const size_t W = (size_t)qk + 1;
    q.assign(((size_t)qn + 1) * W, -INFINITY);
    for (size_t n = 1; n <= qn; ++n) {
        double* row = q.data() + n * W;
        row[1] = 0.0;
        size_t kend = n < qk ? n : qk;
        for (size_t k = 2; k <= kend; ++k) {
            auto lsum = [](double a, double b) { double mx = a > b ? a : b; return mx + log1p(exp(-fabs(a - b))); };
            double v = lsum(row[k], row[k - 1]);
            if (n > k) v = lsum(v, q[(n - k) * W + k]);
            row[k] = v;
        }
    }
}

static void rebuild(Emul* s) {
    const uint32_t KA = s->ka, KB = s->kb;
    std::fill(s->m.begin(), s->m.end(), 0); std::fill(s->e.begin(), s->e.end(), 0);
    std::fill(s->nr.begin(), s->nr.end(), 0); std::fill(s->eta.begin(), s->eta.end(), 0);
    for (uint32_t v = 0; v < s->n; ++v) {
        uint32_t b = s->labels[(size_t)v * s->C + s->chain];
        uint32_t slot = v < s->na ? b : KA + b;
        s->nr[slot]++;
        s->eta[(size_t)slot * s->W + s->degidx[v]]++;
        if (v < s->na)
            for (uint32_t x = s->row_ptr[v]; x < s->row_ptr[v + 1]; ++x)
                s->m[(size_t)b * KB + s->labels[(size_t)s->col[x] * s->C + s->chain]]++;
    }
    for (uint32_t a = 0; a < KA; ++a) for (uint32_t b = 0; b < KB; ++b) {
        s->e[a] += s->m[(size_t)a * KB + b]; s->e[KA + b] += s->m[(size_t)a * KB + b];
    }
}

static ReplayCtx ctx(Emul* s) {
    ReplayCtx x;
    x.g.n = s->n; x.g.na = s->na; x.g.nb = s->nb; x.g.n_edges = s->E;
    x.g.row_ptr = s->row_ptr.data(); x.g.col = s->col.data(); x.g.degidx = s->degidx.data();
    x.g.W = s->W; x.g.max_degree = s->maxdeg;
    x.c.labels = s->labels.data() + s->chain; x.c.C = s->C;
    x.c.m = s->m.data(); x.c.e = s->e.data(); x.c.nr = s->nr.data(); x.c.eta = s->eta.data();
    x.c.ka = s->ka; x.c.kb = s->kb; x.c.KA = s->ka; x.c.KB = s->kb; x.c.W = s->W; x.c.cs = 1;
    x.tb.lg = s->lg.data(); x.tb.lg_n = s->lg.size(); x.tb.qtab = s->qtab.data(); x.tb.qn = s->qn; x.tb.qk = s->qk;
    x.rs = &s->rs; x.vlist = s->vlist.data(); x.kh = s->kh.data(); x.kt = s->kt.data(); x.eps = s->eps;
    return x;
}

extern "C" {

void* emul_create(uint32_t na, uint32_t nb, uint64_t E, const uint32_t* ea, const uint32_t* eb, const uint32_t* labels,
                  uint32_t ka, uint32_t kb, double eps, uint32_t engine_seed, uint32_t gen_seed, int randomize) {
    Emul* s = new Emul();
    s->n = na + nb; s->na = na; s->nb = nb; s->ka = ka; s->kb = kb; s->E = E; s->eps = eps;
    s->C = 32; s->chain = 3;
    s->row_ptr.assign(s->n + 1, 0);
    for (uint64_t i = 0; i < E; ++i) { s->row_ptr[ea[i] + 1]++; s->row_ptr[eb[i] + 1]++; }
    for (uint32_t v = 0; v < s->n; ++v) s->row_ptr[v + 1] += s->row_ptr[v];
    std::vector<uint32_t> fill(s->row_ptr.begin(), s->row_ptr.end() - 1);
    s->col.resize(2 * E);
    for (uint64_t i = 0; i < E; ++i) { s->col[fill[ea[i]]++] = eb[i]; s->col[fill[eb[i]]++] = ea[i]; }
    std::vector<uint32_t> deg(s->n);
    s->maxdeg = 0;
    for (uint32_t v = 0; v < s->n; ++v) { deg[v] = s->row_ptr[v + 1] - s->row_ptr[v]; if (deg[v] > s->maxdeg) s->maxdeg = deg[v]; }
    s->degvals = deg;
    std::sort(s->degvals.begin(), s->degvals.end());
    s->degvals.erase(std::unique(s->degvals.begin(), s->degvals.end()), s->degvals.end());
    s->W = s->degvals.size();
    s->degidx.resize(s->n);
    for (uint32_t v = 0; v < s->n; ++v) s->degidx[v] = std::lower_bound(s->degvals.begin(), s->degvals.end(), deg[v]) - s->degvals.begin();
    s->labels.assign((size_t)s->n * s->C, 0);
    for (uint32_t v = 0; v < s->n; ++v) s->labels[(size_t)v * s->C + s->chain] = v < na ? labels[v] : labels[v] - ka;
    s->m.assign((size_t)ka * kb, 0); s->e.assign(ka + kb, 0); s->nr.assign(ka + kb, 0);
    s->eta.assign((size_t)(ka + kb) * s->W, 0); s->kh.assign(std::max(ka, kb), 0); s->kt.assign(1 + RP_KT_MAX, 0);
    s->vlist.resize(s->n);
    for (uint32_t v = 0; v < s->n; ++v) s->vlist[v] = v;
    s->lg.resize(2 * E + 2 + s->maxdeg);
    s->lg[0] = INFINITY;
    for (size_t i = 1; i < s->lg.size(); ++i) s->lg[i] = lgamma((double)i);
    s->qn = E < 10000 ? E : 10000;
    uint32_t kmax = na > nb ? na : nb;
    s->qk = kmax < s->qn ? kmax : s->qn;
    if (s->qk < 1) s->qk = 1;
    build_qtab(s->qtab, s->qn, s->qk);
    mt_seed(s->rs.engine, engine_seed); mt_seed(s->rs.gen, gen_seed);
    s->rs.entropy_accum = 0; s->rs.entropy_min = INFINITY; s->rs.accu_r = 0;
    s->rs.accepted = 0; s->rs.u = 0; s->rs.sweeps_done = 0; s->rs.stopped = 0;
    if (randomize) rp_shuffle_labels(ctx(s));
    rebuild(s);
    return s;
}

void emul_destroy(void* p) { delete (Emul*)p; }

double emul_anneal(void* p, int schedule, float p0, float p1, const double* temps, uint64_t duration, uint64_t steps_await) {
    Emul* s = (Emul*)p;
    s->rs.entropy_min = INFINITY; s->rs.accepted = 0; s->rs.u = 0; s->rs.sweeps_done = 0; s->rs.stopped = 0;
    uint64_t sweeps = duration / s->n;
    rp_anneal_sweeps(ctx(s), schedule, p0, p1, temps, 0, sweeps, steps_await);
    if (s->rs.stopped) return (double)s->rs.accepted / (double)(s->rs.sweeps_done * s->n);
    return (double)s->rs.accepted / (double)duration;
}

void emul_transition(void* p, uint32_t v, uint32_t sg, double* dS, double* accu) {
    Emul* s = (Emul*)p;
    ReplayCtx x = ctx(s);
    rp_hist(x, v);
    *dS = rp_transition(x, v, rp_label(x, v), sg);
    *accu = s->rs.accu_r;
}

// the parallel kernel's arithmetic for the same move: dS via log-products / Stirling / logq_delta
void emul_par_dS(void* p, uint32_t v, uint32_t sg, int use_taylor, double* dS, double* accu) {
    Emul* s = (Emul*)p;
    ReplayCtx x = ctx(s);
    rp_hist(x, v);
    const bool va = v < s->na;
    const uint32_t r = s->labels[(size_t)v * s->C + s->chain];
    const uint32_t sl = va ? sg : sg - s->ka;
    const uint32_t kopp = va ? s->kb : s->ka, KB = s->kb, KA = s->ka;
    const uint32_t own_off = va ? 0 : KA, opp_off = va ? KA : 0;
    const uint32_t sx = va ? KB : 1, st = va ? 1 : KB;
    const int32_t* Mr = s->m.data() + (size_t)r * sx; const int32_t* Ms = s->m.data() + (size_t)sl * sx;
    const double eps = s->eps, epsK = eps * (double)(KA + KB);
    const uint32_t d = s->row_ptr[v + 1] - s->row_ptr[v];
    // the kernel's one pass over v's neighbours (sweep.cuh: acc_edge / acc_guard)
    MoveAcc A; acc_init(A);
    std::vector<int> cnt(kopp, 0);
    for (uint32_t e = 0; e < d; ++e) {
        uint32_t nb = s->col[s->row_ptr[v] + e];
        uint32_t t = s->labels[(size_t)nb * s->C + s->chain];
        int c = cnt[t]++;
        double inv = 1.0 / ((double)s->e[opp_off + t] + epsK);
        acc_edge(A, Mr[(size_t)t * st], Ms[(size_t)t * st], c, inv, eps);
        if ((e & 7u) == 7u) acc_guard(A);
    }
    int e_r = s->e[own_off + r], e_s = s->e[own_off + sl], n_r = s->nr[own_off + r], n_s = s->nr[own_off + sl];
    uint32_t didx = s->degidx[v];
    int eta_r = s->eta[(size_t)(own_off + r) * s->W + didx], eta_s = s->eta[(size_t)(own_off + sl) * s->W + didx];
    double eta_ratio = (double)(eta_r > 0 ? eta_r : 1) / (double)(eta_s + 1);
    double out = A.logacc + log(A.num / A.den * eta_ratio);
    out += block_degree_delta(e_r, e_s, (int)d);
    double a0 = A.a0, a1 = A.a1;
    Tables tb = x.tb; tb.lg = nullptr; tb.lg_n = 0;
    LogqExp qr, qs;
    memset(&qr, 0, sizeof qr); memset(&qs, 0, sizeof qs);
    if (use_taylor) {
        auto mk = [&](int e0, int n0) { return logq_expand(tb, e0, n0); };
        // expansion point deliberately off the current point, to exercise the drift terms
        qr = mk(e_r + e_r / 40, n_r - n_r / 50); qs = mk(e_s - e_s / 40, n_s + n_s / 50);
    }
    out += logq_delta(tb, qr, e_r, n_r, -(int)d, -1);
    out += logq_delta(tb, qs, e_s, n_s, (int)d, 1);
    *dS = out;
    *accu = d == 0 ? 1.0 : a1 / a0;
}

}  // extern "C"

// the staged kernel's arithmetic (sweep2.cuh: macc_edge / macc_fold / bdd_fast / logq_fast / move_finish) for the same
// move, in R = float or double; *fell_back counts the terms that took the out-of-line double completion
template <typename R>
static void par2_dS(Emul* s, uint32_t v, uint32_t sg, int use_taylor, double* dS, double* log_accu, int* fell_back) {
    ReplayCtx x = ctx(s);
    const bool va = v < s->na;
    const uint32_t r = s->labels[(size_t)v * s->C + s->chain];
    const uint32_t sl = va ? sg : sg - s->ka;
    const uint32_t kopp = va ? s->kb : s->ka, KB = s->kb, KA = s->ka;
    const uint32_t own_off = va ? 0 : KA, opp_off = va ? KA : 0;
    const uint32_t sx = va ? KB : 1, st = va ? 1 : KB;
    const int32_t* Mr = s->m.data() + (size_t)r * sx; const int32_t* Ms = s->m.data() + (size_t)sl * sx;
    const uint32_t d = s->row_ptr[v + 1] - s->row_ptr[v];
    MAcc<R> A; macc_init(A);
    std::vector<int> cnt(kopp, 0);
    for (uint32_t e = 0; e < d; ++e) {
        uint32_t nb = s->col[s->row_ptr[v] + e];
        uint32_t t = s->labels[(size_t)nb * s->C + s->chain];
        int c = cnt[t]++;
        R inv = (R)(1.0 / ((double)s->e[opp_off + t] + s->eps * (double)(KA + KB)));
        macc_edge<R>(A, Mr[(size_t)t * st], Ms[(size_t)t * st], (uint32_t)c, inv);
        // the kernel folds after every chunk of 4 edges (float) or after every block of 32 (double)
        if (Ar<R>::FOLD < 32 ? ((e & 3u) == 3u || e + 1 == d) : ((e & 31u) == 31u && e + 1 < d)) macc_fold(A);
    }
    int e_r = s->e[own_off + r], e_s = s->e[own_off + sl], n_r = s->nr[own_off + r], n_s = s->nr[own_off + sl];
    uint32_t didx = s->degidx[v];
    int eta_r = s->eta[(size_t)(own_off + r) * s->W + didx], eta_s = s->eta[(size_t)(own_off + sl) * s->W + didx];
    Tables tb = x.tb; tb.lg = nullptr; tb.lg_n = 0;
    LogqExp qr, qs;
    memset(&qr, 0, sizeof qr); memset(&qs, 0, sizeof qs);
    if (use_taylor) {
        auto mk = [&](int e0, int n0) { return logq_expand(tb, e0, n0); };
        qr = mk(e_r + e_r / 40, n_r - n_r / 50); qs = mk(e_s - e_s / 40, n_s + n_s / 50);
    }
    bool ok_b, ok_r, ok_s;
    R bdd = bdd_fast<R>(e_r, e_s, (int)d, &ok_b);
    R lqr = logq_fast<R>(qr, e_r, n_r, -(int)d, -1, &ok_r);
    R lqs = logq_fast<R>(qs, e_s, n_s, (int)d, 1, &ok_s);
    *fell_back = (!ok_b) + (!ok_r) + (!ok_s);
    if (!ok_b) bdd = (R)block_degree_delta(e_r, e_s, (int)d);
    if (!ok_r) lqr = (R)logq_delta_exact(tb, e_r, n_r, -(int)d, -1);
    if (!ok_s) lqs = (R)logq_delta_exact(tb, e_s, n_s, (int)d, 1);
    R out, lh;
    move_finish<R>(A, eta_r, eta_s, bdd, lqr, lqs, d, (R)s->eps, &out, &lh);
    *dS = (double)out;
    *log_accu = (double)(lh * Ar<R>::unit());
}

extern "C" {

void emul_par2_dS(void* p, uint32_t v, uint32_t sg, int use_taylor, int fp32, double* dS, double* log_accu, int* fell_back) {
    if (fp32) par2_dS<float>((Emul*)p, v, sg, use_taylor, dS, log_accu, fell_back);
    else par2_dS<double>((Emul*)p, v, sg, use_taylor, dS, log_accu, fell_back);
}

double emul_entropy_accum(void* p) { return ((Emul*)p)->rs.entropy_accum; }
void emul_get_labels(void* p, uint32_t* out) {
    Emul* s = (Emul*)p;
    for (uint32_t v = 0; v < s->n; ++v) { uint32_t l = s->labels[(size_t)v * s->C + s->chain]; out[v] = v < s->na ? l : s->ka + l; }
}
void emul_get_counts(void* p, int32_t* m, int32_t* e, int32_t* nr) {
    Emul* s = (Emul*)p;
    memcpy(m, s->m.data(), s->m.size() * 4); memcpy(e, s->e.data(), s->e.size() * 4); memcpy(nr, s->nr.data(), s->nr.size() * 4);
}
void emul_words(void* p, uint64_t* ew, uint64_t* gw) {
    Emul* s = (Emul*)p;
    *ew = ((uint64_t)s->rs.engine[626] << 32) | s->rs.engine[625];
    *gw = ((uint64_t)s->rs.gen[626] << 32) | s->rs.gen[625];
}
double emul_lgamma_diff(double x, double d) { return lgamma_diff(x, d); }
// log q expansion about (e0, n0) evaluated at (e0 + x, n0 + y) for the move (de, dn), and the exact difference
void emul_logq_expansion(int e0, int n0, int x, int y, int de, int dn, double* approx, double* exact, int* ok) {
    Tables tb; memset(&tb, 0, sizeof tb);
    LogqExp q = logq_expand(tb, e0, n0);
    bool o;
    *approx = logq_fast<double>(q, e0 + x, n0 + y, de, dn, &o);
    *ok = o ? 1 : 0;
    *exact = log_q_approx(tb, e0 + x + de, n0 + y + dn) - log_q_approx(tb, e0 + x, n0 + y);
}
double emul_dlog(double x) { return dlog(x); }
double emul_dexp(double x) { return dexp(x); }
double emul_block_degree_delta(int e_r, int e_s, int d) { return block_degree_delta(e_r, e_s, d); }
double emul_log_q_approx(uint64_t n, uint64_t k) { Tables tb; memset(&tb, 0, sizeof tb); return log_q_approx(tb, n, k); }
// log_q_approx through the tabulated g(u), lf(u) (what the parallel kernel's blocks without an expansion use) next to the formula
double emul_log_q_approx_tab(uint64_t n, uint64_t k) {
    static std::vector<GlNode> tab;
    const double t0 = -2.0, inv_h = 1024.0, t1 = 11.6;
    if (tab.empty()) build_gl_table(tab, t0, inv_h, t1);
    Tables tb; memset(&tb, 0, sizeof tb);
    tb.gl = tab.data(); tb.gl_t0 = t0; tb.gl_inv_h = inv_h; tb.gl_n = (uint32_t)tab.size();
    return log_q_approx_tab(tb, n, k);
}
// how the table is organised: nodes, runs of equal iteration count, flagged intervals
void emul_gl_table_stats(uint32_t* nodes, uint32_t* runs, uint32_t* short_runs, uint32_t* flagged) {
    std::vector<GlNode> tab;
    build_gl_table(tab, -2.0, 1024.0, 11.6);
    *nodes = (uint32_t)tab.size(); *runs = 0; *short_runs = 0; *flagged = 0;
    for (size_t i = 0; i < tab.size(); ++i) {
        if (tab[i].first == i) { ++*runs; if (tab[i].last - tab[i].first < 5) ++*short_runs; }
        if (tab[i].flags & 1u) ++*flagged;
    }
}
uint32_t emul_feistel(uint32_t i, uint32_t n, uint64_t key) { return feistel_perm(i, n, feistel_half_bits(n), key); }

}  // extern "C"

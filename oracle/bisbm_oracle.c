/* oracle/bisbm_oracle.c -- plain-C restatement of the reference's Metropolis-Hastings sweep.
 *
 * TEST INFRASTRUCTURE, NOT A PRODUCT PATH.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load liboracle.so.  The product
 * (libbisbm.so, bin/mcmc) never links or calls anything in this file.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here, bit for bit,
 * against the unmodified reference compiled in place (oracle/_ref/libref.so, built by
 * oracle/Makefile from /root/reference/src) and against the fixtures in tests/golden/
 * that tests/golden/make_golden.py generated from that same build.
 *
 * Written from the behaviour of the reference (paths below are under /root/reference) and
 * of libstdc++ 13 <random>/<algorithm>; no reference source is copied.  Floating point:
 * every accumulation is a separate IEEE double operation in the reference's order; build
 * with -ffp-contract=off (the reference is built for baseline x86-64, no FMA).
 */
#include "bisbm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ mt19937 */
/* std::mt19937 (32-bit Mersenne twister, libstdc++ bits/random.tcc seed()/_M_gen_rand). */
void ora_mt_seed(ora_mt19937* g, uint32_t seed) {
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
    g->words = 0;
}

static void mt_twist(ora_mt19937* g) {
    uint32_t* mt = g->mt;
    for (int i = 0; i < 624; ++i) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
        mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
}

uint32_t ora_mt_next(ora_mt19937* g) {
    if (g->idx >= 624) mt_twist(g);
    uint32_t y = g->mt[g->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    g->words++;
    return y;
}

/* std::generate_canonical<double,53>(mt19937): two words, low word first
 * (libstdc++ bits/random.tcc:3349-3381); uniform_real_distribution<double>(0,1) and
 * discrete_distribution both draw through it. */
double ora_canon(ora_mt19937* g) {
    double w0 = (double)ora_mt_next(g);
    double w1 = (double)ora_mt_next(g);
    double r = (w0 + w1 * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

/* uniform_int_distribution::_S_nd<uint64_t> (Lemire), range < 2^32
 * (libstdc++ bits/uniform_int_dist.h:257-281). */
uint32_t ora_nd(ora_mt19937* g, uint32_t range) {
    uint64_t prod = (uint64_t)ora_mt_next(g) * (uint64_t)range;
    uint32_t low = (uint32_t)prod;
    if (low < range) {
        uint32_t thr = (uint32_t)(0u - range) % range;
        while (low < thr) {
            prod = (uint64_t)ora_mt_next(g) * (uint64_t)range;
            low = (uint32_t)prod;
        }
    }
    return (uint32_t)(prod >> 32);
}

/* uniform_int_distribution<size_t>(0, hi) on a 32-bit engine when hi+1 <= 2^32. */
static uint64_t uid(ora_mt19937* g, uint64_t hi) {
    if (hi == 0xffffffffull) return ora_mt_next(g);
    return ora_nd(g, (uint32_t)(hi + 1));
}

/* std::shuffle (libstdc++ bits/stl_algo.h:3742-3805): two swaps per draw while
 * n*n fits the engine range, else one draw per element. */
void ora_shuffle(uint32_t* x, uint64_t n, ora_mt19937* g) {
    if (n == 0) return;
    uint32_t tmp;
    if (0xffffffffull / n >= n) {
        uint64_t i = 1;
        if ((n % 2) == 0) {
            uint64_t j = uid(g, 1);
            tmp = x[i]; x[i] = x[j]; x[j] = tmp;
            ++i;
        }
        while (i != n) {
            uint64_t sr = i + 1;
            uint64_t b1 = sr + 1;
            uint64_t v = uid(g, sr * b1 - 1);
            uint64_t p0 = v / b1, p1 = v % b1;
            tmp = x[i]; x[i] = x[p0]; x[p0] = tmp;
            ++i;
            tmp = x[i]; x[i] = x[p1]; x[p1] = tmp;
            ++i;
        }
        return;
    }
    for (uint64_t i = 1; i < n; ++i) {
        uint64_t j = uid(g, i);
        tmp = x[i]; x[i] = x[j]; x[j] = tmp;
    }
}

/* std::discrete_distribution<size_t>(w, w+k)(g) (libstdc++ bits/random.tcc:2657-2714):
 * probabilities p_i = w_i / sum, cumulative by partial_sum, last forced to 1,
 * lower_bound(cp, canon). */
uint32_t ora_categorical(const int32_t* w, uint32_t k, ora_mt19937* g) {
    if (k < 2) return 0; /* empty _M_cp: returns 0 without drawing */
    double sum = 0.0;
    for (uint32_t i = 0; i < k; ++i) sum += (double)w[i];
    double* cp = (double*)malloc(sizeof(double) * k);
    double acc = 0.0;
    for (uint32_t i = 0; i < k; ++i) {
        double p = (double)w[i] / sum;
        acc = (i == 0) ? p : acc + p;
        cp[i] = acc;
    }
    cp[k - 1] = 1.0;
    double u = ora_canon(g);
    uint32_t lo = 0, len = k; /* std::lower_bound: first cp[i] >= u */
    while (len > 0) {
        uint32_t half = len >> 1;
        if (cp[lo + half] < u) { lo = lo + half + 1; len = len - half - 1; }
        else len = half;
    }
    free(cp);
    return lo;
}

/* ------------------------------------------------------------------ math tables */
/* Dilogarithm, Cephes spence() (S. L. Moshier), as used by src/support/spence.cc:108-154.
 * The rational-approximation coefficients are the published Cephes constants. */
static const double SP_A[8] = {4.65128586073990045278E-5, 7.31589045238094711071E-3, 1.33847639578309018650E-1,
                               8.79691311754530315341E-1, 2.71149851196553469920E0,  4.25697156008121755724E0,
                               3.29771340985225106936E0,  1.00000000000000000126E0};
static const double SP_B[8] = {6.90990488912553276999E-4, 2.54043763932544379113E-2, 2.82974860602568089943E-1,
                               1.41172597751831069617E0,  3.63800533345137075418E0,  5.03278880143316990390E0,
                               3.54771340985225096217E0,  9.99999999999999998740E-1};

static double horner7(double x, const double* c) { /* src/support/spence.cc:91-106 */
    double ans = c[0];
    for (int i = 1; i <= 7; ++i) ans = ans * x + c[i];
    return ans;
}

double ora_spence(double x) {
    if (x < 0.0) return NAN;
    if (x == 1.0) return 0.0;
    if (x == 0.0) return M_PI * M_PI / 6.0;
    int flag = 0;
    double w;
    if (x > 2.0) { x = 1.0 / x; flag |= 2; }
    if (x > 1.5) { w = (1.0 / x) - 1.0; flag |= 2; }
    else if (x < 0.5) { w = -x; flag |= 1; }
    else w = x - 1.0;
    double y = -w * horner7(w, SP_A) / horner7(w, SP_B);
    if (flag & 1) y = (M_PI * M_PI) / 6.0 - log(x) * log1p(-x) - y;
    if (flag & 2) { double z = log(x); y = -0.5 * z * z - y; }
    return y;
}

/* lgamma_fast: table of glibc lgamma(i), entry 0 = +inf (src/support/cache.cc:64-79,
 * src/support/cache.hh:82-93).  The table is only a cache of lgamma(double(i)). */
static double lg(int64_t i) {
    if (i == 0) return INFINITY;
    return lgamma((double)i);
}

/* src/support/int_part.cc:30-32 */
static double log_sum(double a, double b) {
    double mx = a > b ? a : b;
    return mx + log1p(exp(-fabs(a - b)));
}

/* init_q_cache (src/support/int_part.cc:34-51) restricted to the rows/columns asked for.
 * Row-major [n_max+1][k_max+1]; untouched cells stay -inf exactly as in the reference. */
double* ora_build_log_q_table(uint32_t n_max, uint32_t k_max) {
    size_t W = (size_t)k_max + 1;
    double* q = (double*)malloc(sizeof(double) * ((size_t)n_max + 1) * W);
    if (!q) return NULL;
    for (size_t i = 0; i < ((size_t)n_max + 1) * W; ++i) q[i] = -INFINITY;
    for (size_t n = 1; n <= n_max; ++n) {
        if (k_max >= 1) q[n * W + 1] = 0;
        size_t kend = n < k_max ? n : k_max;
        for (size_t k = 2; k <= kend; ++k) {
            double v = log_sum(q[n * W + k], q[n * W + k - 1]);
            if (n > k) v = log_sum(v, q[(n - k) * W + k]);
            q[n * W + k] = v;
        }
    }
    return q;
}

void ora_free(void* p) { free(p); }

/* lbinom_fast (src/support/util.hh:41-47) */
static double lbinom_fast(uint64_t N, uint64_t k) {
    if (N == 0 || k == 0 || k > N) return 0;
    return (lg((int64_t)(N + 1)) - lg((int64_t)(k + 1))) - lg((int64_t)(N - k + 1));
}

/* get_v (src/support/int_part.cc:77-86) */
static double get_v(double u) {
    double v = u, delta = 1;
    while (delta > 1e-8) {
        double n_v = u * sqrt(ora_spence(exp(-v)));
        delta = fabs(n_v - v);
        v = n_v;
    }
    return v;
}

/* log_q_approx (src/support/int_part.cc:73-75, 88-98) */
double ora_log_q_approx(uint64_t n, uint64_t k) {
    if ((double)k < pow((double)n, 1 / 4.)) return lbinom_fast(n - 1, k - 1) - lg((int64_t)(k + 1));
    double u = (double)k / sqrt((double)n);
    double v = get_v(u);
    double lf = log(v) - log1p(-exp(-v) * (1 + u * u / 2)) / 2 - log(2) * 3 / 2. - log(u) - log(M_PI);
    double g = 2 * v / u - u * log1p(-exp(-v));
    return lf - log((double)n) + sqrt((double)n) * g;
}

/* The five cooling schedules (src/metropolis_hasting.cc:10-37) with the reference's
 * float-typed parameters and mixed float/double arithmetic. */
double ora_schedule(int schedule, float p0, float p1, uint64_t t) {
    switch (schedule) {
        case ORA_EXPONENTIAL: return (double)p0 * pow((double)p1, (double)t);
        case ORA_LINEAR: { float r = p0 - p1 * (float)t; return (double)r; }
        case ORA_LOGARITHMIC: {
            float x = (float)t + p1;
            uint64_t i = (uint64_t)x; /* safelog_fast: table index = truncation */
            double l = (i == 0) ? 0.0 : log((double)i);
            return (double)p0 / l;
        }
        case ORA_CONSTANT: return (double)p0;
        default: return ((float)t < p0) ? 1.0 : 0.0;
    }
}

/* ------------------------------------------------------------------ chain */
struct ora_chain {
    uint32_t n, na, nb, ka, kb, K;
    uint64_t n_edges;
    double eps;
    uint64_t* adj_off; /* n+1 */
    uint32_t* adj;     /* 2E, file order, multi-edges kept (src/graph_utilities.cc:36-49) */
    uint32_t* deg;
    uint32_t max_degree;
    uint32_t* labels;
    uint32_t* vlist;
    int32_t* n_r;
    int32_t* e_r; /* reference m_r_ */
    int32_t* m;   /* K x K symmetric */
    int32_t* k;   /* n x K */
    uint32_t* eta; /* K x (max_degree+1) */
    double entropy_accum, entropy_min, accu_r;
    uint64_t sweeps_done;
    ora_mt19937 engine, gen;
    double* qtab; /* [qn+1][qk+1] */
    uint32_t qn, qk;
};

ora_chain* ora_create(uint32_t n, uint32_t na, uint32_t nb, uint64_t n_edges, const uint32_t* ea,
                      const uint32_t* eb, const uint32_t* labels, uint32_t ka, uint32_t kb, double eps,
                      uint32_t engine_seed, uint32_t gen_seed) {
    ora_chain* c = (ora_chain*)calloc(1, sizeof(ora_chain));
    c->n = n; c->na = na; c->nb = nb; c->ka = ka; c->kb = kb; c->K = ka + kb;
    c->n_edges = n_edges; c->eps = eps;
    c->deg = (uint32_t*)calloc(n, sizeof(uint32_t));
    for (uint64_t i = 0; i < n_edges; ++i) { c->deg[ea[i]]++; c->deg[eb[i]]++; }
    c->adj_off = (uint64_t*)calloc((size_t)n + 1, sizeof(uint64_t));
    for (uint32_t v = 0; v < n; ++v) {
        c->adj_off[v + 1] = c->adj_off[v] + c->deg[v];
        if (c->deg[v] > c->max_degree) c->max_degree = c->deg[v];
    }
    c->adj = (uint32_t*)malloc(sizeof(uint32_t) * (2 * n_edges + 1));
    uint64_t* fill = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)n + 1));
    memcpy(fill, c->adj_off, sizeof(uint64_t) * ((size_t)n + 1));
    for (uint64_t i = 0; i < n_edges; ++i) { c->adj[fill[ea[i]]++] = eb[i]; c->adj[fill[eb[i]]++] = ea[i]; }
    free(fill);
    c->labels = (uint32_t*)malloc(sizeof(uint32_t) * n);
    memcpy(c->labels, labels, sizeof(uint32_t) * n);
    c->vlist = (uint32_t*)malloc(sizeof(uint32_t) * n);
    for (uint32_t v = 0; v < n; ++v) c->vlist[v] = v; /* src/blockmodel.cc:41 */
    c->n_r = (int32_t*)calloc(c->K, sizeof(int32_t));
    c->e_r = (int32_t*)calloc(c->K, sizeof(int32_t));
    c->m = (int32_t*)calloc((size_t)c->K * c->K, sizeof(int32_t));
    c->k = (int32_t*)calloc((size_t)n * c->K, sizeof(int32_t));
    c->eta = (uint32_t*)calloc((size_t)c->K * (c->max_degree + 1), sizeof(uint32_t));
    c->entropy_accum = 0.0; c->entropy_min = INFINITY; c->accu_r = 0.0;
    ora_mt_seed(&c->engine, engine_seed);
    ora_mt_seed(&c->gen, gen_seed);
    /* the reference builds a 10001^2 table (src/blockmodel.cc:48); only rows n <= E and
     * columns k <= max(na, nb) can ever be looked up for this graph */
    c->qn = n_edges < 10000 ? (uint32_t)n_edges : 10000;
    uint32_t kmax = na > nb ? na : nb;
    c->qk = kmax < c->qn ? kmax : c->qn;
    if (c->qk < 1) c->qk = 1;
    c->qtab = ora_build_log_q_table(c->qn, c->qk);
    return c;
}

void ora_destroy(ora_chain* c) {
    if (!c) return;
    free(c->adj_off); free(c->adj); free(c->deg); free(c->labels); free(c->vlist);
    free(c->n_r); free(c->e_r); free(c->m); free(c->k); free(c->eta); free(c->qtab);
    free(c);
}

/* compute_n_r/k/m/m_r/eta_rk (src/blockmodel.cc:691-746) */
static void rebuild(ora_chain* c) {
    uint32_t K = c->K, W = c->max_degree + 1;
    memset(c->n_r, 0, sizeof(int32_t) * K);
    memset(c->e_r, 0, sizeof(int32_t) * K);
    memset(c->m, 0, sizeof(int32_t) * (size_t)K * K);
    memset(c->k, 0, sizeof(int32_t) * (size_t)c->n * K);
    memset(c->eta, 0, sizeof(uint32_t) * (size_t)K * W);
    for (uint32_t v = 0; v < c->n; ++v) {
        uint32_t b = c->labels[v];
        c->n_r[b]++;
        c->eta[(size_t)b * W + c->deg[v]]++;
        for (uint64_t e = c->adj_off[v]; e < c->adj_off[v + 1]; ++e) {
            uint32_t bn = c->labels[c->adj[e]];
            c->k[(size_t)v * K + bn]++;
            c->m[(size_t)b * K + bn]++;
        }
    }
    for (uint32_t r = 0; r < K; ++r) {
        int64_t s = 0;
        for (uint32_t t = 0; t < K; ++t) s += c->m[(size_t)r * K + t];
        c->e_r[r] = (int32_t)s;
    }
}

/* shuffle_bisbm / init_bisbm (src/blockmodel.cc:672-688) */
void ora_init(ora_chain* c, int randomize) {
    if (randomize) {
        ora_shuffle(c->labels, c->na, &c->engine);
        ora_shuffle(c->labels + c->na, c->nb, &c->engine);
    }
    rebuild(c);
}

/* log_q<int> (src/support/int_part.hh:27-37) with the reference's 10001-row table */
double ora_log_q(const ora_chain* c, int n, int k) {
    if (n <= 0 || k < 1) return 0;
    if (k > n) k = n;
    if (n < 10001) {
        if ((uint32_t)n <= c->qn && (uint32_t)k <= c->qk) return c->qtab[(size_t)n * (c->qk + 1) + k];
        /* outside the graph-restricted table: rebuild the reference value directly */
        double* t = ora_build_log_q_table((uint32_t)n, (uint32_t)k);
        double v = t[(size_t)n * ((size_t)k + 1) + k];
        free(t);
        return v;
    }
    return ora_log_q_approx((uint64_t)n, (uint64_t)k);
}

/* transition_ratio (src/metropolis_hasting.cc:103-192).  Leaves accu_r untouched on the
 * cross-type early return, exactly like the reference member accu_r_. */
static double transition(ora_chain* c, uint32_t v, uint32_t r, uint32_t s) {
    if (r == s) { c->accu_r = 1.; return 0.; }
    uint32_t KA = c->ka, K = c->K;
    double Kd = (double)K;
    if ((r < KA && s >= KA) || (r >= KA && s < KA)) return INFINITY;
    double eps = c->eps;
    double a0 = 0., a1 = 0., S0 = 0., S1 = 0.;
    const int32_t* kv = c->k + (size_t)v * K;
    const int32_t* mr = c->m + (size_t)r * K;
    const int32_t* ms = c->m + (size_t)s * K;
    int deg = (int)c->deg[v];
    uint32_t W = c->max_degree + 1;
    int n_rr = c->n_r[r], n_rs = c->n_r[s];
    int eta_r = (int)c->eta[(size_t)r * W + deg], eta_s = (int)c->eta[(size_t)s * W + deg];
    int e0r = c->e_r[r], e1r = e0r - deg, e0s = c->e_r[s], e1s = e0s + deg;
    uint32_t lo = (r < KA) ? KA : 0, hi = (r < KA) ? K : KA;
    for (uint32_t i = lo; i < hi; ++i) {
        int kk = kv[i];
        if (kk == 0) continue;
        a0 += kk * (ms[i] + eps) / (c->e_r[i] + eps * Kd) / deg;
        a1 += kk * (mr[i] - kk + eps) / (c->e_r[i] + eps * Kd) / deg;
        S0 -= lg(mr[i] + 1);
        S0 -= lg(ms[i] + 1);
        S1 -= lg(mr[i] - kk + 1);
        S1 -= lg(ms[i] + kk + 1);
    }
    S0 -= -lg(e0r + 1);
    S0 -= -lg(e0s + 1);
    S1 -= -lg(e1r + 1);
    S1 -= -lg(e1s + 1);
    S0 += -lg(eta_r + 1);
    S0 += -lg(eta_s + 1);
    S1 += -lg(eta_r - 1 + 1);
    S1 += -lg(eta_s + 1 + 1);
    S0 += ora_log_q(c, e0r, n_rr);
    S0 += ora_log_q(c, e0s, n_rs);
    S1 += ora_log_q(c, e1r, n_rr - 1);
    S1 += ora_log_q(c, e1s, n_rs + 1);
    c->accu_r = (deg == 0) ? 1 : a1 / a0;
    return S1 - S0;
}

void ora_transition(ora_chain* c, uint32_t v, uint32_t s, double* dS, double* accu_r) {
    *dS = transition(c, v, c->labels[v], s);
    *accu_r = c->accu_r;
}

/* single_vertex_change (src/blockmodel.cc:613-637) */
static uint32_t propose(ora_chain* c, uint32_t v) {
    uint32_t K = c->K;
    int type_b = v >= c->na;
    if ((!type_b && c->ka == 1) || (type_b && c->kb == 1)) return c->labels[v];
    uint32_t d = c->deg[v];
    if (d == 0) return (uint32_t)(uint64_t)(ora_canon(&c->engine) * (double)K);
    uint64_t which = (uint64_t)(ora_canon(&c->engine) * (double)d);
    uint32_t j = c->adj[c->adj_off[v] + which];
    uint32_t t = c->labels[j];
    double R = c->eps * (double)K / (c->e_r[t] + c->eps * (double)K);
    if (ora_canon(&c->engine) < R) return (uint32_t)(uint64_t)(ora_canon(&c->engine) * (double)K);
    return ora_categorical(c->m + (size_t)t * K, K, &c->gen);
}

/* apply_mcmc_moves (src/blockmodel.cc:461-503) */
static int apply(ora_chain* c, uint32_t v, uint32_t r, uint32_t s, double dS) {
    uint32_t K = c->K, W = c->max_degree + 1;
    if (c->n_r[r] - 1 == 0) return 0;
    c->n_r[r]--; c->n_r[s]++;
    uint32_t d = c->deg[v];
    c->eta[(size_t)r * W + d]--;
    c->eta[(size_t)s * W + d]++;
    const int32_t* kv = c->k + (size_t)v * K;
    for (uint32_t i = 0; i < K; ++i) {
        int kk = kv[i];
        if (kk != 0) {
            c->m[(size_t)r * K + i] -= kk;
            c->m[(size_t)s * K + i] += kk;
            c->m[(size_t)i * K + r] = c->m[(size_t)r * K + i];
            c->m[(size_t)i * K + s] = c->m[(size_t)s * K + i];
        }
    }
    c->e_r[r] -= (int32_t)d;
    c->e_r[s] += (int32_t)d;
    for (uint64_t e = c->adj_off[v]; e < c->adj_off[v + 1]; ++e) {
        uint32_t nb = c->adj[e];
        c->k[(size_t)nb * K + r]--;
        c->k[(size_t)nb * K + s]++;
    }
    c->labels[v] = s;
    c->entropy_accum += dS;
    return 1;
}

/* step (src/metropolis_hasting.cc:42-62) */
int ora_step(ora_chain* c, uint32_t v, double T) {
    uint32_t r = c->labels[v];
    uint32_t s = propose(c, v);
    double dS = transition(c, v, r, s);
    if (T == 0.) {
        if (dS < 0) return apply(c, v, r, s, dS);
        return 0;
    }
    double a = -1. / T * dS + log(c->accu_r);
    if (a > 0.) return apply(c, v, r, s, dS);
    if (ora_canon(&c->engine) < exp(a)) return apply(c, v, r, s, dS);
    return 0;
}

/* anneal (src/metropolis_hasting.cc:64-101) */
double ora_anneal(ora_chain* c, int schedule, float p0, float p1, uint64_t duration, uint64_t steps_await) {
    uint64_t N = c->n, accepted = 0, u = 0;
    c->entropy_min = INFINITY;
    uint64_t all_sweeps = duration / N;
    c->sweeps_done = 0;
    for (uint64_t sweep = 0; sweep < all_sweeps; ++sweep) {
        ora_shuffle(c->vlist, N, &c->engine);
        uint64_t base = N * sweep;
        for (uint64_t vi = 0; vi < N; ++vi) {
            double T = ora_schedule(schedule, p0, p1, base + vi);
            if (ora_step(c, c->vlist[vi], T)) {
                ++accepted;
                if (c->entropy_accum < c->entropy_min) { c->entropy_min = c->entropy_accum; u = 0; }
            }
            if (T < 1.) ++u;
        }
        c->sweeps_done = sweep + 1;
        if (u >= steps_await) return (double)accepted / (double)((sweep + 1) * N);
    }
    return (double)accepted / (double)duration;
}

/* NOT a reference function: anneal() with the visiting order of the GPU's parallel mode -- after the reference's
 * shuffle of vlist, the type-a vertices are visited first (in their shuffled order, steps base .. base + na - 1),
 * then the type-b vertices.  Everything else is ora_anneal.  The statistical-parity tests use it to tell the effect
 * of that documented deviation (type-alternating half sweeps) from everything else. */
double ora_anneal_alternating(ora_chain* c, int schedule, float p0, float p1, uint64_t duration, uint64_t steps_await) {
    uint64_t N = c->n, accepted = 0, u = 0;
    c->entropy_min = INFINITY;
    uint64_t all_sweeps = duration / N;
    c->sweeps_done = 0;
    uint32_t* order = (uint32_t*)malloc(sizeof(uint32_t) * (N ? N : 1));
    for (uint64_t sweep = 0; sweep < all_sweeps; ++sweep) {
        ora_shuffle(c->vlist, N, &c->engine);
        uint64_t k = 0;
        for (uint64_t vi = 0; vi < N; ++vi) if (c->vlist[vi] < c->na) order[k++] = c->vlist[vi];
        for (uint64_t vi = 0; vi < N; ++vi) if (c->vlist[vi] >= c->na) order[k++] = c->vlist[vi];
        uint64_t base = N * sweep;
        for (uint64_t vi = 0; vi < N; ++vi) {
            double T = ora_schedule(schedule, p0, p1, base + vi);
            if (ora_step(c, order[vi], T)) {
                ++accepted;
                if (c->entropy_accum < c->entropy_min) { c->entropy_min = c->entropy_accum; u = 0; }
            }
            if (T < 1.) ++u;
        }
        c->sweeps_done = sweep + 1;
        if (u >= steps_await) { free(order); return (double)accepted / (double)((sweep + 1) * N); }
    }
    free(order);
    return (double)accepted / (double)duration;
}

/* entropy (src/blockmodel.cc:753-787).  adj_map_ (std::map per node, ascending neighbour
 * id) is restated by sorting a copy of each adjacency row. */
static int cmp_u32(const void* a, const void* b) {
    uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return (x > y) - (x < y);
}

double ora_entropy(const ora_chain* c) {
    uint32_t K = c->K, W = c->max_degree + 1;
    double ent = 0;
    for (uint32_t v = 0; v < c->n; ++v) ent -= lg((int64_t)c->deg[v] + 1);
    for (uint32_t r = 0; r < K; ++r) {
        for (uint32_t s = r + 1; s < K; ++s) ent -= lg((int64_t)c->m[(size_t)r * K + s] + 1);
        for (uint32_t d = 0; d < W; ++d) ent -= lg((int64_t)c->eta[(size_t)r * W + d] + 1);
        ent += lg((int64_t)c->e_r[r] + 1);
        ent += ora_log_q(c, c->e_r[r], c->n_r[r]);
    }
    uint32_t* row = (uint32_t*)malloc(sizeof(uint32_t) * (c->max_degree + 1));
    for (uint32_t v = 0; v < c->n; ++v) {
        uint32_t d = c->deg[v];
        if (d < 2) continue;
        memcpy(row, c->adj + c->adj_off[v], sizeof(uint32_t) * d);
        qsort(row, d, sizeof(uint32_t), cmp_u32);
        uint32_t i = 0;
        while (i < d) {
            uint32_t j = i;
            while (j < d && row[j] == row[i]) ++j;
            uint32_t mult = j - i;
            if (mult > 1 && v > row[i]) ent += lg((int64_t)mult + 1);
            i = j;
        }
    }
    free(row);
    ent += lbinom_fast((uint64_t)c->ka * c->kb + c->n_edges - 1, c->n_edges);
    ent += lbinom_fast((uint64_t)c->na - 1, (uint64_t)c->ka - 1);
    ent += lbinom_fast((uint64_t)c->nb - 1, (uint64_t)c->kb - 1);
    {
        uint64_t p = (uint64_t)c->na * c->nb; /* safelog_fast(na*nb) */
        ent += (p == 0) ? 0.0 : log((double)p);
    }
    ent += lg((int64_t)c->na + 1);
    ent += lg((int64_t)c->nb + 1);
    return ent;
}

double ora_entropy_accum(const ora_chain* c) { return c->entropy_accum; }
uint64_t ora_sweeps_done(const ora_chain* c) { return c->sweeps_done; }
void ora_get_labels(const ora_chain* c, uint32_t* out) { memcpy(out, c->labels, sizeof(uint32_t) * c->n); }
void ora_get_vlist(const ora_chain* c, uint32_t* out) { memcpy(out, c->vlist, sizeof(uint32_t) * c->n); }
void ora_get_m(const ora_chain* c, int32_t* out) { memcpy(out, c->m, sizeof(int32_t) * (size_t)c->K * c->K); }
void ora_get_m_r(const ora_chain* c, int32_t* out) { memcpy(out, c->e_r, sizeof(int32_t) * c->K); }
void ora_get_n_r(const ora_chain* c, int32_t* out) { memcpy(out, c->n_r, sizeof(int32_t) * c->K); }
uint32_t ora_eta_width(const ora_chain* c) { return c->max_degree + 1; }
void ora_get_eta(const ora_chain* c, uint32_t* out) {
    memcpy(out, c->eta, sizeof(uint32_t) * (size_t)c->K * (c->max_degree + 1));
}
void ora_get_k(const ora_chain* c, uint32_t v, int32_t* out) {
    memcpy(out, c->k + (size_t)v * c->K, sizeof(int32_t) * c->K);
}
void ora_rng_words(const ora_chain* c, uint64_t* engine_words, uint64_t* gen_words) {
    *engine_words = c->engine.words;
    *gen_words = c->gen.words;
}

// oracle/ref_driver.cc -- C entry points over the UNMODIFIED reference classes.
//
// TEST INFRASTRUCTURE.  Built only by oracle/Makefile into oracle/_ref/libref*.so, from
// the reference sources where they lie under /root/reference (never copied).  Only
// tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the result.
//
// It replaces the reference's src/mcmc_main.cc (which needs boost::program_options):
// it mirrors the straight path of main -- labels, types, edge_to_adj, blockmodel_t ctor,
// shuffle_bisbm/init_bisbm, metropolis_hasting::anneal (reference src/mcmc_main.cc:
// 121-130, 336-339, 453-485) -- and exposes the state getters through a C ABI for ctypes.
//
// `step` / `transition_ratio` are `inline` members defined only in the reference's
// metropolis_hasting.cc, so this TU includes that .cc file (in place) instead of
// linking it.
#include "metropolis_hasting.cc"  // reference src/metropolis_hasting.cc, in place
#include "graph_utilities.hh"
#include "support/util.hh"       // geospace (reference src/support/util.hh:99-146), used by ref_merge_path

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>

extern "C" { unsigned int oracle_fake_rd_value = 12345u; }
double spence(double);  // reference src/support/spence.cc:108

namespace {

struct bm_open : public blockmodel_t {
    using blockmodel_t::blockmodel_t;
    std::mt19937& gen_ref() { return gen; }
};

struct mh_open : public metropolis_hasting {
    double accu_r() const { return accu_r_; }
    double entropy_min() const { return entropy_min_; }
    double dS_of(const blockmodel_t& bm, size_t v, size_t r, size_t s) {
        std::vector<mcmc_move_t> mv(1);
        mv[0].vertex = v; mv[0].source = r; mv[0].target = s;
        return transition_ratio(bm, mv);
    }
};

struct ref_handle {
    adj_list_t adj;
    std::unique_ptr<bm_open> bm;
    mh_open mh;
    std::mt19937 engine;
    size_t n = 0, na = 0, nb = 0, ka = 0, kb = 0;
    double last_seconds = 0.0;
};

}  // namespace

extern "C" {

// edges: n_edges pairs (a[i], b[i]) in file order; labels: n global block ids.
void* ref_create(uint64_t n, uint64_t na, uint64_t nb, uint64_t n_edges, const uint32_t* ea,
                 const uint32_t* eb, const uint32_t* labels, uint64_t ka, uint64_t kb, double eps,
                 uint32_t gen_seed, uint64_t engine_seed) {
    auto* h = new ref_handle();
    h->n = n; h->na = na; h->nb = nb; h->ka = ka; h->kb = kb;
    edge_list_t el;
    el.reserve(n_edges);
    for (uint64_t i = 0; i < n_edges; ++i) el.push_back(std::make_pair(size_t(ea[i]), size_t(eb[i])));
    h->adj = edge_to_adj(el, n);  // reference src/graph_utilities.cc:36-49
    uint_vec_t mb(labels, labels + n);
    uint_vec_t types(n, 0);       // reference src/mcmc_main.cc:121-130
    for (uint64_t i = na; i < n; ++i) types[i] = 1;
    oracle_fake_rd_value = gen_seed;
    h->bm.reset(new bm_open(mb, types, ka + kb, ka, kb, eps, &h->adj));
    h->engine.seed(engine_seed);  // reference src/mcmc_main.cc:242
    return h;
}

void ref_destroy(void* p) { delete static_cast<ref_handle*>(p); }

// reference src/mcmc_main.cc:457-461
void ref_init(void* p, int randomize) {
    auto* h = static_cast<ref_handle*>(p);
    if (randomize) h->bm->shuffle_bisbm(h->engine, h->na, h->nb);
    else h->bm->init_bisbm();
}

// schedule ids: 0 exponential, 1 linear, 2 logarithmic, 3 constant, 4 abrupt_cool
// (reference src/mcmc_main.cc:463-482).  Returns the accepted fraction.
double ref_anneal(void* p, int schedule, float p0, float p1, uint64_t duration, uint64_t steps_await) {
    auto* h = static_cast<ref_handle*>(p);
    float_vec_t kw(2, 0);
    kw[0] = p0; kw[1] = p1;
    double (*fn)(size_t, float_vec_t) = nullptr;
    switch (schedule) {
        case 0: fn = &exponential_schedule; break;
        case 1: fn = &linear_schedule; break;
        case 2: fn = &logarithmic_schedule; break;
        case 3: fn = &constant_schedule; break;
        default: fn = &abrupt_cool_schedule; break;
    }
    auto t0 = std::chrono::steady_clock::now();
    double rate = h->mh.anneal(*h->bm, fn, kw, duration, steps_await, h->engine);
    auto t1 = std::chrono::steady_clock::now();
    h->last_seconds = std::chrono::duration<double>(t1 - t0).count();
    return rate;
}

double ref_last_anneal_seconds(void* p) { return static_cast<ref_handle*>(p)->last_seconds; }

// The reference's merge path for initial labels with MORE blocks than -z asks for (src/mcmc_main.cc:406-451, the
// diff_a >= 0 && diff_b >= 0 branch): geospace ladder, agg_merge(engine, diff_a, diff_b, 10) per rung, one greedy sweep
// (abrupt_cool with kwargs {0}) between rungs, then the final anneal(abrupt_cool, {p0}, sampling_steps).  The handle
// must have been created with the labels' own (ka, kb) and initialised (ref_init(h, 0)).  Returns the final entropy().
double ref_merge_path(void* p, uint64_t KA, uint64_t KB, float p0, uint64_t sampling_steps, uint64_t steps_await) {
    auto* h = static_cast<ref_handle*>(p);
    blockmodel_t& bm = *h->bm;
    const double sigma = 1.01;
    int diff_a = (int)bm.get_KA() - (int)KA, diff_b = (int)bm.get_KB() - (int)KB;
    float_vec_t agg_kw(1, 0.);
    const size_t N = h->n;
    if (diff_a != 0 || diff_b != 0) {
        int_vec_t ka_s, kb_s;
        std::tie(ka_s, kb_s) = geospace(KA + diff_a, KA, KB + diff_b, KB, sigma);
        if (ka_s.size() == 1) bm.agg_merge(h->engine, diff_a, diff_b, 10);
        for (size_t i = 0; i + 1 < ka_s.size(); ++i) {
            diff_a = -(ka_s[i + 1] - ka_s[i]);
            diff_b = -(kb_s[i + 1] - kb_s[i]);
            bm.agg_merge(h->engine, diff_a, diff_b, 10);
            if (i != ka_s.size() - 2) h->mh.anneal(bm, &abrupt_cool_schedule, agg_kw, N * 1, steps_await, h->engine);
        }
    }
    float_vec_t kw(2, 0);
    kw[0] = p0;
    h->mh.anneal(bm, &abrupt_cool_schedule, kw, sampling_steps, steps_await, h->engine);
    h->ka = bm.get_KA(); h->kb = bm.get_KB();
    return bm.entropy();
}
// blockmodel_t::agg_merge(engine, diff_a, diff_b, nm) (src/blockmodel.cc:109-204), negative diffs -> agg_split
void ref_agg_merge(void* p, int diff_a, int diff_b, int nm) {
    auto* h = static_cast<ref_handle*>(p);
    h->bm->agg_merge(h->engine, diff_a, diff_b, nm);
    h->ka = h->bm->get_KA(); h->kb = h->bm->get_KB();
}
// blockmodel_t::agg_merge(engine, diff, nm) (src/blockmodel.cc:206-256), the --nature form
void ref_agg_merge_total(void* p, int diff, int nm) {
    auto* h = static_cast<ref_handle*>(p);
    h->bm->agg_merge(h->engine, diff, nm);
    h->ka = h->bm->get_KA(); h->kb = h->bm->get_KB();
}
// The reference's -g -u path (src/mcmc_main.cc:360-379, 397-402): from singleton blocks, agg_merge(engine, ceil(K (sigma - 1) /
// sigma), 10) + one greedy sweep until one type has fewer than ceil(sqrt(2E) / 2) blocks, then the final anneal.
double ref_nature_path(void* p, float p0, uint64_t sampling_steps, uint64_t steps_await) {
    auto* h = static_cast<ref_handle*>(p);
    blockmodel_t& bm = *h->bm;
    const double sigma = 1.01;
    float_vec_t agg_kw(1, 0.);
    size_t tKA = h->na, tKB = h->nb, tGroups = h->na + h->nb;
    size_t num_edges = bm.get_num_edges();
    size_t ceiling = ceil(sqrt(2 * num_edges) / 2);
    while (tKA >= ceiling && tKB >= ceiling) {
        bm.agg_merge(h->engine, ceil(tGroups * (sigma - 1) / sigma), 10);
        tKA = bm.get_KA(); tKB = bm.get_KB(); tGroups = tKA + tKB;
        h->mh.anneal(bm, &abrupt_cool_schedule, agg_kw, (h->na + h->nb) * 1, steps_await, h->engine);
    }
    float_vec_t kw(2, 0);
    kw[0] = p0;
    h->mh.anneal(bm, &abrupt_cool_schedule, kw, sampling_steps, steps_await, h->engine);
    h->ka = bm.get_KA(); h->kb = bm.get_KB();
    return bm.entropy();
}
uint64_t ref_get_ka(void* p) { return static_cast<ref_handle*>(p)->bm->get_KA(); }
uint64_t ref_get_kb(void* p) { return static_cast<ref_handle*>(p)->bm->get_KB(); }

double ref_schedule(int schedule, float p0, float p1, uint64_t t) {
    float_vec_t kw(2, 0);
    kw[0] = p0; kw[1] = p1;
    switch (schedule) {
        case 0: return exponential_schedule(t, kw);
        case 1: return linear_schedule(t, kw);
        case 2: return logarithmic_schedule(t, kw);
        case 3: return constant_schedule(t, kw);
        default: return abrupt_cool_schedule(t, kw);
    }
}

// One MH step on vertex v at temperature T (reference src/metropolis_hasting.cc:42-62).
int ref_step(void* p, uint64_t v, double T) {
    auto* h = static_cast<ref_handle*>(p);
    return h->mh.step(*h->bm, v, T, h->engine) ? 1 : 0;
}

// dS and Hastings factor of moving v to s (reference src/metropolis_hasting.cc:103-192).
// accu_r is the member left behind by the call (stale on the cross-type early return).
void ref_transition(void* p, uint64_t v, uint64_t s, double* dS, double* accu_r) {
    auto* h = static_cast<ref_handle*>(p);
    size_t r = h->bm->get_memberships()->at(v);
    *dS = h->mh.dS_of(*h->bm, v, r, s);
    *accu_r = h->mh.accu_r();
}

double ref_entropy(void* p) { return static_cast<ref_handle*>(p)->bm->entropy(); }  // blockmodel.cc:753-787
double ref_entropy_accum(void* p) { return static_cast<ref_handle*>(p)->bm->get_entropy(); }
double ref_entropy_min(void* p) { return static_cast<ref_handle*>(p)->mh.entropy_min(); }

void ref_get_labels(void* p, uint32_t* out) {
    auto* h = static_cast<ref_handle*>(p);
    const uint_vec_t* mb = h->bm->get_memberships();
    for (size_t i = 0; i < mb->size(); ++i) out[i] = (*mb)[i];
}
void ref_get_vlist(void* p, uint32_t* out) {
    auto* h = static_cast<ref_handle*>(p);
    uint_vec_t& vl = h->bm->get_vlist();
    for (size_t i = 0; i < vl.size(); ++i) out[i] = vl[i];
}
void ref_get_m(void* p, int32_t* out) {  // K x K row-major, symmetric
    auto* h = static_cast<ref_handle*>(p);
    const int_mat_t* m = h->bm->get_m();
    size_t K = m->size();
    for (size_t r = 0; r < K; ++r)
        for (size_t s = 0; s < K; ++s) out[r * K + s] = (*m)[r][s];
}
void ref_get_m_r(void* p, int32_t* out) {
    auto* h = static_cast<ref_handle*>(p);
    const int_vec_t* v = h->bm->get_m_r();
    for (size_t i = 0; i < v->size(); ++i) out[i] = (*v)[i];
}
void ref_get_n_r(void* p, int32_t* out) {
    auto* h = static_cast<ref_handle*>(p);
    const int_vec_t* v = h->bm->get_n_r();
    for (size_t i = 0; i < v->size(); ++i) out[i] = (*v)[i];
}
uint64_t ref_eta_width(void* p) {
    auto* h = static_cast<ref_handle*>(p);
    return h->bm->get_eta_rk_()->at(0).size();
}
void ref_get_eta(void* p, uint32_t* out) {  // K x (max_degree+1) row-major
    auto* h = static_cast<ref_handle*>(p);
    const uint_mat_t* e = h->bm->get_eta_rk_();
    size_t W = e->at(0).size();
    for (size_t r = 0; r < e->size(); ++r)
        for (size_t k = 0; k < W; ++k) out[r * W + k] = (*e)[r][k];
}
void ref_get_k(void* p, uint64_t v, int32_t* out) {
    auto* h = static_cast<ref_handle*>(p);
    const int_vec_t* k = h->bm->get_k(v);
    for (size_t i = 0; i < k->size(); ++i) out[i] = (*k)[i];
}

// Number of 32-bit words drawn so far from `engine` and from `gen` (ORACLE_LOG_RNG builds;
// 0 otherwise).
void ref_rng_words(void* p, uint64_t* engine_words, uint64_t* gen_words) {
#ifdef ORACLE_LOG_RNG
    auto* h = static_cast<ref_handle*>(p);
    *engine_words = h->engine.words_drawn;
    *gen_words = h->bm->gen_ref().words_drawn;
#else
    (void)p;
    *engine_words = 0;
    *gen_words = 0;
#endif
}

// Known-answer hooks for the math tables (reference src/support/*.hh).
double ref_log_q(int n, int k) { return log_q<int>(n, k); }               // int_part.hh:27-37
double ref_log_q_approx(uint64_t n, uint64_t k) { return log_q_approx(n, k); }  // int_part.cc:88-98
double ref_lgamma_fast(uint64_t x) { return lgamma_fast(x); }             // cache.hh:82-93
double ref_safelog_fast(uint64_t x) { return safelog_fast(x); }           // cache.hh:46-57
double ref_spence(double x) { return spence(x); }                         // spence.cc:108-154
void ref_init_tables(uint64_t num_edges) {                                // blockmodel.cc:47-48
    init_cache(num_edges);
    init_q_cache(10000);
}

}  // extern "C"

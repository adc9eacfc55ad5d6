/* oracle/bisbm_oracle.h -- plain-C restatement of the reference's Metropolis-Hastings sweep.
 *
 * TEST INFRASTRUCTURE.  This is the CPU checker of the CUDA path, not a product path and not
 * a fallback: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may load
 * liboracle.so.  Parity status: PINNED -- tests/test_oracle_vs_reference.py checks every
 * function here against the unmodified reference compiled in place (oracle/_ref, see
 * oracle/Makefile) and against the committed fixtures under tests/golden/.
 *
 * Each function cites the reference file:line it restates (paths under /root/reference).
 */
#ifndef BISBM_ORACLE_H
#define BISBM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ora_chain ora_chain;

enum { ORA_EXPONENTIAL = 0, ORA_LINEAR = 1, ORA_LOGARITHMIC = 2, ORA_CONSTANT = 3, ORA_ABRUPT_COOL = 4 };

/* --- random primitives: libstdc++ 13 semantics --- */
typedef struct {
    uint32_t mt[624];
    int idx;
    uint64_t words; /* 32-bit words drawn so far */
} ora_mt19937;
void ora_mt_seed(ora_mt19937* g, uint32_t seed);
uint32_t ora_mt_next(ora_mt19937* g);
double ora_canon(ora_mt19937* g);
uint32_t ora_nd(ora_mt19937* g, uint32_t range);
void ora_shuffle(uint32_t* x, uint64_t n, ora_mt19937* g);
uint32_t ora_categorical(const int32_t* w, uint32_t k, ora_mt19937* g);

/* --- math tables (reference src/support) --- */
double ora_spence(double x);
double ora_log_q_approx(uint64_t n, uint64_t k);
/* exact table value for 1 <= n <= n_max (<= 10000), 1 <= k <= min(n, k_max); -inf where the
 * reference leaves its table untouched */
double* ora_build_log_q_table(uint32_t n_max, uint32_t k_max);
void ora_free(void* p);
double ora_schedule(int schedule, float p0, float p1, uint64_t t);

/* --- chain --- */
ora_chain* ora_create(uint32_t n, uint32_t na, uint32_t nb, uint64_t n_edges, const uint32_t* ea,
                      const uint32_t* eb, const uint32_t* labels, uint32_t ka, uint32_t kb, double eps,
                      uint32_t engine_seed, uint32_t gen_seed);
void ora_destroy(ora_chain* c);
void ora_init(ora_chain* c, int randomize);
double ora_anneal(ora_chain* c, int schedule, float p0, float p1, uint64_t duration, uint64_t steps_await);
/* anneal() with the type-alternating visiting order of the GPU's parallel mode (test aid, not a reference function) */
double ora_anneal_alternating(ora_chain* c, int schedule, float p0, float p1, uint64_t duration, uint64_t steps_await);
int ora_step(ora_chain* c, uint32_t v, double T);
void ora_transition(ora_chain* c, uint32_t v, uint32_t s, double* dS, double* accu_r);
double ora_log_q(const ora_chain* c, int n, int k);
double ora_entropy(const ora_chain* c);
double ora_entropy_accum(const ora_chain* c);
uint64_t ora_sweeps_done(const ora_chain* c);
void ora_get_labels(const ora_chain* c, uint32_t* out);
void ora_get_vlist(const ora_chain* c, uint32_t* out);
void ora_get_m(const ora_chain* c, int32_t* out);   /* K x K */
void ora_get_m_r(const ora_chain* c, int32_t* out); /* K */
void ora_get_n_r(const ora_chain* c, int32_t* out); /* K */
uint32_t ora_eta_width(const ora_chain* c);
void ora_get_eta(const ora_chain* c, uint32_t* out); /* K x (max_degree+1) */
void ora_get_k(const ora_chain* c, uint32_t v, int32_t* out);
void ora_rng_words(const ora_chain* c, uint64_t* engine_words, uint64_t* gen_words);

#ifdef __cplusplus
}
#endif
#endif

// devmath.cuh -- device-side arithmetic shared by the replay and the parallel kernels.
//
//  * strict IEEE helpers (no FMA contraction) for the bit-exact replay path;
//  * std::mt19937 and the libstdc++ 13 distribution transforms the reference draws through
//    (reference call sites: src/metropolis_hasting.cc:57,80; src/blockmodel.cc:617-628,673-674);
//  * the partition-count function log q(n,k) of src/support/int_part.{hh,cc} and the Cephes
//    dilogarithm it needs (src/support/spence.cc);
//  * Philox4x32-10 and a Feistel index permutation for the parallel kernels.
//
// Everything is BISBM_HD so the same text also compiles for the host-side emulation
// harness in tests/emul (a debugging aid; never part of libbisbm.so's execution path).
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define BISBM_HD __host__ __device__ __forceinline__
#define BISBM_D __device__ __forceinline__
#define BISBM_NOINLINE_HD __host__ __device__ __noinline__
#else
#define BISBM_HD inline
#define BISBM_D inline
#define BISBM_NOINLINE_HD
#endif

namespace bisbm {

// ---- strict double arithmetic (replay): every operation rounds separately ---------------
BISBM_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
BISBM_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    volatile double r = a - b; return r;
#endif
}
BISBM_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
BISBM_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b; return r;
#endif
}

#define BISBM_INF (__builtin_huge_val())
#define BISBM_PI 3.14159265358979323846

// ---- std::mt19937 -----------------------------------------------------------------------
// state layout: s[0..623] = words, s[624] = index, s[625..626] = 64-bit draw counter
enum { MT_N = 624, MT_STATE_WORDS = 628 };

BISBM_HD void mt_seed(uint32_t* s, uint32_t seed) {
    s[0] = seed;
    for (int i = 1; i < MT_N; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
    s[624] = MT_N;
    s[625] = 0; s[626] = 0; s[627] = 0;
}

BISBM_HD uint32_t mt_next(uint32_t* s) {
    uint32_t idx = s[624];
    if (idx >= MT_N) {
        for (int i = 0; i < MT_N; ++i) {
            int i1 = (i + 1 == MT_N) ? 0 : i + 1;
            int im = (i + 397 >= MT_N) ? i + 397 - MT_N : i + 397;
            uint32_t y = (s[i] & 0x80000000u) | (s[i1] & 0x7fffffffu);
            s[i] = s[im] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx = 0;
    }
    uint32_t y = s[idx];
    s[624] = idx + 1;
    if (++s[625] == 0) ++s[626];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

// std::generate_canonical<double,53> on a 32-bit engine: two words, low first
// (libstdc++ bits/random.tcc:3349-3381)
BISBM_HD double mt_canon(uint32_t* s) {
    double w0 = (double)mt_next(s);
    double w1 = (double)mt_next(s);
    double r = ddiv(dadd(w0, dmul(w1, 4294967296.0)), 18446744073709551616.0);
    if (r >= 1.0) r = 0.99999999999999988897769753748;  // nextafter(1, 0)
    return r;
}

// uniform_int_distribution::_S_nd (Lemire) for range < 2^32 (bits/uniform_int_dist.h:257-281)
BISBM_HD uint32_t mt_nd(uint32_t* s, uint32_t range) {
    uint64_t prod = (uint64_t)mt_next(s) * (uint64_t)range;
    uint32_t low = (uint32_t)prod;
    if (low < range) {
        uint32_t thr = (uint32_t)(0u - range) % range;
        while (low < thr) {
            prod = (uint64_t)mt_next(s) * (uint64_t)range;
            low = (uint32_t)prod;
        }
    }
    return (uint32_t)(prod >> 32);
}

BISBM_HD uint64_t mt_uid(uint32_t* s, uint64_t hi) {  // uniform_int_distribution<size_t>(0, hi)
    if (hi == 0xffffffffull) return mt_next(s);
    return mt_nd(s, (uint32_t)(hi + 1));
}

// std::shuffle (bits/stl_algo.h:3742-3805) over x[i*stride], i < n
BISBM_HD void mt_shuffle(uint32_t* x, uint64_t n, uint64_t stride, uint32_t* s) {
    if (n == 0) return;
    uint32_t tmp;
    if (0xffffffffull / n >= n) {
        uint64_t i = 1;
        if ((n % 2) == 0) {
            uint64_t j = mt_uid(s, 1);
            tmp = x[i * stride]; x[i * stride] = x[j * stride]; x[j * stride] = tmp;
            ++i;
        }
        while (i != n) {
            uint64_t b1 = i + 2;
            uint64_t v = mt_uid(s, (i + 1) * b1 - 1);
            uint64_t q0 = v / b1, q1 = v % b1;
            tmp = x[i * stride]; x[i * stride] = x[q0 * stride]; x[q0 * stride] = tmp;
            ++i;
            tmp = x[i * stride]; x[i * stride] = x[q1 * stride]; x[q1 * stride] = tmp;
            ++i;
        }
        return;
    }
    for (uint64_t i = 1; i < n; ++i) {
        uint64_t j = mt_uid(s, i);
        tmp = x[i * stride]; x[i * stride] = x[j * stride]; x[j * stride] = tmp;
    }
}

// ---- Cephes dilogarithm (src/support/spence.cc:91-154); published Cephes coefficients ----
BISBM_HD double spence(double x) {
    if (x < 0.0) return NAN;
    if (x == 1.0) return 0.0;
    if (x == 0.0) return ddiv(dmul(BISBM_PI, BISBM_PI), 6.0);
    int flag = 0;
    double w;
    if (x > 2.0) { x = ddiv(1.0, x); flag |= 2; }
    if (x > 1.5) { w = dsub(ddiv(1.0, x), 1.0); flag |= 2; }
    else if (x < 0.5) { w = -x; flag |= 1; }
    else w = dsub(x, 1.0);
    const double kSpA[8] = {4.65128586073990045278E-5, 7.31589045238094711071E-3, 1.33847639578309018650E-1,
                            8.79691311754530315341E-1, 2.71149851196553469920E0,  4.25697156008121755724E0,
                            3.29771340985225106936E0,  1.00000000000000000126E0};
    const double kSpB[8] = {6.90990488912553276999E-4, 2.54043763932544379113E-2, 2.82974860602568089943E-1,
                            1.41172597751831069617E0,  3.63800533345137075418E0,  5.03278880143316990390E0,
                            3.54771340985225096217E0,  9.99999999999999998740E-1};
    double pa = kSpA[0], pb = kSpB[0];
#pragma unroll
    for (int i = 1; i <= 7; ++i) { pa = dadd(dmul(pa, w), kSpA[i]); pb = dadd(dmul(pb, w), kSpB[i]); }
    double y = ddiv(dmul(-w, pa), pb);
    if (flag & 1) y = dsub(dsub(ddiv(dmul(BISBM_PI, BISBM_PI), 6.0), dmul(log(x), log1p(-x))), y);
    if (flag & 2) { double z = log(x); y = dsub(dmul(dmul(-0.5, z), z), y); }
    return y;
}

// Host-built exact tables (glibc values, uploaded once per handle).
struct Tables {
    const double* lg;    // lg[i] = lgamma(double(i)), lg[0] = +inf; i < lg_n   (src/support/cache.cc:64-79)
    uint64_t lg_n;
    const double* qtab;  // exact log q(n,k) table, row-major [qn+1][qk+1]     (src/support/int_part.cc:34-51)
    uint32_t qn, qk;
    // the asymptotic formula of log_q_approx is lf(u) - log(n) + sqrt(n) g(u) with u = k / sqrt(n): g and lf tabulated by
    // the host (the reference's own iteration, glibc) at u_i = exp(gl_t0 + i / gl_inv_h); see GlNode
    const struct GlNode* gl;
    double gl_t0, gl_inv_h;
    uint32_t gl_n;
};

// One node of the g(u), lf(u) table.  The reference stops get_v's fixed-point iteration at |v_N - v_(N-1)| <= 1e-8, so its
// g and lf are smooth only between the values of u where the iteration count N changes (steps of up to ~1e-6 in log q there).
// The table therefore interpolates inside one RUN of nodes with equal N: [first, last] are the run's node indices, `cross` is
// the u where the count switches between this node and the next (0 when they are in the same run), `flags` bit 0 marks an
// interval where the host found the count changing back and forth (the device then iterates like the reference).
struct GlNode {
    double g, lf, cross;
    uint32_t first, last;
    uint32_t flags, pad;
};

// lgamma_fast (src/support/cache.hh:82-93): table value when in range, else libm lgamma
BISBM_HD double lgamma_int(const Tables& tb, int64_t i) {
    if (i <= 0) return BISBM_INF;
    if ((uint64_t)i < tb.lg_n) return tb.lg[i];
    return lgamma((double)i);
}

// the u-dependent part of log_q_approx's large-k branch (src/support/int_part.cc:77-97): v from get_v's fixed-point iteration,
// then lf (without the - log n) and g; the reference's operation order
BISBM_HD int logq_gl(double u, double* g_out, double* lf_out) {   // returns the number of iterations get_v took
    double v = u, delta = 1.0;
    int guard = 0, iters = 0;
    while (delta > 1e-8 && guard++ < 10000) {  // get_v (int_part.cc:77-86)
        double nv = dmul(u, sqrt(spence(exp(-v))));
        delta = fabs(dsub(nv, v));
        v = nv;
        ++iters;
    }
    double emv = exp(-v);
    double t1 = ddiv(log1p(dmul(-emv, dadd(1.0, ddiv(dmul(u, u), 2.0)))), 2.0);
    *lf_out = dsub(dsub(dsub(dsub(log(v), t1), ddiv(dmul(log(2.0), 3.0), 2.0)), log(u)), log(BISBM_PI));
    *g_out = dsub(ddiv(dmul(2.0, v), u), dmul(u, log1p(-emv)));
    return iters;
}


// log_q_approx (src/support/int_part.cc:73-98).  Branch test k < n^(1/4) done in exact
// integer arithmetic (k^4 < n), which agrees with the reference's pow() for every n < 2^53.
BISBM_HD double log_q_approx(const Tables& tb, uint64_t n, uint64_t k) {
    bool small = (k < 65536ull) && (k * k * k * k < n);
    if (small) {
        // lbinom_fast(n-1, k-1) - lgamma_fast(k+1)   (src/support/util.hh:41-47)
        uint64_t N = n - 1, kk = k - 1;
        double lb = 0.0;
        if (!(N == 0 || kk == 0 || kk > N))
            lb = dsub(dsub(lgamma_int(tb, (int64_t)N + 1), lgamma_int(tb, (int64_t)kk + 1)), lgamma_int(tb, (int64_t)(N - kk) + 1));
        return dsub(lb, lgamma_int(tb, (int64_t)k + 1));
    }
    double sn = sqrt((double)n);
    double u = ddiv((double)k, sn);
    double g, lf;
    logq_gl(u, &g, &lf);
    return dadd(dsub(lf, log((double)n)), dmul(sn, g));
}

// 6-point Lagrange weights for nodes 0 .. 5 and a position p (inside [0, 5], or just outside)
BISBM_HD void lagrange6(double p, double* w) {
    const double a = p, b = p - 1.0, c = p - 2.0, d = p - 3.0, e = p - 4.0, g = p - 5.0;
    w[0] = b * c * d * e * g * (-1.0 / 120.0);
    w[1] = a * c * d * e * g * (1.0 / 24.0);
    w[2] = a * b * d * e * g * (-1.0 / 12.0);
    w[3] = a * b * c * e * g * (1.0 / 12.0);
    w[4] = a * b * c * d * g * (-1.0 / 24.0);
    w[5] = a * b * c * d * e * (1.0 / 120.0);
}

// The same value as log_q_approx's large-k branch from the tabulated g(u), lf(u): 6-point Lagrange in log u (node spacing
// 2^-10: interpolation error ~1e-17, rounding a few ulp of g) over nodes of ONE run of equal iteration count -- ~80 flops and
// 6 table rows instead of 2 .. 23 iterations of exp / spence / sqrt.  For the parallel sweep's blocks that have no valid
// expansion (small, or drifted out of its range).  Outside the table, in runs shorter than the stencil and in the intervals
// the host flagged, it evaluates the formula itself.
BISBM_HD double log_q_approx_tab(const Tables& tb, uint64_t n, uint64_t k) {
    const bool small = (k < 65536ull) && (k * k * k * k < n);
    if (small || tb.gl == nullptr) return log_q_approx(tb, n, k);
    const double sn = sqrt((double)n);
    const double u = ddiv((double)k, sn);
    const double x = (log(u) - tb.gl_t0) * tb.gl_inv_h;
    if (!(x >= 0.0) || !(x < (double)tb.gl_n - 1.0)) return log_q_approx(tb, n, k);
    const uint32_t i = (uint32_t)x;
    const GlNode a = tb.gl[i];
    if (a.flags & 1u) return log_q_approx(tb, n, k);
    uint32_t first = a.first, last = a.last;
    if (a.cross != 0.0 && !(u < a.cross)) { const GlNode b = tb.gl[i + 1]; first = b.first; last = b.last; }
    if (last - first < 5u) return log_q_approx(tb, n, k);
    uint32_t s0 = i >= first + 2u ? i - 2u : first;
    if (s0 + 5u > last) s0 = last - 5u;
    double w[6];
    lagrange6(x - (double)s0, w);
    const GlNode* p = tb.gl + s0;
    double g = 0.0, lf = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) { g = fma(w[j], p[j].g, g); lf = fma(w[j], p[j].lf, lf); }
    return (lf - log((double)n)) + sn * g;
}
BISBM_HD double log_q_tab(const Tables& tb, int n, int k) {     // log_q with the tabulated asymptotic branch
    if (n <= 0 || k < 1) return 0.0;
    if (k > n) k = n;
    if (n < 10001) {
        if ((uint32_t)n <= tb.qn && (uint32_t)k <= tb.qk) return tb.qtab[(size_t)n * (tb.qk + 1) + k];
        return NAN;
    }
    return log_q_approx_tab(tb, (uint64_t)n, (uint64_t)k);
}

// log_q<int> (src/support/int_part.hh:27-37).  qtab is the host-built exact table
// (rows n <= qn, columns k <= qk, row-major [qn+1][qk+1]); it covers every (n,k) this graph
// can reach with n < 10001.
BISBM_HD double log_q(const Tables& tb, int n, int k) {
    if (n <= 0 || k < 1) return 0.0;
    if (k > n) k = n;
    if (n < 10001) {
        if ((uint32_t)n <= tb.qn && (uint32_t)k <= tb.qk) return tb.qtab[(size_t)n * (tb.qk + 1) + k];
        return NAN;  // outside the table the host sized for this graph: cannot happen
    }
    return log_q_approx(tb, (uint64_t)n, (uint64_t)k);
}

// ---- Philox4x32-10 (Salmon et al. 2011) ---------------------------------------------------
struct u32x4 { uint32_t x, y, z, w; };

BISBM_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

BISBM_HD u32x4 philox4x32(u32x4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        u32x4 n;
        n.x = hi1 ^ c.y ^ k0; n.y = lo1; n.z = hi0 ^ c.w ^ k1; n.w = lo0;
        c = n;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c;
}

BISBM_HD double u53(uint32_t lo, uint32_t hi) {  // uniform in [0,1) with 53 random bits
    uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)x * (1.0 / 9007199254740992.0);
}

// ---- Feistel permutation of [0, n) (cycle walking over the enclosing 2^(2b) domain) --------
BISBM_HD uint32_t feistel_round(uint32_t x, uint32_t key) {
    x ^= key;
    x *= 0x9E3779B1u; x ^= x >> 15;
    x *= 0x85EBCA77u; x ^= x >> 13;
    x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}

BISBM_HD uint32_t feistel_perm(uint32_t i, uint32_t n, uint32_t half_bits, uint64_t key) {
    if (n <= 1) return 0;  // degenerate
    uint32_t mask = (1u << half_bits) - 1u;
    uint32_t x = i;
    do {
        uint32_t l = x >> half_bits, r = x & mask;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t f = feistel_round(r, (uint32_t)(key >> (16 * k)) ^ (0xA511E9B3u * (k + 1)) ^ (uint32_t)(key >> 32)) & mask;
            uint32_t nl = r; r = l ^ f; l = nl;
        }
        x = (l << half_bits) | r;
    } while (x >= n);
    return x;
}

BISBM_HD uint32_t feistel_half_bits(uint32_t n) {
    uint32_t b = 1;
    while (b < 16 && (1ull << (2 * b)) < (uint64_t)n) ++b;
    return b;
}

}  // namespace bisbm

// gltable.h -- host side of devmath.cuh's log_q_approx_tab: the table of g(u), lf(u) (the u-dependent parts of the reference's
// asymptotic log q, src/support/int_part.cc:73-98) at u_i = exp(t0 + i / inv_h), built with the reference's own fixed-point
// iteration (logq_gl) and glibc, organised in runs of equal iteration count (see GlNode).
#pragma once
#include <cmath>
#include <vector>
#include "devmath.cuh"

namespace bisbm {

inline void build_gl_table(std::vector<GlNode>& tab, double t0, double inv_h, double t1) {
    const size_t n = (size_t)std::ceil((t1 - t0) * inv_h) + 1;
    tab.assign(n, GlNode());
    std::vector<int> iters(n);
    auto u_of = [&](double x) { return std::exp(t0 + x / inv_h); };
    for (size_t i = 0; i < n; ++i) {
        iters[i] = logq_gl(u_of((double)i), &tab[i].g, &tab[i].lf);
        tab[i].cross = 0.0; tab[i].flags = 0; tab[i].pad = 0;
    }
    // runs of equal iteration count
    for (size_t i = 0; i < n;) {
        size_t j = i;
        while (j + 1 < n && iters[j + 1] == iters[i]) ++j;
        for (size_t k = i; k <= j; ++k) { tab[k].first = (uint32_t)i; tab[k].last = (uint32_t)j; }
        i = j + 1;
    }
    double g, lf;
    for (size_t i = 0; i + 1 < n; ++i) {
        const double ua = u_of((double)i), ub = u_of((double)i + 1.0);
        if (iters[i] != iters[i + 1]) {
            // the count switches once between the two nodes: bisect for the first u that takes node i+1's count ...
            double lo = ua, hi = ub;
            for (int it = 0; it < 80 && lo < hi; ++it) {
                const double mid = 0.5 * (lo + hi);
                if (mid <= lo || mid >= hi) break;
                if (logq_gl(mid, &g, &lf) == iters[i]) lo = mid; else hi = mid;
            }
            tab[i].cross = hi;
            // ... and make sure it is a single switch: probes on both sides must carry the counts of their nodes
            for (int p = 1; p < 8; ++p) {
                const double up = ua + (hi - ua) * p / 8.0, uq = hi + (ub - hi) * p / 8.0;
                if (up > ua && up < lo && logq_gl(up, &g, &lf) != iters[i]) tab[i].flags |= 1u;
                if (uq > hi && uq < ub && logq_gl(uq, &g, &lf) != iters[i + 1]) tab[i].flags |= 1u;
            }
        } else {
            for (int p = 1; p < 4; ++p)
                if (logq_gl(ua + (ub - ua) * p / 4.0, &g, &lf) != iters[i]) tab[i].flags |= 1u;
        }
    }
}

}  // namespace bisbm

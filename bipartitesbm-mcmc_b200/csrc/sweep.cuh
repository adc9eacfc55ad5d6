// sweep.cuh -- parallel mode: many independent chains per launch.
//
// Mapping: LANE = CHAIN.  A warp takes one vertex v and evaluates v's move for 32 chains at
// once; labels are chain-minor (labels[v][C]) so "neighbour j's label for 32 chains" is one
// coalesced 128-byte load and the CSR row of v is read once for all of them.  A half sweep
// moves only the vertices of one type: their neighbours (all of the other type) are frozen,
// so every neighbour-block histogram is exact and the committed m_rs / e_r / n_r / eta
// deltas keep the counts exactly consistent with the labels; concurrent moves of one chain
// couple only through slightly stale count READS (bounded by `tiles`, the number of warps
// working on one chain group).  Counts live in HBM/L2 and are committed with atomics.
//
// Per move this is the same arithmetic as the reference's step():
//   proposal   single_vertex_change      reference src/blockmodel.cc:613-637
//   dS, accu_r transition_ratio          reference src/metropolis_hasting.cc:103-192
//   accept     step                      reference src/metropolis_hasting.cc:42-62
//   commit     apply_mcmc_moves          reference src/blockmodel.cc:461-503
// with the lgamma differences written as log-products / Stirling differences (no 2E-entry
// table in HBM) and log q(n,k) from the exact table (n < 10001), a per-block second-order
// expansion refreshed every half sweep (large blocks), or the full formula.
#pragma once
#include "state.cuh"

namespace bisbm {

// second-order expansion of f(e,n) = log q(e,n) about (e0,n0) for one (chain, block slot)
struct LogqExp {
    int32_t e0, n0;
    float fe, fn, fee, fen, fnn;
    uint32_t valid;
};

struct SweepParams {
    GraphView g;
    StateView s;
    Tables tb;
    const uint64_t* seeds;           // [C]
    const uint8_t* active;           // [C]
    unsigned long long* accepted;    // [C]
    double* dS_accum;                // [C]
    const LogqExp* lq;               // [C][KA+KB]
    uint32_t n_chains;               // real chains (<= C)
    uint32_t type;                   // 0: move type-a vertices, 1: type-b
    uint32_t tiles;                  // warps per chain group
    uint32_t n_groups;               // C / 32
    uint32_t half_bits;              // Feistel domain of this half sweep
    uint32_t hist_stride;            // max blocks of the opposite type
    uint64_t sweep;                  // sweep index: RNG counter and permutation key
    uint64_t step_base;              // global step index of the first move of this half sweep
    int schedule;
    float p0, p1;
};

// temperature of global step t (same five schedules as src/metropolis_hasting.cc:10-37,
// device libm for pow / log)
BISBM_NOINLINE_HD static double par_temperature(int schedule, float p0, float p1, uint64_t t) {
    switch (schedule) {
        case 0: return (double)p0 * pow((double)p1, (double)t);
        case 1: return (double)(p0 - p1 * (float)t);
        case 2: {
            uint64_t i = (uint64_t)((float)t + p1);
            double l = (i == 0) ? 0.0 : log((double)i);
            return (double)p0 / l;
        }
        case 3: return (double)p0;
        default: return ((float)t < p0) ? 1.0 : 0.0;
    }
}

// lgamma(x + d) - lgamma(x) for integers x >= 1, d >= 0, without the 2E-entry table:
// Stirling difference with log1p for large x, direct lgamma for small x.  Kept out of line:
// the sweep kernel must stay small enough for the instruction cache.
BISBM_NOINLINE_HD static double lgamma_diff(double x, double d) {
    if (d == 0.0) return 0.0;
    if (x < 32.0) return lgamma(x + d) - lgamma(x);
    double y = x + d;
    double ix = 1.0 / x, iy = 1.0 / y;
    double ser = (iy - ix) * (1.0 / 12.0) - (iy * iy * iy - ix * ix * ix) * (1.0 / 360.0) +
                 (iy * iy * iy * iy * iy - ix * ix * ix * ix * ix) * (1.0 / 1260.0);
    return (x - 0.5) * log1p(d * ix) + d * (log(y) - 1.0) + ser;
}

// exact (table / full asymptotic formula) difference: the rare slow path, out of line
BISBM_NOINLINE_HD static double logq_delta_exact(const Tables& tb, int e, int n, int de, int dn) {
    return log_q(tb, e + de, n + dn) - log_q(tb, e, n);
}

// log q difference f(e+de, n+dn) - f(e, n) for one block
BISBM_HD double logq_delta(const Tables& tb, const LogqExp& q, int e, int n, int de, int dn) {
    int x = e - q.e0, y = n - q.n0;
    int ax = x < 0 ? -x : x, ay = y < 0 ? -y : y, ad = de < 0 ? -de : de;
    if (q.valid && ax <= (q.e0 >> 4) && ay <= (q.n0 >> 4) && ad <= (q.e0 >> 4)) {
        double dx = (double)x, dy = (double)y, De = (double)de, Dn = (double)dn;
        return (double)q.fe * De + (double)q.fn * Dn + 0.5 * (double)q.fee * (De * De + 2.0 * dx * De) +
               (double)q.fen * (dx * Dn + dy * De + De * Dn) + 0.5 * (double)q.fnn * (Dn * Dn + 2.0 * dy * Dn);
    }
    return logq_delta_exact(tb, e, n, de, dn);
}

#ifdef __CUDACC__

// Block counts are updated with L2 atomics by every SM; L1 is not coherent, so every count
// READ goes to L2 (ld.global.cg).
__device__ __forceinline__ int ldc(const int32_t* p) { return __ldcg(p); }

// Refresh the log q expansions of the blocks of one type (they only change during that
// type's half sweep).  One thread per (chain, block).
__global__ void logq_refresh_kernel(StateView s, Tables tb, LogqExp* lq, uint32_t n_chains, uint32_t type) {
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t kmax = type ? s.KB : s.KA;
    if (idx >= n_chains * kmax) return;
    uint32_t c = idx / kmax, b = idx % kmax;
    uint32_t kc = type ? s.kb[c] : s.ka[c];
    uint32_t slot = (type ? s.KA : 0) + b;
    LogqExp q;
    q.e0 = 0; q.n0 = 0; q.fe = q.fn = q.fee = q.fen = q.fnn = 0.f; q.valid = 0;
    if (b < kc) {
        int e0 = s.e[(size_t)c * (s.KA + s.KB) + slot];
        int n0 = s.nr[(size_t)c * (s.KA + s.KB) + slot];
        q.e0 = e0; q.n0 = n0;
        if (e0 >= 16384 && n0 >= 1024 && 2 * (int64_t)n0 <= (int64_t)e0) {
            int he = e0 >> 10, hn = n0 >> 8;
            double f00 = log_q_approx(tb, e0, n0);
            double fp0 = log_q_approx(tb, e0 + he, n0), fm0 = log_q_approx(tb, e0 - he, n0);
            double f0p = log_q_approx(tb, e0, n0 + hn), f0m = log_q_approx(tb, e0, n0 - hn);
            double fpp = log_q_approx(tb, e0 + he, n0 + hn), fpm = log_q_approx(tb, e0 + he, n0 - hn);
            double fmp = log_q_approx(tb, e0 - he, n0 + hn), fmm = log_q_approx(tb, e0 - he, n0 - hn);
            double He = (double)he, Hn = (double)hn;
            q.fe = (float)((fp0 - fm0) / (2.0 * He));
            q.fn = (float)((f0p - f0m) / (2.0 * Hn));
            q.fee = (float)((fp0 - 2.0 * f00 + fm0) / (He * He));
            q.fnn = (float)((f0p - 2.0 * f00 + f0m) / (Hn * Hn));
            q.fen = (float)((fpp - fpm - fmp + fmm) / (4.0 * He * Hn));
            q.valid = 1;
        }
    }
    lq[(size_t)c * (s.KA + s.KB) + slot] = q;
}

template <typename HistT>
__global__ void __launch_bounds__(512) sweep_kernel(SweepParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
    HistT* hist = reinterpret_cast<HistT*>(smem_raw) + (size_t)warp * P.hist_stride * 32 + lane;

    const uint32_t group = blockIdx.x % P.n_groups;
    const uint32_t tile = (blockIdx.x / P.n_groups) * wpc + warp;
    if (tile >= P.tiles) return;
    const uint32_t c = group * 32 + lane;
    const bool live = (c < P.n_chains) && P.active[c];

    const GraphView& G = P.g;
    const uint32_t type = P.type;
    const uint32_t v0 = type ? G.na : 0, nv = type ? G.nb : G.na;
    const uint32_t C = P.s.C, KB = P.s.KB, KA = P.s.KA, W = P.s.W;
    const uint32_t cc = live ? c : 0;
    const uint32_t ka = P.s.ka[cc], kb = P.s.kb[cc], K = ka + kb;
    const uint32_t kown = type ? kb : ka, kopp = type ? ka : kb;
    int32_t* const M = P.s.m + (size_t)cc * KA * KB;
    int32_t* const E = P.s.e + (size_t)cc * (KA + KB);
    int32_t* const NR = P.s.nr + (size_t)cc * (KA + KB);
    int32_t* const ETA = P.s.eta + (size_t)cc * (KA + KB) * W;
    const LogqExp* const LQ = P.lq + (size_t)cc * (KA + KB);
    const uint32_t own_off = type ? KA : 0, opp_off = type ? 0 : KA;
    // m(x_own, t_opp) = M[x*sx + t*st]
    const uint32_t sx = type ? 1 : KB, st = type ? KB : 1;
    int32_t* const LAB = P.s.labels + cc;
    const double eps = P.s.eps, epsK = eps * (double)K;
    const uint64_t seed = P.seeds[cc];
    const uint32_t key0 = (uint32_t)seed, key1 = (uint32_t)(seed >> 32);
    const uint64_t pkey = (P.sweep * 2 + type) * 0x9E3779B97F4A7C15ull + (uint64_t)group * 0xD1B54A32D192ED03ull;

    unsigned long long n_acc = 0;
    double ds_sum = 0.0;

    for (uint32_t i = tile; i < nv; i += P.tiles) {
        // The trip count is warp-uniform: reconverge all 32 lanes (= chains) at every vertex so
        // the neighbour gathers below stay coalesced 128-byte loads.
        __syncwarp();
        const uint32_t v = v0 + feistel_perm(i, nv, P.half_bits, pkey);
        const uint32_t row = G.row_ptr[v];
        const uint32_t d = G.row_ptr[v + 1] - row;
        const double T = (P.schedule == 3) ? (double)P.p0 : par_temperature(P.schedule, P.p0, P.p1, P.step_base + i);
        const uint32_t r = live ? (uint32_t)LAB[(size_t)v * C] : 0u;

        // ---- proposal (single_vertex_change) ----
        u32x4 ctr; ctr.x = v; ctr.y = (uint32_t)P.sweep; ctr.z = (uint32_t)(P.sweep >> 32); ctr.w = 0;
        const u32x4 ra = philox4x32(ctr, key0, key1);
        uint32_t s = r;        // own-type local index of the target block
        bool cross = false;    // proposal fell on a block of the other type
        if (live && kown != 1) {
            bool uniform_pick = (d == 0);
            uint32_t t = 0;
            int e_t = 0;
            if (d != 0) {
                const uint32_t j = G.col[row + mulhi32(ra.x, d)];
                t = (uint32_t)LAB[(size_t)j * C];
                e_t = ldc(&E[opp_off + t]);
                const double R = epsK / ((double)e_t + epsK);
                uniform_pick = ((double)ra.y * (1.0 / 4294967296.0)) < R;
            }
            if (uniform_pick) {
                const uint32_t sg = mulhi32(ra.z, K);  // uniform over ALL K blocks (either type)
                const bool sg_a = sg < ka;
                cross = (sg_a != (type == 0));
                s = sg_a ? sg : sg - ka;
            } else {
                // categorical over row m[t][.]: block x of the own type w.p. m(x,t)/e_t
                const uint64_t z = (uint64_t)(u53(ra.z, ra.w) * (double)e_t);
                uint64_t cum = 0;
                s = kown - 1;
                const int32_t* col_t = M + (size_t)t * st;
                bool found = false;
                for (uint32_t x = 0; x < kown; ++x) {   // no early exit: keeps the lanes in step
                    cum += (uint32_t)ldc(&col_t[(size_t)x * sx]);
                    if (!found && cum > z) { s = x; found = true; }
                }
            }
        }
        // dS = +inf for a cross-type target (rejected); dS = 0, accu_r = 1 for s == r: at T > 0
        // accepted unless the block would empty, at T == 0 the reference requires dS < 0
        // (src/metropolis_hasting.cc:47-52)
        const bool eval = live && !cross && (s != r);
        if (live && !cross && s == r) {
            if (T != 0.0 && ldc(&NR[own_off + r]) != 1) ++n_acc;
        }
        __syncwarp();
        if (!__any_sync(0xffffffffu, eval)) continue;

        // ---- neighbour-block histogram (exact: neighbours are frozen in this half sweep) ----
        if (eval) {
            for (uint32_t t = 0; t < kopp; ++t) hist[t * 32] = 0;
#pragma unroll 4
            for (uint32_t e = 0; e < d; ++e) {
                const uint32_t nb = G.col[row + e];
                const uint32_t t = (uint32_t)LAB[(size_t)nb * C];
                hist[t * 32] += 1;
            }
        }
        __syncwarp();

        // ---- dS and Hastings factor (transition_ratio) ----
        bool go = false;
        double dS = 0.0;
        const uint32_t didx = G.degidx[v];
        if (eval) {
            const int32_t* Mr = M + (size_t)r * sx;
            const int32_t* Ms = M + (size_t)s * sx;
            double a0 = 0.0, a1 = 0.0, ratio = 1.0, logacc = 0.0;
            for (uint32_t t = 0; t < kopp; ++t) {
                const int kk = (int)hist[t * 32];
                if (kk == 0) continue;
                const int m_r = ldc(&Mr[(size_t)t * st]), m_s = ldc(&Ms[(size_t)t * st]);
                const double inv = 1.0 / ((double)ldc(&E[opp_off + t]) + epsK);
                a0 += (double)kk * ((double)m_s + eps) * inv;
                a1 += (double)kk * ((double)(m_r - kk) + eps) * inv;
                if (kk <= 8) {
                    double num = 1.0, den = 1.0;
                    for (int q = 0; q < kk; ++q) { num *= (double)(m_r - q); den *= (double)(m_s + 1 + q); }
                    ratio *= num / den;
                    if (ratio > 1e100 || ratio < 1e-100) { logacc += log(ratio); ratio = 1.0; }
                } else {
                    logacc += lgamma_diff((double)(m_r - kk + 1), (double)kk) - lgamma_diff((double)(m_s + 1), (double)kk);
                }
            }
            const int e_r = ldc(&E[own_off + r]), e_s = ldc(&E[own_off + s]);
            const int n_r = ldc(&NR[own_off + r]), n_s = ldc(&NR[own_off + s]);
            const int eta_r = ldc(&ETA[(size_t)(own_off + r) * W + didx]), eta_s = ldc(&ETA[(size_t)(own_off + s) * W + didx]);
            ratio *= (double)(eta_r > 0 ? eta_r : 1) / (double)(eta_s + 1);
            dS = logacc + log(ratio);
            dS += lgamma_diff((double)(e_s + 1), (double)d) - lgamma_diff((double)(e_r - (int)d + 1), (double)d);
            dS += logq_delta(P.tb, LQ[own_off + r], e_r, n_r, -(int)d, -1);
            dS += logq_delta(P.tb, LQ[own_off + s], e_s, n_s, (int)d, 1);

            // ---- accept (step) ----
            if (T == 0.0) {
                go = dS < 0.0;
            } else {
                const double a = -dS / T + ((d == 0) ? 0.0 : log(a1 / a0));
                if (a > 0.0) go = true;
                else {
                    ctr.w = 1;
                    const u32x4 rb = philox4x32(ctr, key0, key1);
                    go = u53(rb.x, rb.y) < exp(a);
                }
            }
        }
        __syncwarp();

        // ---- commit (apply_mcmc_moves) ----
        if (go) {
            const int old = atomicSub(&NR[own_off + r], 1);
            if (old <= 1) {
                atomicAdd(&NR[own_off + r], 1);  // would empty block r: vetoed
            } else {
                atomicAdd(&NR[own_off + s], 1);
                atomicSub(&ETA[(size_t)(own_off + r) * W + didx], 1);
                atomicAdd(&ETA[(size_t)(own_off + s) * W + didx], 1);
                int32_t* Mrw = M + (size_t)r * sx;
                int32_t* Msw = M + (size_t)s * sx;
                for (uint32_t t = 0; t < kopp; ++t) {
                    const int kk = (int)hist[t * 32];
                    if (kk == 0) continue;
                    atomicSub(&Mrw[(size_t)t * st], kk);
                    atomicAdd(&Msw[(size_t)t * st], kk);
                }
                atomicSub(&E[own_off + r], (int)d);
                atomicAdd(&E[own_off + s], (int)d);
                LAB[(size_t)v * C] = (int32_t)s;
                ++n_acc;
                ds_sum += dS;
            }
        }
    }
    if (live) {
        if (n_acc) atomicAdd(&P.accepted[c], n_acc);
        if (ds_sum != 0.0) atomicAdd(&P.dS_accum[c], ds_sum);
    }
}

// ---- state construction (init_bisbm: compute_n_r / compute_m / compute_m_r / compute_eta_rk,
//      reference src/blockmodel.cc:681-746).  lane = chain, one warp per vertex. ----
__global__ void build_counts_kernel(GraphView G, StateView S, uint32_t n_chains) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wpc = blockDim.x >> 5;
    const uint32_t n_groups = S.C / 32;
    const uint64_t gw = (uint64_t)blockIdx.x * wpc + (threadIdx.x >> 5);
    const uint32_t group = gw % n_groups;
    const uint64_t v64 = gw / n_groups;
    if (v64 >= G.n) return;
    const uint32_t v = (uint32_t)v64;
    const uint32_t c = group * 32 + lane;
    if (c >= n_chains) return;
    const uint32_t KA = S.KA, KB = S.KB, W = S.W;
    const bool tb = v >= G.na;
    const uint32_t b = (uint32_t)S.labels[(size_t)v * S.C + c];
    const uint32_t slot = (tb ? KA : 0) + b;
    atomicAdd(&S.nr[(size_t)c * (KA + KB) + slot], 1);
    atomicAdd(&S.eta[((size_t)c * (KA + KB) + slot) * W + G.degidx[v]], 1);
    if (!tb) {
        int32_t* M = S.m + (size_t)c * KA * KB + (size_t)b * KB;
        for (uint32_t e = G.row_ptr[v]; e < G.row_ptr[v + 1]; ++e) {
            const uint32_t t = (uint32_t)S.labels[(size_t)G.col[e] * S.C + c];
            atomicAdd(&M[t], 1);
        }
    }
}

__global__ void build_e_kernel(StateView S, uint32_t n_chains) {  // compute_m_r
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t KA = S.KA, KB = S.KB;
    if (idx >= n_chains * (KA + KB)) return;
    const uint32_t c = idx / (KA + KB), slot = idx % (KA + KB);
    const int32_t* M = S.m + (size_t)c * KA * KB;
    int64_t sum = 0;
    if (slot < KA) for (uint32_t b = 0; b < KB; ++b) sum += M[(size_t)slot * KB + b];
    else for (uint32_t a = 0; a < KA; ++a) sum += M[(size_t)a * KB + (slot - KA)];
    S.e[(size_t)c * (KA + KB) + slot] = (int32_t)sum;
}

// parallel-mode --randomize: per chain, permute the labels of each type with a keyed
// Feistel permutation (keeps block sizes, like shuffle_bisbm)
__global__ void randomize_kernel(GraphView G, const int32_t* in, int32_t* out, uint32_t C, uint32_t n_chains,
                                 const uint64_t* seeds, uint32_t hb_a, uint32_t hb_b) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)G.n * C) return;
    const uint32_t v = (uint32_t)(idx / C), c = (uint32_t)(idx % C);
    if (c >= n_chains) { out[idx] = in[idx]; return; }
    const bool tb = v >= G.na;
    const uint32_t v0 = tb ? G.na : 0, nv = tb ? G.nb : G.na;
    const uint64_t key = seeds[c] * 0x9E3779B97F4A7C15ull + (tb ? 0x632BE59BD9B4E019ull : 0x2545F4914F6CDD1Dull);
    const uint32_t src = v0 + feistel_perm(v - v0, nv, tb ? hb_b : hb_a, key);
    out[idx] = in[(size_t)src * C + c];
}


// ---- label import / export: host layout [chain][node] with GLOBAL block ids  <->  device
//      layout [node][C] chain-minor, type-local.  32x32 tiles through shared memory so both
//      sides are coalesced.  `bad` receives 1 + (chain * n + node) of the first invalid label. ----
__global__ void import_labels_kernel(const uint32_t* __restrict__ in, int32_t* __restrict__ out, uint32_t n,
                                     uint32_t na, uint32_t n_chains, uint32_t C, const uint32_t* __restrict__ ka,
                                     const uint32_t* __restrict__ kb, unsigned long long* bad) {
    __shared__ uint32_t tile[32][33];
    const uint32_t v0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint32_t c = c0 + j, v = v0 + threadIdx.x;
        tile[j][threadIdx.x] = (c < n_chains && v < n) ? in[(size_t)c * n + v] : 0u;
    }
    __syncthreads();
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint32_t v = v0 + j, c = c0 + threadIdx.x;
        if (v >= n) continue;
        int32_t l = 0;
        if (c < n_chains) {
            const uint32_t g = tile[threadIdx.x][j];
            const uint32_t kac = ka[c], kbc = kb[c];
            bool ok;
            if (v < na) { ok = g < kac; l = (int32_t)g; }
            else { ok = (g >= kac) && (g < kac + kbc); l = (int32_t)(g - kac); }
            if (!ok) { atomicMin(bad, 1ull + (unsigned long long)c * n + v); l = 0; }
        }
        out[(size_t)v * C + c] = l;
    }
}

__global__ void export_labels_kernel(const int32_t* __restrict__ in, uint32_t* __restrict__ out, uint32_t n,
                                     uint32_t na, uint32_t n_chains, uint32_t C, const uint32_t* __restrict__ ka) {
    __shared__ uint32_t tile[32][33];
    const uint32_t v0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint32_t v = v0 + j, c = c0 + threadIdx.x;
        uint32_t g = 0;
        if (v < n && c < n_chains) {
            const uint32_t l = (uint32_t)in[(size_t)v * C + c];
            g = v < na ? l : ka[c] + l;
        }
        tile[j][threadIdx.x] = g;
    }
    __syncthreads();
    for (uint32_t j = threadIdx.y; j < 32; j += blockDim.y) {
        const uint32_t c = c0 + j, v = v0 + threadIdx.x;
        if (c < n_chains && v < n) out[(size_t)c * n + v] = tile[threadIdx.x][j];
    }
}

// per-sweep bookkeeping of anneal (src/metropolis_hasting.cc:86-98) at sweep granularity
__global__ void bookkeep_kernel(uint32_t n_chains, uint8_t* active, const double* dS_accum, double* ent_min,
                                unsigned long long* u, unsigned long long* sweeps_done, uint64_t sweep,
                                uint64_t cold_steps, uint64_t steps_await, uint32_t* n_active) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chains || !active[c]) return;
    const double ent = dS_accum[c];
    if (ent < ent_min[c]) { ent_min[c] = ent; u[c] = 0; }
    else u[c] += cold_steps;
    sweeps_done[c] = sweep + 1;
    if (u[c] >= steps_await) { active[c] = 0; atomicSub(n_active, 1u); }
}

// marginal accumulation: hist[v][g] += #chains with global label g at v
__global__ void marginal_kernel(GraphView G, StateView S, uint32_t n_chains, uint32_t* hist, uint32_t width) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (uint64_t)G.n * S.C) return;
    const uint32_t v = (uint32_t)(idx / S.C), c = (uint32_t)(idx % S.C);
    if (c >= n_chains) return;
    const uint32_t l = (uint32_t)S.labels[idx];
    const uint32_t g = v < G.na ? l : S.ka[c] + l;
    atomicAdd(&hist[(size_t)v * width + g], 1u);
}

__global__ void marginal_argmax_kernel(uint32_t n, const uint32_t* hist, uint32_t width, uint32_t* out) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    uint32_t best = 0, bc = 0;
    for (uint32_t g = 0; g < width; ++g) {
        const uint32_t x = hist[(size_t)v * width + g];
        if (x > bc) { bc = x; best = g; }
    }
    out[v] = best;
}

// blockmodel_t::entropy (src/blockmodel.cc:753-787) for every chain: one CTA per chain,
// fixed-order block reduction.  `base` holds the label-independent terms
// -sum_v lgamma(d_v+1) + sum_{i>j, A_ij>1} lgamma(A_ij+1), computed once per graph.
__global__ void entropy_kernel(GraphView G, StateView S, Tables tb, double base, uint32_t n_chains, double* out) {
    const uint32_t c = blockIdx.x;
    if (c >= n_chains) return;
    const uint32_t KA = S.KA, KB = S.KB, W = S.W;
    const uint32_t ka = S.ka[c], kb = S.kb[c];
    const int32_t* M = S.m + (size_t)c * KA * KB;
    const int32_t* E = S.e + (size_t)c * (KA + KB);
    const int32_t* NR = S.nr + (size_t)c * (KA + KB);
    const int32_t* ETA = S.eta + (size_t)c * (KA + KB) * W;
    double acc = 0.0;
    for (uint32_t i = threadIdx.x; i < ka * kb; i += blockDim.x) {
        const uint32_t a = i / kb, b = i % kb;
        acc -= lgamma((double)M[(size_t)a * KB + b] + 1.0);
    }
    for (uint32_t i = threadIdx.x; i < (ka + kb) * W; i += blockDim.x) {
        const uint32_t q = i / W, w = i % W;
        const uint32_t slot = q < ka ? q : KA + (q - ka);
        acc -= lgamma((double)ETA[(size_t)slot * W + w] + 1.0);
    }
    for (uint32_t q = threadIdx.x; q < ka + kb; q += blockDim.x) {
        const uint32_t slot = q < ka ? q : KA + (q - ka);
        acc += lgamma((double)E[slot] + 1.0);
        acc += log_q(tb, E[slot], NR[slot]);
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double ent = base + red[0];
        const double Ed = (double)G.n_edges, na = (double)G.na, nb = (double)G.nb;
        const double kab = (double)ka * (double)kb;
        // lbinom_fast(N, k) = lgamma(N+1) - lgamma(k+1) - lgamma(N-k+1), 0 if N==0, k==0 or k>N
        {
            const double N = kab + Ed - 1.0, k = Ed;
            if (!(N == 0.0 || k == 0.0 || k > N)) ent += lgamma(N + 1.0) - lgamma(k + 1.0) - lgamma(N - k + 1.0);
        }
        {
            const double N = na - 1.0, k = (double)ka - 1.0;
            if (!(N == 0.0 || k == 0.0 || k > N)) ent += lgamma(N + 1.0) - lgamma(k + 1.0) - lgamma(N - k + 1.0);
        }
        {
            const double N = nb - 1.0, k = (double)kb - 1.0;
            if (!(N == 0.0 || k == 0.0 || k > N)) ent += lgamma(N + 1.0) - lgamma(k + 1.0) - lgamma(N - k + 1.0);
        }
        ent += (na * nb == 0.0) ? 0.0 : log(na * nb);
        ent += lgamma(na + 1.0);
        ent += lgamma(nb + 1.0);
        out[c] = ent;
    }
}

#endif  // __CUDACC__

}  // namespace bisbm

// sweep_fast.cuh -- the throughput variant of the parallel sweep (same mapping as sweep.cuh:
// lane = chain, one warp per vertex, counts staged in shared memory), with the per-move
// arithmetic in fp32 and the special functions on the MUFU unit.
//
// Why fp32 is enough here.  The only consumer of dS and of the Hastings factor is the accept test
//     a = log(accu1/accu0) - dS/T;   accept iff a > 0 or U < exp(a)        (reference step(),
// src/metropolis_hasting.cc:42-62), so an absolute error delta in `a` changes one acceptance
// probability by the factor exp(delta).  The fp32 evaluation below has |delta| <~ 2e-5 (products of
// at most four exactly converted integers between logarithms, lg2/ex2/rcp.approx at 2^-22 relative;
// checked against the reference's transition_ratio known answers in tests/test_emul.py), far below
// the count staleness parallel mode already accepts and below what any test with a
// feasible number of samples can resolve.  Everything that must be exact stays exact: the counts
// are int32, the commit is integer atomics, the "would empty block r" veto is an exact atomic, and
// the accumulated dS of a chain is summed in double.  Rare cases outside the fast formulas' range
// (small blocks: exact log q table, Stirling form of the e_r terms) take the double-precision
// routines of sweep.cuh, out of line.  Sequential replay (replay.cuh) is untouched: strict double.
//
// Per move this is still the arithmetic of the reference's step():
//   proposal   single_vertex_change      reference src/blockmodel.cc:613-637
//   dS, accu_r transition_ratio          reference src/metropolis_hasting.cc:103-192
//   accept     step                      reference src/metropolis_hasting.cc:42-62
//   commit     apply_mcmc_moves          reference src/blockmodel.cc:461-503
#pragma once
#include "sweep.cuh"

namespace bisbm {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float f_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float f_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
inline float f_lg2(float x) { return log2f(x); }
inline float f_ex2(float x) { return exp2f(x); }
inline float f_rcp(float x) { return 1.0f / x; }
#endif

#define BISBM_LN2F 0.69314718055994530942f
#define BISBM_LOG2EF 1.44269504088896340736f

// running sums of the one-pass evaluation of transition_ratio (see the header of sweep.cuh):
//   a0 = sum_edges m_st * inv_t,   a1 = sum_edges (m_rt - 2c - 1) * inv_t,   w = sum_edges inv_t
//   (accu0 = a0 + eps w, accu1 = a1 + eps w),   num/den = running products of (m_rt - c), (m_st + 1 + c),
//   lg = log2 of the products folded so far
struct FAcc {
    float a0, a1, w, num, den, lg;
};
BISBM_HD void facc_init(FAcc& A) { A.a0 = 0.f; A.a1 = 0.f; A.w = 0.f; A.num = 1.f; A.den = 1.f; A.lg = 0.f; }
// one edge of v into block t: c earlier edges of v into t (c1 = c + 1 <= 255), counts m_rt, m_st, inv = 1/(e_t + eps K).
// Two int->float conversions per edge (the XU pipe converts one warp per clock per SM); c + 1 becomes a float by
// splicing its 8 bits under the exponent of 2^23, and the remaining terms are float additions (exact below 2^24).
BISBM_HD void facc_edge(FAcc& A, int m_r, int m_s, int c1, float inv) {
    const float fr = (float)m_r, fs = (float)m_s;
#ifdef __CUDA_ARCH__
    const float fc1 = __uint_as_float(0x4B000000u | (uint32_t)c1) - 8388608.0f;
#else
    const float fc1 = (float)c1;
#endif
    const float x2 = fr - (fc1 + fc1);        // m_rt - 2c - 1 = (m_rt + 1) - 2 (c + 1)
    A.a0 = fmaf(fs, inv, A.a0);
    A.a1 = fmaf(x2 + 1.0f, inv, A.a1);
    A.w += inv;
    A.num *= (fr + 1.0f) - fc1;               // m_rt - c
    A.den *= fs + fc1;                        // m_st + 1 + c
}
// fold the products into the logarithm; called at least every 4 edges (4 factors < 2^31 stay inside fp32 range)
BISBM_HD void facc_fold(FAcc& A) {
    A.lg += f_lg2(A.num * f_rcp(A.den));
    A.num = 1.f; A.den = 1.f;
}

// fp32 form of block_degree_delta (sweep.cuh): D log(cs/cr) - k3 (1/cs^2 - 1/cr^2); the next term is
// < D / 3.4e8 when both midpoints are >= 32 D.  *ok = false -> caller takes the double routine.
BISBM_HD float f_block_degree_delta(int e_r, int e_s, int d, bool* ok) {
    const float D = (float)d;
    const float cs = (float)e_s + 0.5f * (D + 1.0f), cr = (float)e_r - 0.5f * (D - 1.0f);
    *ok = (cs >= 32.0f * D) && (cr >= 32.0f * D);
    const float is = f_rcp(cs), ir = f_rcp(cr);
    const float k3 = D * (D * D - 1.0f) * (1.0f / 24.0f);
    return D * BISBM_LN2F * f_lg2(cs * ir) - k3 * (is * is - ir * ir);
}

// fp32 form of logq_delta (sweep.cuh): second-order expansion about (e0, n0); *ok = false when the
// block has no expansion or has drifted out of its range
BISBM_HD float f_logq_delta(const LogqExp& q, int e, int n, int de, int dn, bool* ok) {
    const int x = e - q.e0, y = n - q.n0;
    const int ax = x < 0 ? -x : x, ay = y < 0 ? -y : y, ad = de < 0 ? -de : de;
    const int re = q.e0 >> 4, rn = q.n0 >> 4;
    *ok = (q.valid != 0u) && ax <= re && ay <= rn && ad <= re;
    const float dx = (float)x, dy = (float)y, De = (float)de, Dn = (float)dn;
    return q.fe * De + q.fn * Dn + 0.5f * q.fee * (De * De + 2.0f * dx * De) + q.fen * (dx * Dn + dy * De + De * Dn) +
           0.5f * q.fnn * (Dn * Dn + 2.0f * dy * Dn);
}

#ifdef __CUDACC__

// shared memory of the fast kernel, in this order (KA, KB, kopp/kown = padded maxima):
//   int32  sM  [KA*KB*32]      m_rs of the group, [a][b][lane]
//   int32  sEo [kown*32]       e_r of the moving type's blocks
//   int32  sEp [kopp*32]       e_t of the frozen type's blocks
//   uint4  batch[warps][32]    ring of the warp's next vertices: {vertex, CSR row offset, degree, degree index}
//   u8     hist[warps][ceil(kopp/4)][32 lanes][4]   (bin t of lane l: word t/4, byte t%4 -- lane l only touches bank l)
__host__ __device__ inline size_t sweep_fast_smem_bytes(uint32_t KA, uint32_t KB, uint32_t type, uint32_t warps) {
    const uint32_t kown = type ? KB : KA, kopp = type ? KA : KB;
    return (size_t)KA * KB * 128 + (size_t)kown * 128 + (size_t)kopp * 128 + (size_t)warps * 512 +
           (size_t)warps * ((kopp + 3) / 4) * 128;
}

__device__ __forceinline__ float sh_ld_f32(uint32_t a) { float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t sh_ld_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sh_st_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t sh_ld_u32v(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sh_st_u32v(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 sh_ld_v4(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sh_st_v4(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// the rare double-precision completions, out of line (arguments by value: nothing of the kernel's parameter
// block has its address taken)
__device__ __noinline__ static float slow_block_degree_delta(int e_r, int e_s, int d) {
    return (float)block_degree_delta(e_r, e_s, d);
}
__device__ __noinline__ static float slow_logq_delta(const double* qtab, uint32_t qn, uint32_t qk, int e, int n, int de, int dn) {
    Tables tb; tb.lg = nullptr; tb.lg_n = 0; tb.qtab = qtab; tb.qn = qn; tb.qk = qk;
    return (float)logq_delta_exact(tb, e, n, de, dn);
}
// 1/T of global step t as fp32; T == 0 is returned as a negative value
__device__ __noinline__ static float slow_beta(int schedule, float p0, float p1, uint64_t t) {
    const double T = par_temperature(schedule, p0, p1, t);
    return (T == 0.0) ? -1.f : (float)(1.0 / T);
}
// commit one histogram bin: m(r,t) -= k, m(s,t) += k.  Unconditional: with 32 chains per warp nearly every bin is
// non-zero in some lane, so a test per bin only adds branches; lanes that do not move pass k = 0.
__device__ __forceinline__ void commit_bin(uint32_t ar, uint32_t as, int k) {
    sh_red_add(ar, -k);
    sh_red_add(as, k);
}

// KF > 0: Ka == Kb == KF padded strides and the moving type TYPE fixed at compile time; KF == 0: from SweepParams.
// NT threads per CTA (one CTA per SM: 1024 -> 64 registers per thread, 768 -> 80, 512 -> 128).
// Preconditions (plan_sweep): max degree <= 255 (u8 histogram bins), K per type <= 256 (u8 labels).
template <int KF, int TYPE, int NT>
__global__ void __launch_bounds__(NT, 1) sweep_fast_kernel(const __grid_constant__ SweepParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t lane, warp;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(warp));
    warp >>= 5;
    const uint32_t wpc = blockDim.x >> 5;
    const GraphView& G = P.g;
    const uint32_t type = KF ? (uint32_t)TYPE : P.type;
    const uint32_t C = P.s.C, KB = KF ? (uint32_t)KF : P.s.KB, KA = KF ? (uint32_t)KF : P.s.KA, W = P.s.W, KK = KA + KB;
    const uint32_t kown_max = type ? KB : KA, kopp_max = type ? KA : KB;
    const uint32_t group = blockIdx.x % P.n_groups;
    const uint32_t cta_in_group = blockIdx.x / P.n_groups;
    const uint32_t own_off = type ? KA : 0, opp_off = type ? 0 : KA;

    int32_t* const gM = P.s.m + (size_t)group * KA * KB * GROUP;
    int32_t* const gE = P.s.e + (size_t)group * KK * GROUP;
    int32_t* const gNR = P.s.nr + (size_t)group * KK * GROUP + (size_t)own_off * 32;
    int32_t* const gETA = P.s.eta + (size_t)group * KK * W * GROUP + (size_t)own_off * W * 32;
    const uint32_t* const gLQ = P.lq_soa + ((size_t)group * KK + own_off) * 8 * GROUP;   // [slot][field][lane]

    const uint32_t c = group * 32 + lane;
    const bool live = (c < P.n_chains) && P.active[c];
    const uint32_t cc = (c < P.s.C) ? c : 0;
    const uint32_t ka = P.s.ka[cc], kb = P.s.kb[cc], K = ka + kb;
    const uint32_t kown = type ? kb : ka;
    const float eps = (float)P.s.eps;
    const float epsK32 = (float)(P.s.eps * (double)K) * 4294967296.0f;   // threshold scale of the uniform-vs-categorical test

    // ---- stage the group's counts ----
    int32_t* const sM = reinterpret_cast<int32_t*>(smem_raw);
    int32_t* const sEo = sM + KA * KB * 32;
    int32_t* const sEp = sEo + kown_max * 32;
    uint32_t* const sBatch = reinterpret_cast<uint32_t*>(sEp + kopp_max * 32);
    uint32_t* const hist_all = sBatch + wpc * 128;
    const uint32_t hist_words = (kopp_max + 3u) / 4u;
    copy_i4(sM, gM, KA * KB * 32);
    copy_i4(sEo, gE + own_off * 32, kown_max * 32);
    copy_i4(sEp, gE + opp_off * 32, kopp_max * 32);
    for (uint32_t i = threadIdx.x; i < wpc * hist_words * 32u; i += blockDim.x) hist_all[i] = 0u;
    __syncthreads();

    // lane-private shared byte addresses (entry j of this chain at +j*128)
    const uint32_t lane4 = lane * 4u;
    const uint32_t M_base = (uint32_t)__cvta_generic_to_shared(sM) + lane4;
    const uint32_t Eo_base = (uint32_t)__cvta_generic_to_shared(sEo) + lane4;
    const uint32_t Ep_base = (uint32_t)__cvta_generic_to_shared(sEp) + lane4;
    const uint32_t batch_base = (uint32_t)__cvta_generic_to_shared(sBatch) + warp * 512u;
    const float epsK = (float)(P.s.eps * (double)K);
    const uint32_t hist_base = (uint32_t)__cvta_generic_to_shared(hist_all) + warp * hist_words * 128u + lane4;
    // m(x_own, t_opp) at M_base + x*SX + t*ST  (bytes)
    const uint32_t SX = (type ? 1u : KB) * 128u, ST = (type ? KB : 1u) * 128u;
    // the lane's label column as an opaque global address (kept in a register pair, not re-derived)
    uint64_t LAB8 = (uint64_t)__cvta_generic_to_global(P.lab8 + cc);
    asm volatile("" : "+l"(LAB8));
    auto lab_ld = [&](uint32_t vtx) -> uint32_t {     // label of vertex vtx in this lane's chain
        uint32_t x;
        // .cg: label rows have no reuse in L1; keeping them out leaves the L1 to the log q expansions and the CSR rows
        asm("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %1, %2, %3;\n\tld.global.cg.u8 %0, [a];\n\t}" : "=r"(x) : "r"(vtx), "r"(C), "l"(LAB8));
        return x;
    };
    auto lab_st = [&](uint32_t vtx, uint32_t x) {
        asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %0, %1, %2;\n\tst.global.u8 [a], %3;\n\t}" :: "r"(vtx), "r"(C), "l"(LAB8), "r"(x) : "memory");
    };
    const uint32_t rot4 = (lane + 4u) & 31u;
    const uint32_t key0 = (uint32_t)P.seeds[cc], key1 = (uint32_t)(P.seeds[cc] >> 32);
    const uint32_t nv = type ? G.nb : G.na, v0 = type ? G.na : 0;
    constexpr int SAFE_NR = 8192;   // > any number of concurrently evaluated moves of one chain (<= 148 SMs * 32 warps)
    const bool const_T = (P.schedule == 3);
    const bool movable_chain = live && (kown != 1);

    uint32_t n_acc = 0;
    double ds_sum = 0.0;

    if (warp < P.warps_used) {
        const uint32_t stride = P.ctas_per_group * P.warps_used;
        const uint32_t i_first = P.pos_begin + cta_in_group * P.warps_used + warp;
        // The warp's vertices are prepared 32 at a time: lane l computes the l-th one (place in the permuted
        // visiting order, CSR row offset, degree) and parks it in a 64-slot ring in shared memory.
        // While vertex k is evaluated the warp loads the neighbour ids of vertex k+2 (one coalesced load, lane e =
        // neighbour e) and issues one L2 prefetch per lane for the label rows vertex k+1 will gather (lane e -> row
        // of neighbour e; lanes past the degree -> vertex k+1's own label).  Pipeline state: two registers.
        // (Prefetching into L1 was measured useless: with 190 KB of shared memory the L1 holds ~500 lines and every
        // prefetched 32-byte row occupies a whole line.  See DESIGN.md for the other things that were tried.)
        auto refill = [&](uint32_t k0) {     // slots k0 .. k0+15 of the ring, prepared by lanes 0..15
            const uint64_t il = (uint64_t)i_first + (uint64_t)(k0 + lane) * stride;
            uint4 b; b.x = 0; b.y = 0; b.z = 0; b.w = 0;
            if (lane < 16u && il < P.pos_end) {
                const uint64_t pkey = (P.sweep * 2 + type) * 0x9E3779B97F4A7C15ull + (uint64_t)group * 0xD1B54A32D192ED03ull;
                b.x = v0 + feistel_perm((uint32_t)il, nv, P.half_bits, pkey);
                b.y = G.row_ptr[b.x];
                b.z = G.row_ptr[b.x + 1] - b.y;
                b.w = G.degidx[b.x];
            }
            __syncwarp();
            if (lane < 16u) sh_st_v4(batch_base + ((k0 + lane) & 31u) * 16u, b);
            __syncwarp();
        };
        auto load_nbr = [&](uint32_t k) -> uint32_t {   // neighbour ids of the vertex in slot k (lane e = neighbour e, 0 past the degree)
            const uint4 b = sh_ld_v4(batch_base + (k & 31u) * 16u);
            return (lane < b.z) ? G.col[b.y + lane] : 0u;
        };
        auto prefetch_rows = [&](uint32_t k, uint32_t nbr) {   // label rows the vertex in slot k will gather
            const uint4 b = sh_ld_v4(batch_base + (k & 31u) * 16u);
            const uint32_t vtx = (lane < b.z) ? nbr : b.x;     // lanes past the degree: the vertex's own label
            asm volatile("{\n\t.reg .u64 a;\n\tmad.wide.u32 a, %0, %1, %2;\n\tprefetch.global.L2 [a];\n\t}" :: "r"(vtx), "r"(C), "l"(LAB8));
        };
        uint32_t nbr1 = 0, nbr2 = 0, nbr0 = 0;
        if (i_first < P.pos_end) {
            refill(0);
            refill(16);
            nbr0 = load_nbr(0);
            nbr1 = load_nbr(1);
            prefetch_rows(0, nbr0);
        }
        for (uint32_t ib = i_first, k = 0; ib < P.pos_end; ib += stride, ++k, nbr0 = nbr1, nbr1 = nbr2) {
            if (k >= 16u && (k & 15u) == 0) refill(k + 16u);   // the ring half holding slots k-16 .. k-1 is free
            const uint4 cur = sh_ld_v4(batch_base + (k & 31u) * 16u);
            const uint32_t v = cur.x, d = cur.z, didx = cur.w;
            nbr2 = load_nbr(k + 2);
            prefetch_rows(k + 1, nbr1);
            // ---- the draw of this move: Philox4x32-10 ----
            uint32_t ry, rz, rw, jv;
            {
                // counter = (vertex, sweep, chain seed); the key is common to the pool, so the ten round keys are
                // warp-uniform constants instead of twenty per-lane registers
                u32x4 ctr; ctr.x = v; ctr.y = (uint32_t)P.sweep; ctr.z = key0; ctr.w = key1;
                const u32x4 ra = philox4x32(ctr, (uint32_t)(P.sweep >> 32) ^ 0xA4093822u, 0x299F31D0u);
                ry = ra.y; rz = ra.z; rw = ra.w;
                const uint32_t e = mulhi32(ra.x, d);       // the proposal's random neighbour (differs per lane)
                jv = __shfl_sync(0xffffffffu, nbr0, e & 31u);
                if (d > 32u) jv = G.col[cur.y + e];
                if (d == 0u) jv = v;                      // isolated vertex: no neighbour, any valid label
            }
            const uint32_t r = lab_ld(v);
            const uint32_t tq = min(lab_ld(jv), kopp_max - 1u);
            float beta = 1.0f / P.p0;
            if (!const_T) beta = slow_beta(P.schedule, P.p0, P.p1, P.step_base + ib);
            const bool T_zero = beta < 0.f;

            // ---- proposal (single_vertex_change), branch-free ----
            const int e_t = sh_ld(Ep_base + tq * 128u);
            const float inv_t = f_rcp((float)e_t + epsK);
            // U < eps K / (e_t + eps K), both sides scaled by 2^32 (the conversion saturates at 2^32 - 1)
            const bool uniform_pick = (d == 0) || (ry < __float2uint_rz(epsK32 * inv_t));
            const uint32_t sg = mulhi32(rz, K);      // uniform over ALL K blocks (either type)
            const bool sg_a = sg < ka;
            const uint32_t s_uni = sg_a ? sg : sg - ka;
            // categorical over row m[t][.]: s = #{x : cum_x <= z}; wz tracks cum - z - 1 (negative while cum <= z)
            int wz = -(int)mulhi32(rz, (uint32_t)e_t) - 1;
            uint32_t cnt_le = 0;
            {
                uint32_t a = M_base + tq * ST;
#pragma unroll 8
                for (uint32_t x = 0; x < kown_max; ++x, a += SX) {   // uniform bound; blocks >= kown hold 0
                    wz += sh_ld(a);
                    cnt_le += ((uint32_t)wz) >> 31;
                }
            }
            const uint32_t s_cat = cnt_le < kown ? cnt_le : kown - 1;
            // a uniform draw that falls on a block of the other type is rejected (dS = +inf): s stays r for it
            const bool cross = movable_chain && uniform_pick && (sg_a != (type == 0));
            const uint32_t s = (movable_chain && !cross) ? (uniform_pick ? s_uni : s_cat) : r;   // own-type local index
            const bool eval = live && (s != r);
            if (!__any_sync(0xffffffffu, eval)) {
                // s == r (dS = 0, accu_r = 1): accepted at T > 0 unless the block would empty, rejected at T == 0
                // (src/metropolis_hasting.cc:47-52)
                if (live && !cross && !T_zero && ldc(&gNR[r * 32 + lane]) != 1) ++n_acc;
                continue;
            }

            // ---- one pass over v's neighbours: dS and the Hastings factor (transition_ratio).  Every lane
            //      runs it (masked lanes cost the same issue slots); only `eval` lanes may commit. ----
            FAcc A; facc_init(A);
            const uint32_t Mr = M_base + r * SX, Ms = M_base + s * SX;
            // lane e holds the id of neighbour (block of 32) + e, rotated so that the next 4 to gather sit in lanes 0..3
            uint32_t nbr = nbr0;
            uint32_t t0, t1, t2, t3;
            auto gather4 = [&](uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) {
                const uint32_t i0 = __shfl_sync(0xffffffffu, nbr, 0), i1 = __shfl_sync(0xffffffffu, nbr, 1);
                const uint32_t i2 = __shfl_sync(0xffffffffu, nbr, 2), i3 = __shfl_sync(0xffffffffu, nbr, 3);
                nbr = __shfl_sync(0xffffffffu, nbr, rot4);
                // (ids past the degree are 0: a valid, unused load)
                a0 = lab_ld(i0); a1 = lab_ld(i1); a2 = lab_ld(i2); a3 = lab_ld(i3);
            };
            auto consume = [&](uint32_t t) {
                const uint32_t ha = (t << 5) + ((t & 3u) * 0xffffffe1u + hist_base);   // (t/4)*128 + t%4 = 32 t - 31 (t%4)
                const uint32_t cnt = sh_ld_u8(ha);
                const uint32_t cnt1 = cnt + 1u;
                sh_st_u8(ha, cnt1);
                const int m_r = sh_ld(Mr + t * ST), m_s = sh_ld(Ms + t * ST);
                facc_edge(A, m_r, m_s, (int)cnt1, f_rcp((float)sh_ld(Ep_base + t * 128u) + epsK));
            };
            const uint32_t n_ch = (d + 3u) >> 2;
            if (n_ch) gather4(t0, t1, t2, t3);
            for (uint32_t ch = 0, rem = d; ch < n_ch; ++ch, rem -= 4u) {
                uint32_t u0 = 0, u1 = 0, u2 = 0, u3 = 0;
                if (ch + 1 < n_ch) {   // next chunk's labels in flight while this one is consumed
                    if (((ch + 1) & 7u) == 0) {
                        const uint32_t e0 = ch * 4u + 4u;
                        nbr = (e0 + lane < d) ? G.col[cur.y + e0 + lane] : 0u;
                    }
                    gather4(u0, u1, u2, u3);
                }
                if (rem >= 4u) {
                    consume(t0); consume(t1); consume(t2); consume(t3);
                } else {
                    consume(t0);
                    if (rem > 1u) consume(t1);
                    if (rem > 2u) consume(t2);
                }
                facc_fold(A);
                t0 = u0; t1 = u1; t2 = u2; t3 = u3;
            }
            // the global (L2) counts of the two blocks
            const int n_r = ldc(&gNR[r * 32 + lane]);
            const int n_s = ldc(&gNR[s * 32 + lane]);
            const int eta_r = ldc(&gETA[(r * W + didx) * 32 + lane]);
            const int eta_s = ldc(&gETA[(s * W + didx) * 32 + lane]);
            __syncwarp();

            // ---- dS, accept (step) ----
            bool go;
            float dS;
            {
                const int e_r = sh_ld(Eo_base + r * 128u), e_s = sh_ld(Eo_base + s * 128u);
                bool ok_b, ok_r, ok_s;
                float bdd = f_block_degree_delta(e_r, e_s, (int)d, &ok_b);
                float lqr, lqs;
                auto load_q = [&](uint32_t slot) -> LogqExp {   // seven coalesced 4-byte loads (constant during the launch)
                    const uint32_t* p = gLQ + slot * 256u + lane;
                    LogqExp q;
                    q.e0 = (int)__ldg(p); q.n0 = (int)__ldg(p + 32);
                    q.fe = __uint_as_float(__ldg(p + 64)); q.fn = __uint_as_float(__ldg(p + 96));
                    q.fee = __uint_as_float(__ldg(p + 128)); q.fen = __uint_as_float(__ldg(p + 160));
                    q.fnn = __uint_as_float(__ldg(p + 192));
                    q.valid = (uint32_t)q.e0;     // blocks without expansion are stored with e0 = 0
                    return q;
                };
                lqr = f_logq_delta(load_q(r), e_r, n_r, -(int)d, -1, &ok_r);
                lqs = f_logq_delta(load_q(s), e_s, n_s, (int)d, 1, &ok_s);
                if (__any_sync(0xffffffffu, eval && !(ok_b && ok_r && ok_s))) {
                    if (eval && !ok_b) bdd = slow_block_degree_delta(e_r, e_s, (int)d);
                    if (eval && !ok_r) lqr = slow_logq_delta(P.tb.qtab, P.tb.qn, P.tb.qk, e_r, n_r, -(int)d, -1);
                    if (eval && !ok_s) lqs = slow_logq_delta(P.tb.qtab, P.tb.qn, P.tb.qk, e_s, n_s, (int)d, 1);
                    __syncwarp();
                }
                // log( prod (m_rt-c)/(m_st+1+c) * eta_r/(eta_s+1) ) + e_r terms + log q terms
                dS = BISBM_LN2F * (A.lg + f_lg2((float)(eta_r > 0 ? eta_r : 1) * f_rcp((float)(eta_s + 1))));
                dS += ((d == 0) ? 0.f : bdd) + lqr + lqs;
                // log2 of the Hastings factor (accu1 / accu0), 0 for an isolated vertex
                const float ew = eps * A.w;
                const float lh = (d == 0) ? 0.f : f_lg2((A.a1 + ew) * f_rcp(A.a0 + ew));
                const float a2 = lh - dS * (beta * BISBM_LOG2EF);          // log2 of the acceptance ratio
                const float u = ((float)(rw >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const bool go_hot = (a2 > 0.f) || (u < f_ex2(a2));
                go = eval && (T_zero ? (dS < 0.f) : go_hot);
                // s == r: see above
                if (live && !cross && s == r && !T_zero && n_r != 1) ++n_acc;
                if (go) {
                    // the exact "would empty block r" veto of apply_mcmc_moves.  n_r was read from L2 a moment
                    // ago; fewer than SAFE_NR moves of this chain can be in flight, so a block that large
                    // cannot empty and a fire-and-forget reduction is enough
                    if (n_r > SAFE_NR) {
                        atomicSub(&gNR[r * 32 + lane], 1);
                    } else {
                        const int old = atomicSub(&gNR[r * 32 + lane], 1);
                        if (old <= 1) { atomicAdd(&gNR[r * 32 + lane], 1); go = false; }
                    }
                }
            }
            __syncwarp();

            // ---- clear the histogram and commit (apply_mcmc_moves): k_t is the histogram ----
            {
                uint32_t ha = hist_base, ar = Mr, as = Ms;
#pragma unroll 2
                for (uint32_t w = 0; w < hist_words; ++w, ha += 128u) {   // uniform bound; bins >= kopp stay 0
                    uint32_t word = sh_ld_u32v(ha);
                    sh_st_u32v(ha, 0u);
                    word = go ? word : 0u;
#pragma unroll
                    for (uint32_t b = 0; b < 4; ++b, ar += ST, as += ST)
                        commit_bin(ar, as, (int)__byte_perm(word, 0u, 0x4440u + b));
                }
                if (go) {
                    sh_red_add(Eo_base + r * 128u, -(int)d);
                    sh_red_add(Eo_base + s * 128u, (int)d);
                    atomicAdd(&gNR[s * 32 + lane], 1);
                    atomicSub(&gETA[(r * W + didx) * 32 + lane], 1);
                    atomicAdd(&gETA[(s * W + didx) * 32 + lane], 1);
                    lab_st(v, s);      // the i32 labels are refreshed from the shadow after the call
                    ++n_acc;
                    ds_sum += (double)dS;
                }
            }
        }
        if (live) {
            if (n_acc) atomicAdd(&P.accepted[c], (unsigned long long)n_acc);
            if (ds_sum != 0.0) atomicAdd(&P.dS_accum[c], ds_sum);
        }
    }

    // ---- publish the staged counts ----
    __syncthreads();
    if (P.exclusive) {
        copy_i4(gM, sM, KA * KB * 32);
        copy_i4(gE + own_off * 32, sEo, kown_max * 32);
    } else {
        int32_t* const nM = P.m_next + (size_t)group * KA * KB * GROUP;
        int32_t* const nE = P.e_next + (size_t)group * KK * GROUP + own_off * 32;
        // (staged - base) of every entry, 16 bytes at a time.  Two neighbouring int32 deltas go out as ONE 64-bit
        // reduction, value d0 + d1 * 2^32 in two's complement: a borrow out of the low word while another CTA's
        // positive delta is still on its way is repaid by that delta's carry, so the sums are exact in any order
        // (every final count is >= 0) -- half the L2 atomics of the end-of-launch burst.
        const int4* const s4 = reinterpret_cast<const int4*>(sM);
        const int4* const g4 = reinterpret_cast<const int4*>(gM);
        unsigned long long* const n8 = reinterpret_cast<unsigned long long*>(nM);
        const uint32_t n4 = KA * KB * 8;
        // four base loads in flight per thread (the loop is a chain of L2 round trips otherwise)
        for (uint32_t i0 = threadIdx.x; i0 < n4; i0 += 4 * blockDim.x) {
            int4 a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * blockDim.x;
                a[u] = (i < n4) ? __ldcg(g4 + i) : make_int4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t i = i0 + u * blockDim.x;
                if (i < n4) {
                    const int4 b = s4[i];
                    const int d0 = b.x - a[u].x, d1 = b.y - a[u].y, d2 = b.z - a[u].z, d3 = b.w - a[u].w;
                    if (d0 | d1) atomicAdd(&n8[2 * i], (unsigned long long)((long long)d0 + ((long long)d1 << 32)));
                    if (d2 | d3) atomicAdd(&n8[2 * i + 1], (unsigned long long)((long long)d2 + ((long long)d3 << 32)));
                }
            }
        }
        for (uint32_t i = threadIdx.x; i < kown_max * 32; i += blockDim.x) {
            const int dlt = sEo[i] - gE[own_off * 32 + i];
            if (dlt) atomicAdd(&nE[i], dlt);
        }
    }
}

#endif  // __CUDACC__

}  // namespace bisbm
